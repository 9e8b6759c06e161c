// K14: SparseLoRA masked merge  W <- (W + scaling * B A) (.) M  in one pass.
//
// Replaces lavis/peft/src/peft/tuners/lora.py:384-387 (Linear.merge, sparse branch)
//   self.weight.data += (B @ A * scaling) * mask        (fp32 math, one rounding to W's dtype)
// fused with the re-mask of train.py:634-637
//   module.weight.data[~module.mask] = 0
// The reference materialises an fp32 [R, C] temporary and makes ~5 passes; here a thread
// owns 16 B of one row, the A columns it needs sit in registers for all rows of its tile,
// the B row is a broadcast load.  HBM-bound: 5 B/weight (read W + mask, write W).
#include "common.cuh"

namespace vlmc {

constexpr int kMergeThreads = 128;
constexpr int kMaxRankRegs = 8;

template <typename T, int RK>  // RK = padded rank held in registers (<= kMaxRankRegs), 0 = generic
__global__ void __launch_bounds__(kMergeThreads)
lora_merge_kernel(T* __restrict__ W, int64_t ldw, int R, int C,
                  const float* __restrict__ A, const float* __restrict__ B, int rank, float scaling,
                  const uint8_t* __restrict__ mask, int64_t ldm, int remask) {
  constexpr int V = Elem<T>::kVec;
  const int col = (blockIdx.x * kMergeThreads + threadIdx.x) * V;
  if (col >= C) return;
  float a[RK > 0 ? RK : 1][V];
  if (RK > 0) {
#pragma unroll
    for (int kk = 0; kk < RK; ++kk)
#pragma unroll
      for (int e = 0; e < V; ++e) a[kk][e] = kk < rank ? A[(int64_t)kk * C + col + e] : 0.f;
  }
  for (int row = blockIdx.y; row < R; row += gridDim.y) {
    T* wp = W + (int64_t)row * ldw + col;
    const uint8_t* mp = mask + (int64_t)row * ldm + col;
    uint4 wv = ld_stream(wp);
    uint32_t mb[2];
    if (V == 8) { uint2 t = *reinterpret_cast<const uint2*>(mp); mb[0] = t.x; mb[1] = t.y; }
    else { mb[0] = *reinterpret_cast<const uint32_t*>(mp); mb[1] = 0; }
    float f[V], acc[V];
    Elem<T>::unpack(wv, f);
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = 0.f;
    const float* brow = B + (int64_t)row * rank;
    if (RK > 0) {
#pragma unroll
      for (int kk = 0; kk < RK; ++kk) {
        const float b = kk < rank ? __ldg(brow + kk) : 0.f;
#pragma unroll
        for (int e = 0; e < V; ++e) acc[e] = fmaf(b, a[kk][e], acc[e]);   // k ascending, like SGEMM
      }
    } else {
      for (int kk = 0; kk < rank; ++kk) {
        const float b = __ldg(brow + kk);
#pragma unroll
        for (int e = 0; e < V; ++e) acc[e] = fmaf(b, __ldg(A + (int64_t)kk * C + col + e), acc[e]);
      }
    }
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const bool keep = (mb[e / 4] >> (8 * (e % 4))) & 0xffu;
      // (BA * scaling) * mask, then W += : two separately rounded fp32 ops, no FMA contraction
      const float delta = __fmul_rn(acc[e], scaling);
      const float merged = __fadd_rn(f[e], delta);
      f[e] = keep ? merged : (remask ? 0.f : f[e]);
    }
    st_stream(wp, Elem<T>::pack(f));
  }
}

// ---- all LoRA linears of a block (or of the model, 16 at a time) in ONE launch -------------------------------------
// train.py:626-637 merges module by module; the matrices are 34-90 MB, so one launch per linear spends a third of its
// time ramping up and draining (measured: 7 launches, 2.6 TB/s).  Same walk as nm_batch_kernel: the linears are one
// list of work units (kMbRows rows x 1024 columns) that a fully resident grid strides over.
constexpr int kMbMax = 16;
constexpr int kMbRows = 64;       // rows per unit: the rank x 8 A values a thread keeps in registers are reloaded per unit
struct MergeBatchItem {
  void* W; int64_t ldw; int R, C; const float* A; const float* B; int rank; float scaling;
  const uint8_t* mask; int64_t ldm; int coltiles;
};
struct MergeBatch { MergeBatchItem it[kMbMax]; int unit_begin[kMbMax + 1]; int count; };

template <typename T>
__global__ void __launch_bounds__(kMergeThreads)
lora_merge_batch_kernel(const __grid_constant__ MergeBatch b, int remask) {
  constexpr int V = Elem<T>::kVec;
  constexpr int RK = kMaxRankRegs;
  const int total = b.unit_begin[b.count];
  for (int unit = blockIdx.x; unit < total; unit += gridDim.x) {
    int p = 0;
    while (unit >= b.unit_begin[p + 1]) ++p;
    const MergeBatchItem& it = b.it[p];
    const int local = unit - b.unit_begin[p];
    const int ct = local % it.coltiles, rb = local / it.coltiles;
    const int col = (ct * kMergeThreads + threadIdx.x) * V;
    if (col >= it.C) continue;
    float a[RK][V];
#pragma unroll
    for (int kk = 0; kk < RK; ++kk)
#pragma unroll
      for (int q = 0; q < V; q += 4) {
        float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kk < it.rank) av = __ldg(reinterpret_cast<const float4*>(it.A + (int64_t)kk * it.C + col + q));
        a[kk][q] = av.x; a[kk][q + 1] = av.y; a[kk][q + 2] = av.z; a[kk][q + 3] = av.w;
      }
    T* W = reinterpret_cast<T*>(it.W);
    const int row_end = (rb + 1) * kMbRows < it.R ? (rb + 1) * kMbRows : it.R;
    constexpr int kRows = 4;                 // rows in flight per thread
    for (int row0 = rb * kMbRows; row0 < row_end; row0 += kRows) {
      uint4 wv[kRows];
      uint2 mv[kRows];
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        if (row0 + r < row_end) {
          wv[r] = ld_stream(W + (int64_t)(row0 + r) * it.ldw + col);
          const uint8_t* mp = it.mask + (int64_t)(row0 + r) * it.ldm + col;
          if (V == 8) mv[r] = *reinterpret_cast<const uint2*>(mp);
          else mv[r] = make_uint2(*reinterpret_cast<const uint32_t*>(mp), 0u);
        }
      }
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        const int row = row0 + r;
        if (row >= row_end) break;
        const uint32_t mb[2] = {mv[r].x, mv[r].y};
        float f[V], acc[V];
        Elem<T>::unpack(wv[r], f);
#pragma unroll
        for (int e = 0; e < V; ++e) acc[e] = 0.f;
        const float* brow = it.B + (int64_t)row * it.rank;
#pragma unroll
        for (int kk = 0; kk < RK; ++kk) {
          const float bv = kk < it.rank ? __ldg(brow + kk) : 0.f;
#pragma unroll
          for (int e = 0; e < V; ++e) acc[e] = fmaf(bv, a[kk][e], acc[e]);   // k ascending, like SGEMM
        }
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const bool keep = (mb[e / 4] >> (8 * (e % 4))) & 0xffu;
          const float delta = __fmul_rn(acc[e], it.scaling);
          const float merged = __fadd_rn(f[e], delta);
          f[e] = keep ? merged : (remask ? 0.f : f[e]);
        }
        st_stream(W + (int64_t)row * it.ldw + col, Elem<T>::pack(f));
      }
    }
  }
}

}  // namespace vlmc

extern "C" int vlmc_sparselora_merge(void* W, int dtype, int R, int C, int64_t ldw,
                                     const float* A, const float* B, int rank, float scaling,
                                     const uint8_t* keep_mask, int64_t ldm, int remask, void* stream) {
  using namespace vlmc;
  if (!W || !A || !B || !keep_mask || R < 1 || C < 1 || rank < 1 || ldw < C || ldm < C) return VLMC_ERR_BAD_ARG;
  if (dtype != VLMC_F32 && dtype != VLMC_F16 && dtype != VLMC_BF16) return VLMC_ERR_BAD_ARG;
  const int V = dtype == VLMC_F32 ? 4 : 8;
  if (C % V != 0 || ldw % V != 0 || ldm % V != 0 || ((uintptr_t)W & 15) != 0 || ((uintptr_t)keep_mask & 7) != 0)
    return VLMC_ERR_UNSUPPORTED;
  if (!is_device_ptr(W) || !is_device_ptr(A) || !is_device_ptr(B) || !is_device_ptr(keep_mask))
    return VLMC_ERR_NOT_DEVICE;
  const int coltiles = (C / V + kMergeThreads - 1) / kMergeThreads;
  int rowblocks = (kNumSMs * 16 + coltiles - 1) / coltiles;
  if (rowblocks > R) rowblocks = R;
  if (rowblocks > 65535) rowblocks = 65535;
  dim3 grid(coltiles, rowblocks);
  cudaStream_t st = (cudaStream_t)stream;
  if (rank <= 4) {
    VLMC_DISPATCH_DTYPE(dtype, (lora_merge_kernel<scalar_t, 4><<<grid, kMergeThreads, 0, st>>>(
                                   reinterpret_cast<scalar_t*>(W), ldw, R, C, A, B, rank, scaling, keep_mask, ldm, remask)));
  } else if (rank <= kMaxRankRegs) {
    VLMC_DISPATCH_DTYPE(dtype, (lora_merge_kernel<scalar_t, 8><<<grid, kMergeThreads, 0, st>>>(
                                   reinterpret_cast<scalar_t*>(W), ldw, R, C, A, B, rank, scaling, keep_mask, ldm, remask)));
  } else {
    VLMC_DISPATCH_DTYPE(dtype, (lora_merge_kernel<scalar_t, 0><<<grid, kMergeThreads, 0, st>>>(
                                   reinterpret_cast<scalar_t*>(W), ldw, R, C, A, B, rank, scaling, keep_mask, ldm, remask)));
  }
  return check_launch();
}

extern "C" int vlmc_sparselora_merge_batch(const vlmc_merge_item* items, int count, int dtype, int remask, void* stream) {
  using namespace vlmc;
  if (!items || count < 1 || count > kMbMax) return VLMC_ERR_BAD_ARG;
  if (dtype != VLMC_F32 && dtype != VLMC_F16 && dtype != VLMC_BF16) return VLMC_ERR_BAD_ARG;
  const int V = dtype == VLMC_F32 ? 4 : 8;
  MergeBatch b;
  b.count = count;
  b.unit_begin[0] = 0;
  for (int i = 0; i < count; ++i) {
    const vlmc_merge_item& s = items[i];
    if (!s.W || !s.A || !s.B || !s.keep_mask || s.R < 1 || s.C < 1 || s.rank < 1 || s.ldw < s.C || s.ldm < s.C) return VLMC_ERR_BAD_ARG;
    if (s.rank > kMaxRankRegs) return VLMC_ERR_UNSUPPORTED;     // larger ranks: vlmc_sparselora_merge per linear
    if (s.C % V != 0 || s.ldw % V != 0 || s.ldm % V != 0 || ((uintptr_t)s.W & 15) != 0 || ((uintptr_t)s.keep_mask & 7) != 0 ||
        ((uintptr_t)s.A & 15) != 0 || s.C % 4 != 0)
      return VLMC_ERR_UNSUPPORTED;
    if (!is_device_ptr(s.W) || !is_device_ptr(s.A) || !is_device_ptr(s.B) || !is_device_ptr(s.keep_mask)) return VLMC_ERR_NOT_DEVICE;
    MergeBatchItem& it = b.it[i];
    it.W = s.W; it.ldw = s.ldw; it.R = s.R; it.C = s.C; it.A = s.A; it.B = s.B; it.rank = s.rank; it.scaling = s.scaling;
    it.mask = s.keep_mask; it.ldm = s.ldm;
    it.coltiles = (s.C / V + kMergeThreads - 1) / kMergeThreads;
    const int64_t units = (int64_t)it.coltiles * ((s.R + kMbRows - 1) / kMbRows);
    if (b.unit_begin[i] + units > 0x7fffffff) return VLMC_ERR_UNSUPPORTED;
    b.unit_begin[i + 1] = b.unit_begin[i] + (int)units;
  }
  const int total = b.unit_begin[count];
  cudaStream_t st = (cudaStream_t)stream;
  VLMC_DISPATCH_DTYPE(dtype, {
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lora_merge_batch_kernel<scalar_t>, kMergeThreads, 0);
    int grid = kNumSMs * (per_sm < 1 ? 1 : per_sm);
    if (grid > total) grid = total;
    lora_merge_batch_kernel<scalar_t><<<grid, kMergeThreads, 0, st>>>(b, remask);
  });
  return check_launch();
}
