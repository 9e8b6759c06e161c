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
