// K11 + K12 + K13: the SparseGPT column-block OBS sweep.
//
// Replaces lavis/compression/pruners/sparsegpt_pruner.py:160-215 (fasterprune after the factorisation):
//   s = W^2 / diag(U)^2 ; importance_score = mean(s)                                           (:160-165)
//   for each 128-column block:                                                                  (:169)
//     unstructured: thresh = sort((W1^2/diag(U1)^2).flatten())[int(numel*p)] ; mask1 = tmp <= thresh  (:183-185) K11
//     n:m: every m-th column the n smallest of the group, on the already compensated W1         (:193-195)
//     128 sequential steps: q = masked w ; err = (w-q)/d ; W1[:, i:] -= err (x) U1[i, i:]        (:189-205) K12
//     W[:, i2:] -= Err1 @ U[i1:i2, i2:]                                                         (:210)     K13
// The reference issues ~8 launches per COLUMN (~90 k per down_proj) and one flatten-sort per block.  Here one
// cooperative kernel per block does K11 + K12: a thread owns one weight row (rows are independent given U), the
// row and the U1 tile live in shared memory, the block-wide k-th value is found by counting passes with a
// grid barrier (exact, no sort), and the finished columns are written straight to the fp16/bf16 weight.
// The lazy trailing update is the fp32 GEMM of sgemm.cuh.
#include <cooperative_groups.h>
#include "sgemm.cuh"

namespace vlmc {

int launch_mean_finalize(const float* part, int n, double denom, float* out, cudaStream_t st);

constexpr int kOB = 128;            // column block (the reference's blocksize default, the only one the scripts use)
constexpr int kObsThreads = 128;    // rows per CTA
constexpr int kObsPivots = 8;
constexpr int kObsPasses = 20;
constexpr int kObsCap = 512;

typedef unsigned long long ull;

struct ObsSelState {
  ull counts[kObsPasses][kObsPivots];
  unsigned int cand_cnt;
  unsigned int bar;
  uint32_t cand[kObsCap];
};

struct ObsBracket { uint32_t lo, hi; ull glo, ghi; };

__device__ __forceinline__ uint32_t obs_clamp(double x, uint32_t lo, uint32_t hi) {
  if (!(x > (double)lo + 1.0)) return lo + 1;
  if (!(x < (double)hi - 1.0)) return hi - 1;
  return (uint32_t)x;
}

__device__ void obs_pivots(const ObsBracket& b, ull k, uint32_t* p) {
  const double w = (double)(b.hi - b.lo), lo = (double)b.lo;
  p[0] = obs_clamp(lo + 0.25 * w, b.lo, b.hi);
  p[1] = obs_clamp(lo + 0.50 * w, b.lo, b.hi);
  p[2] = obs_clamp(lo + 0.75 * w, b.lo, b.hi);
  const double f = ((double)(k - b.glo) - 0.5) / (double)(b.ghi - b.glo);
  const double e = lo + f * w;
  p[3] = obs_clamp(e - w * 0.0625, b.lo, b.hi);
  p[4] = obs_clamp(e - w * (1.0 / 256.0), b.lo, b.hi);
  p[5] = obs_clamp(e, b.lo, b.hi);
  p[6] = obs_clamp(e + w * (1.0 / 256.0), b.lo, b.hi);
  p[7] = obs_clamp(e + w * 0.0625, b.lo, b.hi);
}

// all CTAs are co-resident (cooperative launch): arrive + spin on a monotonically increasing counter
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int nblocks, unsigned int& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += 1;
    __threadfence();
    atomicAdd(bar, 1u);
    const unsigned int target = epoch * nblocks;
    while (*reinterpret_cast<volatile unsigned int*>(bar) < target) { __nanosleep(32); }
    __threadfence();
  }
  __syncthreads();
}

template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_float<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <typename T> __device__ __forceinline__ float to_float(T v);
template <> __device__ __forceinline__ float to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_float<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// W (dtype) -> fp32 working copy, dead input channels zeroed (sparsegpt_pruner.py:84-97)
template <typename T>
__global__ void obs_upcast_kernel(const T* __restrict__ W, int64_t ldw, float* __restrict__ W32, int R, int C,
                                  const uint8_t* __restrict__ dead) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const bool d = dead ? dead[c] != 0 : false;
  for (int r = blockIdx.y; r < R; r += gridDim.y)
    W32[(int64_t)r * C + c] = d ? 0.f : to_float<T>(W[(int64_t)r * ldw + c]);
}

// sum over the matrix of W^2 / diag(U)^2 (the reference's importance score numerator)
__global__ void obs_importance_kernel(const float* __restrict__ W32, int R, int C, const float* __restrict__ U,
                                      int64_t ldu, float* __restrict__ part) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  float s = 0.f;
  if (c < C) {
    const float d = U[(int64_t)c * ldu + c];
    const float d2 = __fmul_rn(d, d);
    for (int r = blockIdx.y; r < R; r += gridDim.y) {
      const float w = W32[(int64_t)r * C + c];
      s += __fdiv_rn(__fmul_rn(w, w), d2);
    }
  }
  __shared__ float red[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    part[blockIdx.y * gridDim.x + blockIdx.x] = t;
  }
}

struct ObsParams {
  float* W32;          // [R, C] fp32 working copy
  void* Wout;          // weight tensor (dtype), receives the finished columns
  int64_t ldw;
  int R, C;
  const float* U;
  int64_t ldu;
  int i1, bs;          // column block [i1, i1 + bs)
  float* Err;          // [R, kOB]
  uint8_t* keep;       // optional [R, ldm]
  int64_t ldm;
  ull kth;             // unstructured: 1-indexed rank of the threshold value among the R*bs block scores
  int prune_n, prune_m;
  ObsSelState* sel;
  int rows_per_cta;
};

template <typename T>
__global__ void __launch_bounds__(kObsThreads, 1)
obs_block_kernel(const ObsParams p) {
  extern __shared__ __align__(16) float osm[];
  float* Us = osm;                              // [kOB][kOB]  U1 tile (row i broadcast-read)
  float* Ws = Us + kOB * kOB;                   // [kOB][kObsThreads]  row-private: Ws[j * 128 + tid]
  uint32_t* Ks = reinterpret_cast<uint32_t*>(Ws + kOB * kObsThreads);   // [kOB][kObsThreads]  score bits
  __shared__ uint32_t s_piv[kObsPivots];
  __shared__ unsigned int s_cnt[kObsPivots];
  __shared__ ull s_glob[kObsPivots];
  __shared__ uint32_t s_v;
  __shared__ uint32_t s_cand[kObsCap];

  const int tid = threadIdx.x;
  const int bs = p.bs;
  const int row = blockIdx.x * p.rows_per_cta + tid;
  const bool active = tid < p.rows_per_cta && row < p.R;

  for (int idx = tid; idx < bs * bs; idx += kObsThreads) {
    const int i = idx / bs, j = idx % bs;
    Us[i * kOB + j] = p.U[(int64_t)(p.i1 + i) * p.ldu + p.i1 + j];
  }
  if (active) {
    const float* wr = p.W32 + (int64_t)row * p.C + p.i1;
    for (int j = 0; j < bs; j += 4) {
      const float4 v = *reinterpret_cast<const float4*>(wr + j);
      Ws[(j + 0) * kObsThreads + tid] = v.x; Ws[(j + 1) * kObsThreads + tid] = v.y;
      Ws[(j + 2) * kObsThreads + tid] = v.z; Ws[(j + 3) * kObsThreads + tid] = v.w;
    }
  }
  __syncthreads();

  uint32_t mbits[kOB / 32] = {};      // bit j set = column j of this row is pruned
  unsigned int epoch = 0;

  if (p.prune_n == 0) {
    // ---- K11: exact k-th smallest of the R x bs block scores, mask = score <= that value ----
    for (int j = 0; j < bs; ++j) {
      uint32_t key = 0xffffffffu;
      if (active) {
        const float w = Ws[j * kObsThreads + tid], d = Us[j * kOB + j];
        key = __float_as_uint(__fdiv_rn(__fmul_rn(w, w), __fmul_rn(d, d)));
      }
      Ks[j * kObsThreads + tid] = key;
    }
    const ull n = (ull)p.R * (ull)bs;
    ObsBracket b{0u, 0xffffffffu, 0ull, n};
    int pass = 0;
    while (!((b.ghi - b.glo) <= (ull)kObsCap || (b.hi - b.lo) == 1u) && pass < kObsPasses) {
      if (tid == 0) { uint32_t pv[kObsPivots]; obs_pivots(b, p.kth, pv); for (int i = 0; i < kObsPivots; ++i) { s_piv[i] = pv[i]; s_cnt[i] = 0; } }
      __syncthreads();
      uint32_t pv[kObsPivots];
#pragma unroll
      for (int i = 0; i < kObsPivots; ++i) pv[i] = s_piv[i];
      unsigned int cnt[kObsPivots] = {};
      for (int j = 0; j < bs; ++j) {
        const uint32_t key = Ks[j * kObsThreads + tid];
#pragma unroll
        for (int i = 0; i < kObsPivots; ++i) cnt[i] += key < pv[i] ? 1u : 0u;
      }
#pragma unroll
      for (int i = 0; i < kObsPivots; ++i) {
        const unsigned int c = __reduce_add_sync(0xffffffffu, cnt[i]);
        if ((tid & 31) == 0) atomicAdd(&s_cnt[i], c);
      }
      __syncthreads();
      if (tid < kObsPivots) atomicAdd(&p.sel->counts[pass][tid], (ull)s_cnt[tid]);
      grid_barrier(&p.sel->bar, gridDim.x, epoch);
      if (tid < kObsPivots) s_glob[tid] = *reinterpret_cast<volatile ull*>(&p.sel->counts[pass][tid]);
      __syncthreads();
#pragma unroll
      for (int i = 0; i < kObsPivots; ++i) {
        const ull c = s_glob[i];
        if (c <= p.kth - 1) { if (pv[i] > b.lo) { b.lo = pv[i]; b.glo = c; } }
        else                { if (pv[i] < b.hi) { b.hi = pv[i]; b.ghi = c; } }
      }
      ++pass;
      __syncthreads();
    }
    uint32_t v = b.lo;
    if (b.hi - b.lo != 1u) {
      // gather the <= kObsCap candidates in [lo, hi) and rank them (every CTA does the same tiny ranking)
      for (int j = 0; j < bs; ++j) {
        const uint32_t key = Ks[j * kObsThreads + tid];
        if (key >= b.lo && key < b.hi) {
          const unsigned int slot = atomicAdd(&p.sel->cand_cnt, 1u);
          if (slot < (unsigned)kObsCap) p.sel->cand[slot] = key;
        }
      }
      grid_barrier(&p.sel->bar, gridDim.x, epoch);
      const int m = (int)(b.ghi - b.glo);
      for (int i = tid; i < m && i < kObsCap; i += kObsThreads) s_cand[i] = *reinterpret_cast<volatile uint32_t*>(&p.sel->cand[i]);
      __syncthreads();
      const int target = (int)(p.kth - 1 - b.glo);
      for (int t = tid; t < m && t < kObsCap; t += kObsThreads) {
        const uint32_t kt = s_cand[t];
        int less = 0, leq = 0;
        for (int j = 0; j < m; ++j) { less += s_cand[j] < kt ? 1 : 0; leq += s_cand[j] <= kt ? 1 : 0; }
        if (less <= target && target < leq) s_v = kt;
      }
      __syncthreads();
      v = s_v;
    }
    if (active) {
      for (int j = 0; j < bs; ++j)
        if (Ks[j * kObsThreads + tid] <= v) mbits[j >> 5] |= 1u << (j & 31);   // `<=` (:185)
    }
  }

  // ---- K12: 128 sequential column steps, the row stays in this thread's shared-memory column ----
  if (active) {
    float* er = p.Err + (int64_t)row * kOB;
    for (int i = 0; i < bs; ++i) {
      if (p.prune_n != 0 && (i % p.prune_m) == 0) {
        // n:m on the compensated weights: the n smallest w^2/d^2 of columns [i, i+m), ties -> lower column
        const int m = p.prune_m;
        for (int a = 0; a < m && i + a < bs; ++a) {
          const float wa = Ws[(i + a) * kObsThreads + tid], da = Us[(i + a) * kOB + i + a];
          const float ka = __fdiv_rn(__fmul_rn(wa, wa), __fmul_rn(da, da));
          int rank = 0;
          for (int c = 0; c < m && i + c < bs; ++c) {
            if (c == a) continue;
            const float wc = Ws[(i + c) * kObsThreads + tid], dc = Us[(i + c) * kOB + i + c];
            const float kc2 = __fdiv_rn(__fmul_rn(wc, wc), __fmul_rn(dc, dc));
            rank += (kc2 < ka || (kc2 == ka && c < a)) ? 1 : 0;
          }
          if (rank < p.prune_n) mbits[(i + a) >> 5] |= 1u << ((i + a) & 31);
        }
      }
      const float w = Ws[i * kObsThreads + tid];
      const bool pr = (mbits[i >> 5] >> (i & 31)) & 1u;
      const float q = pr ? 0.f : w;
      const float d = Us[i * kOB + i];
      const float err = __fdiv_rn(__fsub_rn(w, q), d);
      er[i] = err;
      Ws[i * kObsThreads + tid] = q;                 // Q1[:, i]
      if (err != 0.f) {
        for (int j = i + 1; j < bs; ++j) {
          const float u = Us[i * kOB + j];
          // product and subtraction rounded separately, like the reference's outer-product matmul + in-place sub
          Ws[j * kObsThreads + tid] = __fsub_rn(Ws[j * kObsThreads + tid], __fmul_rn(err, u));
        }
      }
    }
    for (int i = bs; i < kOB; ++i) er[i] = 0.f;
    // finished columns -> weight tensor (one rounding to its dtype), fp32 working copy, optional keep mask
    T* wo = reinterpret_cast<T*>(p.Wout) + (int64_t)row * p.ldw + p.i1;
    float* w32 = p.W32 + (int64_t)row * p.C + p.i1;
    for (int j = 0; j < bs; ++j) {
      const float q = Ws[j * kObsThreads + tid];
      wo[j] = from_float<T>(q);
      w32[j] = q;
    }
    if (p.keep) {
      uint8_t* kp = p.keep + (int64_t)row * p.ldm + p.i1;
      for (int j = 0; j < bs; ++j) kp[j] = ((mbits[j >> 5] >> (j & 31)) & 1u) ? 0 : 1;
    }
  }
}

size_t obs_workspace_bytes(int R, int C) {
  const int nblk = (C + kOB - 1) / kOB;
  return VLMC_WS_COUNTER_BYTES + align_up((size_t)R * C * sizeof(float), 256) + align_up((size_t)R * kOB * sizeof(float), 256) +
         align_up((size_t)nblk * sizeof(ObsSelState), 256) + align_up((size_t)kNumSMs * 64 * sizeof(float), 256);
}

}  // namespace vlmc

extern "C" int vlmc_obs_sweep(void* W, int dtype, int R, int C, int64_t ldw, const float* U, int64_t ldu,
                              const uint8_t* dead, double sparsity, int prune_n, int prune_m, int blocksize,
                              uint8_t* keep_mask, int64_t ldm, float* importance_score,
                              void* ws, size_t ws_bytes, void* stream) {
  using namespace vlmc;
  if (!W || !U || !ws || R < 1 || C < 1 || ldw < C || ldu < C) return VLMC_ERR_BAD_ARG;
  if (dtype != VLMC_F32 && dtype != VLMC_F16 && dtype != VLMC_BF16) return VLMC_ERR_BAD_ARG;
  if (blocksize != kOB) return VLMC_ERR_UNSUPPORTED;
  if (prune_n < 0 || (prune_n > 0 && (prune_m <= prune_n || prune_m > 32 || kOB % prune_m != 0))) return VLMC_ERR_BAD_ARG;
  if (prune_n == 0 && !(sparsity >= 0.0 && sparsity < 1.0)) return VLMC_ERR_BAD_ARG;
  if ((C & 3) || (ldu & 3) || ((uintptr_t)U & 15)) return VLMC_ERR_UNSUPPORTED;
  if (keep_mask && ldm < C) return VLMC_ERR_BAD_ARG;
  if (!is_device_ptr(W) || !is_device_ptr(U) || !is_device_ptr(ws)) return VLMC_ERR_NOT_DEVICE;
  if (ws_bytes < obs_workspace_bytes(R, C)) return VLMC_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;

  const int nblk = (C + kOB - 1) / kOB;
  char* base = reinterpret_cast<char*>(ws) + VLMC_WS_COUNTER_BYTES;
  float* W32 = reinterpret_cast<float*>(base);
  base += align_up((size_t)R * C * sizeof(float), 256);
  float* Err = reinterpret_cast<float*>(base);
  base += align_up((size_t)R * kOB * sizeof(float), 256);
  ObsSelState* sel = reinterpret_cast<ObsSelState*>(base);
  base += align_up((size_t)nblk * sizeof(ObsSelState), 256);
  float* part = reinterpret_cast<float*>(base);

  if (cudaMemsetAsync(sel, 0, (size_t)nblk * sizeof(ObsSelState), st) != cudaSuccess) return check_launch();
  {
    dim3 grid((C + 255) / 256, R < 64 ? R : 64);
    VLMC_DISPATCH_DTYPE(dtype, (obs_upcast_kernel<scalar_t><<<grid, 256, 0, st>>>(
                                   reinterpret_cast<const scalar_t*>(W), ldw, W32, R, C, dead)));
    if (importance_score) {
      obs_importance_kernel<<<grid, 256, 0, st>>>(W32, R, C, U, ldu, part);
      int rc = launch_mean_finalize(part, grid.x * grid.y, (double)R * (double)C, importance_score, st);
      if (rc) return rc;
    }
  }

  const size_t smem = (size_t)(kOB * kOB + 2 * kOB * kObsThreads) * sizeof(float);
  void* kern = nullptr;
  switch (dtype) {
    case VLMC_F32: kern = (void*)obs_block_kernel<float>; break;
    case VLMC_F16: kern = (void*)obs_block_kernel<__half>; break;
    default: kern = (void*)obs_block_kernel<__nv_bfloat16>; break;
  }
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return check_launch();

  int rows_per_cta = (R + kNumSMs - 1) / kNumSMs;
  if (rows_per_cta > kObsThreads) rows_per_cta = kObsThreads;   // more rows than 148 x 128: see the grid check below
  const int grid = (R + rows_per_cta - 1) / rows_per_cta;
  if (grid > kNumSMs) return VLMC_ERR_UNSUPPORTED;              // the grid barrier needs every CTA resident

  for (int blk = 0; blk < nblk; ++blk) {
    ObsParams p;
    p.W32 = W32; p.Wout = W; p.ldw = ldw; p.R = R; p.C = C; p.U = U; p.ldu = ldu;
    p.i1 = blk * kOB;
    p.bs = (C - p.i1 < kOB) ? (C - p.i1) : kOB;
    p.Err = Err; p.keep = keep_mask; p.ldm = ldm;
    p.kth = (ull)((double)((ull)R * (ull)p.bs) * sparsity) + 1;   // int(numel * sparsity) is the 0-indexed rank (:184)
    p.prune_n = prune_n; p.prune_m = prune_m;
    p.sel = sel + blk;
    p.rows_per_cta = rows_per_cta;
    void* args[] = {&p};
    cudaError_t e = cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(kObsThreads), args, smem, st);
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; return VLMC_ERR_CUDA; }
    const int i2 = p.i1 + p.bs;
    if (i2 < C) {
      // K13: W[:, i2:] -= Err1 @ U[i1:i2, i2:]
      int rc = sgemm(false, R, C - i2, kOB, -1.f, Err, kOB, U + (int64_t)p.i1 * ldu + i2, ldu, 1.f, W32 + i2, C, 0, st);
      if (rc) return rc;
    }
  }
  return check_launch();
}
