// K15 / K16: the SparseLoRA masked training forward and its LoRA gradients.
//
// Replaces lavis/peft/src/peft/tuners/lora.py:359-382 (Linear.forward, r > 0, not merged)
//   sparse:      F.linear(x, (W + (B @ A).to(W.dtype) * scaling) * mask)        (:364-369)
//   not sparse:  F.linear(x,  W * mask + (B @ A).to(W.dtype) * scaling)         (:370-375)
// which re-materialises a dense [R, C] weight EVERY training step through ~5 elementwise passes (fp32 product,
// cast, scale, add, mask) and whose autograd backward makes as many again (mask, scale, cast to fp32) before two
// rank-r GEMMs.  Here
//   K15  vlmc_sparselora_effective_weight   the effective weight in one pass: read W + mask, write W_eff (5 B / weight),
//                                           rank-r dot in registers, the reference's three roundings in W's dtype
//   K16  vlmc_sparselora_lora_grads         dB = E A^T and dA = B^T E with E = (G (.) mask) * scaling rounded like the
//                                           reference's backward (W dtype, then fp32): one read of G + mask per output,
//                                           fp32 accumulation in a fixed order (deterministic)
// The two dense GEMMs of the step (y = x W_eff^T, G = dy^T x) stay library GEMMs (cuBLAS through torch).
#include "common.cuh"

namespace vlmc {

constexpr int kFwdThreads = 128;

template <typename T> __device__ __forceinline__ float round_dt(float v);
template <> __device__ __forceinline__ float round_dt<float>(float v) { return v; }
template <> __device__ __forceinline__ float round_dt<__half>(float v) { return __half2float(__float2half_rn(v)); }
template <> __device__ __forceinline__ float round_dt<__nv_bfloat16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// ---- K15 -------------------------------------------------------------------------------------------------
template <typename T, int RK>
__global__ void __launch_bounds__(kFwdThreads)
lora_effective_weight_kernel(const T* __restrict__ W, int64_t ldw, int R, int C, const float* __restrict__ A,
                             const float* __restrict__ B, int rank, float scaling, const uint8_t* __restrict__ mask,
                             int64_t ldm, int sparse, T* __restrict__ out, int64_t ldo) {
  constexpr int V = Elem<T>::kVec;
  const int col = (blockIdx.x * kFwdThreads + threadIdx.x) * V;
  if (col >= C) return;
  float a[RK > 0 ? RK : 1][V];
  if (RK > 0) {
#pragma unroll
    for (int kk = 0; kk < RK; ++kk)
#pragma unroll
      for (int e = 0; e < V; ++e) a[kk][e] = kk < rank ? A[(int64_t)kk * C + col + e] : 0.f;
  }
  for (int row = blockIdx.y; row < R; row += gridDim.y) {
    const uint4 wv = ld_stream(W + (int64_t)row * ldw + col);
    const uint8_t* mp = mask + (int64_t)row * ldm + col;
    uint32_t mb[2];
    if (V == 8) { const uint2 t = *reinterpret_cast<const uint2*>(mp); mb[0] = t.x; mb[1] = t.y; }
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
      // (B @ A).to(dtype) -> * scaling (rounded in dtype) -> W + . (rounded in dtype) -> * mask (exact)
      const float d = round_dt<T>(__fmul_rn(round_dt<T>(acc[e]), scaling));
      if (sparse) f[e] = keep ? round_dt<T>(__fadd_rn(f[e], d)) : 0.f;
      else f[e] = round_dt<T>(__fadd_rn(keep ? f[e] : 0.f, d));
    }
    st_stream(out + (int64_t)row * ldo + col, Elem<T>::pack(f));
  }
}

// ---- K16 -------------------------------------------------------------------------------------------------
constexpr int kGradMaxRank = 16;

// E[i][j] = float( round_dt( G[i][j] * mask * scaling ) )   (the reference's backward: mask and scale in W's dtype,
// then .to(float32) for the rank-r products)
template <typename T>
__device__ __forceinline__ void grad_e(const uint4& gv, const uint32_t (&mb)[2], int sparse, float scaling, float* e) {
  constexpr int V = Elem<T>::kVec;
  float g[V];
  Elem<T>::unpack(gv, g);
#pragma unroll
  for (int q = 0; q < V; ++q) {
    const bool keep = !sparse || ((mb[q / 4] >> (8 * (q % 4))) & 0xffu);
    e[q] = keep ? round_dt<T>(__fmul_rn(g[q], scaling)) : 0.f;
  }
}

// dB[i][k] = sum_j E[i][j] A[k][j]: one warp per row, lanes stride the 16-byte vectors of the row, fixed-order warp
// reduction.  A (rank x C fp32) is read through L1 / L2 (32-128 KB, shared by every row).
template <typename T, int RK>
__global__ void __launch_bounds__(256)
lora_grad_b_kernel(const T* __restrict__ G, int64_t ldg, int R, int C, const uint8_t* __restrict__ mask, int64_t ldm,
                   int sparse, const float* __restrict__ A, int rank, float scaling, float* __restrict__ dB) {
  constexpr int V = Elem<T>::kVec;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = C / V;
  for (int row = blockIdx.x * 8 + warp; row < R; row += gridDim.x * 8) {
    float acc[RK];
#pragma unroll
    for (int k = 0; k < RK; ++k) acc[k] = 0.f;
    for (int j = lane; j < nvec; j += 32) {
      const uint4 gv = ld_stream(G + (int64_t)row * ldg + j * V);
      uint32_t mb[2] = {0x01010101u, 0x01010101u};
      if (sparse) {
        const uint8_t* mp = mask + (int64_t)row * ldm + j * V;
        if (V == 8) { const uint2 t = *reinterpret_cast<const uint2*>(mp); mb[0] = t.x; mb[1] = t.y; }
        else mb[0] = *reinterpret_cast<const uint32_t*>(mp);
      }
      float e[V];
      grad_e<T>(gv, mb, sparse, scaling, e);
#pragma unroll
      for (int k = 0; k < RK; ++k) {
        if (k < rank) {
          const float* ap = A + (int64_t)k * C + j * V;
#pragma unroll
          for (int q = 0; q < V; q += 4) {
            const float4 av = __ldg(reinterpret_cast<const float4*>(ap + q));
            acc[k] = fmaf(e[q], av.x, acc[k]); acc[k] = fmaf(e[q + 1], av.y, acc[k]);
            acc[k] = fmaf(e[q + 2], av.z, acc[k]); acc[k] = fmaf(e[q + 3], av.w, acc[k]);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < RK; ++k) {
      if (k < rank) {
        const float s = warp_sum(acc[k]);
        if (lane == 0) dB[(int64_t)row * rank + k] = s;
      }
    }
  }
}

// dA[k][j] = sum_i B[i][k] E[i][j]: a thread owns 16 bytes of columns, CTAs split the rows into chunks, partial sums
// go to the workspace [chunk][rank][C] and a second kernel adds the chunks in order.
template <typename T, int RK>
__global__ void __launch_bounds__(kFwdThreads)
lora_grad_a_partial_kernel(const T* __restrict__ G, int64_t ldg, int R, int C, const uint8_t* __restrict__ mask,
                           int64_t ldm, int sparse, const float* __restrict__ B, int rank, float scaling,
                           int rows_per_chunk, float* __restrict__ part) {
  constexpr int V = Elem<T>::kVec;
  const int col = (blockIdx.x * kFwdThreads + threadIdx.x) * V;
  if (col >= C) return;
  const int r0 = blockIdx.y * rows_per_chunk;
  const int r1 = r0 + rows_per_chunk < R ? r0 + rows_per_chunk : R;
  float acc[RK][V];
#pragma unroll
  for (int k = 0; k < RK; ++k)
#pragma unroll
    for (int q = 0; q < V; ++q) acc[k][q] = 0.f;
  for (int row = r0; row < r1; ++row) {
    const uint4 gv = ld_stream(G + (int64_t)row * ldg + col);
    uint32_t mb[2] = {0x01010101u, 0x01010101u};
    if (sparse) {
      const uint8_t* mp = mask + (int64_t)row * ldm + col;
      if (V == 8) { const uint2 t = *reinterpret_cast<const uint2*>(mp); mb[0] = t.x; mb[1] = t.y; }
      else mb[0] = *reinterpret_cast<const uint32_t*>(mp);
    }
    float e[V];
    grad_e<T>(gv, mb, sparse, scaling, e);
    const float* brow = B + (int64_t)row * rank;
#pragma unroll
    for (int k = 0; k < RK; ++k) {
      if (k < rank) {
        const float b = __ldg(brow + k);
#pragma unroll
        for (int q = 0; q < V; ++q) acc[k][q] = fmaf(b, e[q], acc[k][q]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < RK; ++k) {
    if (k < rank) {
      float* pp = part + ((int64_t)blockIdx.y * rank + k) * C + col;
#pragma unroll
      for (int q = 0; q < V; q += 4)
        *reinterpret_cast<float4*>(pp + q) = make_float4(acc[k][q], acc[k][q + 1], acc[k][q + 2], acc[k][q + 3]);
    }
  }
}

__global__ void __launch_bounds__(256)
lora_grad_a_reduce_kernel(const float* __restrict__ part, int nchunks, int64_t n, float* __restrict__ dA) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int c = 0; c < nchunks; ++c) s += part[(int64_t)c * n + i];
  dA[i] = s;
}

static int grad_chunks(int R) {
  int chunks = kNumSMs * 2;                 // with C / (128 * V) column tiles this fills the GPU for every shape on the path
  if (chunks > (R + 15) / 16) chunks = (R + 15) / 16;
  return chunks < 1 ? 1 : chunks;
}

}  // namespace vlmc

extern "C" int vlmc_sparselora_effective_weight(const void* W, int dtype, int R, int C, int64_t ldw, const float* A,
                                                const float* B, int rank, float scaling, const uint8_t* keep_mask,
                                                int64_t ldm, int sparse, void* out, int64_t ldo, void* stream) {
  using namespace vlmc;
  if (!W || !A || !B || !keep_mask || !out || R < 1 || C < 1 || rank < 1 || ldw < C || ldm < C || ldo < C) return VLMC_ERR_BAD_ARG;
  if (dtype != VLMC_F32 && dtype != VLMC_F16 && dtype != VLMC_BF16) return VLMC_ERR_BAD_ARG;
  const int V = dtype == VLMC_F32 ? 4 : 8;
  if (C % V != 0 || ldw % V != 0 || ldm % V != 0 || ldo % V != 0 || ((uintptr_t)W & 15) != 0 || ((uintptr_t)out & 15) != 0 ||
      ((uintptr_t)keep_mask & 7) != 0)
    return VLMC_ERR_UNSUPPORTED;
  if (!is_device_ptr(W) || !is_device_ptr(A) || !is_device_ptr(B) || !is_device_ptr(keep_mask) || !is_device_ptr(out))
    return VLMC_ERR_NOT_DEVICE;
  const int coltiles = (C / V + kFwdThreads - 1) / kFwdThreads;
  int rowblocks = (kNumSMs * 16 + coltiles - 1) / coltiles;
  if (rowblocks > R) rowblocks = R;
  if (rowblocks > 65535) rowblocks = 65535;
  dim3 grid(coltiles, rowblocks);
  cudaStream_t st = (cudaStream_t)stream;
#define VLMC_EFF(RK)                                                                                              \
  VLMC_DISPATCH_DTYPE(dtype, (lora_effective_weight_kernel<scalar_t, RK><<<grid, kFwdThreads, 0, st>>>(          \
                                 reinterpret_cast<const scalar_t*>(W), ldw, R, C, A, B, rank, scaling, keep_mask, \
                                 ldm, sparse, reinterpret_cast<scalar_t*>(out), ldo)))
  if (rank <= 4) { VLMC_EFF(4); }
  else if (rank <= 8) { VLMC_EFF(8); }
  else { VLMC_EFF(0); }
#undef VLMC_EFF
  return check_launch();
}

extern "C" size_t vlmc_sparselora_lora_grads_workspace_bytes(int R, int C, int rank) {
  using namespace vlmc;
  if (R < 1 || C < 1 || rank < 1) return 0;
  return VLMC_WS_COUNTER_BYTES + (size_t)grad_chunks(R) * rank * C * sizeof(float);
}

extern "C" int vlmc_sparselora_lora_grads(const void* G, int dtype, int R, int C, int64_t ldg, const uint8_t* keep_mask,
                                          int64_t ldm, int sparse, const float* A, const float* B, int rank,
                                          float scaling, float* dA, float* dB, void* ws, size_t ws_bytes, void* stream) {
  using namespace vlmc;
  if (!G || !A || !B || !dA || !dB || !ws || R < 1 || C < 1 || rank < 1 || ldg < C) return VLMC_ERR_BAD_ARG;
  if (sparse && (!keep_mask || ldm < C)) return VLMC_ERR_BAD_ARG;
  if (rank > kGradMaxRank) return VLMC_ERR_UNSUPPORTED;
  if (dtype != VLMC_F32 && dtype != VLMC_F16 && dtype != VLMC_BF16) return VLMC_ERR_BAD_ARG;
  const int V = dtype == VLMC_F32 ? 4 : 8;
  if (C % V != 0 || ldg % V != 0 || (sparse && ldm % V != 0) || ((uintptr_t)G & 15) != 0 || ((uintptr_t)A & 15) != 0 ||
      (sparse && ((uintptr_t)keep_mask & 7) != 0))
    return VLMC_ERR_UNSUPPORTED;
  if (!is_device_ptr(G) || !is_device_ptr(A) || !is_device_ptr(B) || !is_device_ptr(dA) || !is_device_ptr(dB) ||
      !is_device_ptr(ws) || (sparse && !is_device_ptr(keep_mask)))
    return VLMC_ERR_NOT_DEVICE;
  if (ws_bytes < vlmc_sparselora_lora_grads_workspace_bytes(R, C, rank)) return VLMC_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  float* part = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + VLMC_WS_COUNTER_BYTES);
  int bgrid = (R + 7) / 8;
  if (bgrid > kNumSMs * 8) bgrid = kNumSMs * 8;
  const int chunks = grad_chunks(R);
  const int rows_per_chunk = (R + chunks - 1) / chunks;
  const int nchunks = (R + rows_per_chunk - 1) / rows_per_chunk;
  dim3 agrid((C / V + kFwdThreads - 1) / kFwdThreads, nchunks);
#define VLMC_GRADS(RK)                                                                                                  \
  VLMC_DISPATCH_DTYPE(dtype, {                                                                                          \
    lora_grad_b_kernel<scalar_t, RK><<<bgrid, 256, 0, st>>>(reinterpret_cast<const scalar_t*>(G), ldg, R, C, keep_mask, \
                                                            ldm, sparse, A, rank, scaling, dB);                         \
    lora_grad_a_partial_kernel<scalar_t, RK><<<agrid, kFwdThreads, 0, st>>>(                                            \
        reinterpret_cast<const scalar_t*>(G), ldg, R, C, keep_mask, ldm, sparse, B, rank, scaling, rows_per_chunk, part); \
  })
  if (rank <= 4) { VLMC_GRADS(4); }
  else if (rank <= 8) { VLMC_GRADS(8); }
  else { VLMC_GRADS(16); }
#undef VLMC_GRADS
  const int64_t n = (int64_t)rank * C;
  lora_grad_a_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(part, nchunks, n, dA);
  return check_launch();
}
