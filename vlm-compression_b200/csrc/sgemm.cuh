// fp32 GEMM building block for the SparseGPT factorisation and OBS sweep (K10, K13):
//   C[M,N] = beta * C + alpha * A[M,K] * op(B)     A row-major; B row-major [N,K] (B_NK: C += A B^T) or [K,N]
// True fp32 FFMA, like the reference's cuBLAS SGEMM with TF32 off (torch default).
// 128x128x16 CTA tile, 256 threads, 8x8 register micro-tile, double-buffered shared memory with
// register prefetch of the next k-slab.  tri = 1 computes only tiles on or below the block diagonal
// (symmetric rank-k updates of a lower triangle).  Round-2 work: move these onto tcgen05 with the 3xTF32 split.
#pragma once
#include "common.cuh"

namespace vlmc {

constexpr int kGM = 128, kGN = 128, kGK = 16, kGThreads = 256;

template <bool B_NK>
__global__ void __launch_bounds__(kGThreads, 2)
sgemm128_kernel(int M, int N, int K, float alpha, const float* A, int64_t lda,
                const float* __restrict__ B, int64_t ldb, float beta, float* C, int64_t ldc, int tri) {
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (tri && bj > bi) return;
  __shared__ __align__(16) float As[2][kGK][kGM + 4];
  __shared__ __align__(16) float Bs[2][kGK][kGN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = bi * kGM, n0 = bj * kGN;

  // global -> register staging: A tile 128 x 16 = 512 float4 (k-contiguous): 2 per thread
  float4 ra[2], rb[2];
  auto load_a = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int idx = tid + q * kGThreads;       // 0..511
      const int r = idx >> 2, kq = (idx & 3) * 4;
      const int gm = m0 + r, gk = k0 + kq;
      ra[q] = (gm < M && gk < K) ? *reinterpret_cast<const float4*>(A + (int64_t)gm * lda + gk) : make_float4(0, 0, 0, 0);
    }
  };
  auto load_b = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int idx = tid + q * kGThreads;
      if (B_NK) {
        const int r = idx >> 2, kq = (idx & 3) * 4;
        const int gn = n0 + r, gk = k0 + kq;
        rb[q] = (gn < N && gk < K) ? *reinterpret_cast<const float4*>(B + (int64_t)gn * ldb + gk) : make_float4(0, 0, 0, 0);
      } else {
        const int kr = idx >> 5, nq = (idx & 31) * 4;   // 16 k-rows x 32 float4
        const int gk = k0 + kr, gn = n0 + nq;
        rb[q] = (gk < K && gn < N) ? *reinterpret_cast<const float4*>(B + (int64_t)gk * ldb + gn) : make_float4(0, 0, 0, 0);
      }
    }
  };
  auto store_ab = [&](int buf) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int idx = tid + q * kGThreads;
      const int r = idx >> 2, kq = (idx & 3) * 4;
      As[buf][kq + 0][r] = ra[q].x; As[buf][kq + 1][r] = ra[q].y; As[buf][kq + 2][r] = ra[q].z; As[buf][kq + 3][r] = ra[q].w;
      if (B_NK) {
        Bs[buf][kq + 0][r] = rb[q].x; Bs[buf][kq + 1][r] = rb[q].y; Bs[buf][kq + 2][r] = rb[q].z; Bs[buf][kq + 3][r] = rb[q].w;
      } else {
        const int kr = idx >> 5, nq = (idx & 31) * 4;
        *reinterpret_cast<float4*>(&Bs[buf][kr][nq]) = rb[q];
      }
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  load_a(0); load_b(0);
  store_ab(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < K; k0 += kGK) {
    const bool more = k0 + kGK < K;
    if (more) { load_a(k0 + kGK); load_b(k0 + kGK); }
#pragma unroll
    for (int kk = 0; kk < kGK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      store_ab(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= M) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int gn = n0 + h * 64 + tx * 4;
      if (gn >= N) continue;                   // N % 4 == 0: a float4 is entirely in or out
      float* cp = C + (int64_t)gm * ldc + gn;
      float4 o;
      if (beta != 0.f) {
        const float4 c = *reinterpret_cast<const float4*>(cp);
        o.x = fmaf(alpha, acc[i][h * 4 + 0], beta * c.x); o.y = fmaf(alpha, acc[i][h * 4 + 1], beta * c.y);
        o.z = fmaf(alpha, acc[i][h * 4 + 2], beta * c.z); o.w = fmaf(alpha, acc[i][h * 4 + 3], beta * c.w);
      } else {
        o.x = alpha * acc[i][h * 4 + 0]; o.y = alpha * acc[i][h * 4 + 1];
        o.z = alpha * acc[i][h * 4 + 2]; o.w = alpha * acc[i][h * 4 + 3];
      }
      *reinterpret_cast<float4*>(cp) = o;
    }
  }
}

// host launcher; all sizes / leading dimensions must be multiples of 4 and pointers 16-byte aligned
inline int sgemm(bool b_nk, int M, int N, int K, float alpha, const float* A, int64_t lda, const float* B, int64_t ldb,
                 float beta, float* C, int64_t ldc, int tri, cudaStream_t st) {
  if (M <= 0 || N <= 0) return VLMC_OK;
  if (K <= 0) return VLMC_ERR_BAD_ARG;
  if ((K & 3) || (N & 3) || (lda & 3) || (ldb & 3) || (ldc & 3) || ((uintptr_t)A & 15) || ((uintptr_t)B & 15) ||
      ((uintptr_t)C & 15))
    return VLMC_ERR_UNSUPPORTED;
  dim3 grid((N + kGN - 1) / kGN, (M + kGM - 1) / kGM);
  if (b_nk) sgemm128_kernel<true><<<grid, kGThreads, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, tri);
  else sgemm128_kernel<false><<<grid, kGThreads, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, tri);
  return check_launch();
}

}  // namespace vlmc
