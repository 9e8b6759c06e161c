// K10: U = upper Cholesky factor of H^-1 (H^-1 = U^T U), fp32, without ever forming H^-1.
//
// Replaces the chain at lavis/compression/pruners/sparsegpt_pruner.py:114-157
//   L = cholesky(H) ; Hinv = cholesky_inverse(L) ; U = cholesky(Hinv, upper=True)      (cuSOLVER potrf/potri/potrf)
// Identity used: H^-1 = U^T U  <=>  H = V V^T with V = U^-1 upper triangular, and with J the exchange matrix
// V = J * chol_lower(J H J) * J.  So ONE Cholesky of the flipped matrix and ONE triangular inverse give U:
//   F = J H J ; F = L L^T ; U = J L^-1 J                      (2/3 C^3 flops instead of 4/3 C^3 + a full inverse)
//
//   chol_lower   right-looking blocked (128): diagonal block factor + its inverse in one CTA (shared memory),
//                panel = GEMM with the block inverse, trailing symmetric update = GEMM on lower tiles only
//   tri inverse  recursive doubling: [[A,0],[B,D]]^-1 = [[A^-1,0],[-D^-1 B A^-1, D^-1]], two GEMMs per merge
// GEMMs run on the tensor cores with the 3xTF32 split (gemm3x.cu): fp32-grade accuracy, like the reference's fp32
// cuSOLVER/cuBLAS path.  A non-positive or NaN pivot
// sets *status = VLMC_NOT_POSDEF; the caller adds percdamp * mean(diag H) and retries (reference :114-128).
#include "gemm3x.cuh"

namespace vlmc {

constexpr int kNB = 128;          // block size
constexpr int kPotrfThreads = 128;

__global__ void flip_copy_kernel(const float* __restrict__ H, int64_t ldh, float* __restrict__ F, int64_t ldf, int C) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j < C) F[(int64_t)i * ldf + j] = H[(int64_t)(C - 1 - i) * ldh + (C - 1 - j)];
}

// strictly-lower entries of Li move to the mirrored strictly-upper slot; the lower slot is cleared
__global__ void flip_to_upper_kernel(float* __restrict__ U, int64_t ldu, int C) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j < i) {
    const float v = U[(int64_t)i * ldu + j];
    U[(int64_t)i * ldu + j] = 0.f;
    U[(int64_t)(C - 1 - i) * ldu + (C - 1 - j)] = v;
  } else if (j == i && i < C / 2) {
    const int64_t a = (int64_t)i * ldu + i, b = (int64_t)(C - 1 - i) * ldu + (C - 1 - i);
    const float t = U[a]; U[a] = U[b]; U[b] = t;
  }
}

// One CTA: factor the bs x bs diagonal block at (k0,k0) of F in shared memory (left-looking, one thread per
// row), write L_kk back, and write L_kk^-1 (lower, zero above the diagonal) into Li at the same position.
__global__ void __launch_bounds__(kPotrfThreads)
potrf_block_kernel(float* F, int64_t ldf, float* Li, int64_t ldi, int k0, int bs, int* status) {
  extern __shared__ float sm[];
  float (*S)[kNB + 1] = reinterpret_cast<float (*)[kNB + 1]>(sm);
  float (*X)[kNB + 1] = reinterpret_cast<float (*)[kNB + 1]>(sm + kNB * (kNB + 1));
  const int t = threadIdx.x;
  for (int idx = t; idx < bs * bs; idx += kPotrfThreads) {
    const int i = idx / bs, j = idx % bs;
    S[i][j] = (j <= i) ? F[(int64_t)(k0 + i) * ldf + k0 + j] : 0.f;
    X[i][j] = 0.f;
  }
  __syncthreads();
  bool bad = false;
  for (int j = 0; j < bs; ++j) {
    // row t >= j: s = F[t][j] - sum_{k<j} L[t][k] L[j][k]
    float s = 0.f;
    if (t >= j && t < bs) {
      s = S[t][j];
      for (int k = 0; k < j; ++k) s = fmaf(-S[t][k], S[j][k], s);
    }
    __syncthreads();
    if (t == j) {
      if (!(s > 0.f) || !isfinite(s)) { bad = true; s = 1.f; }
      S[j][j] = sqrtf(s);
    }
    __syncthreads();
    if (t > j && t < bs) S[t][j] = s / S[j][j];
    __syncthreads();   // column j is final before any thread starts the dot products of column j + 1
  }
  if (bad) atomicMax(status, (int)VLMC_NOT_POSDEF);
  // inverse: thread c solves L x = e_c by forward substitution (column c of L^-1)
  if (t < bs) {
    const int c = t;
    X[c][c] = 1.f / S[c][c];
    for (int i = c + 1; i < bs; ++i) {
      float s = 0.f;
      for (int k = c; k < i; ++k) s = fmaf(S[i][k], X[k][c], s);
      X[i][c] = -s / S[i][i];
    }
  }
  __syncthreads();
  for (int idx = t; idx < bs * bs; idx += kPotrfThreads) {
    const int i = idx / bs, j = idx % bs;
    if (j <= i) F[(int64_t)(k0 + i) * ldf + k0 + j] = S[i][j];
    Li[(int64_t)(k0 + i) * ldi + k0 + j] = (j <= i) ? X[i][j] : 0.f;
  }
}

__global__ void diag_prepare_kernel(float* H, int64_t ldh, int C, float percdamp, float* damp_out, uint8_t* dead_out) {
  // single CTA: dead columns (diag == 0 -> 1, sparsegpt_pruner.py:95-96) and damp = percdamp * mean(diag) (:111)
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    float d = H[(int64_t)i * ldh + i];
    const bool dead = d == 0.f;
    if (dead) { d = 1.f; H[(int64_t)i * ldh + i] = 1.f; }
    if (dead_out) dead_out[i] = dead ? 1 : 0;
    s += (double)d;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    *damp_out = percdamp * (float)(t / (double)C);
  }
}

__global__ void add_diag_kernel(float* H, int64_t ldh, int C, const float* damp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) H[(int64_t)i * ldh + i] += *damp;
}

size_t chol_workspace_bytes(int C) {
  const int nb = (C + kNB - 1) / kNB;
  const size_t half = (size_t)((nb + 1) / 2) * kNB;
  return VLMC_WS_COUNTER_BYTES + align_up((size_t)C * C * sizeof(float), 256) + align_up(half * half * sizeof(float), 256);
}

}  // namespace vlmc

extern "C" int vlmc_hessian_prepare(float* H, int C, int64_t ldh, float percdamp, float* damp_out,
                                    uint8_t* dead_out, void* stream) {
  using namespace vlmc;
  if (!H || !damp_out || C < 1 || ldh < C) return VLMC_ERR_BAD_ARG;
  if (!is_device_ptr(H) || !is_device_ptr(damp_out)) return VLMC_ERR_NOT_DEVICE;
  diag_prepare_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(H, ldh, C, percdamp, damp_out, dead_out);
  return check_launch();
}

extern "C" int vlmc_hessian_add_damp(float* H, int C, int64_t ldh, const float* damp, void* stream) {
  using namespace vlmc;
  if (!H || !damp || C < 1 || ldh < C) return VLMC_ERR_BAD_ARG;
  if (!is_device_ptr(H) || !is_device_ptr(damp)) return VLMC_ERR_NOT_DEVICE;
  add_diag_kernel<<<(C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(H, ldh, C, damp);
  return check_launch();
}

extern "C" int vlmc_chol_inv_upper(const float* H, int C, int64_t ldh, float* U, int64_t ldu, int* status,
                                   void* ws, size_t ws_bytes, void* stream) {
  using namespace vlmc;
  if (!H || !U || !status || !ws || C < 1 || ldh < C || ldu < C) return VLMC_ERR_BAD_ARG;
  if ((C & 3) || (ldh & 3) || (ldu & 3) || ((uintptr_t)H & 15) || ((uintptr_t)U & 15)) return VLMC_ERR_UNSUPPORTED;
  if (!is_device_ptr(H) || !is_device_ptr(U) || !is_device_ptr(status) || !is_device_ptr(ws)) return VLMC_ERR_NOT_DEVICE;
  if (ws_bytes < chol_workspace_bytes(C)) return VLMC_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  char* base = reinterpret_cast<char*>(ws) + VLMC_WS_COUNTER_BYTES;
  float* F = reinterpret_cast<float*>(base);
  const int64_t ldf = C;
  float* X = reinterpret_cast<float*>(base + align_up((size_t)C * C * sizeof(float), 256));
  const int nb = (C + kNB - 1) / kNB;

  static bool attr_set = false;
  const int potrf_smem = 2 * kNB * (kNB + 1) * sizeof(float);
  if (!attr_set) {
    if (cudaFuncSetAttribute(potrf_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, potrf_smem) != cudaSuccess)
      return check_launch();
    attr_set = true;
  }
  if (cudaMemsetAsync(status, 0, sizeof(int), st) != cudaSuccess) return check_launch();
  if (cudaMemset2DAsync(U, ldu * sizeof(float), 0, (size_t)C * sizeof(float), C, st) != cudaSuccess) return check_launch();
  flip_copy_kernel<<<dim3((C + 255) / 256, C), 256, 0, st>>>(H, ldh, F, ldf, C);

  // ---- blocked Cholesky of F (lower), diagonal-block inverses go straight into U (used as Li) ----
  int rc;
  for (int k = 0; k < nb; ++k) {
    const int k0 = k * kNB;
    const int bs = (C - k0 < kNB) ? (C - k0) : kNB;
    potrf_block_kernel<<<1, kPotrfThreads, potrf_smem, st>>>(F, ldf, U, ldu, k0, bs, status);
    const int below = C - k0 - bs;
    if (below > 0) {
      float* panel = F + (int64_t)(k0 + bs) * ldf + k0;
      // panel <- panel * L_kk^-T   (C = A B^T with B = L_kk^-1 stored [N,K])
      rc = gemm3x(true, below, bs, bs, 1.f, panel, ldf, U + (int64_t)k0 * ldu + k0, ldu, 0.f, panel, ldf, 0, 0, st);
      if (rc) return rc;
      // trailing <- trailing - panel panel^T on the lower tiles
      float* trail = F + (int64_t)(k0 + bs) * ldf + (k0 + bs);
      rc = gemm3x(true, below, below, bs, -1.f, panel, ldf, panel, ldf, 1.f, trail, ldf, 1, 0, st);
      if (rc) return rc;
    }
  }
  rc = check_launch();
  if (rc) return rc;

  // ---- triangular inverse by recursive doubling: nodes are [start, end) column ranges with a known inverse ----
  int starts[512], ends[512];
  if (nb > 512) return VLMC_ERR_UNSUPPORTED;
  int nn = nb;
  for (int k = 0; k < nb; ++k) { starts[k] = k * kNB; ends[k] = (k + 1) * kNB < C ? (k + 1) * kNB : C; }
  while (nn > 1) {
    int out = 0;
    for (int q = 0; q + 1 < nn; q += 2) {
      const int a = starts[q], b = ends[q], c = ends[q + 1];   // left [a,b), right [b,c)
      const int m = c - b, n = b - a;
      // X = L[b:c, a:b] * Li[a:b, a:b]
      rc = gemm3x(false, m, n, n, 1.f, F + (int64_t)b * ldf + a, ldf, U + (int64_t)a * ldu + a, ldu, 0.f, X, n, 0, 0, st);
      if (rc) return rc;
      // Li[b:c, a:b] = -Li[b:c, b:c] * X
      rc = gemm3x(false, m, n, m, -1.f, U + (int64_t)b * ldu + b, ldu, X, n, 0.f, U + (int64_t)b * ldu + a, ldu, 0, 0, st);
      if (rc) return rc;
      starts[out] = a; ends[out] = c; ++out;
    }
    if (nn & 1) { starts[out] = starts[nn - 1]; ends[out] = ends[nn - 1]; ++out; }
    nn = out;
  }
  // ---- U = J Li J ----
  flip_to_upper_kernel<<<dim3((C + 255) / 256, C), 256, 0, st>>>(U, ldu, C);
  return check_launch();
}
