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
#include <stdlib.h>
#include <atomic>
#include <map>
#include <mutex>
#include <utility>
#include "gemm3x.cuh"

namespace vlmc {

constexpr int kNB = 128;          // block size
constexpr int kPotrfThreads = 128;

// Both flips walk rows with a grid-stride loop (a few hundred CTAs in all): one CTA per row meant 473 k CTAs of 256
// threads at C = 11008 and 1 ms for a pass that moves 0.7 GB.
constexpr int kFlipUnroll = 4;    // rows in flight per thread: one 4-byte load per thread left the flips latency-bound

// flip != 0: F = J H J (the exchange-matrix flip of vlmc_chol_inv_upper), else F = H.  Only the lower triangle is moved
// (plus the rest of the 128-wide diagonal blocks, so that every element a diagonal GEMM tile touches is initialised):
// nothing of the blocked Cholesky reads above it.  A +-inf / NaN entry sets VLMC_NONFINITE in *status: the host then
// takes the reference's clamp path (sparsegpt_pruner.py:101-109) instead of damping a matrix that can never factorise.
__global__ void __launch_bounds__(256)
flip_copy_kernel(const float* __restrict__ H, int64_t ldh, float* __restrict__ F, int64_t ldf, int C, int flip,
                 int* __restrict__ status) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= C) return;
  bool bad = false;
  const int first = j & ~127;            // rows above the diagonal block of column j are never read
  for (int i0 = first + blockIdx.y; i0 < C; i0 += gridDim.y * kFlipUnroll) {
    float v[kFlipUnroll];
#pragma unroll
    for (int u = 0; u < kFlipUnroll; ++u) {
      const int i = i0 + u * (int)gridDim.y;
      if (i < C) v[u] = flip ? H[(int64_t)(C - 1 - i) * ldh + (C - 1 - j)] : H[(int64_t)i * ldh + j];
    }
#pragma unroll
    for (int u = 0; u < kFlipUnroll; ++u) {
      const int i = i0 + u * (int)gridDim.y;
      if (i < C) {
        F[(int64_t)i * ldf + j] = v[u];
        bad |= !isfinite(v[u]);
      }
    }
  }
  if (bad) atomicOr(status, (int)VLMC_NONFINITE);
}

// U = L^T: U[i][j] = L[j][i] for j >= i, zero below (vlmc_chol_upper); 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256)
lower_to_upper_transpose_kernel(const float* __restrict__ L, int64_t ldl, float* __restrict__ U, int64_t ldu, int C) {
  __shared__ float tile[32][33];
  const int bi = blockIdx.y * 32, bj = blockIdx.x * 32;          // output tile rows bi.., columns bj..
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (bj + 31 >= bi) {
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
      const int li = bj + r, lj = bi + tx;                         // L[bj + r][bi + tx]
      tile[r][tx] = (li < C && lj < C && lj <= li) ? L[(int64_t)li * ldl + lj] : 0.f;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int i = bi + r, j = bj + tx;
    if (i < C && j < C) U[(int64_t)i * ldu + j] = (j >= i && bj + 31 >= bi) ? tile[tx][r] : 0.f;
  }
}

// strictly-lower entries of Li move to the mirrored strictly-upper slot; the lower slot is cleared.  In place without a
// hazard: only strictly-lower entries (and the diagonal) are read, only strictly-upper slots receive values.
// An entry above kHugeFactor sets VLMC_HUGE_FACTOR: diag(H^-1) = column sums of squares of U may then overflow fp32, which
// is where the reference's second clamp / damping stage (:133-157) can differ from the fused factorisation.
constexpr float kHugeFactor = 1e15f;

__global__ void __launch_bounds__(256)
flip_to_upper_kernel(float* __restrict__ U, int64_t ldu, int C, int* __restrict__ status) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  bool huge = false;
  for (int i0 = blockIdx.y; i0 < C; i0 += gridDim.y * kFlipUnroll) {
    float v[kFlipUnroll];
#pragma unroll
    for (int u = 0; u < kFlipUnroll; ++u) {
      const int i = i0 + u * (int)gridDim.y;
      v[u] = (i < C && j < i) ? U[(int64_t)i * ldu + j] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kFlipUnroll; ++u) {
      const int i = i0 + u * (int)gridDim.y;
      if (i >= C) continue;
      if (j < i) {
        U[(int64_t)i * ldu + j] = 0.f;
        U[(int64_t)(C - 1 - i) * ldu + (C - 1 - j)] = v[u];
        huge |= !(fabsf(v[u]) <= kHugeFactor);
      } else if (j == i && i < (C + 1) / 2) {
        const int64_t a = (int64_t)i * ldu + i, b = (int64_t)(C - 1 - i) * ldu + (C - 1 - i);
        const float t = U[a], w = U[b];
        U[a] = w; U[b] = t;
        huge |= !(fabsf(t) <= kHugeFactor) || !(fabsf(w) <= kHugeFactor);
      }
    }
  }
  if (huge) atomicOr(status, (int)VLMC_HUGE_FACTOR);
}

// One CTA factors the bs x bs diagonal block at (k0,k0) of F and inverts the factor, all in shared memory:
// writes L_kk back to F and L_kk^-1 (lower, zero above the diagonal) to Li at the same position.
// Thread t owns ROW t of the block.  Columns are processed 16 at a time: the thread's 16 entries live in registers,
// the update with all previous columns is a left-looking pass over a transposed copy of L (one private value + four
// broadcast 16-byte loads per k, 16 independent FMAs), the 16 columns themselves are a right-looking sweep with ONE
// __syncthreads per column (the pivot column is exchanged through a double-buffered 16-float line).  The inverse
// needs no synchronisation at all: thread c owns COLUMN c of L^-1 and forward-substitutes 16 rows at a time.
// Blocks smaller than 128 are padded with the identity.
constexpr int kPB = 16;
constexpr int kLdS = kNB + 1;
constexpr size_t kPotrfSmem = ((size_t)kNB * kLdS + 2 * (size_t)kNB * kNB + 2 * kPB + kNB) * sizeof(float);

constexpr int kPotrfIoThreads = 512;     // all warps move the block in and out; the first 128 threads compute

__device__ __forceinline__ void potrf_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
// named barriers of the factor kernel: 1 = F's column barrier, 2-3 "block b - 2 final" (F arrives, U1 + U2 wait), 4-5 "look-ahead
// of block b in S" (U1 + U2 arrive, F waits), 6-13 "row block b may be inverted" (F arrives, I waits; one id per block: I may
// trail F by several blocks)
constexpr int kBarDone0 = 2, kBarReady0 = 4, kBarInv0 = 6;
__device__ __forceinline__ void potrf_named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void potrf_named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(count) : "memory"); }

__global__ void __launch_bounds__(kPotrfIoThreads)
potrf_block_kernel(float* F, int64_t ldf, float* Li, int64_t ldi, int k0, int bs, int* status) {
  extern __shared__ __align__(16) float sm[];
  float* S = sm;                          // [128][129]  A, then L, row-major
  float* St = S + kNB * kLdS;             // [128][128]  St[k][row] = L[row][k]
  float* Xs = St + kNB * kNB;             // [128][128]  Xs[k][c] = (L^-1)[k][c]
  float* colbuf = Xs + kNB * kNB;         // [2][16]
  float* invd = colbuf + 2 * kPB;         // [128]  1 / L[k][k]
  const int t = threadIdx.x;
  {
    // 128 x 128 block = 4096 float4, 8 per thread, all loads in flight together (k0, ldf are multiples of 4)
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = t + u * kPotrfIoThreads, i = idx >> 5, j4 = (idx & 31) * 4;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < bs && j4 <= i) v[u] = *reinterpret_cast<const float4*>(F + (int64_t)(k0 + i) * ldf + k0 + j4);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = t + u * kPotrfIoThreads, i = idx >> 5, j4 = (idx & 31) * 4;
      const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = j4 + q;
        S[i * kLdS + j] = (i < bs && j <= i) ? e[q] : (i == j ? 1.f : 0.f);
      }
      *reinterpret_cast<float4*>(Xs + idx * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __syncthreads();

  // Four groups of four warps (named barriers, no CTA-wide synchronisation inside the factorisation):
  //   F   threads   0-127  the 128 dependent pivots: per 16-column block the update with the PREVIOUS block's columns, then the
  //                        right-looking sweep (one 128-thread barrier per column)
  //   U1  threads 128-255  look-ahead: while F sweeps block b - 1 they apply the columns of blocks 0 .. b - 2 to block b
  //   U2  threads 384-511  (columns 0-7 / 8-15 of the block), in the same ascending-k fma order as F would - the factor is
  //                        bit-identical to the one-group kernel this replaces
  //   I   threads 256-383  the inverse, row block b as soon as F has finished column block b (it trails F and ends ~one
  //                        block after it)
  // Before: F did all of it in sequence (55 us per block at the chain's clock, 44 % of a C = 4096 factorisation).
  const int grp = t >> 7, tt = t & 127;
  if (grp == 0) {
    bool bad = false;
    const int wend = (tt | 31) + 1;          // rows of this warp end here: nothing to update once c0 >= wend
#pragma unroll 1
    for (int c0 = 0; c0 < kNB; c0 += kPB) {
      const int b = c0 / kPB;
      if (b >= 2) potrf_named_sync(kBarReady0 + (b & 1), 384);          // the look-ahead for this block is in S
      float a[kPB];
#pragma unroll
      for (int c = 0; c < kPB; ++c) a[c] = S[tt * kLdS + c0 + c];
      // left-looking, the previous block's columns only: a[c] -= sum_{c0 - 16 <= k < c0} L[t][k] * L[c0 + c][k]
      if (b >= 1 && c0 < wend) {
#pragma unroll 4
        for (int k = c0 - kPB; k < c0; ++k) {
          const float lk = St[k * kNB + tt];
          const float4* rp = reinterpret_cast<const float4*>(St + k * kNB + c0);
          const float4 r0 = rp[0], r1 = rp[1], r2 = rp[2], r3 = rp[3];
          const float r[kPB] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w, r3.x, r3.y, r3.z, r3.w};
#pragma unroll
          for (int c = 0; c < kPB; ++c) a[c] = fmaf(-lk, r[c], a[c]);
        }
      }
      // the 16 columns of this block, right-looking, one barrier per column
#pragma unroll
      for (int j = 0; j < kPB; ++j) {
        float* buf = colbuf + (j & 1) * kPB;
        if (tt >= c0 + j && tt < c0 + kPB) buf[tt - c0] = a[j];
        potrf_bar();
        float d = buf[j];
        if (!(d > 0.f) || !isfinite(d)) { bad = true; d = 1.f; }
        const float piv = sqrtf(d);
        const float inv = 1.f / piv;
        if (tt == c0 + j) { a[j] = piv; invd[c0 + j] = inv; }
        else a[j] *= inv;                                   // L[t][c0 + j] for t > c0 + j (rows above are never stored)
#pragma unroll
        for (int c = j + 1; c < kPB; ++c) a[c] = fmaf(-a[j], buf[c] * inv, a[c]);
      }
#pragma unroll
      for (int c = 0; c < kPB; ++c) {
        const bool low = tt >= c0 + c;
        if (low) S[tt * kLdS + c0 + c] = a[c];
        St[(c0 + c) * kNB + tt] = low ? a[c] : 0.f;
      }
      potrf_bar();
      if (b <= kNB / kPB - 3) potrf_named_arrive(kBarDone0 + (b & 1), 384);   // U1 / U2 may use these columns
      potrf_named_arrive(kBarInv0 + b, 256);                                   // I may invert row block b
    }
    if (bad && tt == 0) atomicOr(status, (int)VLMC_NOT_POSDEF);
  } else if (grp == 1 || grp == 3) {
    // look-ahead for block b: S[t][c0 + 8h + c] -= sum_{k < c0 - 16} L[t][k] * L[c0 + 8h + c][k], k ascending
    const int h = grp == 1 ? 0 : 1;
    const int wend = (tt | 31) + 1;
#pragma unroll 1
    for (int b = 2; b < kNB / kPB; ++b) {
      const int c0 = b * kPB, cb = c0 + 8 * h;
      potrf_named_sync(kBarDone0 + (b & 1), 384);                              // block b - 2 is final
      if (c0 < wend) {
        float a[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) a[c] = S[tt * kLdS + cb + c];
#pragma unroll 4
        for (int k = 0; k < c0 - kPB; ++k) {
          const float lk = St[k * kNB + tt];
          const float4* rp = reinterpret_cast<const float4*>(St + k * kNB + cb);
          const float4 r0 = rp[0], r1 = rp[1];
          const float r[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
          for (int c = 0; c < 8; ++c) a[c] = fmaf(-lk, r[c], a[c]);
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) S[tt * kLdS + cb + c] = a[c];
      }
      potrf_named_arrive(kBarReady0 + (b & 1), 384);
    }
  } else {
    // inverse: column c of X = L^-1 by blocked forward substitution; reads only L and this thread's own column
    const int c = tt;
    const int kbeg = (tt >> 5) * 32;                      // X[k][c] = 0 for k < c; uniform per warp
#pragma unroll 1
    for (int r0 = 0; r0 < kNB; r0 += kPB) {
      potrf_named_sync(kBarInv0 + r0 / kPB, 256);         // rows r0 .. r0 + 15 of L and their 1 / diagonal are final
      if (r0 < kbeg) continue;
      float acc[kPB];
#pragma unroll
      for (int r = 0; r < kPB; ++r) acc[r] = 0.f;
#pragma unroll 4
      for (int k = kbeg; k < r0; ++k) {
        const float xk = Xs[k * kNB + c];
        const float4* rp = reinterpret_cast<const float4*>(St + k * kNB + r0);
        const float4 q0 = rp[0], q1 = rp[1], q2 = rp[2], q3 = rp[3];
        const float l[kPB] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
#pragma unroll
        for (int r = 0; r < kPB; ++r) acc[r] = fmaf(l[r], xk, acc[r]);
      }
      float x[kPB];
#pragma unroll
      for (int r = 0; r < kPB; ++r) {
        float sacc = ((r0 + r == c) ? 1.f : 0.f) - acc[r];
#pragma unroll
        for (int q = 0; q < r; ++q) sacc = fmaf(-S[(r0 + r) * kLdS + r0 + q], x[q], sacc);
        x[r] = sacc * invd[r0 + r];
      }
#pragma unroll
      for (int r = 0; r < kPB; ++r) Xs[(r0 + r) * kNB + c] = x[r];
    }
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int idx = t + u * kPotrfIoThreads, i = idx >> 5, j4 = (idx & 31) * 4;
    if (i < bs && j4 < bs) {                              // bs is a multiple of 4
      float4 lo, xo;
      float* lp = &lo.x;
      float* xp = &xo.x;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = j4 + q;
        lp[q] = S[i * kLdS + j];
        xp[q] = (j <= i) ? Xs[i * kNB + j] : 0.f;
      }
      if (j4 <= i) {
        float* fp = F + (int64_t)(k0 + i) * ldf + k0 + j4;
        if (j4 + 3 <= i) *reinterpret_cast<float4*>(fp) = lo;
        else {
#pragma unroll
          for (int q = 0; q < 4; ++q) if (j4 + q <= i) fp[q] = lp[q];
        }
      }
      *reinterpret_cast<float4*>(Li + (int64_t)(k0 + i) * ldi + k0 + j4) = xo;
    }
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

// Side stream + fork/join events of the look-ahead, one set per (device, caller stream): chains that run concurrently on
// different streams must not share a side stream (their trailing updates would queue behind each other).
static std::atomic<int> g_chol_lookahead{-1};             // -1: follow VLMC_CHOL_LOOKAHEAD (default on); 0 / 1: set by the host

bool chain_lookahead_enabled() {
  const int m = g_chol_lookahead.load();
  if (m >= 0) return m != 0;
  const char* e = getenv("VLMC_CHOL_LOOKAHEAD");          // read per call: tests flip it inside one process
  return !(e && e[0] == '0');
}

ChainSide* chain_side_for(cudaStream_t st, int user) {
  static std::mutex mu;
  static std::map<std::pair<std::pair<int, int>, cudaStream_t>, ChainSide> sides;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_pair(std::make_pair(dev, user), st);
  auto it = sides.find(key);
  if (it != sides.end()) return &it->second;
  ChainSide s;
  if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (cudaEventCreateWithFlags(&s.solved, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&s.updated, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return &sides.emplace(key, s).first->second;
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

extern "C" int vlmc_chol_set_lookahead(int mode) {
  if (mode < -1 || mode > 1) return VLMC_ERR_BAD_ARG;
  const int prev = vlmc::g_chol_lookahead.exchange(mode);
  return prev < 0 ? 2 : prev;
}

namespace vlmc {

// Panels per super-panel of the two-level schedule below (VLMC_CHOL_SUPERPANEL = 1 / 2 / 4 / 8, read per call; 1 = the plain
// right-looking loop with one K = 128 trailing update per panel).
static int chol_superpanel() {
  if (const char* e = getenv("VLMC_CHOL_SUPERPANEL")) {
    const int v = atoi(e);
    if (v == 1 || v == 2 || v == 4 || v == 8) return v;
  }
  return 1;       // measured on B200 (scripts/chol_sp_probe.py): 2 / 4 / 8 are SLOWER than the plain loop, see below
}

// Two-level right-looking schedule (experiment, off by default: C = 4096 3.20 -> 3.38 ms, C = 11008 15.0-15.7 -> 16.4-16.7 ms
// for sp = 2 / 4 / 8, profiles/r02ax_chol_sp_probe.log - the K = sp * 128 update runs on the register-summing 128-wide tiles,
// which cost more than the passes over the trailing matrix they save; results differ from the plain loop by 4e-7).  The plain loop read-modify-writes the whole fp32 trailing matrix once per 128-wide
// panel (K = 128: 13.9 GB of traffic and 5.7 of the 17.3 ms at C = 11008, a 3xTF32 GEMM that is all epilogue).  Here `sp`
// panels form a super-panel: a panel's update goes at once only to the remaining columns of ITS super-panel (<= (sp - 1) * 128
// wide), and the matrix behind the super-panel receives all of its panels in ONE update with K = sp * 128 (a quarter of the
// passes at sp = 4; the same products per output element, the chunks of 128 summed round-to-nearest in fp32 registers before
// the one subtraction instead of four subtractions).  Look-ahead as in the plain loop, at super-panel granularity: the columns of the
// NEXT super-panel are updated on the caller's stream, the rest on the side stream under the next super-panel's factor.
static int chol_lower_superpanel(float* F, int64_t ldf, float* Li, int64_t ldi, int C, int* status, cudaStream_t st, int sp,
                                 int potrf_smem) {
  const int nb = (C + kNB - 1) / kNB;
  int rc;
  ChainSide* side = chain_lookahead_enabled() && nb > 2 * sp ? chain_side_for(st, 0) : nullptr;
  bool pending_b = false;
  for (int s0 = 0; s0 < nb; s0 += sp) {
    const int s1 = s0 + sp < nb ? s0 + sp : nb;                  // panels [s0, s1)
    const int e = s1 * kNB < C ? s1 * kNB : C;                   // one past the super-panel's last column
    for (int k = s0; k < s1; ++k) {
      const int k0 = k * kNB;
      const int bs = (C - k0 < kNB) ? (C - k0) : kNB;
      potrf_block_kernel<<<1, kPotrfIoThreads, potrf_smem, st>>>(F, ldf, Li, ldi, k0, bs, status);
      const int below = C - k0 - bs;
      if (below <= 0) continue;
      float* panel = F + (int64_t)(k0 + bs) * ldf + k0;
      rc = gemm3x(true, below, bs, bs, 1.f, panel, ldf, Li + (int64_t)k0 * ldi + k0, ldi, 0.f, panel, ldf, 0, 0, st);
      if (rc) return rc;
      const int cols_in = e - (k0 + bs);                          // what is left of this super-panel
      if (cols_in > 0) {
        rc = gemm3x(true, below, cols_in, bs, -1.f, panel, ldf, panel, ldf, 1.f, F + (int64_t)(k0 + bs) * ldf + (k0 + bs), ldf,
                    1, 0, st);
        if (rc) return rc;
      }
    }
    const int rows_rem = C - e;
    if (rows_rem <= 0) continue;
    const float* P = F + (int64_t)e * ldf + (int64_t)s0 * kNB;    // L[e:, super-panel columns]
    const int K = e - s0 * kNB;
    float* trail = F + (int64_t)e * ldf + e;
    const int next_cols = sp * kNB < rows_rem ? sp * kNB : rows_rem;
    const int rest = rows_rem - next_cols;
    if (!side || rest <= 0) {
      if (pending_b) {
        if (cudaStreamWaitEvent(st, side->updated, 0) != cudaSuccess) return check_launch();
        pending_b = false;
      }
      rc = gemm3x(true, rows_rem, rows_rem, K, -1.f, P, ldf, P, ldf, 1.f, trail, ldf, 1, 0, st);
      if (rc) return rc;
      continue;
    }
    if (cudaEventRecord(side->solved, st) != cudaSuccess) return check_launch();
    if (pending_b) {                                             // the previous rest-update wrote what this one updates
      if (cudaStreamWaitEvent(st, side->updated, 0) != cudaSuccess) return check_launch();
      pending_b = false;
    }
    rc = gemm3x(true, rows_rem, next_cols, K, -1.f, P, ldf, P, ldf, 1.f, trail, ldf, 1, 0, st);
    if (rc) return rc;
    if (cudaStreamWaitEvent(side->stream, side->solved, 0) != cudaSuccess) return check_launch();
    const float* Prest = P + (int64_t)next_cols * ldf;
    rc = gemm3x(true, rest, rest, K, -1.f, Prest, ldf, Prest, ldf, 1.f, trail + (int64_t)next_cols * ldf + next_cols, ldf, 1, 0,
                side->stream, 0, false, kNumSMs - 1);
    if (rc) return rc;
    if (cudaEventRecord(side->updated, side->stream) != cudaSuccess) return check_launch();
    pending_b = true;
  }
  if (pending_b && cudaStreamWaitEvent(st, side->updated, 0) != cudaSuccess) return check_launch();
  return check_launch();
}

// Blocked lower Cholesky of F in place (F = L L^T on the lower triangle), 128-wide panels; the inverse of every diagonal
// block goes to the same position of Li.  A non-positive / NaN pivot sets VLMC_NOT_POSDEF in *status.
// Look-ahead (default; VLMC_CHOL_LOOKAHEAD=0 restores the plain right-looking loop): the trailing update of step k is
// split into the next block column (A_k, on the caller's stream: all that the next diagonal factor and panel solve
// need) and the rest (B_k, on a side stream forked from and joined to the caller's stream with events).  B_k leaves
// one SM free, so the one-CTA diagonal factor of step k+1 (55 us, a fifth of a step) runs under it instead of after it:
//   caller's stream:  P_k  S_k  [wait B_k-1]  A_k        P = diagonal factor + inverse, S = panel solve
//   side stream:      [wait S_k]  B_k                    (B_k after B_k-1 by stream order: same output region)
// Same GEMM per output element either way (K = 128, one chunk).
static int chol_lower_blocked(float* F, int64_t ldf, float* Li, int64_t ldi, int C, int* status, cudaStream_t st) {
  static bool attr_set = false;
  const int potrf_smem = (int)kPotrfSmem;
  if (!attr_set) {
    if (cudaFuncSetAttribute(potrf_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, potrf_smem) != cudaSuccess)
      return check_launch();
    attr_set = true;
  }
  const int nb = (C + kNB - 1) / kNB;
  int rc;
  const int sp = chol_superpanel();
  if (sp > 1 && nb > sp) return chol_lower_superpanel(F, ldf, Li, ldi, C, status, st, sp, potrf_smem);
  ChainSide* side = chain_lookahead_enabled() && nb > 2 ? chain_side_for(st, 0) : nullptr;
  bool pending_b = false;
  for (int k = 0; k < nb; ++k) {
    const int k0 = k * kNB;
    const int bs = (C - k0 < kNB) ? (C - k0) : kNB;
    potrf_block_kernel<<<1, kPotrfIoThreads, potrf_smem, st>>>(F, ldf, Li, ldi, k0, bs, status);
    const int below = C - k0 - bs;
    if (below > 0) {
      float* panel = F + (int64_t)(k0 + bs) * ldf + k0;
      // panel <- panel * L_kk^-T   (C = A B^T with B = L_kk^-1 stored [N,K])
      rc = gemm3x(true, below, bs, bs, 1.f, panel, ldf, Li + (int64_t)k0 * ldi + k0, ldi, 0.f, panel, ldf, 0, 0, st);
      if (rc) return rc;
      // trailing <- trailing - panel panel^T on the lower tiles
      float* trail = F + (int64_t)(k0 + bs) * ldf + (k0 + bs);
      const int bs2 = below < kNB ? below : kNB;          // width of the next block column
      if (!side) {
        rc = gemm3x(true, below, below, bs, -1.f, panel, ldf, panel, ldf, 1.f, trail, ldf, 1, 0, st);
        if (rc) return rc;
        continue;
      }
      const int rest = below - bs2;
      if (rest > 0 && cudaEventRecord(side->solved, st) != cudaSuccess) return check_launch();
      if (pending_b) {                                     // B_k-1 wrote the block column A_k is about to update
        if (cudaStreamWaitEvent(st, side->updated, 0) != cudaSuccess) return check_launch();
        pending_b = false;
      }
      rc = gemm3x(true, below, bs2, bs, -1.f, panel, ldf, panel, ldf, 1.f, trail, ldf, 0, 0, st);
      if (rc) return rc;
      if (rest > 0) {
        if (cudaStreamWaitEvent(side->stream, side->solved, 0) != cudaSuccess) return check_launch();
        const float* prest = panel + (int64_t)bs2 * ldf;
        rc = gemm3x(true, rest, rest, bs, -1.f, prest, ldf, prest, ldf, 1.f, trail + (int64_t)bs2 * ldf + bs2, ldf, 1, 0,
                    side->stream, 0, false, kNumSMs - 1);
        if (rc) return rc;
        if (cudaEventRecord(side->updated, side->stream) != cudaSuccess) return check_launch();
        pending_b = true;
      }
    }
  }
  if (pending_b && cudaStreamWaitEvent(st, side->updated, 0) != cudaSuccess) return check_launch();
  return check_launch();
}

static int chol_arg_checks(const float* H, int C, int64_t ldh, float* U, int64_t ldu, int* status, void* ws, size_t ws_bytes) {
  if (!H || !U || !status || !ws || C < 1 || ldh < C || ldu < C) return VLMC_ERR_BAD_ARG;
  if ((C & 3) || (ldh & 3) || (ldu & 3) || ((uintptr_t)H & 15) || ((uintptr_t)U & 15)) return VLMC_ERR_UNSUPPORTED;
  if (!is_device_ptr(H) || !is_device_ptr(U) || !is_device_ptr(status) || !is_device_ptr(ws)) return VLMC_ERR_NOT_DEVICE;
  if (ws_bytes < chol_workspace_bytes(C)) return VLMC_ERR_WORKSPACE;
  return VLMC_OK;
}

}  // namespace vlmc

extern "C" int vlmc_chol_inv_upper(const float* H, int C, int64_t ldh, float* U, int64_t ldu, int* status,
                                   void* ws, size_t ws_bytes, void* stream) {
  using namespace vlmc;
  int rc = chol_arg_checks(H, C, ldh, U, ldu, status, ws, ws_bytes);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  char* base = reinterpret_cast<char*>(ws) + VLMC_WS_COUNTER_BYTES;
  float* F = reinterpret_cast<float*>(base);
  const int64_t ldf = C;
  float* X = reinterpret_cast<float*>(base + align_up((size_t)C * C * sizeof(float), 256));
  const int nb = (C + kNB - 1) / kNB;

  if (cudaMemsetAsync(status, 0, sizeof(int), st) != cudaSuccess) return check_launch();
  if (cudaMemset2DAsync(U, ldu * sizeof(float), 0, (size_t)C * sizeof(float), C, st) != cudaSuccess) return check_launch();
  const int flip_rows = C < kNumSMs * 4 ? C : kNumSMs * 4;
  flip_copy_kernel<<<dim3((C + 255) / 256, flip_rows), 256, 0, st>>>(H, ldh, F, ldf, C, 1, status);

  // ---- blocked Cholesky of F (lower), diagonal-block inverses go straight into U (used as Li) ----
  rc = chol_lower_blocked(F, ldf, U, ldu, C, status, st);
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
      rc = gemm3x(false, m, n, n, 1.f, F + (int64_t)b * ldf + a, ldf, U + (int64_t)a * ldu + a, ldu, 0.f, X, n, 0, 0, st, 1);
      if (rc) return rc;
      // Li[b:c, a:b] = -Li[b:c, b:c] * X
      rc = gemm3x(false, m, n, m, -1.f, U + (int64_t)b * ldu + b, ldu, X, n, 0.f, U + (int64_t)b * ldu + a, ldu, 0, 0, st, 2);
      if (rc) return rc;
      starts[out] = a; ends[out] = c; ++out;
    }
    if (nn & 1) { starts[out] = starts[nn - 1]; ends[out] = ends[nn - 1]; ++out; }
    nn = out;
  }
  // ---- U = J Li J ----
  flip_to_upper_kernel<<<dim3((C + 255) / 256, flip_rows), 256, 0, st>>>(U, ldu, C, status);
  return check_launch();
}

// ---------------------------------------------------------------------------------------------------------------------
// The reference's SECOND stage in its own order (sparsegpt_pruner.py:131-157), for the Hessians the fused factorisation
// flags (VLMC_HUGE_FACTOR) or on request: Hinv = U0^T U0 (= cholesky_inverse), the +-inf clamp and the damp-and-retry
// loop on cholesky(Hinv, upper=True) are then the caller's, built from these pieces.
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int vlmc_gram_upper(const float* U, int C, int64_t ldu, float* Hinv, int64_t ldh, void* stream) {
  using namespace vlmc;
  if (!U || !Hinv || C < 1 || ldu < C || ldh < C) return VLMC_ERR_BAD_ARG;
  if ((C & 3) || (ldu & 3) || (ldh & 3) || ((uintptr_t)U & 15) || ((uintptr_t)Hinv & 15)) return VLMC_ERR_UNSUPPORTED;
  if (!is_device_ptr(U) || !is_device_ptr(Hinv)) return VLMC_ERR_NOT_DEVICE;
  // Hinv[i][j] = sum_k U[k][i] U[k][j]: A given as [K, M] (a_km), B as [K, N]
  return gemm3x(false, C, C, C, 1.f, U, ldu, U, ldu, 0.f, Hinv, ldh, 0, 0, (cudaStream_t)stream, 0, true);
}

extern "C" int vlmc_chol_upper(const float* A, int C, int64_t lda, float* U, int64_t ldu, int* status,
                               void* ws, size_t ws_bytes, void* stream) {
  using namespace vlmc;
  int rc = chol_arg_checks(A, C, lda, U, ldu, status, ws, ws_bytes);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  float* F = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + VLMC_WS_COUNTER_BYTES);
  const int64_t ldf = C;
  if (cudaMemsetAsync(status, 0, sizeof(int), st) != cudaSuccess) return check_launch();
  const int flip_rows = C < kNumSMs * 4 ? C : kNumSMs * 4;
  flip_copy_kernel<<<dim3((C + 255) / 256, flip_rows), 256, 0, st>>>(A, lda, F, ldf, C, 0, status);
  rc = chol_lower_blocked(F, ldf, U, ldu, C, status, st);      // U's diagonal blocks hold L_kk^-1 until the transpose
  if (rc) return rc;
  lower_to_upper_transpose_kernel<<<dim3((C + 31) / 32, (C + 31) / 32), 256, 0, st>>>(F, ldf, U, ldu, C);
  return check_launch();
}

namespace vlmc {

__global__ void __launch_bounds__(256)
nonfinite_count_kernel(const float* __restrict__ A, int rows, int cols, int64_t lda, unsigned long long* __restrict__ out) {
  unsigned int pos = 0, neg = 0, nan = 0;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < (int64_t)rows * cols;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const float v = A[(idx / cols) * lda + (idx % cols)];
    if (isinf(v)) { if (v > 0.f) ++pos; else ++neg; }
    else if (v != v) ++nan;
  }
  if (pos) atomicAdd(out + 0, (unsigned long long)pos);
  if (neg) atomicAdd(out + 1, (unsigned long long)neg);
  if (nan) atomicAdd(out + 2, (unsigned long long)nan);
}

__global__ void __launch_bounds__(256)
replace_inf_kernel(float* __restrict__ A, int rows, int cols, int64_t lda, float value, int negative) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < (int64_t)rows * cols;
       idx += (int64_t)gridDim.x * blockDim.x) {
    float* p = A + (idx / cols) * lda + (idx % cols);
    const float v = *p;
    if (isinf(v) && ((v < 0.f) == (negative != 0))) *p = value;
  }
}

__global__ void diag_abs_mean_kernel(const float* A, int64_t lda, int C, float scale, float* out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < C; i += blockDim.x) s += (double)fabsf(A[(int64_t)i * lda + i]);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    *out = scale * (float)(t / (double)C);
  }
}

}  // namespace vlmc

// out[0..2] = number of +inf, -inf, NaN entries of A (the tests of sparsegpt_pruner.py:101,106,133,138 in one pass)
extern "C" int vlmc_matrix_nonfinite_count(const float* A, int rows, int cols, int64_t lda, unsigned long long* out3,
                                           void* stream) {
  using namespace vlmc;
  if (!A || !out3 || rows < 1 || cols < 1 || lda < cols) return VLMC_ERR_BAD_ARG;
  if (!is_device_ptr(A) || !is_device_ptr(out3)) return VLMC_ERR_NOT_DEVICE;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(out3, 0, 3 * sizeof(unsigned long long), st) != cudaSuccess) return check_launch();
  nonfinite_count_kernel<<<kNumSMs * 8, 256, 0, st>>>(A, rows, cols, lda, out3);
  return check_launch();
}

// A[A == +inf] = value (negative == 0) or A[A == -inf] = value (negative != 0): sparsegpt_pruner.py:103-104 / :108-109
extern "C" int vlmc_matrix_replace_inf(float* A, int rows, int cols, int64_t lda, float value, int negative, void* stream) {
  using namespace vlmc;
  if (!A || rows < 1 || cols < 1 || lda < cols) return VLMC_ERR_BAD_ARG;
  if (!is_device_ptr(A)) return VLMC_ERR_NOT_DEVICE;
  replace_inf_kernel<<<kNumSMs * 8, 256, 0, (cudaStream_t)stream>>>(A, rows, cols, lda, value, negative);
  return check_launch();
}

// *out = scale * mean(|diag(A)|): the second-stage damp of sparsegpt_pruner.py:143 (and :111 after a clamp)
extern "C" int vlmc_diag_abs_mean(const float* A, int C, int64_t lda, float scale, float* out, void* stream) {
  using namespace vlmc;
  if (!A || !out || C < 1 || lda < C) return VLMC_ERR_BAD_ARG;
  if (!is_device_ptr(A) || !is_device_ptr(out)) return VLMC_ERR_NOT_DEVICE;
  diag_abs_mean_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(A, lda, C, scale, out);
  return check_launch();
}
