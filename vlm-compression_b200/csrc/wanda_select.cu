// K4/K5/K6: Wanda score fused with mask selection (score is never materialised).
//
// Replaces lavis/compression/pruners/wanda_pruner.py:318-341:
//   W_metric = |W| * sqrt(scaler_row)                     (:318)  fp32, one IEEE multiply
//   importance_score = mean(W_metric)                     (:320)
//   unstructured: stable sort per row, first int(C*p) pruned   (:332-337)   -> rowselect_kernel
//   n:m: python loop over C/m groups of torch.topk          (:323-329)   -> nm_kernel
//   module.mask = ~W_mask ; W[W_mask] = 0                  (:339-341)
//
// rowselect: one CTA (128 threads) owns a row; the fp32 score bits live in REGISTERS
// (non-negative floats order like their uint32 bit patterns).  The k-th smallest
// (score, column) pair is found by counting passes: 4 pivots per pass, counts packed in
// 16-bit lanes and reduced with redux.sync + one __syncthreads.  The first pass of a row
// is seeded from the previous row's threshold rescaled by the row mean, so iid rows
// converge in 1-2 passes; the bracket is closed exactly by ranking the <=128 remaining
// candidates on (score, column), which also gives the stable-sort tie-break.
// HBM-bound: 5 B/weight for fp16/bf16 (read W, write W, write mask).
#include "common.cuh"

namespace vlmc {

constexpr int kSelThreads = 128;
constexpr int kSelWarps = kSelThreads / 32;
constexpr int kCap = 128;  // candidates ranked exhaustively

__device__ __forceinline__ uint32_t redux_add(uint32_t v) {
  return __reduce_add_sync(0xffffffffu, v);
}

struct SelShared {
  uint32_t red[2][kSelWarps][2];
  float fred[kSelWarps];
  uint32_t cand_key[kCap];
  int cand_idx[kCap];
  int cand_cnt;
  uint32_t v;
  int iv;
};

// counts of keys < p[0..3] over the whole CTA; every thread gets all four
template <int KPT>
__device__ __forceinline__ void count4(const uint32_t (&keys)[KPT], const uint32_t (&p)[4],
                                       int (&c)[4], SelShared& sh, int& parity) {
  uint32_t a = 0, b = 0;
#pragma unroll
  for (int i = 0; i < KPT; ++i) {
    const uint32_t key = keys[i];
    a += (key < p[0] ? 1u : 0u) + (key < p[1] ? 0x10000u : 0u);
    b += (key < p[2] ? 1u : 0u) + (key < p[3] ? 0x10000u : 0u);
  }
  a = redux_add(a);
  b = redux_add(b);
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { sh.red[parity][warp][0] = a; sh.red[parity][warp][1] = b; }
  __syncthreads();
  a = 0; b = 0;
#pragma unroll
  for (int w = 0; w < kSelWarps; ++w) { a += sh.red[parity][w][0]; b += sh.red[parity][w][1]; }
  parity ^= 1;
  c[0] = a & 0xffff; c[1] = a >> 16; c[2] = b & 0xffff; c[3] = b >> 16;
}

__device__ __forceinline__ uint32_t clampu(double x, uint32_t lo, uint32_t hi) {
  // pivot into [lo+1, hi-1] (caller guarantees hi - lo >= 2)
  if (!(x > (double)lo + 1.0)) return lo + 1;
  if (!(x < (double)hi - 1.0)) return hi - 1;
  return (uint32_t)x;
}

template <typename T, int KPT>
__global__ void __launch_bounds__(kSelThreads)
rowselect_kernel(T* __restrict__ W, int64_t ldw, int R, int C,
                 const float* __restrict__ scaler_row, int k, int zero_w,
                 uint8_t* __restrict__ mask, int64_t ldm, float* __restrict__ row_sum) {
  constexpr int V = Elem<T>::kVec;
  constexpr int NV = KPT / V;
  extern __shared__ float sq[];  // sqrt(scaler_row), [C]
  __shared__ SelShared sh;

  for (int c = threadIdx.x; c < C; c += kSelThreads) sq[c] = __fsqrt_rn(scaler_row[c]);
  __syncthreads();

  int parity = 0;
  float hint_ratio = -1.f;  // previous threshold / previous row mean

  for (int row = blockIdx.x; row < R; row += gridDim.x) {
    T* wrow = W + (int64_t)row * ldw;
    uint32_t keys[KPT];
    float lsum = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int col = (j * kSelThreads + threadIdx.x) * V;
      if (col < C) {
        uint4 v = ld_stream(wrow + col);
        float f[V];
        Elem<T>::unpack(v, f);
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const float s = __fmul_rn(fabsf(f[e]), sq[col + e]);
          keys[j * V + e] = __float_as_uint(s);
          lsum += s;
        }
      } else {
#pragma unroll
        for (int e = 0; e < V; ++e) keys[j * V + e] = 0xffffffffu;
      }
    }
    // row sum (importance score + pivot hint)
    lsum = warp_sum(lsum);
    if ((threadIdx.x & 31) == 0) sh.fred[threadIdx.x >> 5] = lsum;
    if (threadIdx.x == 0) sh.cand_cnt = 0;
    __syncthreads();
    float rsum = 0.f;
#pragma unroll
    for (int w = 0; w < kSelWarps; ++w) rsum += sh.fred[w];
    if (threadIdx.x == 0 && row_sum) row_sum[row] = rsum;
    const float rmean = rsum / (float)C;

    uint32_t v = 0;
    int iv = -1;  // prune (key < v) || (key == v && col <= iv)
    if (k >= C) {
      v = 0xffffffffu; iv = 0x7fffffff;
    } else if (k > 0) {
      uint32_t lo = 0, hi = 0xffffffffu;
      int glo = 0, ghi = C;
      bool first = true;
      while (true) {
        const int m = ghi - glo;
        if (m <= kCap || hi - lo == 1u) break;
        const uint32_t w = hi - lo;
        uint32_t p[4];
        const float hv = hint_ratio * rmean;
        if (first && hint_ratio > 0.f && hv > 0.f && hv < 3.0e38f) {
          const double e = (double)__float_as_uint(hv);
          p[0] = clampu(e - 524288.0, lo, hi); p[1] = clampu(e - 65536.0, lo, hi);
          p[2] = clampu(e + 65536.0, lo, hi);  p[3] = clampu(e + 524288.0, lo, hi);
        } else if (m > C / 2 && w > (1u << 26)) {
          const double q = (double)w * 0.2;
          p[0] = clampu(lo + q, lo, hi);       p[1] = clampu(lo + 2.0 * q, lo, hi);
          p[2] = clampu(lo + 3.0 * q, lo, hi); p[3] = clampu(lo + 4.0 * q, lo, hi);
        } else {
          const double f = ((double)(k - glo) - 0.5) / (double)m;
          const double e = (double)lo + f * (double)w;
          const double d1 = (double)w * (1.0 / 64.0), d2 = (double)w * 0.125;
          p[0] = clampu(e - d2, lo, hi); p[1] = clampu(e - d1, lo, hi);
          p[2] = clampu(e + d1, lo, hi); p[3] = clampu(e + d2, lo, hi);
        }
        first = false;
        int c[4];
        count4<KPT>(keys, p, c, sh, parity);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (c[i] <= k - 1) { if (p[i] > lo) { lo = p[i]; glo = c[i]; } }
          else               { if (p[i] < hi) { hi = p[i]; ghi = c[i]; } }
        }
      }
      const int m = ghi - glo;
      if (m <= kCap) {
        // gather the candidates lo <= key < hi and rank them on (key, column)
#pragma unroll
        for (int i = 0; i < KPT; ++i) {
          const uint32_t key = keys[i];
          if (key >= lo && key < hi) {
            const int slot = atomicAdd(&sh.cand_cnt, 1);
            sh.cand_key[slot] = key;
            sh.cand_idx[slot] = ((i / V) * kSelThreads + threadIdx.x) * V + (i % V);
          }
        }
        __syncthreads();
        if (threadIdx.x < m) {
          const uint32_t kt = sh.cand_key[threadIdx.x];
          const int it = sh.cand_idx[threadIdx.x];
          int rank = 0;
          for (int j = 0; j < m; ++j) {
            const uint32_t kj = sh.cand_key[j];
            rank += (kj < kt || (kj == kt && sh.cand_idx[j] < it)) ? 1 : 0;
          }
          if (rank == k - 1 - glo) { sh.v = kt; sh.iv = it; }
        }
        __syncthreads();
        v = sh.v; iv = sh.iv;
      } else {
        // more than kCap exact ties at the threshold value: bisect on the column index
        v = lo;
        const int r = k - glo;  // how many of the ties are pruned (lowest columns first)
        int qlo = 0, qhi = C;   // count(key==v && col < qlo) < r <= count(key==v && col < qhi)
        while (qhi - qlo > 1) {
          const int q = (qlo + qhi) >> 1;
          uint32_t cnt = 0;
#pragma unroll
          for (int i = 0; i < KPT; ++i) {
            const int col = ((i / V) * kSelThreads + threadIdx.x) * V + (i % V);
            cnt += (keys[i] == v && col < q) ? 1u : 0u;
          }
          cnt = redux_add(cnt);
          if ((threadIdx.x & 31) == 0) sh.red[parity][threadIdx.x >> 5][0] = cnt;
          __syncthreads();
          cnt = 0;
#pragma unroll
          for (int w2 = 0; w2 < kSelWarps; ++w2) cnt += sh.red[parity][w2][0];
          parity ^= 1;
          if ((int)cnt < r) qlo = q; else qhi = q;
        }
        iv = qlo;  // columns <= qlo with key == v are exactly the r lowest ties
      }
      const float vf = __uint_as_float(v);
      hint_ratio = (rmean > 0.f && vf > 0.f && vf < 3.0e38f) ? vf / rmean : -1.f;
    }

    // apply: mask bytes (1 = keep) and zeroed weights
    uint8_t* mrow = mask + (int64_t)row * ldm;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int col = (j * kSelThreads + threadIdx.x) * V;
      if (col < C) {
        uint32_t mb[V / 4] = {};
        bool any = false;
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const uint32_t key = keys[j * V + e];
          const bool pr = (key < v) || (key == v && (col + e) <= iv);
          any |= pr;
          mb[e / 4] |= (pr ? 0u : 1u) << (8 * (e % 4));
        }
        if (V == 8) st_stream8(mrow + col, make_uint2(mb[0], mb[V / 4 - 1]));
        else st_stream4(mrow + col, mb[0]);
        if (zero_w && any) {
          uint4 wv = *reinterpret_cast<const uint4*>(wrow + col);  // L2 hit: the row was just streamed
          float f[V];
          Elem<T>::unpack(wv, f);
          uint32_t* wr = reinterpret_cast<uint32_t*>(&wv);
          if (sizeof(T) == 4) {
#pragma unroll
            for (int e = 0; e < V; ++e) if (!((mb[e / 4] >> (8 * (e % 4))) & 1u)) wr[e] = 0u;
          } else {
#pragma unroll
            for (int e = 0; e < V; ++e)
              if (!((mb[e / 4] >> (8 * (e % 4))) & 1u)) wr[e / 2] &= (e & 1) ? 0x0000ffffu : 0xffff0000u;
          }
          st_stream(wrow + col, wv);
        }
      }
    }
    __syncthreads();  // sh.cand_cnt / sh.v reuse
  }
}

// ---- n:m ---------------------------------------------------------------------------------
template <typename T, int M>
__global__ void __launch_bounds__(kSelThreads)
nm_kernel(T* __restrict__ W, int64_t ldw, int R, int C, const float* __restrict__ scaler_row,
          int n, int zero_w, uint8_t* __restrict__ mask, int64_t ldm, float* __restrict__ part_sum) {
  constexpr int V = Elem<T>::kVec;
  constexpr int E = (M > V) ? M : V;   // elements per thread step
  constexpr int NVEC = E / V;
  const int col = (blockIdx.x * kSelThreads + threadIdx.x) * E;
  const bool ok = col < C;
  float sq[E];
#pragma unroll
  for (int e = 0; e < E; ++e) sq[e] = ok ? __fsqrt_rn(scaler_row[col + e]) : 0.f;
  float lsum = 0.f;
  if (ok) {
    for (int row = blockIdx.y; row < R; row += gridDim.y) {
      T* wp = W + (int64_t)row * ldw + col;
      uint4 wv[NVEC];
#pragma unroll
      for (int q = 0; q < NVEC; ++q) wv[q] = ld_stream(wp + q * V);
      uint32_t keys[E];
#pragma unroll
      for (int q = 0; q < NVEC; ++q) {
        float f[V];
        Elem<T>::unpack(wv[q], f);
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const float s = __fmul_rn(fabsf(f[e]), sq[q * V + e]);
          keys[q * V + e] = __float_as_uint(s);
          lsum += s;
        }
      }
      bool pr[E];
#pragma unroll
      for (int g = 0; g < E / M; ++g) {
#pragma unroll
        for (int a = 0; a < M; ++a) {
          int rank = 0;
#pragma unroll
          for (int b = 0; b < M; ++b) {
            if (b < a) rank += keys[g * M + b] <= keys[g * M + a] ? 1 : 0;
            if (b > a) rank += keys[g * M + b] < keys[g * M + a] ? 1 : 0;
          }
          pr[g * M + a] = rank < n;
        }
      }
      uint8_t* mp = mask + (int64_t)row * ldm + col;
#pragma unroll
      for (int q = 0; q < NVEC; ++q) {
        uint32_t mb[V / 4] = {};
#pragma unroll
        for (int e = 0; e < V; ++e) mb[e / 4] |= (pr[q * V + e] ? 0u : 1u) << (8 * (e % 4));
        if (V == 8) st_stream8(mp + q * V, make_uint2(mb[0], mb[V / 4 - 1]));
        else st_stream4(mp + q * V, mb[0]);
        if (zero_w) {
          uint32_t* wr = reinterpret_cast<uint32_t*>(&wv[q]);
          if (sizeof(T) == 4) {
#pragma unroll
            for (int e = 0; e < V; ++e) if (pr[q * V + e]) wr[e] = 0u;
          } else {
#pragma unroll
            for (int e = 0; e < V; ++e)
              if (pr[q * V + e]) wr[e / 2] &= (e & 1) ? 0x0000ffffu : 0xffff0000u;
          }
          st_stream(wp + q * V, wv[q]);
        }
      }
    }
  }
  __shared__ float fred[kSelWarps];
  lsum = warp_sum(lsum);
  if ((threadIdx.x & 31) == 0) fred[threadIdx.x >> 5] = lsum;
  __syncthreads();
  if (threadIdx.x == 0 && part_sum) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kSelWarps; ++w) s += fred[w];
    part_sum[blockIdx.y * gridDim.x + blockIdx.x] = s;
  }
}

// partial sums -> mean, single CTA, fixed order => deterministic
__global__ void __launch_bounds__(256)
mean_finalize_kernel(const float* __restrict__ part, int n, double denom, float* __restrict__ out) {
  __shared__ double sred[8];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += (double)part[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sred[w];
    *out = (float)(t / denom);
  }
}

int launch_mean_finalize(const float* part, int n, double denom, float* out, cudaStream_t st) {
  mean_finalize_kernel<<<1, 256, 0, st>>>(part, n, denom, out);
  return check_launch();
}

static int sel_common_checks(void* W, int dtype, int R, int C, int64_t ldw, const float* scaler_row,
                             uint8_t* keep_mask, int64_t ldm, void* ws) {
  if (!W || !scaler_row || !keep_mask || !ws || R < 1 || C < 1 || ldw < C || ldm < C) return VLMC_ERR_BAD_ARG;
  if (dtype != VLMC_F32 && dtype != VLMC_F16 && dtype != VLMC_BF16) return VLMC_ERR_BAD_ARG;
  const int V = dtype == VLMC_F32 ? 4 : 8;
  if (C % V != 0 || ldw % V != 0 || ldm % V != 0 || ((uintptr_t)W & 15) != 0 ||
      ((uintptr_t)keep_mask & 7) != 0)
    return VLMC_ERR_UNSUPPORTED;
  if (!is_device_ptr(W) || !is_device_ptr(scaler_row) || !is_device_ptr(keep_mask) || !is_device_ptr(ws))
    return VLMC_ERR_NOT_DEVICE;
  return VLMC_OK;
}

template <typename T, int KPT>
static int launch_rowselect(void* W, int R, int C, int64_t ldw, const float* scaler_row, int k, int zero_w,
                            uint8_t* mask, int64_t ldm, float* row_sum, cudaStream_t st) {
  const size_t smem = (size_t)C * sizeof(float);
  auto kern = rowselect_kernel<T, KPT>;
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return check_launch();
  }
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSelThreads, smem);
  if (per_sm < 1) per_sm = 1;
  int grid = kNumSMs * per_sm;
  if (grid > R) grid = R;
  kern<<<grid, kSelThreads, smem, st>>>(reinterpret_cast<T*>(W), ldw, R, C, scaler_row, k, zero_w,
                                        mask, ldm, row_sum);
  return check_launch();
}

}  // namespace vlmc

extern "C" int vlmc_wanda_rowselect(void* W, int dtype, int R, int C, int64_t ldw,
                                    const float* scaler_row, int k, int zero_w,
                                    uint8_t* keep_mask, int64_t ldm, float* score_mean,
                                    void* ws, size_t ws_bytes, void* stream) {
  using namespace vlmc;
  int rc = sel_common_checks(W, dtype, R, C, ldw, scaler_row, keep_mask, ldm, ws);
  if (rc) return rc;
  if (k < 0) return VLMC_ERR_BAD_ARG;
  if (C > kSelThreads * 128) return VLMC_ERR_UNSUPPORTED;  // row must fit the register file of one CTA
  if (ws_bytes < VLMC_WS_COUNTER_BYTES + (size_t)R * sizeof(float)) return VLMC_ERR_WORKSPACE;
  float* row_sum = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + VLMC_WS_COUNTER_BYTES);
  cudaStream_t st = (cudaStream_t)stream;
  const int kpt_needed = (C + kSelThreads - 1) / kSelThreads;
#define VLMC_ROWSEL(KPT)                                                                          \
  VLMC_DISPATCH_DTYPE(dtype, rc = (launch_rowselect<scalar_t, KPT>(W, R, C, ldw, scaler_row, k,   \
                                                                   zero_w, keep_mask, ldm, row_sum, st)))
  if (kpt_needed <= 16) { VLMC_ROWSEL(16); }
  else if (kpt_needed <= 32) { VLMC_ROWSEL(32); }
  else if (kpt_needed <= 48) { VLMC_ROWSEL(48); }
  else if (kpt_needed <= 88) { VLMC_ROWSEL(88); }
  else { VLMC_ROWSEL(128); }
#undef VLMC_ROWSEL
  if (rc) return rc;
  if (score_mean) return launch_mean_finalize(row_sum, R, (double)R * (double)C, score_mean, st);
  return VLMC_OK;
}

extern "C" int vlmc_wanda_nm(void* W, int dtype, int R, int C, int64_t ldw,
                             const float* scaler_row, int n, int m, int zero_w,
                             uint8_t* keep_mask, int64_t ldm, float* score_mean,
                             void* ws, size_t ws_bytes, void* stream) {
  using namespace vlmc;
  int rc = sel_common_checks(W, dtype, R, C, ldw, scaler_row, keep_mask, ldm, ws);
  if (rc) return rc;
  if (!(m == 2 || m == 4 || m == 8 || m == 16)) return VLMC_ERR_UNSUPPORTED;
  if (n <= 0 || n >= m) return VLMC_ERR_BAD_ARG;
  if (C % m != 0) return VLMC_ERR_UNSUPPORTED;
  const int V = dtype == VLMC_F32 ? 4 : 8;
  const int E = m > V ? m : V;
  if (C % E != 0) return VLMC_ERR_UNSUPPORTED;
  const int coltiles = (C / E + kSelThreads - 1) / kSelThreads;
  int rowblocks = (kNumSMs * 16 + coltiles - 1) / coltiles;
  if (rowblocks > R) rowblocks = R;
  if (rowblocks > 65535) rowblocks = 65535;
  const int nparts = coltiles * rowblocks;
  if (ws_bytes < VLMC_WS_COUNTER_BYTES + (size_t)nparts * sizeof(float)) return VLMC_ERR_WORKSPACE;
  float* part = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + VLMC_WS_COUNTER_BYTES);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(coltiles, rowblocks);
#define VLMC_NM(MM)                                                                                 \
  VLMC_DISPATCH_DTYPE(dtype, (nm_kernel<scalar_t, MM><<<grid, kSelThreads, 0, st>>>(                \
                                 reinterpret_cast<scalar_t*>(W), ldw, R, C, scaler_row, n, zero_w,  \
                                 keep_mask, ldm, part)))
  switch (m) {
    case 2: VLMC_NM(2); break;
    case 4: VLMC_NM(4); break;
    case 8: VLMC_NM(8); break;
    default: VLMC_NM(16); break;
  }
#undef VLMC_NM
  rc = check_launch();
  if (rc) return rc;
  if (score_mean) return launch_mean_finalize(part, nparts, (double)R * (double)C, score_mean, st);
  return VLMC_OK;
}
