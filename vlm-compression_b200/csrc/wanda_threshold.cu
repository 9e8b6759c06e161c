// K7: Wanda score + whole-matrix threshold (ViT path).
//
// Replaces lavis/compression/pruners/wanda_pruner.py:682-683
//   thres = torch.sort(W_metric.flatten())[0][int(numel * p)] ;  W_mask = W_metric < thres
// i.e. a full device sort of up to 8.65 M keys, by an exact counting select: every pass counts the
// scores below 16 pivots (7 octile points that guarantee an 8x shrink of the bracket + 9 around the
// expected position of the target rank: a sampled estimate on the first pass, interpolation
// afterwards); W (<= 17 MB) stays L2-resident between passes.  The last CTA of a pass folds the counts
// into the bracket kept in the workspace, so there is no host synchronisation between launches and
// passes after convergence exit at once (typically 2 passes do work).
#include <stdlib.h>
#include "common.cuh"

namespace vlmc {

int launch_mean_finalize(const float* part, int n, double denom, float* out, cudaStream_t st);

constexpr int kThrThreads = 256;
constexpr int kThrPasses = 12;       // worst case: 8x shrink of a 32-bit range per pass (octile pivots) -> 11 passes
constexpr int kThrPivots = 16;
constexpr int kThrCap = 2048;

typedef unsigned long long ull;

struct Bracket { uint32_t lo, hi; ull glo, ghi; };

// search state in the workspace.  The LAST CTA of a counting pass folds the pass's counts into the bracket, so the next
// pass (and the gather / resolve / apply kernels) read the bracket instead of replaying the pivot history.
struct ThrState {
  ull counts[kThrPivots];
  Bracket b;
  int converged;
  int passes;                 // counting passes that did work (diagnostic)
  unsigned int ticket;
  uint32_t seed;              // sampled estimate of the threshold key (0: none)
  unsigned int cand_cnt;
  uint32_t v;
  uint32_t cand[kThrCap];
};

__device__ __forceinline__ bool thr_done(const Bracket& b) {
  return (b.ghi - b.glo) <= (ull)kThrCap || (b.hi - b.lo) == 1u;
}

__device__ __forceinline__ uint32_t thr_clamp(double x, uint32_t lo, uint32_t hi) {
  if (!(x > (double)lo + 1.0)) return lo + 1;
  if (!(x < (double)hi - 1.0)) return hi - 1;
  return (uint32_t)x;
}

// 16 pivots inside the bracket (lo, hi): 7 octile points of the bit range (an 8x shrink whatever the data) and 9 points
// around the expected position of the target rank: the sampled estimate on the first pass, linear interpolation of the
// rank inside the bracket afterwards, at geometrically growing offsets.
__device__ void thr_pivots(const Bracket& b, ull k, uint32_t seed, bool first, uint32_t* p) {
  const double w = (double)(b.hi - b.lo), lo = (double)b.lo;
#pragma unroll
  for (int i = 0; i < 7; ++i) p[i] = thr_clamp(lo + (double)(i + 1) * 0.125 * w, b.lo, b.hi);
  double e, unit;
  if (first && seed != 0u) {
    e = (double)seed;
    unit = 4096.0;                    // key bits: 2^-11 of a binade = 0.034 % in value; offsets up to +-2^21 bits = +-19 %
  } else {
    const double f = ((double)(k - b.glo) - 0.5) / (double)(b.ghi - b.glo);
    e = lo + f * w;
    unit = w * (1.0 / 4096.0);
  }
  p[7] = thr_clamp(e, b.lo, b.hi);
  p[8] = thr_clamp(e - unit, b.lo, b.hi);        p[9] = thr_clamp(e + unit, b.lo, b.hi);
  p[10] = thr_clamp(e - 8.0 * unit, b.lo, b.hi);  p[11] = thr_clamp(e + 8.0 * unit, b.lo, b.hi);
  p[12] = thr_clamp(e - 64.0 * unit, b.lo, b.hi); p[13] = thr_clamp(e + 64.0 * unit, b.lo, b.hi);
  p[14] = thr_clamp(e - 512.0 * unit, b.lo, b.hi); p[15] = thr_clamp(e + 512.0 * unit, b.lo, b.hi);
}

__global__ void thr_init_kernel(ThrState* st, ull n) {
  if (threadIdx.x < kThrPivots) st->counts[threadIdx.x] = 0ull;
  if (threadIdx.x == 0) {
    st->b = Bracket{0u, 0xffffffffu, 0ull, n};
    st->converged = 0; st->passes = 0; st->ticket = 0u; st->seed = 0u; st->cand_cnt = 0u; st->v = 0u;
  }
}

__global__ void sqrt_kernel(const float* __restrict__ s, float* __restrict__ sq, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) sq[c] = __fsqrt_rn(s[c]);
}

template <typename T>
__device__ __forceinline__ void load_keys(const T* W, int64_t ldw, int cvecs, int64_t vec,
                                          const float* __restrict__ sq, uint32_t* keys, float* fsum) {
  constexpr int V = Elem<T>::kVec;
  const int64_t row = vec / cvecs;
  const int col = (int)(vec % cvecs) * V;
  uint4 v = *reinterpret_cast<const uint4*>(W + row * ldw + col);
  float f[V];
  Elem<T>::unpack(v, f);
  const float4* sp = reinterpret_cast<const float4*>(sq + col);
  float sv[V];
#pragma unroll
  for (int q = 0; q < V / 4; ++q) {
    float4 t = __ldg(sp + q);
    sv[4 * q] = t.x; sv[4 * q + 1] = t.y; sv[4 * q + 2] = t.z; sv[4 * q + 3] = t.w;
  }
#pragma unroll
  for (int e = 0; e < V; ++e) {
    const float s = __fmul_rn(fabsf(f[e]), sv[e]);
    keys[e] = __float_as_uint(s);
    if (fsum) *fsum += s;
  }
}

// one warp: the k/n quantile of 1024 (16-bit) / 512 (fp32) sampled scores, by bisection in key-bit space
template <typename T>
__global__ void __launch_bounds__(32)
thr_seed_kernel(const T* __restrict__ W, int64_t ldw, int R, int C, const float* __restrict__ sq, ull k, ThrState* st) {
  constexpr int V = Elem<T>::kVec;
  constexpr int kVecs = 4;
  const int lane = threadIdx.x;
  const int cvecs = C / V;
  uint32_t sk[kVecs * V];
#pragma unroll
  for (int u = 0; u < kVecs; ++u) {
    const int sidx = lane * kVecs + u;
    const int64_t row = ((int64_t)((sidx * 37) % 128) * R) / 128;
    const int64_t cv = ((int64_t)sidx * cvecs) / 128;
    load_keys<T>(W, ldw, cvecs, row * cvecs + cv, sq, sk + u * V, nullptr);
  }
  const int ns = 32 * kVecs * V;
  const ull n = (ull)R * (ull)C;
  int ks = (int)(((double)k / (double)n) * ns + 0.5);
  if (ks < 1) ks = 1;
  if (ks > ns) ks = ns;
  uint32_t slo = 0u, shi = 0x7f800000u;
  while (shi - slo > 4096u) {
    const double q = (double)(shi - slo) * 0.2;
    const uint32_t sp[4] = {thr_clamp(slo + q, slo, shi), thr_clamp(slo + 2.0 * q, slo, shi), thr_clamp(slo + 3.0 * q, slo, shi),
                            thr_clamp(slo + 4.0 * q, slo, shi)};
    uint32_t c[4] = {0, 0, 0, 0};
#pragma unroll
    for (int e = 0; e < kVecs * V; ++e)
#pragma unroll
      for (int i = 0; i < 4; ++i) c[i] += sk[e] < sp[i] ? 1u : 0u;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int sc = (int)__reduce_add_sync(0xffffffffu, c[i]);
      if (sc <= ks - 1) { if (sp[i] > slo) slo = sp[i]; }
      else              { if (sp[i] < shi) shi = sp[i]; }
    }
  }
  if (lane == 0) st->seed = slo + ((shi - slo) >> 1);
}

template <typename T>
__global__ void __launch_bounds__(kThrThreads)
thr_count_kernel(const T* __restrict__ W, int64_t ldw, int R, int C, const float* __restrict__ sq,
                 ull k, ThrState* st, int pass) {
  constexpr int V = Elem<T>::kVec;
  __shared__ uint32_t sp[kThrPivots];
  __shared__ int s_done;
  __shared__ unsigned int s_cnt[kThrPivots];
  __shared__ unsigned int s_last;
  if (threadIdx.x == 0) {
    s_done = *reinterpret_cast<volatile int*>(&st->converged);
    if (!s_done) {
      const Bracket b = st->b;
      uint32_t p[kThrPivots];
      thr_pivots(b, k, st->seed, pass == 0, p);
      for (int i = 0; i < kThrPivots; ++i) sp[i] = p[i];
    }
  }
  if (threadIdx.x < kThrPivots) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  if (s_done) return;
  uint32_t p[kThrPivots];
#pragma unroll
  for (int i = 0; i < kThrPivots; ++i) p[i] = sp[i];
  unsigned int cnt[kThrPivots] = {};
  const int cvecs = C / V;
  const int64_t nvec = (int64_t)R * cvecs;
  for (int64_t vec = (int64_t)blockIdx.x * kThrThreads + threadIdx.x; vec < nvec;
       vec += (int64_t)gridDim.x * kThrThreads) {
    uint32_t keys[V];
    load_keys<T>(W, ldw, cvecs, vec, sq, keys, nullptr);
#pragma unroll
    for (int e = 0; e < V; ++e)
#pragma unroll
      for (int i = 0; i < kThrPivots; ++i) cnt[i] += keys[e] < p[i] ? 1u : 0u;
  }
#pragma unroll
  for (int i = 0; i < kThrPivots; ++i) {
    unsigned int c = __reduce_add_sync(0xffffffffu, cnt[i]);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_cnt[i], c);
  }
  __syncthreads();
  if (threadIdx.x < kThrPivots) atomicAdd(&st->counts[threadIdx.x], (ull)s_cnt[threadIdx.x]);
  // last CTA of the pass: fold the counts into the bracket, clear them for the next pass
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&st->ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x == 0) {
    Bracket b = st->b;
    for (int i = 0; i < kThrPivots; ++i) {
      const ull c = *reinterpret_cast<volatile ull*>(&st->counts[i]);
      if (c <= k - 1) { if (p[i] > b.lo) { b.lo = p[i]; b.glo = c; } }
      else            { if (p[i] < b.hi) { b.hi = p[i]; b.ghi = c; } }
      st->counts[i] = 0ull;
    }
    st->b = b;
    st->passes = pass + 1;
    st->converged = thr_done(b) ? 1 : 0;
    st->ticket = 0u;
  }
}

template <typename T>
__global__ void __launch_bounds__(kThrThreads)
thr_gather_kernel(const T* __restrict__ W, int64_t ldw, int R, int C, const float* __restrict__ sq,
                  ull k, ThrState* st) {
  constexpr int V = Elem<T>::kVec;
  __shared__ uint32_t s_lo, s_hi;
  __shared__ int s_skip;
  if (threadIdx.x == 0) {
    const Bracket b = st->b;
    s_lo = b.lo; s_hi = b.hi;
    s_skip = (b.hi - b.lo == 1u) ? 1 : 0;
  }
  __syncthreads();
  if (s_skip) return;
  const uint32_t lo = s_lo, hi = s_hi;
  const int cvecs = C / V;
  const int64_t nvec = (int64_t)R * cvecs;
  for (int64_t vec = (int64_t)blockIdx.x * kThrThreads + threadIdx.x; vec < nvec;
       vec += (int64_t)gridDim.x * kThrThreads) {
    uint32_t keys[V];
    load_keys<T>(W, ldw, cvecs, vec, sq, keys, nullptr);
#pragma unroll
    for (int e = 0; e < V; ++e) {
      if (keys[e] >= lo && keys[e] < hi) {
        const unsigned int slot = atomicAdd(&st->cand_cnt, 1u);
        if (slot < (unsigned)kThrCap) st->cand[slot] = keys[e];
      }
    }
  }
}

// one CTA: the (k-1-glo)-th smallest candidate is the threshold value
__global__ void __launch_bounds__(1024)
thr_resolve_kernel(int R, int C, ull k, ThrState* st) {
  __shared__ uint32_t s_cand[kThrCap];
  __shared__ Bracket s_b;
  if (threadIdx.x == 0) s_b = st->b;
  __syncthreads();
  const Bracket b = s_b;
  if (b.hi - b.lo == 1u) {
    if (threadIdx.x == 0) st->v = b.lo;
    return;
  }
  const int m = (int)(b.ghi - b.glo);
  for (int i = threadIdx.x; i < m; i += blockDim.x) s_cand[i] = st->cand[i];
  __syncthreads();
  const int target = (int)(k - 1 - b.glo);
  for (int t = threadIdx.x; t < m; t += blockDim.x) {
    const uint32_t kt = s_cand[t];
    int less = 0, leq = 0;
    for (int j = 0; j < m; ++j) { less += s_cand[j] < kt ? 1 : 0; leq += s_cand[j] <= kt ? 1 : 0; }
    if (less <= target && target < leq) st->v = kt;  // all writers store the same value
  }
}

template <typename T>
__global__ void __launch_bounds__(kThrThreads)
thr_apply_kernel(T* __restrict__ W, int64_t ldw, int R, int C, const float* __restrict__ sq,
                 const uint32_t* __restrict__ vptr, int zero_w, uint8_t* __restrict__ mask, int64_t ldm,
                 float* __restrict__ part_sum) {
  constexpr int V = Elem<T>::kVec;
  // a NaN threshold (the rank falls among NaN scores, which sort last) prunes nothing: `W_metric < nan` is all False (:683)
  const uint32_t v = *vptr > 0x7f800000u ? 0u : *vptr;
  const int cvecs = C / V;
  const int64_t nvec = (int64_t)R * cvecs;
  float lsum = 0.f;
  for (int64_t vec = (int64_t)blockIdx.x * kThrThreads + threadIdx.x; vec < nvec;
       vec += (int64_t)gridDim.x * kThrThreads) {
    uint32_t keys[V];
    load_keys<T>(W, ldw, cvecs, vec, sq, keys, &lsum);
    const int64_t row = vec / cvecs;
    const int col = (int)(vec % cvecs) * V;
    uint32_t mb[V / 4] = {};
    bool any = false;
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const bool pr = keys[e] < v;   // strict: ties with the threshold are kept (:683)
      any |= pr;
      mb[e / 4] |= (pr ? 0u : 1u) << (8 * (e % 4));
    }
    uint8_t* mp = mask + row * ldm + col;
    if (V == 8) st_stream8(mp, make_uint2(mb[0], mb[V / 4 - 1]));
    else st_stream4(mp, mb[0]);
    if (zero_w && any) {
      T* wp = W + row * ldw + col;
      uint4 wv = *reinterpret_cast<const uint4*>(wp);
      uint32_t* wr = reinterpret_cast<uint32_t*>(&wv);
      if (sizeof(T) == 4) {
#pragma unroll
        for (int e = 0; e < V; ++e) if (keys[e] < v) wr[e] = 0u;
      } else {
#pragma unroll
        for (int e = 0; e < V; ++e) if (keys[e] < v) wr[e / 2] &= (e & 1) ? 0x0000ffffu : 0xffff0000u;
      }
      st_stream(wp, wv);
    }
  }
  __shared__ float fred[kThrThreads / 32];
  lsum = warp_sum(lsum);
  if ((threadIdx.x & 31) == 0) fred[threadIdx.x >> 5] = lsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < kThrThreads / 32; ++w) s += fred[w];
    part_sum[blockIdx.x] = s;
  }
}

// ---- radix form of the select (default): scores are non-negative floats, so their bit patterns ARE the sort keys
// (31 bits; NaN above inf like torch.sort).  Three histogram passes (11 / 11 / 9 key bits, shared-memory histogram per CTA
// flushed into the workspace) find the k-th smallest key exactly; the LAST CTA of a pass (ticket) picks the bin that holds
// the rank, narrows the prefix and clears the histogram, so a call is init + 3 passes + apply + mean: 6 stream-ordered
// launches instead of the ~19 of the counting form above (kept behind VLMC_THRESHOLD_COUNTING=1 for A/B runs).
constexpr int kRdxBins = 2048;
struct RadixState {
  ull k_rem;
  uint32_t prefix;
  unsigned int ticket;
  uint32_t v;
  uint32_t pad;
  ull hist[kRdxBins];
};

__global__ void __launch_bounds__(256)
thr_radix_init_kernel(const float* __restrict__ s, float* __restrict__ sq, int C, RadixState* st, ull k) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) sq[i] = __fsqrt_rn(s[i]);
  if (i < kRdxBins) st->hist[i] = 0ull;
  if (i == 0) { st->k_rem = k; st->prefix = 0u; st->ticket = 0u; st->v = 0u; st->pad = 0u; }
}

template <int PASS>
__device__ __forceinline__ bool rdx_bin(uint32_t key, uint32_t prefix, uint32_t& bin) {
  if (PASS == 0) { bin = key >> 20; return true; }                              // keys < 2^31: 11 bits
  if (PASS == 1) { bin = (key >> 9) & 0x7ffu; return (key >> 20) == prefix; }
  bin = key & 0x1ffu;
  return (key >> 9) == prefix;
}

template <typename T, int PASS>
__global__ void __launch_bounds__(kThrThreads)
thr_radix_kernel(const T* __restrict__ W, int64_t ldw, int R, int C, const float* __restrict__ sq, RadixState* st) {
  constexpr int V = Elem<T>::kVec;
  constexpr int kPer = kRdxBins / kThrThreads;          // 8 consecutive bins per thread in the resolve step
  __shared__ uint32_t h[kRdxBins];
  __shared__ ull part[kThrThreads];
  __shared__ unsigned int s_last;
  __shared__ int s_owner;
  __shared__ ull s_before;
  for (int i = threadIdx.x; i < kRdxBins; i += kThrThreads) h[i] = 0u;
  const uint32_t prefix = PASS > 0 ? st->prefix : 0u;
  __syncthreads();
  const int cvecs = C / V;
  const int64_t nvec = (int64_t)R * cvecs;
  for (int64_t vec = (int64_t)blockIdx.x * kThrThreads + threadIdx.x; vec < nvec;
       vec += (int64_t)gridDim.x * kThrThreads) {
    uint32_t keys[V];
    load_keys<T>(W, ldw, cvecs, vec, sq, keys, nullptr);
    uint32_t bin;
#pragma unroll
    for (int e = 0; e < V; ++e)
      if (rdx_bin<PASS>(keys[e], prefix, bin)) atomicAdd(&h[bin], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kRdxBins; i += kThrThreads) {
    const uint32_t c = h[i];
    if (c) atomicAdd(&st->hist[i], (ull)c);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&st->ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  // last CTA of the pass: the bin that holds rank k_rem
  __threadfence();
  ull c[kPer];
  ull mine = 0;
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    c[j] = *reinterpret_cast<volatile ull*>(&st->hist[threadIdx.x * kPer + j]);
    mine += c[j];
  }
  part[threadIdx.x] = mine;
  __syncthreads();
  const ull k = st->k_rem;
  if (threadIdx.x == 0) {
    ull acc = 0;
    int o = -1;
    for (int t = 0; t < kThrThreads; ++t) {
      if (acc + part[t] >= k) { o = t; break; }
      acc += part[t];
    }
    s_owner = o;            // -1 cannot happen: k <= R * C = total count (checked by the caller)
    s_before = acc;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kPer; ++j) st->hist[threadIdx.x * kPer + j] = 0ull;
  if ((int)threadIdx.x == s_owner) {
    ull acc = s_before;
    int bin = -1;
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      if (bin < 0) {
        if (acc + c[j] >= k) bin = j;
        else acc += c[j];
      }
    }
    if (bin < 0) bin = kPer - 1;                        // unreachable: the owner's chunk holds the rank
    const uint32_t b = (uint32_t)(threadIdx.x * kPer + bin);
    const uint32_t np = PASS == 0 ? b : PASS == 1 ? ((prefix << 11) | b) : ((prefix << 9) | b);
    st->prefix = np;
    st->k_rem = k - acc;
    if (PASS == 2) st->v = np;
  }
  if (threadIdx.x == 0) st->ticket = 0u;
}

static bool thr_use_counting() {
  const char* e = getenv("VLMC_THRESHOLD_COUNTING");
  return e && e[0] == '1';
}

size_t threshold_workspace_bytes(int R, int C) {
  return VLMC_WS_COUNTER_BYTES + align_up((size_t)C * sizeof(float), 256) +
         align_up((size_t)kNumSMs * 8 * sizeof(float), 256) + align_up(sizeof(ThrState), 256) +
         align_up(sizeof(RadixState), 256);
}

}  // namespace vlmc

extern "C" int vlmc_wanda_threshold(void* W, int dtype, int R, int C, int64_t ldw,
                                    const float* scaler_row, int64_t k_global, int zero_w,
                                    uint8_t* keep_mask, int64_t ldm, float* score_mean,
                                    void* ws, size_t ws_bytes, void* stream) {
  using namespace vlmc;
  if (!W || !scaler_row || !keep_mask || !ws || R < 1 || C < 1 || ldw < C || ldm < C) return VLMC_ERR_BAD_ARG;
  if (dtype != VLMC_F32 && dtype != VLMC_F16 && dtype != VLMC_BF16) return VLMC_ERR_BAD_ARG;
  if (k_global < 0 || k_global >= (int64_t)R * C) return VLMC_ERR_BAD_ARG;  // sorted[k] must exist
  const int V = dtype == VLMC_F32 ? 4 : 8;
  if (C % 8 != 0 || ldw % V != 0 || ldm % V != 0 || ((uintptr_t)W & 15) != 0 || ((uintptr_t)keep_mask & 7) != 0)
    return VLMC_ERR_UNSUPPORTED;
  if (!is_device_ptr(W) || !is_device_ptr(scaler_row) || !is_device_ptr(keep_mask) || !is_device_ptr(ws))
    return VLMC_ERR_NOT_DEVICE;
  if (ws_bytes < threshold_workspace_bytes(R, C)) return VLMC_ERR_WORKSPACE;
  char* base = reinterpret_cast<char*>(ws) + VLMC_WS_COUNTER_BYTES;
  float* sq = reinterpret_cast<float*>(base);
  base += align_up((size_t)C * sizeof(float), 256);
  float* part = reinterpret_cast<float*>(base);
  base += align_up((size_t)kNumSMs * 8 * sizeof(float), 256);
  ThrState* st = reinterpret_cast<ThrState*>(base);
  cudaStream_t s = (cudaStream_t)stream;
  const ull k = (ull)k_global + 1;  // 1-indexed rank of the threshold value

  RadixState* rst = reinterpret_cast<RadixState*>(base + align_up(sizeof(ThrState), 256));
  const int64_t nvec = (int64_t)R * (C / V);
  int grid = kNumSMs * 8;
  if ((int64_t)grid * kThrThreads > nvec) grid = (int)((nvec + kThrThreads - 1) / kThrThreads);
  const uint32_t* vptr;
  if (!thr_use_counting()) {
    const int n_init = C > kRdxBins ? C : kRdxBins;
    thr_radix_init_kernel<<<(n_init + 255) / 256, 256, 0, s>>>(scaler_row, sq, C, rst, k);
    VLMC_DISPATCH_DTYPE(dtype, (thr_radix_kernel<scalar_t, 0><<<grid, kThrThreads, 0, s>>>(
                                   reinterpret_cast<const scalar_t*>(W), ldw, R, C, sq, rst)));
    VLMC_DISPATCH_DTYPE(dtype, (thr_radix_kernel<scalar_t, 1><<<grid, kThrThreads, 0, s>>>(
                                   reinterpret_cast<const scalar_t*>(W), ldw, R, C, sq, rst)));
    VLMC_DISPATCH_DTYPE(dtype, (thr_radix_kernel<scalar_t, 2><<<grid, kThrThreads, 0, s>>>(
                                   reinterpret_cast<const scalar_t*>(W), ldw, R, C, sq, rst)));
    vptr = &rst->v;
  } else {
    thr_init_kernel<<<1, 32, 0, s>>>(st, (ull)R * (ull)C);
    sqrt_kernel<<<(C + 255) / 256, 256, 0, s>>>(scaler_row, sq, C);
    VLMC_DISPATCH_DTYPE(dtype, (thr_seed_kernel<scalar_t><<<1, 32, 0, s>>>(reinterpret_cast<const scalar_t*>(W), ldw, R, C, sq, k, st)));
    for (int pass = 0; pass < kThrPasses; ++pass) {
      VLMC_DISPATCH_DTYPE(dtype, (thr_count_kernel<scalar_t><<<grid, kThrThreads, 0, s>>>(
                                     reinterpret_cast<const scalar_t*>(W), ldw, R, C, sq, k, st, pass)));
    }
    VLMC_DISPATCH_DTYPE(dtype, (thr_gather_kernel<scalar_t><<<grid, kThrThreads, 0, s>>>(
                                   reinterpret_cast<const scalar_t*>(W), ldw, R, C, sq, k, st)));
    thr_resolve_kernel<<<1, 1024, 0, s>>>(R, C, k, st);
    vptr = &st->v;
  }
  VLMC_DISPATCH_DTYPE(dtype, (thr_apply_kernel<scalar_t><<<grid, kThrThreads, 0, s>>>(
                                 reinterpret_cast<scalar_t*>(W), ldw, R, C, sq, vptr, zero_w, keep_mask, ldm, part)));
  int rc = check_launch();
  if (rc) return rc;
  if (score_mean) return launch_mean_finalize(part, grid, (double)R * (double)C, score_mean, s);
  return VLMC_OK;
}
