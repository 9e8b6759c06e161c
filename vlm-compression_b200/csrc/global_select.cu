// K18-K22: global sparsity allocation on fp32 importance scores.
//
// Replaces, in lavis/compression/pruners/layer_single_base_pruner.py (same code in global_pruner.py:108-148):
//   :157-160  threshold = topk(scores[k].flatten(), j, largest=True)[0][-1]; scores[k][v >= threshold] = finfo.max
//   :165-169  all = cat([t.flatten() ...]); threshold = topk(all, int(p * numel), largest=False)[0][-1]
//   :172-174  masks[k] = (v > threshold).type(v.dtype)          :223-225  v.data *= masks[k]
//   :296      group_scores[g] += importance_measure[l].sum()
//   :456-473  gradients += grad.float() ** 2 ; importance = (w.float() ** 2) * (gradients / n_batches)
// The reference moves every score to the CPU and runs torch.topk on the concatenation of the whole model (6.5e9 scores
// for Vicuna-7B).  Here nothing is concatenated or sorted: the k-th score is found by an exact radix select - three
// streaming passes (11 / 11 / 10 key bits) that histogram the scores of every tensor into one histogram per segment -
// and the dependent passes are stream ordered (a one-CTA-per-segment resolve kernel narrows the prefix on the device).
// All kernels walk up to 64 tensors per launch as one list of 64 K-element units with a grid-stride loop.
#include <float.h>
#include "common.cuh"

namespace vlmc {

constexpr int kSelMax = 64;
constexpr int kSelThreads = 256;
constexpr int64_t kSelUnit = 65536;          // elements per work unit (256 KB of scores)
constexpr int kSelBins = 2048;

typedef unsigned long long ull;

struct ScoreBatch {
  float* s[kSelMax];
  void* aux[kSelMax];
  float* out[kSelMax];
  int64_t numel[kSelMax];
  int64_t unit_begin[kSelMax + 1];
  int seg[kSelMax];
  signed char adt[kSelMax];
  int count;
};

struct SelState { ull k_rem; uint32_t prefix; uint32_t pad; };

// order-preserving key of a score under torch.topk's total order: NaN above everything, -0.0 == +0.0
__device__ __forceinline__ uint32_t score_key(float v) {
  if (v != v) return 0xffffffffu;
  const uint32_t u = __float_as_uint(v + 0.0f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_score(uint32_t k) {
  const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);                 // key 0xffffffff -> 0x7fffffff, a NaN
}

template <int PASS>
__device__ __forceinline__ bool key_bin(uint32_t key, uint32_t prefix, uint32_t& bin) {
  if (PASS == 0) { bin = key >> 21; return true; }
  if (PASS == 1) { bin = (key >> 10) & 0x7ffu; return (key >> 21) == prefix; }
  bin = key & 0x3ffu;
  return (key >> 10) == prefix;
}

__device__ __forceinline__ int find_item(const ScoreBatch& b, int64_t unit) {
  int p = 0;
  while (unit >= b.unit_begin[p + 1]) ++p;
  return p;
}

// ---- K18: histogram pass of the radix select -------------------------------------------------------------------------
template <int PASS>
__global__ void __launch_bounds__(kSelThreads)
scores_hist_kernel(const __grid_constant__ ScoreBatch b, const SelState* __restrict__ state, ull* __restrict__ hist) {
  __shared__ uint32_t h[kSelBins];
  const int64_t total = b.unit_begin[b.count];
  int cur = -1;
  uint32_t prefix = 0;
  for (int i = threadIdx.x; i < kSelBins; i += kSelThreads) h[i] = 0u;
  __syncthreads();
  for (int64_t unit = blockIdx.x; unit < total; unit += gridDim.x) {
    const int p = find_item(b, unit);
    const int seg = b.seg[p];
    if (seg != cur) {
      if (cur >= 0) {
        __syncthreads();
        for (int i = threadIdx.x; i < kSelBins; i += kSelThreads) {
          const uint32_t c = h[i];
          if (c) { atomicAdd(hist + (size_t)cur * kSelBins + i, (ull)c); h[i] = 0u; }
        }
        __syncthreads();
      }
      cur = seg;
      if (PASS > 0) prefix = state[seg].prefix;
    }
    const int64_t e0 = (unit - b.unit_begin[p]) * kSelUnit;
    const int64_t e1 = e0 + kSelUnit < b.numel[p] ? e0 + kSelUnit : b.numel[p];
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(b.s[p]) + (uintptr_t)e0 * 4;
    const uintptr_t a1 = reinterpret_cast<uintptr_t>(b.s[p]) + (uintptr_t)e1 * 4;
    uintptr_t v0 = (a0 + 15) & ~(uintptr_t)15, v1 = a1 & ~(uintptr_t)15;
    if (v0 > v1) { v0 = a1; v1 = a1; }
    uint32_t bin;
    for (uintptr_t a = a0 + (uintptr_t)threadIdx.x * 4; a < v0; a += (uintptr_t)kSelThreads * 4)
      if (key_bin<PASS>(score_key(*reinterpret_cast<const float*>(a)), prefix, bin)) atomicAdd(&h[bin], 1u);
#pragma unroll 4
    for (uintptr_t a = v0 + (uintptr_t)threadIdx.x * 16; a < v1; a += (uintptr_t)kSelThreads * 16) {
      const uint4 v = ld_stream(reinterpret_cast<const void*>(a));
      const uint32_t k0 = score_key(__uint_as_float(v.x)), k1 = score_key(__uint_as_float(v.y));
      const uint32_t k2 = score_key(__uint_as_float(v.z)), k3 = score_key(__uint_as_float(v.w));
      if (PASS == 0) {
        // neighbouring scores often share a bin (the top 11 bits are sign, exponent and two mantissa bits)
        const uint32_t b0 = k0 >> 21, b1 = k1 >> 21, b2 = k2 >> 21, b3 = k3 >> 21;
        if (b0 == b1 && b2 == b3) {
          if (b0 == b2) atomicAdd(&h[b0], 4u);
          else { atomicAdd(&h[b0], 2u); atomicAdd(&h[b2], 2u); }
        } else {
          atomicAdd(&h[b0], 1u); atomicAdd(&h[b1], 1u); atomicAdd(&h[b2], 1u); atomicAdd(&h[b3], 1u);
        }
      } else {
        if (key_bin<PASS>(k0, prefix, bin)) atomicAdd(&h[bin], 1u);
        if (key_bin<PASS>(k1, prefix, bin)) atomicAdd(&h[bin], 1u);
        if (key_bin<PASS>(k2, prefix, bin)) atomicAdd(&h[bin], 1u);
        if (key_bin<PASS>(k3, prefix, bin)) atomicAdd(&h[bin], 1u);
      }
    }
    for (uintptr_t a = v1 + (uintptr_t)threadIdx.x * 4; a < a1; a += (uintptr_t)kSelThreads * 4)
      if (key_bin<PASS>(score_key(*reinterpret_cast<const float*>(a)), prefix, bin)) atomicAdd(&h[bin], 1u);
  }
  if (cur >= 0) {
    __syncthreads();
    for (int i = threadIdx.x; i < kSelBins; i += kSelThreads) {
      const uint32_t c = h[i];
      if (c) atomicAdd(hist + (size_t)cur * kSelBins + i, (ull)c);
    }
  }
}

__global__ void scores_init_kernel(SelState* __restrict__ state, const int64_t* __restrict__ k, int nseg) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nseg) {
    const int64_t ks = k[s];
    state[s].k_rem = ks < 1 ? 1ull : (ull)ks;
    state[s].prefix = 0u;
    state[s].pad = 0u;
  }
}

// one CTA per segment: finds the bin that holds rank k_rem, narrows the prefix, clears the histogram for the next pass
template <int PASS>
__global__ void __launch_bounds__(kSelThreads)
scores_resolve_kernel(SelState* __restrict__ state, ull* __restrict__ hist, float* __restrict__ kth_out) {
  constexpr int kBins = PASS == 2 ? 1024 : kSelBins;
  constexpr int kPer = kSelBins / kSelThreads;     // 8 consecutive bins per thread
  __shared__ ull part[kSelThreads];
  __shared__ int owner;
  __shared__ ull before_owner;
  const int seg = blockIdx.x;
  ull* hs = hist + (size_t)seg * kSelBins;
  ull c[kPer];
  ull mine = 0;
#pragma unroll
  for (int j = 0; j < kPer; ++j) { c[j] = hs[threadIdx.x * kPer + j]; mine += c[j]; }
  part[threadIdx.x] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    const ull k = state[seg].k_rem;
    ull acc = 0;
    int o = kSelThreads - 1;
    for (int t = 0; t < kSelThreads; ++t) {
      if (acc + part[t] >= k) { o = t; break; }
      acc += part[t];
    }
    if (o == kSelThreads - 1 && acc + part[o] < k) {          // k beyond the population: clamp to the last non-empty chunk
      acc = 0; o = 0;
      for (int t = 0; t < kSelThreads; ++t) if (part[t]) o = t;
      for (int t = 0; t < o; ++t) acc += part[t];
    }
    owner = o;
    before_owner = acc;
  }
  __syncthreads();
  if (threadIdx.x == owner) {
    const ull k = state[seg].k_rem;
    ull acc = before_owner;
    int bin = kPer - 1;
    bool found = false;
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      if (!found) {
        if (acc + c[j] >= k) { bin = j; found = true; }
        else acc += c[j];
      }
    }
    if (!found) {                                               // clamped case: last non-empty bin of the chunk
      acc = before_owner; bin = 0;
      for (int j = 0; j < kPer; ++j) if (c[j]) bin = j;
      for (int j = 0; j < bin; ++j) acc += c[j];
      state[seg].k_rem = c[bin] ? c[bin] : 1ull;
    } else {
      state[seg].k_rem = k - acc;
    }
    const uint32_t b = (uint32_t)(threadIdx.x * kPer + bin);
    const uint32_t prefix = PASS == 0 ? b : PASS == 1 ? ((state[seg].prefix << 11) | b) : ((state[seg].prefix << 10) | (b & (kBins - 1)));
    state[seg].prefix = prefix;
    if (PASS == 2) kth_out[seg] = key_score(prefix);
  }
#pragma unroll
  for (int j = 0; j < kPer; ++j) hs[threadIdx.x * kPer + j] = 0ull;
}

// ---- elementwise passes ------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float aux_load(const void* p, int64_t i);
template <> __device__ __forceinline__ float aux_load<float>(const void* p, int64_t i) { return reinterpret_cast<const float*>(p)[i]; }
template <> __device__ __forceinline__ float aux_load<__half>(const void* p, int64_t i) { return __half2float(reinterpret_cast<const __half*>(p)[i]); }
template <> __device__ __forceinline__ float aux_load<__nv_bfloat16>(const void* p, int64_t i) {
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
template <typename T> __device__ __forceinline__ void aux_store(void* p, int64_t i, float v);
template <> __device__ __forceinline__ void aux_store<float>(void* p, int64_t i, float v) { reinterpret_cast<float*>(p)[i] = v; }
template <> __device__ __forceinline__ void aux_store<__half>(void* p, int64_t i, float v) { reinterpret_cast<__half*>(p)[i] = __float2half_rn(v); }
template <> __device__ __forceinline__ void aux_store<__nv_bfloat16>(void* p, int64_t i, float v) {
  reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}

enum { kOpProtect = 0, kOpMask = 1, kOpAccum = 2, kOpFinalize = 3 };

// four consecutive elements of the parameter / gradient tensor as floats (16-byte or 8-byte vectors)
template <typename T> struct Aux4;
template <> struct Aux4<float> {
  static __device__ __forceinline__ void load(const void* p, int64_t i, float* f) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i));
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  static __device__ __forceinline__ void store(void* p, int64_t i, const float* f) {
    __stcs(reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i), make_float4(f[0], f[1], f[2], f[3]));
  }
};
template <> struct Aux4<__half> {
  static __device__ __forceinline__ void load(const void* p, int64_t i, float* f) {
    const uint2 v = __ldcs(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(p) + i));
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
  }
  static __device__ __forceinline__ void store(void* p, int64_t i, const float* f) {
    uint2 v;
    *reinterpret_cast<__half2*>(&v.x) = __floats2half2_rn(f[0], f[1]);
    *reinterpret_cast<__half2*>(&v.y) = __floats2half2_rn(f[2], f[3]);
    __stcs(reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p) + i), v);
  }
};
template <> struct Aux4<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const void* p, int64_t i, float* f) {
    const uint2 v = __ldcs(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p) + i));
    f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
    f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
  }
  static __device__ __forceinline__ void store(void* p, int64_t i, const float* f) {
    uint2 v;
    *reinterpret_cast<__nv_bfloat162*>(&v.x) = __floats2bfloat162_rn(f[0], f[1]);
    *reinterpret_cast<__nv_bfloat162*>(&v.y) = __floats2bfloat162_rn(f[2], f[3]);
    __stcs(reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p) + i), v);
  }
};

// one element: sv = score (in/out), av = parameter / gradient (in/out), ov = output
template <int OP>
__device__ __forceinline__ void score_elem(float& sv, float& av, float& ov, float thr, int mode, float nb) {
  if (OP == kOpProtect) {
    if (sv >= thr) sv = FLT_MAX;
  } else if (OP == kOpMask) {
    ov = sv > thr ? 1.0f : 0.0f;
    av = __fmul_rn(av, ov);
  } else if (OP == kOpAccum) {
    sv = __fadd_rn(sv, mode == 0 ? __fmul_rn(av, av) : fabsf(av));
  } else {
    const float q = __fdiv_rn(sv, nb);
    ov = mode == 2 ? fabsf(q) : mode == 0 ? __fmul_rn(__fmul_rn(av, av), q) : __fmul_rn(fabsf(av), fabsf(q));
  }
}

template <int OP, typename T>
__device__ __forceinline__ void score_unit(const ScoreBatch& b, int p, int64_t e0, int64_t e1, float thr, int mode, float nb) {
  float* s = b.s[p];
  void* aux = b.aux[p];
  float* out = b.out[p];
  constexpr bool kWriteS = OP == kOpProtect || OP == kOpAccum;
  constexpr bool kWriteAux = OP == kOpMask;
  constexpr bool kWriteOut = OP == kOpMask || OP == kOpFinalize;
  // 16-byte vectors when every array of this tensor allows it (torch allocations do); units start at multiples of 64 K
  const bool vec = ((reinterpret_cast<uintptr_t>(s) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0) &&
                   ((reinterpret_cast<uintptr_t>(aux) & (4 * sizeof(T) - 1)) == 0);
  int64_t done = e0;
  if (vec) {
    const int64_t n4 = (e1 - e0) >> 2;
#pragma unroll 4
    for (int64_t v = threadIdx.x; v < n4; v += kSelThreads) {
      const int64_t i = e0 + 4 * v;
      const float4 s4 = __ldcs(reinterpret_cast<const float4*>(s + i));
      float sv[4] = {s4.x, s4.y, s4.z, s4.w};
      float av[4] = {0.f, 0.f, 0.f, 0.f}, ov[4];
      if (aux) Aux4<T>::load(aux, i, av);
#pragma unroll
      for (int e = 0; e < 4; ++e) score_elem<OP>(sv[e], av[e], ov[e], thr, mode, nb);
      if (kWriteS) {
        if (OP != kOpProtect || sv[0] != s4.x || sv[1] != s4.y || sv[2] != s4.z || sv[3] != s4.w)
          *reinterpret_cast<float4*>(s + i) = make_float4(sv[0], sv[1], sv[2], sv[3]);
      }
      if (kWriteAux && aux) Aux4<T>::store(aux, i, av);
      if (kWriteOut && out) __stcs(reinterpret_cast<float4*>(out + i), make_float4(ov[0], ov[1], ov[2], ov[3]));
    }
    done = e0 + 4 * n4;
  }
  for (int64_t i = done + threadIdx.x; i < e1; i += kSelThreads) {
    float sv = s[i], av = aux ? aux_load<T>(aux, i) : 0.f, ov;
    score_elem<OP>(sv, av, ov, thr, mode, nb);
    if (kWriteS) s[i] = sv;
    if (kWriteAux && aux) aux_store<T>(aux, i, av);
    if (kWriteOut && out) out[i] = ov;
  }
}

template <int OP>
__global__ void __launch_bounds__(kSelThreads)
scores_elementwise_kernel(const __grid_constant__ ScoreBatch b, const float* __restrict__ thrs, int mode, float nb) {
  const int64_t total = b.unit_begin[b.count];
  for (int64_t unit = blockIdx.x; unit < total; unit += gridDim.x) {
    const int p = find_item(b, unit);
    const int64_t e0 = (unit - b.unit_begin[p]) * kSelUnit;
    const int64_t e1 = e0 + kSelUnit < b.numel[p] ? e0 + kSelUnit : b.numel[p];
    const float thr = (OP == kOpProtect || OP == kOpMask) ? thrs[b.seg[p]] : 0.0f;
    if (OP == kOpProtect) { score_unit<OP, float>(b, p, e0, e1, thr, mode, nb); continue; }
    switch (b.adt[p]) {
      case VLMC_F16: score_unit<OP, __half>(b, p, e0, e1, thr, mode, nb); break;
      case VLMC_BF16: score_unit<OP, __nv_bfloat16>(b, p, e0, e1, thr, mode, nb); break;
      default: score_unit<OP, float>(b, p, e0, e1, thr, mode, nb); break;
    }
  }
}

// ---- K21: per-tensor sums, fixed order -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSelThreads)
scores_sum_partial_kernel(const __grid_constant__ ScoreBatch b, double* __restrict__ part) {
  __shared__ double red[kSelThreads / 32];
  const int64_t total = b.unit_begin[b.count];
  for (int64_t unit = blockIdx.x; unit < total; unit += gridDim.x) {
    const int p = find_item(b, unit);
    const int64_t e0 = (unit - b.unit_begin[p]) * kSelUnit;
    const int64_t e1 = e0 + kSelUnit < b.numel[p] ? e0 + kSelUnit : b.numel[p];
    const float* __restrict__ s = b.s[p];
    double acc = 0.0;
#pragma unroll 4
    for (int64_t i = e0 + threadIdx.x; i < e1; i += kSelThreads) acc += (double)s[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < kSelThreads / 32; ++w) t += red[w];
      part[unit] = t;
    }
    __syncthreads();
  }
}

__global__ void scores_sum_final_kernel(const __grid_constant__ ScoreBatch b, const double* __restrict__ part,
                                        double* __restrict__ out) {
  const int p = blockIdx.x;
  double acc = 0.0;
  for (int64_t u = b.unit_begin[p] + threadIdx.x; u < b.unit_begin[p + 1]; u += 32) acc += part[u];
  acc = warp_sum(acc);
  if (threadIdx.x == 0) out[p] = acc;
}

// ---- host side ---------------------------------------------------------------------------------------------------------------
static int64_t item_units(int64_t numel) { return (numel + kSelUnit - 1) / kSelUnit; }

// need: bit 0 aux required, bit 1 out required, bit 2 aux optional (validated when present)
static int fill_batch(ScoreBatch& b, const vlmc_score_item* items, int n, int nseg, int need) {
  b.count = n;
  b.unit_begin[0] = 0;
  for (int i = 0; i < n; ++i) {
    const vlmc_score_item& it = items[i];
    if (it.numel < 0) return VLMC_ERR_BAD_ARG;
    if (nseg > 0 && (it.segment < 0 || it.segment >= nseg)) return VLMC_ERR_BAD_ARG;
    if (it.numel > 0) {
      if (!it.scores) return VLMC_ERR_BAD_ARG;
      if (!is_device_ptr(it.scores)) return VLMC_ERR_NOT_DEVICE;
      if ((uintptr_t)it.scores & 3u) return VLMC_ERR_UNSUPPORTED;
      if ((need & 1) && !it.aux) return VLMC_ERR_BAD_ARG;
      if ((need & 2) && !it.out) return VLMC_ERR_BAD_ARG;
      if (it.aux && (need & 5)) {
        if (it.aux_dtype != VLMC_F32 && it.aux_dtype != VLMC_F16 && it.aux_dtype != VLMC_BF16) return VLMC_ERR_BAD_ARG;
        if (!is_device_ptr(it.aux)) return VLMC_ERR_NOT_DEVICE;
        if ((uintptr_t)it.aux & (uintptr_t)(elem_size(it.aux_dtype) - 1)) return VLMC_ERR_UNSUPPORTED;
      }
      if (it.out) {
        if (!is_device_ptr(it.out)) return VLMC_ERR_NOT_DEVICE;
        if ((uintptr_t)it.out & 3u) return VLMC_ERR_UNSUPPORTED;
      }
    }
    b.s[i] = it.scores;
    b.aux[i] = (need & 5) ? it.aux : nullptr;
    b.out[i] = it.out;
    b.numel[i] = it.numel;
    b.seg[i] = it.segment;
    b.adt[i] = (signed char)it.aux_dtype;
    b.unit_begin[i + 1] = b.unit_begin[i] + item_units(it.numel);
  }
  return VLMC_OK;
}

static int grid_for(int64_t units, int per_sm) {
  const int64_t g = (int64_t)kNumSMs * per_sm;
  return (int)(units < g ? units : g);
}

struct SelLayout { size_t state, hist, part, total; };
static SelLayout sel_layout(const vlmc_score_item* items, int count, int nseg) {
  SelLayout l;
  int64_t units = 0;
  for (int i = 0; i < count; ++i) units += items[i].numel > 0 ? item_units(items[i].numel) : 0;
  const size_t ns = (size_t)(nseg < 1 ? 1 : nseg);
  l.state = VLMC_WS_COUNTER_BYTES;
  l.hist = align_up(l.state + ns * sizeof(SelState), 256);
  l.part = align_up(l.hist + ns * kSelBins * sizeof(ull), 256);
  l.total = l.part + (size_t)units * sizeof(double) + 256;
  return l;
}

template <int OP>
static int run_elementwise(const vlmc_score_item* items, int count, int nseg, const float* thrs, int need, int mode,
                           float nb, cudaStream_t st) {
  for (int c0 = 0; c0 < count; c0 += kSelMax) {
    const int n = count - c0 < kSelMax ? count - c0 : kSelMax;
    ScoreBatch b;
    const int rc = fill_batch(b, items + c0, n, nseg, need);
    if (rc != VLMC_OK) return rc;
    const int64_t units = b.unit_begin[n];
    if (units == 0) continue;
    scores_elementwise_kernel<OP><<<grid_for(units, 8), kSelThreads, 0, st>>>(b, thrs, mode, nb);
    const int lc = check_launch();
    if (lc != VLMC_OK) return lc;
  }
  return VLMC_OK;
}

template <int PASS>
static int run_hist_pass(const vlmc_score_item* items, int count, int nseg, SelState* state, ull* hist, float* kth_out,
                         cudaStream_t st) {
  for (int c0 = 0; c0 < count; c0 += kSelMax) {
    const int n = count - c0 < kSelMax ? count - c0 : kSelMax;
    ScoreBatch b;
    const int rc = fill_batch(b, items + c0, n, nseg, 0);
    if (rc != VLMC_OK) return rc;
    const int64_t units = b.unit_begin[n];
    if (units == 0) continue;
    scores_hist_kernel<PASS><<<grid_for(units, 8), kSelThreads, 0, st>>>(b, state, hist);
    const int lc = check_launch();
    if (lc != VLMC_OK) return lc;
  }
  scores_resolve_kernel<PASS><<<nseg, kSelThreads, 0, st>>>(state, hist, kth_out);
  return check_launch();
}

}  // namespace vlmc

extern "C" size_t vlmc_scores_workspace_bytes(const vlmc_score_item* items, int count, int nseg) {
  if (!items || count < 0) return 0;
  return vlmc::sel_layout(items, count, nseg).total;
}

extern "C" int vlmc_scores_kth(const vlmc_score_item* items, int count, int nseg, const int64_t* k, float* kth_out,
                               void* ws, size_t ws_bytes, void* stream) {
  using namespace vlmc;
  if (!items || count < 1 || nseg < 1 || !k || !kth_out || !ws) return VLMC_ERR_BAD_ARG;
  if (!is_device_ptr(k) || !is_device_ptr(kth_out) || !is_device_ptr(ws)) return VLMC_ERR_NOT_DEVICE;
  const SelLayout l = sel_layout(items, count, nseg);
  if (ws_bytes < l.part) return VLMC_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  char* base = reinterpret_cast<char*>(ws);
  SelState* state = reinterpret_cast<SelState*>(base + l.state);
  ull* hist = reinterpret_cast<ull*>(base + l.hist);
  if (cudaMemsetAsync(hist, 0, (size_t)nseg * kSelBins * sizeof(ull), st) != cudaSuccess) return check_launch();
  scores_init_kernel<<<(nseg + 255) / 256, 256, 0, st>>>(state, k, nseg);
  int rc = check_launch();
  if (rc != VLMC_OK) return rc;
  rc = run_hist_pass<0>(items, count, nseg, state, hist, kth_out, st);
  if (rc != VLMC_OK) return rc;
  rc = run_hist_pass<1>(items, count, nseg, state, hist, kth_out, st);
  if (rc != VLMC_OK) return rc;
  return run_hist_pass<2>(items, count, nseg, state, hist, kth_out, st);
}

extern "C" int vlmc_scores_protect(const vlmc_score_item* items, int count, int nseg, const float* thr, void* stream) {
  using namespace vlmc;
  if (!items || count < 1 || nseg < 1 || !thr) return VLMC_ERR_BAD_ARG;
  if (!is_device_ptr(thr)) return VLMC_ERR_NOT_DEVICE;
  return run_elementwise<kOpProtect>(items, count, nseg, thr, 0, 0, 0.0f, (cudaStream_t)stream);
}

extern "C" int vlmc_scores_mask(const vlmc_score_item* items, int count, int nseg, const float* thr, void* stream) {
  using namespace vlmc;
  if (!items || count < 1 || nseg < 1 || !thr) return VLMC_ERR_BAD_ARG;
  if (!is_device_ptr(thr)) return VLMC_ERR_NOT_DEVICE;
  return run_elementwise<kOpMask>(items, count, nseg, thr, 4, 0, 0.0f, (cudaStream_t)stream);
}

extern "C" int vlmc_importance_accum(const vlmc_score_item* items, int count, int mode, void* stream) {
  using namespace vlmc;
  if (!items || count < 1 || mode < 0 || mode > 1) return VLMC_ERR_BAD_ARG;
  return run_elementwise<kOpAccum>(items, count, 0, nullptr, 1, mode, 0.0f, (cudaStream_t)stream);
}

extern "C" int vlmc_importance_finalize(const vlmc_score_item* items, int count, int mode, double num_batches,
                                        void* stream) {
  using namespace vlmc;
  if (!items || count < 1 || mode < 0 || mode > 2 || !(num_batches > 0.0)) return VLMC_ERR_BAD_ARG;
  return run_elementwise<kOpFinalize>(items, count, 0, nullptr, mode == 2 ? 2 : 3, mode, (float)num_batches,
                                      (cudaStream_t)stream);
}

extern "C" int vlmc_scores_sum(const vlmc_score_item* items, int count, double* out, void* ws, size_t ws_bytes,
                               void* stream) {
  using namespace vlmc;
  if (!items || count < 1 || !out || !ws) return VLMC_ERR_BAD_ARG;
  if (!is_device_ptr(out) || !is_device_ptr(ws)) return VLMC_ERR_NOT_DEVICE;
  const SelLayout l = sel_layout(items, count, 1);
  if (ws_bytes < l.total) return VLMC_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  double* part = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + l.part);
  for (int c0 = 0; c0 < count; c0 += kSelMax) {
    const int n = count - c0 < kSelMax ? count - c0 : kSelMax;
    ScoreBatch b;
    const int rc = fill_batch(b, items + c0, n, 0, 0);
    if (rc != VLMC_OK) return rc;
    const int64_t units = b.unit_begin[n];
    if (units > 0) {
      scores_sum_partial_kernel<<<grid_for(units, 8), kSelThreads, 0, st>>>(b, part);
      const int lc = check_launch();
      if (lc != VLMC_OK) return lc;
    }
    scores_sum_final_kernel<<<n, 32, 0, st>>>(b, part, out + c0);
    const int lc = check_launch();
    if (lc != VLMC_OK) return lc;
    part += units;
  }
  return VLMC_OK;
}
