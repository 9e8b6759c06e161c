// K4/K5/K6: Wanda score fused with mask selection (score is never materialised).
//
// Replaces lavis/compression/pruners/wanda_pruner.py:318-341:
//   W_metric = |W| * sqrt(scaler_row)                     (:318)  fp32, one IEEE multiply
//   importance_score = mean(W_metric)                     (:320)
//   unstructured: stable sort per row, first int(C*p) pruned   (:332-337)   -> rowselect_kernel
//   n:m: python loop over C/m groups of torch.topk          (:323-329)   -> nm_kernel
//   module.mask = ~W_mask ; W[W_mask] = 0                  (:339-341)
//
// rowselect: one WARP streams a row; only a short candidate list is kept on chip (see the kernel).
// HBM-bound: 5 B/weight for fp16/bf16 (read W, write W, write mask).
#include <stdlib.h>
#include <type_traits>
#include "common.cuh"

namespace vlmc {

constexpr int kSelThreads = 128;          // n:m kernel
constexpr int kSelWarps = kSelThreads / 32;
constexpr int kRsWarps = 4;               // rowselect: rows in flight per CTA (one per warp)
constexpr int kRsList = 24;               // per-lane candidate list capacity
constexpr int kRsCap = 32;                // candidates ranked exhaustively: one per lane

__device__ __forceinline__ uint32_t redux_add(uint32_t v) {
  return __reduce_add_sync(0xffffffffu, v);
}

__device__ __forceinline__ uint32_t clampu(double x, uint32_t lo, uint32_t hi) {
  // pivot into [lo+1, hi-1] (caller guarantees hi - lo >= 2)
  if (!(x > (double)lo + 1.0)) return lo + 1;
  if (!(x < (double)hi - 1.0)) return hi - 1;
  return (uint32_t)x;
}

// prmt.b32 in its default mode: selector nibble bit 3 replicates the sign bit of the selected byte over the whole byte
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}

__global__ void sqrt_vec_kernel(const float* __restrict__ s, float* __restrict__ out, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) out[i] = __fsqrt_rn(s[i]);
}

// Next four pivots inside the bracket (lo, hi) that holds the k-th smallest key (m keys in it, glo below it).
__device__ __forceinline__ void rs_pivots(uint32_t lo, uint32_t hi, int glo, int m, int k, int C, bool uniform,
                                          uint32_t (&p)[4]) {
  const uint32_t w = hi - lo;
  if (uniform || (m > C / 2 && w > (1u << 26))) {
    // uniform 5-way split of the bit range: shrinks it 5x per pass whatever the distribution
    const double q = (double)w * 0.2;
    p[0] = clampu(lo + q, lo, hi);       p[1] = clampu(lo + 2.0 * q, lo, hi);
    p[2] = clampu(lo + 3.0 * q, lo, hi); p[3] = clampu(lo + 4.0 * q, lo, hi);
  } else {
    // interpolate inside the bracket; inner pivots ~ +-8 keys, outer ~ +-64 keys at uniform density
    const double f = ((double)(k - glo) - 0.5) / (double)m;
    const double e = (double)lo + f * (double)w;
    double d1 = (double)w * (8.0 / (double)m), d2 = (double)w * (64.0 / (double)m);
    if (d2 > (double)w * 0.25) d2 = (double)w * 0.25;
    if (d1 > d2 * 0.25) d1 = d2 * 0.25;
    p[0] = clampu(e - d2, lo, hi); p[1] = clampu(e - d1, lo, hi);
    p[2] = clampu(e + d1, lo, hi); p[3] = clampu(e + d2, lo, hi);
  }
}

__device__ __forceinline__ void rs_update(const uint32_t (&p)[4], const int (&c)[4], int k, uint32_t& lo, uint32_t& hi,
                                          int& glo, int& ghi) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (c[i] <= k - 1) { if (p[i] > lo) { lo = p[i]; glo = c[i]; } }
    else               { if (p[i] < hi) { hi = p[i]; ghi = c[i]; } }
  }
}

// One warp: estimate the per-row threshold from a sample of 128 weight vectors (4 per lane, 1024 scores for 16-bit
// weights) spread over the matrix (row and column strided): the k/C quantile of the sample's scores by bit-space
// bisection.  Seeds the first row of every warp of rowselect_kernel; a miss only costs that row an extra pass.  With
// ~2 rows per warp at R = 4096 more than half of all rows ARE first rows, so the estimate has to be good enough for the
// same narrow bracket the row-to-row seeding uses (1024 samples: ~1.6 % of probability mass at the median).
constexpr int kSeedVecs = 4;
template <typename T>
__global__ void __launch_bounds__(32)
rowselect_seed_kernel(const T* __restrict__ W, int64_t ldw, int R, int C, const float* __restrict__ sq, int k,
                      uint32_t* __restrict__ seed) {
  constexpr int V = Elem<T>::kVec;
  const int lane = threadIdx.x;
  const int nvec = C / V;
  uint32_t sk[kSeedVecs * V];
#pragma unroll
  for (int u = 0; u < kSeedVecs; ++u) {
    const int s = lane * kSeedVecs + u;                       // sample index 0 .. 127
    const int vi = (int)(((int64_t)s * nvec) / (32 * kSeedVecs)) % (nvec > 0 ? nvec : 1);
    const int row = (int)(((int64_t)((s * 37) % (32 * kSeedVecs)) * R) / (32 * kSeedVecs));   // rows decorrelated from columns
    float f[V];
    Elem<T>::unpack(*reinterpret_cast<const uint4*>(W + (int64_t)row * ldw + vi * V), f);
#pragma unroll
    for (int e = 0; e < V; ++e) sk[u * V + e] = __float_as_uint(__fmul_rn(fabsf(f[e]), sq[vi * V + e]));
  }
  const int ns = 32 * kSeedVecs * V;
  int ks = (int)(((int64_t)k * ns + C / 2) / C);
  if (ks < 1) ks = 1;
  if (ks > ns) ks = ns;
  uint32_t slo = 0u, shi = 0x7f800000u;
  while (shi - slo > 65536u) {
    const double q = (double)(shi - slo) * 0.2;
    const uint32_t sp[4] = {clampu(slo + q, slo, shi), clampu(slo + 2.0 * q, slo, shi), clampu(slo + 3.0 * q, slo, shi),
                            clampu(slo + 4.0 * q, slo, shi)};
    uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
    for (int e = 0; e < kSeedVecs * V; ++e) {
      c0 += sk[e] < sp[0] ? 1u : 0u; c1 += sk[e] < sp[1] ? 1u : 0u;
      c2 += sk[e] < sp[2] ? 1u : 0u; c3 += sk[e] < sp[3] ? 1u : 0u;
    }
    const int sc[4] = {(int)redux_add(c0), (int)redux_add(c1), (int)redux_add(c2), (int)redux_add(c3)};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (sc[i] <= ks - 1) { if (sp[i] > slo) slo = sp[i]; }
      else                 { if (sp[i] < shi) shi = sp[i]; }
    }
  }
  if (lane == 0) *seed = slo + ((shi - slo) >> 1);
}

// rowselect: ONE WARP owns a row and STREAMS it; nothing of the row is kept on chip except a short candidate list.
//   pass 1  (HBM)  score every weight (fp32 score bits: non-negative floats order like their uint32 patterns), count the
//                  scores below four pivots seeded from the previous row's threshold, and append the scores that fall
//                  between the outer pivots (~10 % of the row) to per-lane lists in shared memory
//   narrow         counting passes over the lists only (a dozen keys per lane) until <= 32 candidates remain, which are
//                  ranked exactly on (score, column): the stable-sort tie-break of torch.sort(stable=True)
//   apply   (L2)   re-stream the row, recompute the scores, write the mask bytes and the zeroed weights
// If the seeded bracket misses (first row of a warp, outlier rows) or a lane's list overflows, pass 1 is repeated from
// L2 with pivots chosen inside the shrinking bracket.  No CTA barrier, ~6 KB of shared memory per warp: 32 warps per
// SM hide each other's load and reduction latencies.  HBM traffic: 5 B/weight for fp16/bf16.
template <typename T>
__global__ void __launch_bounds__(kRsWarps * 32, 4)
rowselect_kernel(T* __restrict__ W, int64_t ldw, int R, int C, const float* __restrict__ sq, int k, int zero_w,
                 uint8_t* __restrict__ mask, int64_t ldm, float* __restrict__ row_sum, const uint32_t* __restrict__ seed) {
  constexpr int V = Elem<T>::kVec;     // weights per 16-byte vector
  __shared__ uint2 lists[kRsWarps][kRsList][32];     // (score bits, column)
  __shared__ uint2 cand[kRsWarps][kRsCap];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint2 (*mylist)[32] = lists[warp];
  uint2* mycand = cand[warp];
  const int nvec = C / V;              // 16-byte vectors per row
  uint32_t prev_v = seed ? *seed : 0u; // previous row's threshold; a warp's first row starts from the sample estimate
  bool from_sample = true;
  float rd = 786432.f * sqrtf(4096.f / (float)C);
  rd = rd < 200000.f ? 200000.f : (rd > 1400000.f ? 1400000.f : rd);
  const uint32_t row_dlt = (uint32_t)rd;

  // score bits of the V weights of vector `vi`
  auto score = [&](const uint4& wv, int col, uint32_t (&key)[V], float& lsum) {
    float f[V];
    Elem<T>::unpack(wv, f);
#pragma unroll
    for (int q = 0; q < V / 4; ++q) {
      const float4 sv = __ldg(reinterpret_cast<const float4*>(sq + col) + q);
      const float se[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float sc = __fmul_rn(fabsf(f[q * 4 + e]), se[e]);
        key[q * 4 + e] = __float_as_uint(sc);
        lsum += sc;
      }
    }
  };

  for (int row = blockIdx.x * kRsWarps + warp; row < R; row += gridDim.x * kRsWarps) {
    T* wrow = W + (int64_t)row * ldw;
    uint32_t v = 0;
    int iv = -1;  // prune (key < v) || (key == v && col <= iv)
    bool need_sum = true;
    bool ties_all_pruned = false;   // every key equal to v has col <= iv: enables the one-compare apply

    // one streaming pass: counts of keys < p[0..3] -> c, keys in [cl, ch) -> lists; returns false on list overflow.
    // two == true: only the outer pivots p[0], p[3] are counted (the seeded pass: [cl, ch) = [p[0], p[3]))
    auto stream = [&](auto two_tag, const uint32_t (&p)[4], uint32_t cl, uint32_t ch, int (&c)[4], int& lcount) -> bool {
      constexpr bool kTwo = decltype(two_tag)::value;
      uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
      float lsum = 0.f;
      lcount = 0;
      for (int j0 = lane; j0 < nvec; j0 += 128) {
        uint4 wv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (j0 + u * 32 < nvec) wv[u] = need_sum ? ld_stream(wrow + (j0 + u * 32) * V)
                                                   : *reinterpret_cast<const uint4*>(wrow + (j0 + u * 32) * V);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int col = (j0 + u * 32) * V;
          if (j0 + u * 32 < nvec) {
            uint32_t key[V];
            score(wv[u], col, key, lsum);
#pragma unroll
            for (int e = 0; e < V; ++e) {
              const bool l0 = key[e] < p[0], l3 = key[e] < p[3];
              c0 += l0 ? 1u : 0u; c3 += l3 ? 1u : 0u;
              if (!kTwo) { c1 += key[e] < p[1] ? 1u : 0u; c2 += key[e] < p[2] ? 1u : 0u; }
              const bool in = kTwo ? (!l0 && l3) : (key[e] >= cl && key[e] < ch);
              if (in) {
                if (lcount < kRsList) mylist[lcount][lane] = make_uint2(key[e], (uint32_t)(col + e));
                ++lcount;
              }
            }
          }
        }
      }
      c[0] = (int)redux_add(c0); c[3] = (int)redux_add(c3);
      if (kTwo) { c[1] = c[0]; c[2] = c[3]; }
      else { c[1] = (int)redux_add(c1); c[2] = (int)redux_add(c2); }
      if (need_sum) {
        const float rsum = warp_sum(lsum);
        if (lane == 0 && row_sum) row_sum[row] = rsum;
        need_sum = false;
      }
      return __ballot_sync(0xffffffffu, lcount > kRsList) == 0u;
    };
    using Two = std::integral_constant<bool, true>;
    using Four = std::integral_constant<bool, false>;

    if (k >= C || k <= 0) {
      if (k >= C) { v = 0xffffffffu; iv = 0x7fffffff; }
      if (row_sum) {   // the importance score still needs the row sum
        const uint32_t p[4] = {1u, 2u, 3u, 4u};
        int c[4], lc;
        stream(Four{}, p, 0u, 0u, c, lc);
      }
    } else {
      uint32_t lo = 0, hi = 0xffffffffu;
      int glo = 0, ghi = C;      // count(key < lo) = glo <= k - 1 < ghi = count(key < hi)
      uint32_t cl = 0, ch = 0;   // keys in [cl, ch) are on the lists, cb = count(key < cl)
      int cb = 0, lcount = 0;
      bool have = false, uniform = false, first = true, seeded = false;
      int missed = 0;            // after a seeded pass: +1 threshold above the bracket, -1 below
      uint32_t seed_dlt = 0u;
      while (hi - lo > 1u) {
        uint32_t p[4];
        const int m = ghi - glo;
        if (first && prev_v > 4194304u && prev_v < 0x7f000000u) {
          // two pivots around the seed, everything between them collected: +-6.7 % around the previous row's
          // threshold, +-10 % around the matrix-wide sample estimate for a warp's first row
          // (the row-to-row spread of the threshold and the room on the lists both shrink with sqrt(C))
          const uint32_t dlt = from_sample ? row_dlt + (row_dlt >> 1) : row_dlt;
          seed_dlt = dlt;
          p[0] = prev_v - dlt; p[1] = p[0]; p[3] = prev_v + dlt; p[2] = p[3];
          seeded = true;
          cl = p[0]; ch = p[3];
        } else if (missed != 0 && m > 8 * 32) {
          // the seeded bracket missed: exponential search away from the seed on the side the threshold lies (one
          // pass instead of a ~6-pass uniform search of the whole bit range)
          const double d = (double)seed_dlt;
          const double e = (double)prev_v;
          if (missed > 0) {
            p[0] = clampu(e + 2.5 * d, lo, hi); p[1] = clampu(e + 6.0 * d, lo, hi);
            p[2] = clampu(e + 15.0 * d, lo, hi); p[3] = clampu(e + 40.0 * d, lo, hi);
          } else {
            p[3] = clampu(e - 2.5 * d, lo, hi); p[2] = clampu(e - 6.0 * d, lo, hi);
            p[1] = clampu(e - 15.0 * d, lo, hi); p[0] = clampu(e - 40.0 * d, lo, hi);
          }
          cl = 0u; ch = 0u;
          missed = 0;
        } else {
          rs_pivots(lo, hi, glo, m, k, C, uniform, p);
          if (m <= 8 * 32) { cl = lo; ch = hi; }     // the whole bracket fits the lists
          else { cl = p[0]; ch = p[3]; }
        }
        first = false;
        int c[4];
        const bool fits = seeded ? stream(Two{}, p, cl, ch, c, lcount) : stream(Four{}, p, cl, ch, c, lcount);
        if (seeded) missed = (k - 1 >= c[3]) ? 1 : ((c[0] > k - 1) ? -1 : 0);
        seeded = false;
        const bool whole = cl == lo && ch == hi;
        const int below = whole ? glo : c[0];
        const bool inside = whole || (ch != 0u && c[0] <= k - 1 && k - 1 < c[3]);
        rs_update(p, c, k, lo, hi, glo, ghi);
        if (fits && inside) { have = true; cb = below; break; }
        uniform = !uniform && (ghi - glo) * 4 > m;   // interpolation stalled (skewed keys): one uniform pass next
      }
      __syncwarp();
      if (have) {
        // ---- narrow on the lists: entries are (key, col) with key in [cl, ch); the bracket (lo, hi) is inside it
        const int ln = lcount;
        uniform = false;
        while (ghi - glo > kRsCap && hi - lo > 1u) {
          const int m = ghi - glo;
          uint32_t p[4];
          rs_pivots(lo, hi, glo, m, k, C, uniform, p);
          uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
          for (int i = 0; i < ln; ++i) {
            const uint32_t key = mylist[i][lane].x;
            c0 += key < p[0] ? 1u : 0u; c1 += key < p[1] ? 1u : 0u;
            c2 += key < p[2] ? 1u : 0u; c3 += key < p[3] ? 1u : 0u;
          }
          const int c[4] = {cb + (int)redux_add(c0), cb + (int)redux_add(c1), cb + (int)redux_add(c2), cb + (int)redux_add(c3)};
          rs_update(p, c, k, lo, hi, glo, ghi);
          uniform = !uniform && (ghi - glo) * 4 > m;
        }
        const int m = ghi - glo;
        if (m <= kRsCap) {
          // compact the survivors lo <= key < hi and rank them on (key, column), one per lane
          int base = 0;
          const int lmax = __reduce_max_sync(0xffffffffu, ln);
          for (int i = 0; i < lmax; ++i) {
            const uint2 ent = i < ln ? mylist[i][lane] : make_uint2(0xffffffffu, 0u);
            const bool in = ent.x >= lo && ent.x < hi;
            const uint32_t bal = __ballot_sync(0xffffffffu, in);
            if (in) mycand[base + __popc(bal & ((1u << lane) - 1u))] = ent;
            base += __popc(bal);
          }
          __syncwarp();
          bool hit = false;
          uint2 mine = make_uint2(0u, 0u);
          if (lane < m) {
            mine = mycand[lane];
            int rank = 0;
            for (int j = 0; j < m; ++j) {
              const uint2 o = mycand[j];
              rank += (o.x < mine.x || (o.x == mine.x && o.y < mine.y)) ? 1 : 0;
            }
            hit = rank == k - 1 - glo;
          }
          const int src = __ffs(__ballot_sync(0xffffffffu, hit)) - 1;
          v = __shfl_sync(0xffffffffu, mine.x, src);
          iv = (int)__shfl_sync(0xffffffffu, mine.y, src);
          // all keys equal to v are among the m candidates (they lie in [lo, hi)): is any of them kept?
          ties_all_pruned = __ballot_sync(0xffffffffu, lane < m && mine.x == v && (int)mine.y > iv) == 0u;
        } else {
          // more than 32 exact ties at the threshold value, all on the lists: bisect on the column index
          v = lo;
          const int r = k - glo;  // how many of the ties are pruned (lowest columns first)
          int qlo = 0, qhi = C;   // count(key==v && col < qlo) < r <= count(key==v && col < qhi)
          while (qhi - qlo > 1) {
            const int qm = (qlo + qhi) >> 1;
            uint32_t cnt = 0;
            for (int i = 0; i < ln; ++i) {
              const uint2 o = mylist[i][lane];
              cnt += (o.x == v && (int)o.y < qm) ? 1u : 0u;
            }
            cnt = redux_add(cnt);
            if ((int)cnt < r) qlo = qm; else qhi = qm;
          }
          iv = qlo;  // columns <= qlo with key == v are exactly the r lowest ties
        }
      } else {
        // hi - lo == 1: every key of the bracket equals lo and there are too many for the lists (all-zero rows, dead
        // channels): bisect on the column index with streaming counts
        v = lo;
        const int r = k - glo;
        int qlo = 0, qhi = C;
        while (qhi - qlo > 1) {
          const int qm = (qlo + qhi) >> 1;
          uint32_t cnt = 0;
          float dummy = 0.f;
          for (int j = lane; j < nvec; j += 32) {
            const uint4 wv = *reinterpret_cast<const uint4*>(wrow + j * V);
            uint32_t key[V];
            score(wv, j * V, key, dummy);
#pragma unroll
            for (int e = 0; e < V; ++e) cnt += (key[e] == v && j * V + e < qm) ? 1u : 0u;
          }
          cnt = redux_add(cnt);
          if ((int)cnt < r) qlo = qm; else qhi = qm;
        }
        iv = qlo;
      }
      prev_v = v;
      from_sample = false;
    }

    // ---- apply: re-stream the row (L2), mask bytes (1 = keep) and zeroed weights.
    // Fast form: when every tie at the threshold is pruned (nearly always: a single key equals v) the rule
    // "key < v or (key == v and col <= iv)" is "key < v + 1"; keys are < 2^31, so the sign bit of key - (v + 1) is the
    // prune flag and byte permutes turn four of them into mask bytes / 16-bit weight masks without a compare.
    uint8_t* mrow = mask + (int64_t)row * ldm;
    const bool fast = ties_all_pruned && v < 0x7fffffffu;
    const uint32_t vcut = v + 1u;
    for (int j0 = lane; j0 < nvec; j0 += 128) {
      uint4 wv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (j0 + u * 32 < nvec) wv[u] = *reinterpret_cast<const uint4*>(wrow + (j0 + u * 32) * V);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int col = (j0 + u * 32) * V;
        if (j0 + u * 32 < nvec) {
          uint32_t key[V];
          float dummy = 0.f;
          score(wv[u], col, key, dummy);
          uint32_t mb[V / 4];
          uint32_t* wr = reinterpret_cast<uint32_t*>(&wv[u]);
          if (fast) {
            uint32_t d[V];
#pragma unroll
            for (int e = 0; e < V; ++e) d[e] = key[e] - vcut;                       // sign bit set <=> pruned
#pragma unroll
            for (int q = 0; q < V / 4; ++q) {
              const uint32_t t01 = __byte_perm(d[q * 4], d[q * 4 + 1], 0x0073);      // top bytes of d0, d1
              const uint32_t t23 = __byte_perm(d[q * 4 + 2], d[q * 4 + 3], 0x0073);
              const uint32_t tops = __byte_perm(t01, t23, 0x5410);                   // [d0 d1 d2 d3] top bytes
              mb[q] = ((tops >> 7) & 0x01010101u) ^ 0x01010101u;                     // 1 = keep
            }
            if (zero_w) {
              if (sizeof(T) == 4) {
#pragma unroll
                for (int e = 0; e < V; ++e) wr[e] &= ~(uint32_t)((int32_t)d[e] >> 31);
              } else {
#pragma unroll
                for (int e = 0; e < V; e += 2)   // sign-replicated top bytes -> 0xffff per pruned half
                  wr[e / 2] &= ~prmt(d[e], d[e + 1], 0xffbbu);
              }
            }
          } else {
#pragma unroll
            for (int q = 0; q < V / 4; ++q) mb[q] = 0u;
#pragma unroll
            for (int e = 0; e < V; ++e) {
              const bool pr = (key[e] < v) || (key[e] == v && (col + e) <= iv);
              mb[e / 4] |= (pr ? 0u : 1u) << (8 * (e % 4));
            }
            if (zero_w) {
              if (sizeof(T) == 4) {
#pragma unroll
                for (int e = 0; e < V; ++e) if (!((mb[e / 4] >> (8 * (e % 4))) & 1u)) wr[e] = 0u;
              } else {
#pragma unroll
                for (int e = 0; e < V; ++e)
                  if (!((mb[e / 4] >> (8 * (e % 4))) & 1u)) wr[e / 2] &= (e & 1) ? 0x0000ffffu : 0xffff0000u;
              }
            }
          }
          if (V == 8) st_stream8(mrow + col, make_uint2(mb[0], mb[V / 4 - 1]));
          else st_stream4(mrow + col, mb[0]);
          if (zero_w) {
            bool any = false;
#pragma unroll
            for (int q = 0; q < V / 4; ++q) any |= mb[q] != 0x01010101u;
            if (any) st_stream(wrow + col, wv[u]);
          }
        }
      }
    }
    __syncwarp();
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
    constexpr int kRows = 4;               // rows in flight per thread: kRows x NVEC 16-byte loads before any use
    for (int row0 = blockIdx.y; row0 < R; row0 += gridDim.y * kRows) {
      uint4 wv[kRows][NVEC];
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        const int row = row0 + r * gridDim.y;
        if (row < R) {
#pragma unroll
          for (int q = 0; q < NVEC; ++q) wv[r][q] = ld_stream(W + (int64_t)row * ldw + col + q * V);
        }
      }
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        const int row = row0 + r * gridDim.y;
        if (row >= R) break;
        T* wp = W + (int64_t)row * ldw + col;
        uint32_t keys[E];
#pragma unroll
        for (int q = 0; q < NVEC; ++q) {
          float f[V];
          Elem<T>::unpack(wv[r][q], f);
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
            uint32_t* wr = reinterpret_cast<uint32_t*>(&wv[r][q]);
            if (sizeof(T) == 4) {
#pragma unroll
              for (int e = 0; e < V; ++e) if (pr[q * V + e]) wr[e] = 0u;
            } else {
#pragma unroll
              for (int e = 0; e < V; ++e)
                if (pr[q * V + e]) wr[e / 2] &= (e & 1) ? 0x0000ffffu : 0xffff0000u;
            }
            st_stream(wp + q * V, wv[r][q]);
          }
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

// ---- n:m, all linears of a block in ONE launch ----------------------------------------------
// A Vicuna-7B block is 7 matrices of 34-90 MB; launched one by one each kernel lives 15-40 us and its ramp-up and
// drain cost as much as a third of that.  Here the block is one list of work units (kNmbRows rows x 1024 columns of
// one matrix) that a single grid of resident CTAs walks with a fixed stride: one ramp, one drain, every CTA within
// one unit of the others at the end.  Per unit: the same loads, ranks and stores as nm_kernel.
constexpr int kNmbMax = 16;       // matrices per launch
constexpr int kNmbRows = 16;      // rows per work unit (4 in flight x 4)
struct NmBatchItem {
  void* W; int64_t ldw; int R, C; const float* scaler_row; uint8_t* mask; int64_t ldm;
  int coltiles;                   // ceil(C / (kSelThreads * E))
};
struct NmBatch {
  NmBatchItem it[kNmbMax];
  int unit_begin[kNmbMax + 1];    // units of item i: [unit_begin[i], unit_begin[i+1])
  int count;
};

template <typename T, int M>
__global__ void __launch_bounds__(kSelThreads, (M <= 8 && sizeof(T) == 2) ? 8 : 3)
nm_batch_kernel(const __grid_constant__ NmBatch b, int n, int zero_w, float* __restrict__ part_sum) {
  constexpr int V = Elem<T>::kVec;
  constexpr int E = (M > V) ? M : V;
  constexpr int NVEC = E / V;
  const int total = b.unit_begin[b.count];
  const int warp = threadIdx.x >> 5;
  for (int unit = blockIdx.x; unit < total; unit += gridDim.x) {
    int p = 0;
    while (unit >= b.unit_begin[p + 1]) ++p;
    const NmBatchItem& it = b.it[p];
    const int local = unit - b.unit_begin[p];
    const int ct = local % it.coltiles, rb = local / it.coltiles;
    const int col = (ct * kSelThreads + threadIdx.x) * E;
    float lsum = 0.f;
    if (col < it.C) {
      T* W = reinterpret_cast<T*>(it.W);
      float sq[E];
#pragma unroll
      for (int q = 0; q < E / 4; ++q) {
        const float4 sv = __ldg(reinterpret_cast<const float4*>(it.scaler_row + col) + q);
        sq[q * 4] = __fsqrt_rn(sv.x); sq[q * 4 + 1] = __fsqrt_rn(sv.y);
        sq[q * 4 + 2] = __fsqrt_rn(sv.z); sq[q * 4 + 3] = __fsqrt_rn(sv.w);
      }
      const int row_begin = rb * kNmbRows;
      const int row_end = row_begin + kNmbRows < it.R ? row_begin + kNmbRows : it.R;
      constexpr int kRows = 4;
      for (int row0 = row_begin; row0 < row_end; row0 += kRows) {
        uint4 wv[kRows][NVEC];
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
          if (row0 + r < row_end) {
#pragma unroll
            for (int q = 0; q < NVEC; ++q) wv[r][q] = ld_stream(W + (int64_t)(row0 + r) * it.ldw + col + q * V);
          }
        }
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
          const int row = row0 + r;
          if (row >= row_end) break;
          T* wp = W + (int64_t)row * it.ldw + col;
          uint32_t keys[E];
#pragma unroll
          for (int q = 0; q < NVEC; ++q) {
            float f[V];
            Elem<T>::unpack(wv[r][q], f);
#pragma unroll
            for (int e = 0; e < V; ++e) {
              const float sc = __fmul_rn(fabsf(f[e]), sq[q * V + e]);
              keys[q * V + e] = __float_as_uint(sc);
              lsum += sc;
            }
          }
          bool pr[E];
#pragma unroll
          for (int g = 0; g < E / M; ++g) {
#pragma unroll
            for (int a = 0; a < M; ++a) {
              int rank = 0;
#pragma unroll
              for (int c = 0; c < M; ++c) {
                if (c < a) rank += keys[g * M + c] <= keys[g * M + a] ? 1 : 0;
                if (c > a) rank += keys[g * M + c] < keys[g * M + a] ? 1 : 0;
              }
              pr[g * M + a] = rank < n;
            }
          }
          uint8_t* mp = it.mask + (int64_t)row * it.ldm + col;
#pragma unroll
          for (int q = 0; q < NVEC; ++q) {
            uint32_t mb[V / 4] = {};
#pragma unroll
            for (int e = 0; e < V; ++e) mb[e / 4] |= (pr[q * V + e] ? 0u : 1u) << (8 * (e % 4));
            if (V == 8) st_stream8(mp + q * V, make_uint2(mb[0], mb[V / 4 - 1]));
            else st_stream4(mp + q * V, mb[0]);
            if (zero_w) {
              uint32_t* wr = reinterpret_cast<uint32_t*>(&wv[r][q]);
              if (sizeof(T) == 4) {
#pragma unroll
                for (int e = 0; e < V; ++e) if (pr[q * V + e]) wr[e] = 0u;
              } else {
#pragma unroll
                for (int e = 0; e < V; ++e)
                  if (pr[q * V + e]) wr[e / 2] &= (e & 1) ? 0x0000ffffu : 0xffff0000u;
              }
              st_stream(wp + q * V, wv[r][q]);
            }
          }
        }
      }
    }
    // one partial per (unit, warp): no CTA barrier between units, fixed summation order in the finalize
    lsum = warp_sum(lsum);
    if ((threadIdx.x & 31) == 0) part_sum[(int64_t)unit * kSelWarps + warp] = lsum;
  }
}

struct NmBatchOut { float* out[kNmbMax]; double denom[kNmbMax]; int begin[kNmbMax + 1]; };

// grid = items: item i sums its partials [begin[i], begin[i+1]) in a fixed order -> mean
__global__ void __launch_bounds__(256)
mean_finalize_batch_kernel(const float* __restrict__ part, const __grid_constant__ NmBatchOut o) {
  __shared__ double sred[8];
  const int i = blockIdx.x;
  if (!o.out[i]) return;
  double s = 0.0;
  for (int j = o.begin[i] + threadIdx.x; j < o.begin[i + 1]; j += 256) s += (double)part[j];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sred[w];
    *o.out[i] = (float)(t / o.denom[i]);
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

// ---------------------------------------------------------------------------------------------------------------------
// rowselect, CTA-per-row form (default).  The warp-per-row kernel above streams a row several times through registers
// and spends ~38 thread instructions per weight (ncu r01: 109 registers, 24 % of the warps resident, 0.21 of the HBM
// roofline).  Here a 256-thread CTA owns a row and the row is touched in HBM exactly once:
//   P1  one 16-byte load per vector (the row stays in REGISTERS for the apply pass), score bits -> shared memory (each
//       thread only ever reads back its own keys: no barrier), row sum / maximum by shuffle, and - in the same pass - a
//       shared-memory histogram over LINEAR bins of width (1.25 x the previous row's maximum) / 2048.  Binning by
//       floor(score * scale) is monotone in the score, so bin order is score order whatever the scale; unlike the
//       exponent-heavy top bits of the float pattern the bins are evenly loaded (a handful of keys each): no atomic
//       contention, and the bin of the k-th score holds a few candidates only
//   P2  warp 0 scans the 2048 counters -> the bin of the k-th smallest score and the rank inside it (and clears them)
//   P3  the keys of that bin are collected with their columns; EVERY warp then ranks the (<= 32 typical) candidates on
//       (score, column) - the stable-sort tie-break of torch.sort(stable=True), wanda_pruner.py:332 - so the threshold
//       pair is known without another barrier
//   P4  apply from registers + shared memory: prune (key < v) || (key == v && column <= iv); 8-byte mask stores,
//       16-byte weight stores only for vectors that lose a weight
// Three CTA barriers per row.  Degenerate rows (all-zero or non-finite scale, more than kRcCand keys in the chosen bin:
// heavy ties, an outlier row whose k-th score falls in the overflow bin) take an exact radix select on the score bits
// and then on the column index, in the same shared memory.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kRcThreads = 128;           // 4 warps per row: the per-row bookkeeping (scan, ranking, barriers) is amortised
                                          // over 32+ weights per thread, and 5 independent rows are resident per SM
constexpr int kRcWarps = kRcThreads / 32;
constexpr int kRcBins = 1024;           // counters per row: the scan by one warp is on every row's critical path
constexpr int kRcHist = 2048;            // counters allocated: the exact path's radix select uses all of them (11-bit digits)
constexpr int kRcCand = 256;
constexpr int kRcMaxVec = 16;             // 16-byte vectors per thread held in registers: C <= 128 * 16 * V

struct RcShared {
  float wsum[kRcWarps];
  uint32_t wmax[kRcWarps];
  uint32_t scan[kRcWarps + 2];
  uint32_t ncand, sel_bin, sel_before, thr_key;
  int thr_col;
};

// One warp: bin (among NB counters) that holds the kk-th (1-indexed) entry and the count before it -> sh.sel_bin,
// sh.sel_before; the counters are cleared on the way.  64 counters per lane.
template <int NB>
__device__ __forceinline__ void rc_scan_warp(uint32_t* hist, uint32_t kk, RcShared& sh) {
  const int lane = threadIdx.x & 31;
  constexpr int per = NB / 32;                                // lane owns bins [lane * per, lane * per + per)
  // rotated walk in 16-byte steps: at step j lane reads vector (j + lane) % (per / 4) of its segment; the eight lanes of a
  // quarter warp (one LDS.128 phase) then hit eight different 16-byte bank groups: conflict-free
  uint32_t local = 0;
#pragma unroll
  for (int j = 0; j < per / 4; ++j) {
    const uint4 q = *reinterpret_cast<const uint4*>(hist + lane * per + ((j + lane) & (per / 4 - 1)) * 4);
    local += q.x + q.y + q.z + q.w;
  }
  uint32_t incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const uint32_t excl = incl - local;
  // the segment that holds rank kk is walked by ALL lanes together: two bins per lane, one more 32-wide scan each
  const uint32_t owner_mask = __ballot_sync(0xffffffffu, excl < kk && kk <= incl);
  const int owner = __ffs(owner_mask) - 1;                    // exactly one lane when 1 <= kk <= total
  if (owner >= 0) {
    const uint32_t base = __shfl_sync(0xffffffffu, excl, owner);
    uint32_t run = base;
#pragma unroll
    for (int h = 0; h < per / 32; ++h) {
      const uint32_t c = hist[owner * per + h * 32 + lane];
      uint32_t inc = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      const uint32_t before = run + inc - c;
      if (before < kk && kk <= before + c) { sh.sel_bin = owner * per + h * 32 + lane; sh.sel_before = before; }
      run += __shfl_sync(0xffffffffu, inc, 31);
    }
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < per / 4; ++j)
    *reinterpret_cast<uint4*>(hist + lane * per + ((j + lane) & (per / 4 - 1)) * 4) = make_uint4(0u, 0u, 0u, 0u);
}

// All threads: same result in registers (used by the exact path only; two barriers inside)
__device__ __forceinline__ void rc_find_bin(uint32_t* hist, uint32_t kk, RcShared& sh, uint32_t& bin, uint32_t& before) {
  __syncthreads();
  if (threadIdx.x < 32) rc_scan_warp<kRcHist>(hist, kk, sh);
  __syncthreads();
  bin = sh.sel_bin;
  before = sh.sel_before;
}

// bin word of a score: bits(min(fma(score, scale, 2^23), 2^23 + kRcBins - 1)) = kRcMagic | bin.  Round-to-nearest of a
// monotone function of the score (still monotone), one FFMA + one FMNMX; a NaN score lands in the last bin (fminf)
constexpr uint32_t kRcMagic = 0x4B000000u;                   // bits of 8388608.0f
__device__ __forceinline__ uint32_t rc_binbits(uint32_t key, float scale) {
  return __float_as_uint(fminf(__fmaf_rn(__uint_as_float(key), scale, 8388608.0f), 8388608.0f + (float)(kRcBins - 1)));
}

__device__ __forceinline__ uint32_t rc_prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}

// scores of one 16-byte vector of weights: key[e] = bits(|w[e]| * sq[col + e]) (non-negative floats order like their bits)
template <typename T>
__device__ __forceinline__ void rc_keys(const uint4& wv, const float* __restrict__ sqv, uint32_t (&key)[Elem<T>::kVec]) {
  constexpr int V = Elem<T>::kVec;
  float f[V];
  Elem<T>::unpack(wv, f);
#pragma unroll
  for (int q = 0; q < V / 4; ++q) {
    const float4 sv = __ldg(reinterpret_cast<const float4*>(sqv) + q);
    key[4 * q] = __float_as_uint(__fmul_rn(fabsf(f[4 * q]), sv.x));
    key[4 * q + 1] = __float_as_uint(__fmul_rn(fabsf(f[4 * q + 1]), sv.y));
    key[4 * q + 2] = __float_as_uint(__fmul_rn(fabsf(f[4 * q + 2]), sv.z));
    key[4 * q + 3] = __float_as_uint(__fmul_rn(fabsf(f[4 * q + 3]), sv.w));
  }
}

// zero the pruned elements of a packed vector without unpacking it: pm bit e = element e is pruned
template <typename T>
__device__ __forceinline__ uint4 rc_zero(const uint4& wv, uint32_t pm) {
  uint32_t w[4] = {wv.x, wv.y, wv.z, wv.w};
  if (Elem<T>::kVec == 8) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t keep = ((pm >> (2 * i)) & 1u ? 0u : 0x0000ffffu) | ((pm >> (2 * i + 1)) & 1u ? 0u : 0xffff0000u);
      w[i] &= keep;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = (pm >> i) & 1u ? 0u : w[i];
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// Up to kRcBatchMax matrices of ONE row length per launch (vlmc_wanda_rowselect_batch: the linears of a block that share
// C): the CTAs walk the concatenated rows with a fixed stride, so a CTA sees 25-30 rows instead of 5 (the first row of
// a CTA and of every matrix it enters has no predecessor to take its bin scale from), and a block costs 2 launches, not 7.
constexpr int kRcBatchMax = 16;
struct RcItem { void* W; int64_t ldw; const float* sq; int k; uint8_t* mask; int64_t ldm; float* row_sum; };
struct RcBatch { RcItem it[kRcBatchMax]; int row_begin[kRcBatchMax + 1]; int count; };

template <typename T, int NV, bool FULL>      // FULL: C / V == NV * 128, no bounds checks on the vectors
__global__ void __launch_bounds__(kRcThreads, NV <= 4 ? 6 : (NV <= 8 ? 4 : 3))
rowselect_cta_kernel(const __grid_constant__ RcBatch bt, int C, int zero_w) {
  constexpr int V = Elem<T>::kVec;
  extern __shared__ __align__(16) uint32_t rc_smem[];
  const int nvec = C / V;
  // key planes: plane q holds elements 4q..4q+3 of every vector as one uint4 per vector (conflict-free 16-byte accesses)
  uint32_t* keys = rc_smem;                                   // [V / 4][nvec][4]
  uint32_t* hist = keys + (size_t)C;                          // [kRcHist]; the fast path uses the first kRcBins
  uint32_t* cand_key = hist + kRcHist;                        // [kRcCand]
  uint32_t* cand_col = cand_key + kRcCand;                    // [kRcCand]
  __shared__ RcShared sh;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int b = tid; b < kRcHist; b += kRcThreads) hist[b] = 0;
  if (tid == 0) sh.ncand = 0;

#define RC_HAS(u) (FULL || tid + (u) * kRcThreads < nvec)
  // Bin scale.  Rows of one matrix share sq and the weight distribution, so the previous row's threshold predicts this
  // row's: scale = 1024 / previous threshold puts the k-th score near the MIDDLE of the 2048 bins with ~0.1 % of its value
  // per bin (2-3 keys per bin at C = 4096), and everything above twice the threshold in the last bin, which is never
  // counted (the k-th score is found below it, or the row takes the exact path).  The first row a CTA takes of a matrix
  // has no predecessor: its scale comes from its own maximum (coarser: more candidates, same result).
  // (Double-buffering the next row's vectors in registers was measured: 38.0 -> 38.3 us at 4096^2 - with 5-6 CTAs per SM the
  // HBM latency of a row's loads is already covered by the other CTAs - and it costs 10 registers; dropped.)
  uint4 wv[NV];
  float scale = 0.f;
  int item = -1, cur = 0;
  const int total_rows = bt.row_begin[bt.count];

  for (int vrow = blockIdx.x; vrow < total_rows; vrow += gridDim.x) {
    while (vrow >= bt.row_begin[cur + 1]) ++cur;
    const RcItem& im = bt.it[cur];
    const int row = vrow - bt.row_begin[cur];
    const int k = im.k;
    const float* __restrict__ sq = im.sq;
    T* wrow = reinterpret_cast<T*>(im.W) + (int64_t)row * im.ldw;
#pragma unroll
    for (int u = 0; u < NV; ++u)
      if (RC_HAS(u)) wv[u] = ld_stream(wrow + (int64_t)(tid + u * kRcThreads) * V);
    if (cur != item) {                                          // first row of this CTA in this matrix: scale from the row's maximum
      item = cur;
      uint32_t lmax = 0;
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        if (RC_HAS(u)) {
          uint32_t key[V];
          rc_keys<T>(wv[u], sq + (tid + u * kRcThreads) * V, key);
#pragma unroll
          for (int e = 0; e < V; ++e) lmax = key[e] > lmax ? key[e] : lmax;
        }
      }
      lmax = __reduce_max_sync(0xffffffffu, lmax);
      if (lane == 0) sh.wmax[warp] = lmax;
      __syncthreads();
      uint32_t rmax = 0;
#pragma unroll
      for (int w = 0; w < kRcWarps; ++w) rmax = sh.wmax[w] > rmax ? sh.wmax[w] : rmax;
      scale = (rmax > 0u && rmax < 0x7f800000u) ? (float)kRcBins / (1.25f * __uint_as_float(rmax)) : 0.f;
      __syncthreads();
    }
    // ---- P1: score, keys -> smem, row sum, linear-bin histogram
    float lsum = 0.f;
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      if (RC_HAS(u)) {
        const int vi = tid + u * kRcThreads;
        uint32_t key[V];
        rc_keys<T>(wv[u], sq + vi * V, key);
#pragma unroll
        for (int e = 0; e < V; ++e) {
          lsum += __uint_as_float(key[e]);
          atomicAdd(&hist[rc_binbits(key[e], scale) - kRcMagic], 1u);      // no branch: the last bin is counted too
        }
#pragma unroll
        for (int q = 0; q < V / 4; ++q)
          *reinterpret_cast<uint4*>(keys + ((size_t)q * nvec + vi) * 4) = make_uint4(key[4 * q], key[4 * q + 1], key[4 * q + 2], key[4 * q + 3]);
      }
    }
    lsum = warp_sum(lsum);
    if (lane == 0) sh.wsum[warp] = lsum;
    __syncthreads();                                           // barrier A: histogram, partial sums
    const bool select = k > 0 && k < C;
    // ---- P2 (warp 0): the bin of the k-th smallest score
    if (warp == 0) {
      if (lane == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kRcWarps; ++w) t += sh.wsum[w];
        im.row_sum[row] = t;
        sh.sel_bin = (uint32_t)(kRcBins - 1);                  // stays there when the k-th score is not below the last bin
        sh.sel_before = 0;
        sh.ncand = 0;                                          // every thread is past barrier A: done with the previous row's list
      }
      __syncwarp();
      rc_scan_warp<kRcBins>(hist, select ? (uint32_t)k : 1u, sh);
    }
    __syncthreads();                                           // barrier B: sel_bin / sel_before, counters cleared
    uint32_t thr_key = 0;
    int thr_col = -1;                                          // prune (key < thr_key) || (key == thr_key && col <= thr_col)
    bool ties_simple = true;                                   // every key equal to thr_key is pruned: one compare in P4
    if (k >= C) {
      thr_key = 0xffffffffu;
    } else if (k > 0) {
      const uint32_t b_sel = sh.sel_bin;
      const uint32_t want = kRcMagic | b_sel;
      const uint32_t kk = (uint32_t)k - sh.sel_before;         // 1-indexed rank inside the bin
      bool exact_path = b_sel == (uint32_t)(kRcBins - 1);      // the k-th score is in the overflow bin (uniform across the CTA)
      if (!exact_path) {
        // ---- P3: collect the bin's keys (a few per row: one branch per vector, taken by few threads)
#pragma unroll
        for (int u = 0; u < NV; ++u) {
          if (RC_HAS(u)) {
            const int vi = tid + u * kRcThreads;
            uint32_t ke[V];
            bool any = false;
#pragma unroll
            for (int q = 0; q < V / 4; ++q) {
              const uint4 kq = *reinterpret_cast<const uint4*>(keys + ((size_t)q * nvec + vi) * 4);
              ke[4 * q] = kq.x; ke[4 * q + 1] = kq.y; ke[4 * q + 2] = kq.z; ke[4 * q + 3] = kq.w;
            }
            // membership test without the clamp: an unclamped value above the last bin, a clamped one and a NaN all differ
            // from `want` (which is below the last bin here): one FFMA + one FSETP per key, the OR rides on the compare
            const float wantf = __uint_as_float(want);
#pragma unroll
            for (int e = 0; e < V; ++e) any |= __fmaf_rn(__uint_as_float(ke[e]), scale, 8388608.0f) == wantf;
            if (any) {
#pragma unroll
              for (int e = 0; e < V; ++e) {
                if (rc_binbits(ke[e], scale) == want) {
                  const uint32_t pos = atomicAdd(&sh.ncand, 1u);
                  if (pos < (uint32_t)kRcCand) { cand_key[pos] = ke[e]; cand_col[pos] = (uint32_t)(vi * V + e); }
                }
              }
            }
          }
        }
        __syncthreads();                                       // barrier C: candidate list
        const uint32_t nc = sh.ncand;
        if (nc <= (uint32_t)kRcCand) {
          // every warp ranks the (few) candidates itself: the threshold pair lands in registers without another barrier
          uint32_t fk = 0, fc = 0;
          for (uint32_t c0 = 0; c0 < nc; c0 += 32) {
            const uint32_t i = c0 + lane;
            const uint32_t mk = i < nc ? cand_key[i] : 0xffffffffu, mc = i < nc ? cand_col[i] : 0xffffffffu;
            uint32_t rank = 0;
            for (uint32_t j = 0; j < nc; ++j) {
              const uint32_t ok = cand_key[j], oc = cand_col[j];
              rank += (ok < mk || (ok == mk && oc < mc)) ? 1u : 0u;
            }
            const uint32_t ball = __ballot_sync(0xffffffffu, i < nc && rank == kk - 1);
            if (ball) {
              const int src = __ffs(ball) - 1;
              fk = __shfl_sync(0xffffffffu, mk, src);
              fc = __shfl_sync(0xffffffffu, mc, src);
            }
          }
          thr_key = fk;
          thr_col = (int)fc;
          uint32_t kept_eq = 0;                                // an equal key at a higher column stays: P4 needs the column
          for (uint32_t c0 = 0; c0 < nc; c0 += 32) {
            const uint32_t i = c0 + lane;
            kept_eq |= __ballot_sync(0xffffffffu, i < nc && cand_key[i] == fk && cand_col[i] > fc);
          }
          ties_simple = kept_eq == 0;
        } else {
          exact_path = true;
        }
      }
      if (exact_path) {
        // ---- exact radix select on the score bits (31 significant bits: 11 + 11 + 9), then on the column index
        __syncthreads();
        for (int b = tid; b < kRcHist; b += kRcThreads) hist[b] = 0;
        uint32_t prefix = 0, kr = (uint32_t)k, before = 0, bsel = 0;
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass) {
          const int shift = pass == 0 ? 20 : (pass == 1 ? 9 : 0);
          const uint32_t himask = pass == 0 ? 0u : (pass == 1 ? 0xfff00000u : 0xfffffe00u);
          const uint32_t bmask = pass == 2 ? 0x1ffu : 0x7ffu;
          __syncthreads();
          for (int j = tid; j < C; j += kRcThreads) {
            const uint32_t key = keys[j];
            if ((key & himask) == prefix) atomicAdd(&hist[(key >> shift) & bmask], 1u);
          }
          rc_find_bin(hist, kr, sh, bsel, before);
          kr -= before;
          prefix |= bsel << shift;
        }
        // prefix = the k-th smallest key; kr = how many of the keys equal to it are pruned (lowest columns first)
        uint32_t cprefix = 0;
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {                  // columns < 2^22: 11 + 11 bits
          const int shift = pass == 0 ? 11 : 0;
          __syncthreads();
          for (int j = tid; j < C; j += kRcThreads) {
            if (keys[j] == prefix) {
              const int plane = j / (nvec * 4), rem = j - plane * nvec * 4;
              const uint32_t col = (uint32_t)((rem >> 2) * V + plane * 4 + (rem & 3));
              if (pass == 0 || (col >> 11) == cprefix) atomicAdd(&hist[(col >> shift) & 0x7ffu], 1u);
            }
          }
          rc_find_bin(hist, kr, sh, bsel, before);
          kr -= before;
          if (pass == 0) cprefix = bsel;
        }
        thr_key = prefix;
        thr_col = (int)((cprefix << 11) | bsel);
        ties_simple = false;
      }
    }

    // ---- P4: apply (weights are zeroed in their packed form; no unpack / repack)
    uint8_t* mrow = im.mask + (int64_t)row * im.ldm;
    if (ties_simple && k > 0 && k < C) {
      // Fast form: pruned <=> key <= thr_key, both below 2^31, so the sign of (thr_key - key) is the keep bit.  PRMT in its
      // sign-replicating mode (selector nibble 8 | byte) turns that sign straight into the AND mask of the packed weight
      // and into the mask bytes: ~3 instructions per weight instead of a compare / select / shift chain.
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        if (RC_HAS(u)) {
          const int vi = tid + u * kRcThreads;
          uint32_t d[V];
#pragma unroll
          for (int q = 0; q < V / 4; ++q) {
            const uint4 kq = *reinterpret_cast<const uint4*>(keys + ((size_t)q * nvec + vi) * 4);
            d[4 * q] = thr_key - kq.x; d[4 * q + 1] = thr_key - kq.y; d[4 * q + 2] = thr_key - kq.z; d[4 * q + 3] = thr_key - kq.w;
          }
          uint32_t w[4] = {wv[u].x, wv[u].y, wv[u].z, wv[u].w};
          if (V == 8) {
            uint32_t km[4];                                      // bytes {s(2i), s(2i), s(2i+1), s(2i+1)}, s = 0xff kept / 0x00 pruned
#pragma unroll
            for (int i = 0; i < 4; ++i) { km[i] = rc_prmt(d[2 * i], d[2 * i + 1], 0xFFBBu); w[i] &= km[i]; }
            const uint32_t b0 = rc_prmt(km[0], km[1], 0x6420u) & 0x01010101u;
            const uint32_t b1 = rc_prmt(km[2], km[3], 0x6420u) & 0x01010101u;
            st_stream8(mrow + (int64_t)vi * V, make_uint2(b0, b1));
          } else {
            uint32_t km[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { km[i] = rc_prmt(d[i], 0u, 0xBBBBu); w[i] &= km[i]; }
            const uint32_t t01 = rc_prmt(km[0], km[1], 0x0040u), t23 = rc_prmt(km[2], km[3], 0x0040u);
            st_stream4(mrow + (int64_t)vi * V, rc_prmt(t01, t23, 0x5410u) & 0x01010101u);
          }
          if (zero_w) st_stream(wrow + (int64_t)vi * V, make_uint4(w[0], w[1], w[2], w[3]));
        }
      }
    } else {
      // general form: ties at the threshold split by column, k == 0 (nothing pruned), k >= C (everything pruned)
#pragma unroll 1
      for (int u = 0; u < NV; ++u) {
        if (RC_HAS(u)) {
          const int vi = tid + u * kRcThreads;
          uint32_t pm = 0;                                       // bit e: element e pruned
#pragma unroll
          for (int q = 0; q < V / 4; ++q) {
            const uint4 kq = *reinterpret_cast<const uint4*>(keys + ((size_t)q * nvec + vi) * 4);
            const uint32_t ke[4] = {kq.x, kq.y, kq.z, kq.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const bool pruned = k > 0 && (ke[e] < thr_key || (ke[e] == thr_key && vi * V + q * 4 + e <= thr_col));
              pm |= (pruned ? 1u : 0u) << (q * 4 + e);
            }
          }
          const uint32_t m0 = (~pm) & 0xfu, m1 = ((~pm) >> 4) & 0xfu;
          const uint32_t b0 = (m0 & 1u) | ((m0 & 2u) << 7) | ((m0 & 4u) << 14) | ((m0 & 8u) << 21);
          if (V == 8) {
            const uint32_t b1 = (m1 & 1u) | ((m1 & 2u) << 7) | ((m1 & 4u) << 14) | ((m1 & 8u) << 21);
            st_stream8(mrow + (int64_t)vi * V, make_uint2(b0, b1));
          } else {
            st_stream4(mrow + (int64_t)vi * V, b0);
          }
          // wv[u] is indexed dynamically here (no unroll): re-read the vector instead (L2 hit, rare path)
          if (zero_w && pm) st_stream(wrow + (int64_t)vi * V, rc_zero<T>(*reinterpret_cast<const uint4*>(wrow + (int64_t)vi * V), pm));
        }
      }
    }
    // next row's scale: this row's threshold in the middle of the bins
    if (select && thr_key > 0u && thr_key < 0x7f800000u) scale = (float)(kRcBins / 2) / __uint_as_float(thr_key);
  }
#undef RC_HAS
}

template <typename T>
static bool rowselect_cta_fits(int C) {
  constexpr int V = Elem<T>::kVec;
  const size_t smem = ((size_t)C + kRcHist + 2 * kRcCand) * sizeof(uint32_t);
  return C / V <= kRcThreads * kRcMaxVec && smem <= 200 * 1024 && C < (1 << 22);
}

template <typename T, int NV, bool FULL>
static int launch_rowselect_cta_nv(const RcBatch& bt, int C, int zero_w, cudaStream_t st) {
  auto kern = rowselect_cta_kernel<T, NV, FULL>;
  const size_t smem = ((size_t)C + kRcHist + 2 * kRcCand) * sizeof(uint32_t);
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)) != cudaSuccess)
      return check_launch();
    attr_set = true;
  }
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kRcThreads, smem);
  if (per_sm < 1) per_sm = 1;
  // every CTA gets the same number of rows (a ragged last wave would idle part of the grid)
  const int R = bt.row_begin[bt.count];
  const int resident = kNumSMs * per_sm;
  const int rows_per_cta = (R + resident - 1) / resident;
  int grid = (R + rows_per_cta - 1) / rows_per_cta;
  if (grid < 1) grid = 1;
  kern<<<grid, kRcThreads, smem, st>>>(bt, C, zero_w);
  return check_launch();
}

template <typename T>
static int launch_rowselect_cta(const RcBatch& bt, int C, int zero_w, cudaStream_t st) {
  constexpr int V = Elem<T>::kVec;
  const int nv = (C / V + kRcThreads - 1) / kRcThreads;
#define VLMC_RC(NV)                                                                                               \
  return (C / V == (NV) * kRcThreads) ? launch_rowselect_cta_nv<T, NV, true>(bt, C, zero_w, st)                   \
                                      : launch_rowselect_cta_nv<T, NV, false>(bt, C, zero_w, st)
  if (nv <= 1) VLMC_RC(1);
  if (nv <= 2) VLMC_RC(2);
  if (nv <= 4) VLMC_RC(4);
  if (nv <= 5) VLMC_RC(5);
  if (nv <= 8) VLMC_RC(8);
  if (nv <= 11) VLMC_RC(11);
  VLMC_RC(16);
#undef VLMC_RC
}

static bool rowselect_legacy() {
  const char* legacy = getenv("VLMC_ROWSELECT_LEGACY");        // A/B switch: the warp-per-row streaming kernel
  return legacy && legacy[0] == '1';
}

template <typename T>
static int launch_rowselect(void* W, int R, int C, int64_t ldw, const float* sq, int k, int zero_w,
                            uint8_t* mask, int64_t ldm, float* row_sum, uint32_t* seed, cudaStream_t st) {
  if (!rowselect_legacy() && rowselect_cta_fits<T>(C)) {
    RcBatch bt;
    bt.count = 1;
    bt.row_begin[0] = 0; bt.row_begin[1] = R;
    bt.it[0] = RcItem{W, ldw, sq, k, mask, ldm, row_sum};
    return launch_rowselect_cta<T>(bt, C, zero_w, st);
  }
  auto kern = rowselect_kernel<T>;
  rowselect_seed_kernel<T><<<1, 32, 0, st>>>(reinterpret_cast<const T*>(W), ldw, R, C, sq, k, seed);
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kRsWarps * 32, 0);
  if (per_sm < 1) per_sm = 1;
  // every warp gets the same number of rows (the last wave of a 1.7-rows-per-warp split would idle a quarter of the warps)
  const int resident_warps = kNumSMs * per_sm * kRsWarps;
  const int rows_per_warp = (R + resident_warps - 1) / resident_warps;
  int grid = (R + kRsWarps * rows_per_warp - 1) / (kRsWarps * rows_per_warp);
  if (grid < 1) grid = 1;
  kern<<<grid, kRsWarps * 32, 0, st>>>(reinterpret_cast<T*>(W), ldw, R, C, sq, k, zero_w, mask, ldm, row_sum, seed);
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
  if (ws_bytes < VLMC_WS_COUNTER_BYTES + ((size_t)R + (size_t)C + 8) * sizeof(float)) return VLMC_ERR_WORKSPACE;
  float* row_sum = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + VLMC_WS_COUNTER_BYTES);
  uint32_t* seed = reinterpret_cast<uint32_t*>(row_sum + R);
  float* sq = row_sum + R + 4;                 // sqrt(scaler_row), computed once per call instead of once per CTA
  if (((uintptr_t)sq & 15) != 0) sq += 4 - (((uintptr_t)sq >> 2) & 3);
  cudaStream_t st = (cudaStream_t)stream;
  sqrt_vec_kernel<<<(C + 255) / 256, 256, 0, st>>>(scaler_row, sq, C);
  VLMC_DISPATCH_DTYPE(dtype, rc = (launch_rowselect<scalar_t>(W, R, C, ldw, sq, k, zero_w, keep_mask, ldm, row_sum, seed, st)));
  if (rc) return rc;
  if (score_mean) return launch_mean_finalize(row_sum, R, (double)R * (double)C, score_mean, st);
  return VLMC_OK;
}

namespace vlmc {
// sqrt(scaler_row) of every item in one launch: grid (column chunks, items)
struct SqrtBatch { const float* in[kRcBatchMax]; float* out[kRcBatchMax]; int C[kRcBatchMax]; };
__global__ void __launch_bounds__(256) sqrt_vec_batch_kernel(const __grid_constant__ SqrtBatch b) {
  const int i = blockIdx.y, j = blockIdx.x * 256 + threadIdx.x;
  if (j < b.C[i]) b.out[i][j] = sqrtf(b.in[i][j]);
}
static size_t rowselect_item_ws_floats(int R, int C) { return (size_t)((R + 3) & ~3) + (size_t)((C + 3) & ~3) + 8; }
}  // namespace vlmc

extern "C" size_t vlmc_wanda_rowselect_batch_workspace_bytes(const vlmc_select_item* items, int count) {
  using namespace vlmc;
  if (!items || count < 1) return 0;
  size_t fl = 0;
  for (int i = 0; i < count; ++i) fl += rowselect_item_ws_floats(items[i].R, items[i].C);
  return VLMC_WS_COUNTER_BYTES + fl * sizeof(float);
}

extern "C" int vlmc_wanda_rowselect_batch(const vlmc_select_item* items, const int* k, int count, int dtype, int zero_w,
                                          void* ws, size_t ws_bytes, void* stream) {
  using namespace vlmc;
  if (!items || !k || count < 1 || count > kRcBatchMax) return VLMC_ERR_BAD_ARG;
  if (ws_bytes < vlmc_wanda_rowselect_batch_workspace_bytes(items, count)) return VLMC_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  float* base = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + VLMC_WS_COUNTER_BYTES);
  float* row_sum[kRcBatchMax];
  float* sq[kRcBatchMax];
  SqrtBatch sb;
  NmBatchOut o;
  int maxC = 0;
  bool any_mean = false, fits = !rowselect_legacy();
  o.begin[0] = 0;
  for (int i = 0; i < count; ++i) {
    const vlmc_select_item& s = items[i];
    int rc = sel_common_checks(s.W, dtype, s.R, s.C, s.ldw, s.scaler_row, s.keep_mask, s.ldm, ws);
    if (rc) return rc;
    if (k[i] < 0) return VLMC_ERR_BAD_ARG;
    if (s.score_mean && !is_device_ptr(s.score_mean)) return VLMC_ERR_NOT_DEVICE;
    row_sum[i] = base;                                   // the row sums of all items are contiguous: one finalize launch
    base += (s.R + 3) & ~3;
    o.begin[i + 1] = o.begin[i] + ((s.R + 3) & ~3);
    o.out[i] = s.score_mean;
    o.denom[i] = (double)s.R * (double)s.C;
    any_mean |= s.score_mean != nullptr;
    maxC = s.C > maxC ? s.C : maxC;
    VLMC_DISPATCH_DTYPE(dtype, fits = fits && rowselect_cta_fits<scalar_t>(s.C));
  }
  for (int i = 0; i < count; ++i) {
    sq[i] = base;
    base += ((items[i].C + 3) & ~3) + 8;
    sb.in[i] = items[i].scaler_row; sb.out[i] = sq[i]; sb.C[i] = items[i].C;
  }
  if (!fits) {                                           // rows too long for shared memory / the legacy switch: one by one
    for (int i = 0; i < count; ++i) {
      const vlmc_select_item& s = items[i];
      int rc = vlmc_wanda_rowselect(s.W, dtype, s.R, s.C, s.ldw, s.scaler_row, k[i], zero_w, s.keep_mask, s.ldm, s.score_mean,
                                    ws, ws_bytes, stream);
      if (rc) return rc;
    }
    return VLMC_OK;
  }
  if (any_mean) {                                        // padding rows of the contiguous row-sum array must read as zero
    if (cudaMemsetAsync(row_sum[0], 0, (size_t)o.begin[count] * sizeof(float), st) != cudaSuccess) return check_launch();
  }
  sqrt_vec_batch_kernel<<<dim3((maxC + 255) / 256, count), 256, 0, st>>>(sb);
  // one launch per distinct row length (a Vicuna block: C = 4096 x 6, C = 11008 x 1)
  bool done[kRcBatchMax] = {};
  for (int i = 0; i < count; ++i) {
    if (done[i]) continue;
    RcBatch bt;
    bt.count = 0;
    bt.row_begin[0] = 0;
    for (int j = i; j < count; ++j) {
      if (done[j] || items[j].C != items[i].C) continue;
      const vlmc_select_item& s = items[j];
      if ((int64_t)bt.row_begin[bt.count] + s.R > 0x7fffffff) return VLMC_ERR_UNSUPPORTED;
      bt.it[bt.count] = RcItem{s.W, s.ldw, sq[j], k[j], s.keep_mask, s.ldm, row_sum[j]};
      bt.row_begin[bt.count + 1] = bt.row_begin[bt.count] + s.R;
      ++bt.count;
      done[j] = true;
    }
    int rc;
    VLMC_DISPATCH_DTYPE(dtype, rc = (launch_rowselect_cta<scalar_t>(bt, items[i].C, zero_w, st)));
    if (rc) return rc;
  }
  if (any_mean) {
    mean_finalize_batch_kernel<<<count, 256, 0, st>>>(row_sum[0], o);
    return check_launch();
  }
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

extern "C" size_t vlmc_wanda_nm_batch_workspace_bytes(const vlmc_select_item* items, int count, int dtype, int m) {
  using namespace vlmc;
  if (!items || count < 1) return 0;
  const int V = dtype == VLMC_F32 ? 4 : 8;
  const int E = m > V ? m : V;
  size_t units = 0;
  for (int i = 0; i < count; ++i) {
    const size_t coltiles = ((size_t)items[i].C / E + kSelThreads - 1) / kSelThreads;
    units += coltiles * (((size_t)items[i].R + kNmbRows - 1) / kNmbRows);
  }
  return VLMC_WS_COUNTER_BYTES + units * kSelWarps * sizeof(float);
}

extern "C" int vlmc_wanda_nm_batch(const vlmc_select_item* items, int count, int dtype, int n, int m, int zero_w,
                                   void* ws, size_t ws_bytes, void* stream) {
  using namespace vlmc;
  if (!items || count < 1 || count > kNmbMax) return VLMC_ERR_BAD_ARG;
  if (!(m == 2 || m == 4 || m == 8 || m == 16)) return VLMC_ERR_UNSUPPORTED;
  if (n <= 0 || n >= m) return VLMC_ERR_BAD_ARG;
  const int V = dtype == VLMC_F32 ? 4 : 8;
  const int E = m > V ? m : V;
  NmBatch b;
  NmBatchOut o;
  b.count = count;
  b.unit_begin[0] = 0;
  o.begin[0] = 0;
  bool any_mean = false;
  for (int i = 0; i < count; ++i) {
    const vlmc_select_item& s = items[i];
    int rc = sel_common_checks(s.W, dtype, s.R, s.C, s.ldw, s.scaler_row, s.keep_mask, s.ldm, ws);
    if (rc) return rc;
    if (s.C % m != 0 || s.C % E != 0 || ((uintptr_t)s.scaler_row & 15) != 0) return VLMC_ERR_UNSUPPORTED;
    NmBatchItem& it = b.it[i];
    it.W = s.W; it.ldw = s.ldw; it.R = s.R; it.C = s.C; it.scaler_row = s.scaler_row; it.mask = s.keep_mask; it.ldm = s.ldm;
    it.coltiles = (s.C / E + kSelThreads - 1) / kSelThreads;
    const int64_t units = (int64_t)it.coltiles * ((s.R + kNmbRows - 1) / kNmbRows);
    if (b.unit_begin[i] + units > 0x7fffffff / kSelWarps) return VLMC_ERR_UNSUPPORTED;
    b.unit_begin[i + 1] = b.unit_begin[i] + (int)units;
    o.begin[i + 1] = b.unit_begin[i + 1] * kSelWarps;
    o.out[i] = s.score_mean;
    o.denom[i] = (double)s.R * (double)s.C;
    if (s.score_mean) {
      if (!is_device_ptr(s.score_mean)) return VLMC_ERR_NOT_DEVICE;
      any_mean = true;
    }
  }
  const int total = b.unit_begin[count];
  if (ws_bytes < VLMC_WS_COUNTER_BYTES + (size_t)total * kSelWarps * sizeof(float)) return VLMC_ERR_WORKSPACE;
  float* part = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + VLMC_WS_COUNTER_BYTES);
  cudaStream_t st = (cudaStream_t)stream;
  // every CTA resident from the start: the fixed-stride walk then ends within one unit everywhere
#define VLMC_NMB(MM)                                                                                         \
  VLMC_DISPATCH_DTYPE(dtype, {                                                                               \
    int per_sm = 1;                                                                                          \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nm_batch_kernel<scalar_t, MM>, kSelThreads, 0);   \
    int grid = kNumSMs * (per_sm < 1 ? 1 : per_sm);                                                          \
    if (grid > total) grid = total;                                                                          \
    nm_batch_kernel<scalar_t, MM><<<grid, kSelThreads, 0, st>>>(b, n, zero_w, part);                         \
  })
  switch (m) {
    case 2: VLMC_NMB(2); break;
    case 4: VLMC_NMB(4); break;
    case 8: VLMC_NMB(8); break;
    default: VLMC_NMB(16); break;
  }
#undef VLMC_NMB
  int rc = check_launch();
  if (rc) return rc;
  if (any_mean) {
    mean_finalize_batch_kernel<<<count, 256, 0, st>>>(part, o);
    return check_launch();
  }
  return VLMC_OK;
}
