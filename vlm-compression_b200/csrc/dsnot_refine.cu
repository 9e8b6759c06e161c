// K8 + K9: DSnoT mask refinement (training-free prune / regrow swaps) for one linear layer.
//
// Replaces lavis/compression/pruners/dsnot_pruner.py:359-755 (and its ViT copy :1092-1485):
//   DSnoT_metric = W * sum_metric_row ; initial_metric = |W| * sqrt(scaler_row) (or |W|)            (:368-376)
//   unstructured  three stable row sorts + return_reorder_indice (:555-612, :1881-1925), then up to
//                 max_cycle_time cycles of ~25 tiny launches and one host sync each                  (:650-751)
//   n:m           per-group initial mask, one stable row sort, same loop with in-group pruning       (:407-552)
//
// What the loop can touch is tiny: every cycle advances ONE pointer of the regrow ordering and ONE pointer of the
// prune ordering by one step, so after max_cycle_time (<= 128) cycles a row has only looked at the first / last
// <= 128 entries of each ordering.  No row is ever sorted here: one CTA owns a row, keeps two 32-bit keys per
// column in shared memory and extracts just those ordered prefixes with a 3-pass (11/11/10 bit) shared-memory
// radix select on (key, column) - the stable-sort order - followed by a rank sort of the <= 128 survivors.
//   pass 1  dsnot_walk_kernel   per row: initial selection (k-th smallest (score, column) pair), the four / two
//                               ordered prefixes, the cycle loop run by one thread out of shared memory.  Writes
//                               the (pruned column, regrown column) pair of every cycle, the cycle at which the
//                               row stopped updating, and atomically maxes the number of cycles the reference's
//                               `while any(update_mask)` loop would have executed (rows are coupled through it).
//   pass 2  dsnot_apply_kernel  per row: rebuilds the initial mask, replays the recorded writes of the executed
//                               cycles in order (with or without the reference's write-back block :734-740,
//                               SURVEY F4), writes the bool mask and the zeroed weights.
// Latency-bound (one thread walks <= 128 dependent steps per row), not HBM-bound: 2 B/weight read in pass 1,
// 5 B/weight in pass 2.
#include <stdlib.h>
#include <type_traits>
#include "common.cuh"

namespace vlmc {

constexpr int kDsThreads = 256;
constexpr int kDsBins = 2048;
constexpr int kDsCap = 128;          // longest ordered prefix = largest supported max_cycle_time
constexpr int kDsMaxM = 8;           // widest n:m group in the DSnoT branch (scripts use 2:4 and 4:8)
constexpr uint32_t kPrunedKey = 0xffffffffu;
constexpr uint32_t kInfBits = 0x7f800000u;

enum { L_HEAD = 0, L_TAIL = 1, L_NEG = 2, L_POS = 3, L_NEGFAR = 4, L_POSFAR = 5, L_COUNT = 6 };

struct DsParams {
  const void* W; int64_t ldw; int R, C;
  const float* scaler_row; const float* sum_row; const float* var;
  int k, prune_n, prune_m;
  float pow_var; int max_cycle; float thr; int without_same_sign, initial_magnitude, argmin_rule;
  uint32_t* row_v; int* row_iv; int* row_stop; int* walk; int* ncycles;
  const float* sq;                       // sqrt(scaler_row) (fast kernel)
  const int* rows; const int* nrows;     // optional: walk only rows[0 .. *nrows) (the rows the fast kernel handed back)
  int* fb_rows; int* fb_count;           // fast kernel: rows it could not take
};

constexpr int kDsSkipTail = 1 << 30;     // row_stop flag: nothing after the stop cycle is observable, the walk record ends there

struct DsShared {
  uint32_t hist[kDsBins];
  uint32_t warp_tot[32];
  double dred[32];
  uint32_t cand_key[kDsCap];
  int cand_idx[kDsCap];
  int cand_cnt;
  int sel_bin; uint32_t sel_before, sel_count, total;
  int list[2][kDsCap];               // regrow ordering: head / tail prefixes (columns)
  float listD[2][kDsCap];
  union {
    struct { int list[4][kDsCap]; float listD[4][kDsCap]; } u;   // unstructured: neg / pos / negfar / posfar
    float dgrp[2][kDsCap][kDsMaxM];                               // n:m: DSnoT metric of the regrow candidate's group
  } x;
  int nlist[L_COUNT];
  int walk[2 * kDsCap];
  int min_kept;
};

template <typename T> __device__ __forceinline__ float ds_to_float(T v);
template <> __device__ __forceinline__ float ds_to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float ds_to_float<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float ds_to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// order-preserving float -> uint32 with torch.sort semantics: -0 == +0, every NaN equal and greater than +inf
__device__ __forceinline__ uint32_t sortable(float f) {
  if (f != f) return 0xffffffffu;
  if (f == 0.f) return 0x80000000u;
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float sgnf(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }

template <class SH>
__device__ __forceinline__ uint32_t block_sum(uint32_t v, SH& sh) {
  v = __reduce_add_sync(0xffffffffu, v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh.warp_tot[threadIdx.x >> 5] = v;
  __syncthreads();
  uint32_t t = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh.warp_tot[w];
  return t;
}
template <class SH>
__device__ __forceinline__ double block_sum(double v, SH& sh) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh.dred[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh.dred[w];
  return t;
}

// k-th smallest key (1-indexed) among the included items: v, #items with key < v, #items with key == v.
// Returns false (only `total` valid) when fewer than k items are included.
template <class F>
__device__ bool radix_kth(F f, int n, uint32_t k, DsShared& sh, uint32_t& v, uint32_t& cnt_less, uint32_t& cnt_eq,
                          uint32_t& total) {
  const int nthreads = blockDim.x, tid = threadIdx.x;
  const int per = kDsBins / nthreads;
  uint32_t prefix = 0, pmask = 0, kk = k;
#pragma unroll 1
  for (int pass = 0; pass < 3; ++pass) {
    const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
    const uint32_t bmask = pass == 2 ? 0x3ffu : 0x7ffu;
    __syncthreads();
    for (int b = tid; b < kDsBins; b += nthreads) sh.hist[b] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += nthreads) {
      uint32_t key;
      if (f(i, key) && (key & pmask) == prefix) atomicAdd(&sh.hist[(key >> shift) & bmask], 1u);
    }
    __syncthreads();
    uint32_t local = 0;
    for (int j = 0; j < per; ++j) local += sh.hist[tid * per + j];
    uint32_t incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((tid & 31) >= o) incl += t;
    }
    if ((tid & 31) == 31) sh.warp_tot[tid >> 5] = incl;
    __syncthreads();
    uint32_t wbase = 0, tot = 0;
    for (int w = 0; w < (nthreads >> 5); ++w) {
      const uint32_t t = sh.warp_tot[w];
      if (w < (tid >> 5)) wbase += t;
      tot += t;
    }
    incl += wbase;
    const uint32_t excl = incl - local;
    if (pass == 0) {
      total = tot;
      if (tot < k) return false;
    }
    if (excl < kk && kk <= incl) {
      uint32_t run = excl;
      for (int j = 0; j < per; ++j) {
        const uint32_t c = sh.hist[tid * per + j];
        if (kk <= run + c) { sh.sel_bin = tid * per + j; sh.sel_before = run; sh.sel_count = c; break; }
        run += c;
      }
    }
    __syncthreads();
    kk -= sh.sel_before;
    prefix |= (uint32_t)sh.sel_bin << shift;
    pmask |= bmask << shift;
    cnt_eq = sh.sel_count;
  }
  v = prefix;
  cnt_less = k - kk;
  return true;
}

// smallest q with #{included i <= q : key == v} >= need   (stable order among ties: lowest index first)
template <class F, class SH>
__device__ int tie_bound(F f, int n, uint32_t v, uint32_t need, SH& sh) {
  int lo = -1, hi = n - 1;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    uint32_t c = 0;
    for (int i = threadIdx.x; i <= mid; i += blockDim.x) {
      uint32_t key;
      if (f(i, key) && key == v) ++c;
    }
    c = block_sum(c, sh);
    if (c >= need) hi = mid; else lo = mid;
  }
  return hi;
}

// The `want` smallest included items in (key, index) order -> out[0..return); `total` = #included.
template <class F>
__device__ int build_list(F f, int n, int want, DsShared& sh, int* out, uint32_t& total) {
  uint32_t v = 0, cnt_less = 0, cnt_eq = 0;
  const bool found = radix_kth(f, n, (uint32_t)want, sh, v, cnt_less, cnt_eq, total);
  const int take = found ? want : (int)total;
  if (take == 0) return 0;
  int iv = 0x7fffffff;
  if (found) {
    const uint32_t need = (uint32_t)want - cnt_less;
    if (need < cnt_eq) iv = tie_bound(f, n, v, need, sh);
  }
  __syncthreads();
  if (threadIdx.x == 0) sh.cand_cnt = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    uint32_t key;
    if (f(i, key) && (!found || key < v || (key == v && i <= iv))) {
      const int slot = atomicAdd(&sh.cand_cnt, 1);
      if (slot < kDsCap) { sh.cand_key[slot] = key; sh.cand_idx[slot] = i; }
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < take; t += blockDim.x) {
    const uint32_t kt = sh.cand_key[t];
    const int it = sh.cand_idx[t];
    int rank = 0;
    for (int j = 0; j < take; ++j) {
      const uint32_t kj = sh.cand_key[j];
      rank += (kj < kt || (kj == kt && sh.cand_idx[j] < it)) ? 1 : 0;
    }
    out[rank] = it;
  }
  __syncthreads();
  return take;
}

// torch.topk(group, 1, largest=False) (dsnot_pruner.py:517).  rule 0: lowest index among tied minima.
// rule 1: what ATen's CPU kernel returns - std::nth_element(begin, begin, end) on (value, index) pairs, i.e.
// libstdc++ introselect (median-of-3 to front, unguarded partition, insertion sort below 4) - see oracle.torch_cpu_argmin.
__device__ int group_argmin(const uint32_t* keys, int m, int rule) {
  int best = 0, ties = 1;
  for (int j = 1; j < m; ++j) {
    if (keys[j] < keys[best]) { best = j; ties = 1; }
    else if (keys[j] == keys[best]) ++ties;
  }
  if (rule == 0 || ties == 1) return best;
  uint32_t v[kDsMaxM]; int id[kDsMaxM];
  for (int j = 0; j < m; ++j) { v[j] = keys[j]; id[j] = j; }
  auto swp = [&](int a, int b) { const uint32_t tv = v[a]; v[a] = v[b]; v[b] = tv; const int ti = id[a]; id[a] = id[b]; id[b] = ti; };
  int first = 0, last = m;
  while (last - first > 3) {
    const int mid = first + (last - first) / 2;
    const int a = first + 1, b = mid, c = last - 1;
    int pick;
    if (v[a] < v[b]) pick = (v[b] < v[c]) ? b : ((v[a] < v[c]) ? c : a);
    else pick = (v[a] < v[c]) ? a : ((v[b] < v[c]) ? c : b);
    swp(first, pick);
    int lo = first + 1, hi = last;
    while (true) {
      while (v[lo] < v[first]) ++lo;
      --hi;
      while (v[first] < v[hi]) --hi;
      if (!(lo < hi)) break;
      swp(lo, hi);
      ++lo;
    }
    if (lo <= 0) first = lo; else last = lo;
  }
  for (int i = first + 1; i < last; ++i)
    for (int j = i; j > first && v[j] < v[j - 1]; --j) swp(j, j - 1);
  return id[0];
}

template <typename T>
__device__ __forceinline__ float dsnot_metric(const T* wrow, const float* sum_row, int c) {
  return __fmul_rn(ds_to_float<T>(wrow[c]), sum_row[c]);     // :368, one fp32 multiply
}

template <typename T>
__global__ void __launch_bounds__(kDsThreads)
dsnot_walk_kernel(const DsParams p) {
  extern __shared__ __align__(16) uint32_t ds_dyn[];
  uint32_t* keyA = ds_dyn;                                   // [C] initial / wanda score bits, pruned marker
  float* dm = reinterpret_cast<float*>(ds_dyn + p.C);        // [C] DSnoT metric, later the regrow sort key
  uint32_t* keyR = ds_dyn + p.C;
  __shared__ DsShared sh;
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int C = p.C, maxc = p.max_cycle;
  const bool nm = p.prune_n != 0;

  const int nrows = p.rows ? *p.nrows : p.R;
  for (int ri = blockIdx.x; ri < nrows; ri += gridDim.x) {
    const int row = p.rows ? p.rows[ri] : ri;
    const T* wrow = reinterpret_cast<const T*>(p.W) + (int64_t)row * p.ldw;
    __syncthreads();
    for (int c = tid; c < C; c += nthreads) {
      const float w = ds_to_float<T>(wrow[c]);
      const float aw = fabsf(w);
      const float init = p.initial_magnitude ? aw : __fmul_rn(aw, __fsqrt_rn(p.scaler_row[c]));
      keyA[c] = __float_as_uint(init);
      dm[c] = __fmul_rn(w, p.sum_row[c]);
    }
    __syncthreads();

    double esum = 0.0;
    uint32_t v = 0; int iv = -1;
    if (!nm) {
      // ---- initial selection: the k smallest (score, column) pairs are pruned (:555-581) ----
      if (p.k > 0) {
        uint32_t cl, ce, tot;
        auto all = [&](int i, uint32_t& key) { key = keyA[i]; return true; };
        radix_kth(all, C, (uint32_t)p.k, sh, v, cl, ce, tot);
        const uint32_t need = (uint32_t)p.k - cl;
        iv = need < ce ? tie_bound([&](int i, uint32_t& key) { key = keyA[i]; return true; }, C, v, need, sh) : C;
      }
      __syncthreads();
      for (int c = tid; c < C; c += nthreads) {
        const uint32_t key = keyA[c];
        if (key < v || (key == v && c <= iv)) { esum += (double)dm[c]; keyA[c] = kPrunedKey; }
        else if (p.initial_magnitude)   // the prune ordering always uses the Wanda metric (:583)
          keyA[c] = __float_as_uint(__fmul_rn(fabsf(ds_to_float<T>(wrow[c])), __fsqrt_rn(p.scaler_row[c])));
      }
    } else {
      // ---- n:m initial mask: the n smallest of every m consecutive columns, ties -> lower column (:420-438) ----
      const int m = p.prune_m;
      for (int g = tid; g < C / m; g += nthreads) {
        uint32_t kk[kDsMaxM];
        for (int a = 0; a < m; ++a) kk[a] = keyA[g * m + a];
        for (int a = 0; a < m; ++a) {
          int rank = 0;
          for (int b = 0; b < m; ++b) rank += (b < a) ? (kk[b] <= kk[a] ? 1 : 0) : ((b > a && kk[b] < kk[a]) ? 1 : 0);
          if (rank < p.prune_n) { esum += (double)dm[g * m + a]; keyA[g * m + a] = kInfBits; }   // :469
        }
      }
    }
    const float err0 = (float)block_sum(esum, sh);           // :601 / :444
    __syncthreads();

    // ---- prune ordering (unstructured): kept columns by (wanda, column), split by the sign of the DSnoT metric
    //      (return_reorder_indice :1881-1925: negatives first in order, positives reversed at the end) ----
    uint32_t nneg = 0, npos = 0;
    if (!nm) {
      auto fneg = [&](int i, uint32_t& key) { key = keyA[i]; return key != kPrunedKey && dm[i] < 0.f; };
      auto fpos = [&](int i, uint32_t& key) { key = keyA[i]; return key != kPrunedKey && dm[i] > 0.f; };
      const int a = build_list(fneg, C, maxc, sh, sh.x.u.list[L_NEG - 2], nneg);
      const int b = build_list(fpos, C, maxc, sh, sh.x.u.list[L_POS - 2], npos);
      int na = 0, nb = 0;
      if ((int)nneg < maxc || (int)npos < maxc) {
        // a pointer can leave its sign class: needs the far ends and the filler column wanda_res_indices[0]
        auto fnegfar = [&](int i, uint32_t& key) { const int c = C - 1 - i; key = ~keyA[c]; return keyA[c] != kPrunedKey && dm[c] < 0.f; };
        auto fposfar = [&](int i, uint32_t& key) { const int c = C - 1 - i; key = ~keyA[c]; return keyA[c] != kPrunedKey && dm[c] > 0.f; };
        auto fkept = [&](int i, uint32_t& key) { key = keyA[i]; return key != kPrunedKey; };
        uint32_t t0;
        na = build_list(fnegfar, C, maxc, sh, sh.x.u.list[L_NEGFAR - 2], t0);
        nb = build_list(fposfar, C, maxc, sh, sh.x.u.list[L_POSFAR - 2], t0);
        for (int i = tid; i < na; i += nthreads) sh.x.u.list[L_NEGFAR - 2][i] = C - 1 - sh.x.u.list[L_NEGFAR - 2][i];
        for (int i = tid; i < nb; i += nthreads) sh.x.u.list[L_POSFAR - 2][i] = C - 1 - sh.x.u.list[L_POSFAR - 2][i];
        // build_list writes through a shared pointer: reuse walk[] as the 1-entry output
        build_list(fkept, C, 1, sh, sh.walk, t0);
        if (tid == 0) sh.min_kept = sh.walk[0];
      }
      if (tid == 0) { sh.nlist[L_NEG] = a; sh.nlist[L_POS] = b; sh.nlist[L_NEGFAR] = na; sh.nlist[L_POSFAR] = nb; }
      __syncthreads();
    }

    // ---- regrow ordering: (pruned ? DSnoT metric : 0) / var^pow, stable ascending (:599-612 / :440-453) ----
    for (int c = tid; c < C; c += nthreads) {
      const bool pruned = nm ? (keyA[c] == kInfBits) : (keyA[c] == kPrunedKey);
      float mval = pruned ? dm[c] : 0.f;
      if (p.pow_var != 0.f) mval = __fdiv_rn(mval, p.pow_var == 1.f ? p.var[c] : powf(p.var[c], p.pow_var));
      keyR[c] = sortable(mval);
    }
    __syncthreads();
    {
      uint32_t t0;
      auto fhead = [&](int i, uint32_t& key) { key = keyR[i]; return true; };
      auto ftail = [&](int i, uint32_t& key) { key = ~keyR[C - 1 - i]; return true; };
      const int a = build_list(fhead, C, maxc, sh, sh.list[L_HEAD], t0);
      const int b = build_list(ftail, C, maxc, sh, sh.list[L_TAIL], t0);
      for (int i = tid; i < b; i += nthreads) sh.list[L_TAIL][i] = C - 1 - sh.list[L_TAIL][i];
      if (tid == 0) { sh.nlist[L_HEAD] = a; sh.nlist[L_TAIL] = b; }
      __syncthreads();
    }
    // DSnoT metric of every candidate (and of its whole group for n:m)
    for (int e = tid; e < 2 * kDsCap; e += nthreads) {
      const int l = e / kDsCap, i = e % kDsCap;
      if (i < sh.nlist[l]) {
        const int c = sh.list[l][i];
        sh.listD[l][i] = dsnot_metric<T>(wrow, p.sum_row, c);
        if (nm) {
          const int g = c - c % p.prune_m;
          for (int j = 0; j < p.prune_m; ++j) sh.x.dgrp[l][i][j] = dsnot_metric<T>(wrow, p.sum_row, g + j);
        }
      }
    }
    if (!nm) {
      for (int e = tid; e < 4 * kDsCap; e += nthreads) {
        const int l = e / kDsCap, i = e % kDsCap;
        if (i < sh.nlist[2 + l]) sh.x.u.listD[l][i] = dsnot_metric<T>(wrow, p.sum_row, sh.x.u.list[l][i]);
      }
    }
    __syncthreads();

    // ---- the cycle loop, one thread, everything in shared memory ----
    if (tid == 0) {
      float err = err0;
      const float sign0 = sgnf(err0);
      bool upd = true;
      int stop = 0;
      int hR = 0, tR = 0, hP = 0, tP = 0;
      const int Ck = C - p.k;
      for (int c = 1; c <= maxc; ++c) {
        const int l = err > 0.f ? L_TAIL : L_HEAD;            // :654 / :483
        const int e = l == L_TAIL ? tR++ : hR++;
        const int rg = sh.list[l][e];
        int pr; float rm, pm;
        if (!nm) {
          rm = sh.listD[l][e];
          const int j = err < 0.f ? (Ck - 1 - tP++) : hP++;    // :683 position in pruning_indices_block
          if (j < (int)nneg) {
            if (j < sh.nlist[L_NEG]) { pr = sh.x.u.list[L_NEG - 2][j]; pm = sh.x.u.listD[L_NEG - 2][j]; }
            else { const int q = (int)nneg - 1 - j; pr = sh.x.u.list[L_NEGFAR - 2][q]; pm = sh.x.u.listD[L_NEGFAR - 2][q]; }
          } else if (j >= Ck - (int)npos) {
            const int t = Ck - 1 - j;
            if (t < sh.nlist[L_POS]) { pr = sh.x.u.list[L_POS - 2][t]; pm = sh.x.u.listD[L_POS - 2][t]; }
            else { const int q = (int)npos - 1 - t; pr = sh.x.u.list[L_POSFAR - 2][q]; pm = sh.x.u.listD[L_POSFAR - 2][q]; }
          } else {
            pr = sh.min_kept;                                  // reorder filler 0 -> wanda_res_indices[0]
            pm = dsnot_metric<T>(wrow, p.sum_row, pr);
          }
        } else {
          const int m = p.prune_m;
          const int g = rg - rg % m;                           // :502
          const int a = group_argmin(keyA + g, m, p.argmin_rule);   // :513-521
          pr = g + a;
          rm = sh.x.dgrp[l][e][rg - g];
          pm = sh.x.dgrp[l][e][a];
          keyA[pr] = kInfBits;                                 // :529 (unconditional)
        }
        const float after = __fsub_rn(__fadd_rn(err, pm), rm); // :713 / :525
        const bool big = fabsf(err) > p.thr;
        if (nm) upd = upd && (sign0 == sgnf(after)) && big;    // :527
        else if (p.without_same_sign) upd = upd && big;        // :717-720
        else upd = upd && big && (sign0 == sgnf(after));       // :722-729
        sh.walk[2 * (c - 1)] = pr;
        sh.walk[2 * (c - 1) + 1] = rg;
        if (!upd && stop == 0) stop = c;
        if (upd) { err = __fadd_rn(err, pm); err = __fsub_rn(err, rm); }   // :742-751 / :534-543
      }
      p.row_v[row] = v;
      p.row_iv[row] = iv;
      p.row_stop[row] = stop;
      atomicMax(p.ncycles, stop == 0 ? maxc : stop);
    }
    __syncthreads();
    for (int i = tid; i < 2 * maxc; i += nthreads) p.walk[(int64_t)row * 2 * maxc + i] = sh.walk[i];
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// dsnot_walk2_kernel: the unstructured walk with 5 passes over the row instead of ~17.
//
// The old kernel runs one 3-pass radix select + collect + rank sort over ALL C columns for each of the four orderings
// (115 k warp instructions per 4096-column row, ncu).  What the loop can read is tiny and lies at known places:
//   prune orderings   the smallest kept scores, i.e. the ranks k+1 .. k+N of the SAME score order the initial selection
//                     uses: one more 2-level select gives a cut-off with ~3 max_cycle kept columns below it
//   regrow orderings  the two ends of (pruned ? metric / var : 0): the first histogram of that key tells which bins hold
//                     the max_cycle smallest / largest
// so one classification pass appends those few hundred columns to candidate lists, and FOUR WARPS finish the four lists
// on their own (no CTA barrier): a 256-bin linear re-binning of the candidates' narrow key range picks whole bins up to
// max_cycle entries (<= 128), a register bitonic sort on (key, column) pairs puts them in the stable-sort order.
// Rows where any of that does not fit (a sign class with fewer than max_cycle candidates: the pointer would leave its
// class and need the far lists / filler; heavy ties; more candidates than a list holds) are handed back and go
// through dsnot_walk_kernel - same results either way (tests run both on the same inputs).
// The cycle loop stops at the row's stop cycle: with both sign classes >= max_cycle the pruned / regrown columns of
// different cycles are distinct, so the writes of a non-updating cycle (kept stays kept, pruned stays pruned) are
// no-ops and the apply pass may skip them (kDsSkipTail).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kW2Kept = 1024;        // candidate capacity: kept columns just above the threshold
constexpr int kW2End = 512;          // candidate capacity: each end of the regrow ordering
constexpr int kW2Bins = 256;         // per-warp linear re-binning

struct DsShared2 {
  uint32_t hist[kDsBins];            // block histograms; later the four per-warp 256-bin ones
  uint32_t warp_tot[32];
  double dred[32];
  unsigned long long ckey[4][kDsCap];
  int lists[4][kDsCap];              // L_NEG, L_POS (prune), then head, tail (regrow): columns
  float listsD[4][kDsCap];
  int walk[2 * kDsCap];
  uint16_t ckept[kW2Kept];           // column | 0x8000 when the DSnoT metric is negative
  uint16_t chead[kW2End], ctail[kW2End];
  uint32_t nkept, nhead, ntail;
  uint32_t t_bin[2], t_before[2], t_count[2];
  uint32_t bmin, bmax;
  int fail;
};

// bins of sh.hist (2048 counters, 8 per thread) that hold the ranks k0 (and k1 when ntargets == 2), 1-indexed
__device__ void ds2_scan(DsShared2& sh, uint32_t k0, uint32_t k1, int ntargets, bool extent = false) {
  const int tid = threadIdx.x;
  uint32_t c[8], local = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { c[j] = sh.hist[tid * 8 + j]; local += c[j]; }
  if (extent && local) {                                       // first / last non-empty bin -> sh.bmin / sh.bmax
    int first = 7, last = 0;
#pragma unroll
    for (int j = 7; j >= 0; --j) if (c[j]) first = j;
#pragma unroll
    for (int j = 0; j < 8; ++j) if (c[j]) last = j;
    atomicMin(&sh.bmin, (uint32_t)(tid * 8 + first));
    atomicMax(&sh.bmax, (uint32_t)(tid * 8 + last));
  }
  uint32_t incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((tid & 31) >= o) incl += t;
  }
  if ((tid & 31) == 31) sh.warp_tot[tid >> 5] = incl;
  __syncthreads();
  uint32_t wbase = 0;
  for (int w = 0; w < (tid >> 5); ++w) wbase += sh.warp_tot[w];
  incl += wbase;
  const uint32_t excl = incl - local;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const uint32_t kk = t == 0 ? k0 : k1;
    if (t < ntargets && excl < kk && kk <= incl) {
      uint32_t run = excl;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (kk <= run + c[j]) { sh.t_bin[t] = tid * 8 + j; sh.t_before[t] = run; sh.t_count[t] = c[j]; break; }
        run += c[j];
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void ds2_cmpx(unsigned long long& a, unsigned long long& b, bool asc) {
  if ((a > b) == asc) { const unsigned long long t = a; a = b; b = t; }
}

// ascending bitonic sort of 128 keys held four per lane (element index = 4 * lane + j)
__device__ __forceinline__ void ds2_sort128(unsigned long long (&v)[4], int lane) {
#pragma unroll
  for (int k = 2; k <= 128; k <<= 1) {
#pragma unroll
    for (int d = k >> 1; d >= 1; d >>= 1) {
      if (d >= 4) {
        const int ld = d >> 2;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool asc = ((4 * lane + j) & k) == 0;
          const unsigned long long other = __shfl_xor_sync(0xffffffffu, v[j], ld);
          const bool keep_min = ((lane & ld) == 0) == asc;
          v[j] = keep_min ? (v[j] < other ? v[j] : other) : (v[j] > other ? v[j] : other);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if ((j & d) == 0) ds2_cmpx(v[j], v[j | d], ((4 * lane + j) & k) == 0);
      }
    }
  }
}

// One warp: the `want` smallest candidates in (key, ord) order -> out[0 .. want).  f(entry, key, ord) -> included.
// Keys of included entries lie in [lo, lo + (256 << shift)).  Returns false when the list cannot be built here.
template <class F>
__device__ bool ds2_warp_list(F f, const uint16_t* cand, int n, uint32_t lo, int shift, int want, uint32_t* whist,
                              unsigned long long* ckey, int lane) {
  for (int b = lane; b < kW2Bins; b += 32) whist[b] = 0;
  __syncwarp();
  for (int i = lane; i < n; i += 32) {
    uint32_t key, ord;
    if (f(cand[i], key, ord)) {
      uint32_t bin = (key - lo) >> shift;
      if (bin > (uint32_t)(kW2Bins - 1)) bin = kW2Bins - 1;
      atomicAdd(&whist[bin], 1u);
    }
  }
  __syncwarp();
  uint32_t c[8], local = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { c[j] = whist[lane * 8 + j]; local += c[j]; }
  uint32_t incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  if (total < (uint32_t)want) return false;
  const uint32_t excl = incl - local;
  int tb = 0;
  if (excl < (uint32_t)want && (uint32_t)want <= incl) {
    uint32_t run = excl;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if ((uint32_t)want <= run + c[j]) { tb = lane * 8 + j; break; }
      run += c[j];
    }
  }
  const uint32_t owner = __ballot_sync(0xffffffffu, excl < (uint32_t)want && (uint32_t)want <= incl);
  tb = __shfl_sync(0xffffffffu, tb, __ffs(owner) - 1);
  int cnt = 0;
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    uint32_t key = 0, ord = 0;
    bool ok = false;
    if (i < n && f(cand[i], key, ord)) {
      uint32_t bin = (key - lo) >> shift;
      if (bin > (uint32_t)(kW2Bins - 1)) bin = kW2Bins - 1;
      ok = bin <= (uint32_t)tb;
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, ok);
    const int slot = cnt + __popc(bal & ((1u << lane) - 1u));
    if (ok && slot < kDsCap) ckey[slot] = ((unsigned long long)key << 32) | ord;
    cnt += __popc(bal);
  }
  if (cnt > kDsCap) return false;
  __syncwarp();
  unsigned long long v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = (4 * lane + j < cnt) ? ckey[4 * lane + j] : ~0ull;
  ds2_sort128(v, lane);
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; ++j) ckey[4 * lane + j] = v[j];
  __syncwarp();
  return true;
}

__global__ void ds_sqrt_kernel(const float* __restrict__ s, float* __restrict__ out, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) out[i] = __fsqrt_rn(s[i]);
}

template <typename T>
__global__ void __launch_bounds__(kDsThreads, 4)
dsnot_walk2_kernel(const DsParams p) {
  extern __shared__ __align__(16) uint32_t ds_dyn[];
  uint32_t* keyA = ds_dyn;                                   // [C] wanda score bits, pruned marker
  float* dm = reinterpret_cast<float*>(ds_dyn + p.C);        // [C] DSnoT metric, later the regrow sort key
  uint32_t* keyR = ds_dyn + p.C;
  __shared__ DsShared2 sh;
  constexpr int V = Elem<T>::kVec;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = p.C, maxc = p.max_cycle, k = p.k;
  const uint32_t Nk = (uint32_t)(3 * maxc + 16);              // kept candidates wanted: both sign classes then hold >= maxc
  // histogram of (key >> bshift) & bmask over the scores whose bits above pshift equal `prefix` (four keys per load)
  auto light_pass = [&](int pshift, uint32_t prefix, int bshift, uint32_t bmask) {
    for (int c4 = tid * 4; c4 < C; c4 += kDsThreads * 4) {
      const uint4 kq = *reinterpret_cast<const uint4*>(keyA + c4);
      const uint32_t ke[4] = {kq.x, kq.y, kq.z, kq.w};
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if ((ke[e] >> pshift) == prefix) atomicAdd(&sh.hist[(ke[e] >> bshift) & bmask], 1u);
    }
  };

  for (int row = blockIdx.x; row < p.R; row += gridDim.x) {
    const T* wrow = reinterpret_cast<const T*>(p.W) + (int64_t)row * p.ldw;
    __syncthreads();
    for (int b = tid; b < kDsBins; b += kDsThreads) sh.hist[b] = 0;
    if (tid == 0) { sh.fail = 0; sh.nkept = 0; sh.nhead = 0; sh.ntail = 0; sh.bmin = 0xffffffffu; sh.bmax = 0; }
    __syncthreads();
    // ---- pass 1: scores, DSnoT metric, histogram of the score's bits [20, 31) ----
    for (int c0 = tid * V; c0 < C; c0 += kDsThreads * V) {
      float f[V], sqv[V], smv[V];
      Elem<T>::unpack(ld_stream(wrow + c0), f);
#pragma unroll
      for (int q = 0; q < V / 4; ++q) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p.sq + c0) + q);
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.sum_row + c0) + q);
        sqv[4 * q] = a.x; sqv[4 * q + 1] = a.y; sqv[4 * q + 2] = a.z; sqv[4 * q + 3] = a.w;
        smv[4 * q] = b.x; smv[4 * q + 1] = b.y; smv[4 * q + 2] = b.z; smv[4 * q + 3] = b.w;
      }
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const int c = c0 + e;
        const uint32_t key = __float_as_uint(__fmul_rn(fabsf(f[e]), sqv[e]));
        keyA[c] = key;
        dm[c] = __fmul_rn(f[e], smv[e]);
        atomicAdd(&sh.hist[key >> 20], 1u);
      }
    }
    __syncthreads();
    // ---- exact k-th smallest (v, iv) and the 22-bit cut-off of rank k + Nk (three light passes) ----
    const uint32_t kq = (uint32_t)k + Nk < (uint32_t)C ? (uint32_t)k + Nk : (uint32_t)C;
    ds2_scan(sh, (uint32_t)k, kq, 2);
    const uint32_t b0 = sh.t_bin[0], bq0 = sh.t_bin[1];
    uint32_t kk = (uint32_t)k - sh.t_before[0];
    const uint32_t kkq = kq - sh.t_before[1];
    __syncthreads();
    for (int b = tid; b < kDsBins; b += kDsThreads) sh.hist[b] = 0;
    __syncthreads();
    light_pass(20, b0, 9, 0x7ffu);
    __syncthreads();
    uint32_t bq1 = 0;
    if (bq0 == b0) {                                           // both ranks in one top-level bin: one histogram serves both
      ds2_scan(sh, kk, kkq, 2);
      bq1 = sh.t_bin[1];
    } else {
      ds2_scan(sh, kk, 0, 1);
    }
    const uint32_t b1 = sh.t_bin[0];
    kk -= sh.t_before[0];
    __syncthreads();
    if (bq0 != b0) {
      for (int b = tid; b < kDsBins; b += kDsThreads) sh.hist[b] = 0;
      __syncthreads();
      light_pass(20, bq0, 9, 0x7ffu);
      __syncthreads();
      ds2_scan(sh, kkq, 0, 1);
      bq1 = sh.t_bin[0];
      __syncthreads();
    }
    const uint32_t vq22 = (bq0 << 11) | bq1;                   // kept columns with (key >> 9) <= vq22 are the candidates
    for (int b = tid; b < kDsBins; b += kDsThreads) sh.hist[b] = 0;
    __syncthreads();
    const uint32_t pre22 = (b0 << 11) | b1;
    light_pass(9, pre22, 0, 0x1ffu);
    __syncthreads();
    ds2_scan(sh, kk, 0, 1);
    const uint32_t v = (pre22 << 9) | sh.t_bin[0];
    const uint32_t need = kk - sh.t_before[0], ce = sh.t_count[0];
    __syncthreads();
    int iv = C;
    if (need < ce) iv = tie_bound([&](int i, uint32_t& key) { key = keyA[i]; return true; }, C, v, need, sh);
    __syncthreads();
    for (int b = tid; b < kDsBins; b += kDsThreads) sh.hist[b] = 0;
    __syncthreads();

    // ---- pass 2: classify; reconstruction error; regrow key + its histogram; kept candidates ----
    double esum = 0.0;
    // POW: 1 = divide by var (the default pow_of_var_regrowing), 0 = no division, 2 = divide by var^pow
    auto pass2 = [&](auto pow_tag) {
      constexpr int POW = decltype(pow_tag)::value;
      for (int c = tid; c < C; c += kDsThreads) {
        const uint32_t key = keyA[c];
        const float d = dm[c];
        const bool pruned = key < v || (key == v && c <= iv);
        float vv = 1.f;
        if (POW == 1) vv = p.var[c];
        if (POW == 2) vv = powf(p.var[c], p.pow_var);
        uint32_t kr;
        if (pruned) {
          esum += (double)d;
          keyA[c] = kPrunedKey;
          kr = sortable(POW == 0 ? d : __fdiv_rn(d, vv));
        } else {
          // 0 / var without the divide (a zero numerator takes the slow path of the IEEE division): 0 unless var is 0 or NaN
          kr = (POW != 0 && (vv == 0.f || vv != vv)) ? 0xffffffffu : 0x80000000u;
          if ((key >> 9) <= vq22 && d != 0.f) {                // a kept column with a zero metric belongs to neither class
            const uint32_t slot = atomicAdd(&sh.nkept, 1u);
            if (slot < (uint32_t)kW2Kept) sh.ckept[slot] = (uint16_t)((uint32_t)c | (d < 0.f ? 0x8000u : 0u));
          }
        }
        keyR[c] = kr;
        atomicAdd(&sh.hist[kr >> 21], 1u);
      }
    };
    if (p.pow_var == 1.f) pass2(std::integral_constant<int, 1>{});
    else if (p.pow_var == 0.f) pass2(std::integral_constant<int, 0>{});
    else pass2(std::integral_constant<int, 2>{});
    const float err0 = (float)block_sum(esum, sh);             // :601 (barriers inside: histogram and candidates complete)
    __syncthreads();
    // ---- the two ends of the regrow ordering: bins holding the maxc smallest / largest keys ----
    ds2_scan(sh, (uint32_t)maxc, (uint32_t)(C - maxc + 1), 2, true);
    const uint32_t bh = sh.t_bin[0], bt = sh.t_bin[1];
    const uint32_t n_head = sh.t_before[0] + sh.t_count[0];    // columns with bin <= bh
    const uint32_t n_tail = (uint32_t)C - sh.t_before[1];      // columns with bin >= bt
    const uint32_t bmin = sh.bmin, bmax = sh.bmax, nkept = sh.nkept;
    bool ok = nkept <= (uint32_t)kW2Kept && n_head <= (uint32_t)kW2End && n_tail <= (uint32_t)kW2End;
    __syncthreads();
    if (ok) {
      for (int c = tid; c < C; c += kDsThreads) {
        const uint32_t br = keyR[c] >> 21;
        if (br <= bh) sh.chead[atomicAdd(&sh.nhead, 1u)] = (uint16_t)c;
        if (br >= bt) sh.ctail[atomicAdd(&sh.ntail, 1u)] = (uint16_t)c;
      }
    }
    __syncthreads();
    // ---- four warps, four lists ----
    if (ok && warp < 4) {
      uint32_t* whist = sh.hist + warp * kW2Bins;
      bool good;
      if (warp < 2) {
        const uint32_t want_neg = warp == 0 ? 0x8000u : 0u;
        const uint32_t hi = (vq22 << 9) | 0x1ffu;
        const uint32_t range = hi - v;
        const int shift = range < (uint32_t)kW2Bins ? 0 : 32 - __clz(range) - 8;
        good = ds2_warp_list([&](uint16_t e, uint32_t& key, uint32_t& ord) {
                               ord = e & 0x7fffu; key = keyA[ord]; return (uint32_t)(e & 0x8000u) == want_neg; },
                             sh.ckept, (int)nkept, v, shift, maxc, whist, sh.ckey[warp], lane);
      } else if (warp == 2) {
        const uint32_t lo = bmin << 21, hi = (bh << 21) | 0x1fffffu;
        const uint32_t range = hi - lo;
        const int shift = range < (uint32_t)kW2Bins ? 0 : 32 - __clz(range) - 8;
        good = ds2_warp_list([&](uint16_t e, uint32_t& key, uint32_t& ord) { ord = e; key = keyR[e]; return true; },
                             sh.chead, (int)n_head, lo, shift, maxc, whist, sh.ckey[warp], lane);
      } else {
        // descending keys, ties by descending column: ascending order of (~key, C - 1 - column)
        const uint32_t lo = ~((bmax << 21) | 0x1fffffu), hi = ~(bt << 21);
        const uint32_t range = hi - lo;
        const int shift = range < (uint32_t)kW2Bins ? 0 : 32 - __clz(range) - 8;
        good = ds2_warp_list([&](uint16_t e, uint32_t& key, uint32_t& ord) { ord = (uint32_t)(C - 1) - e; key = ~keyR[e]; return true; },
                             sh.ctail, (int)n_tail, lo, shift, maxc, whist, sh.ckey[warp], lane);
      }
      if (!good) atomicOr(&sh.fail, 1);
      else {
        for (int i = lane; i < maxc; i += 32) {
          const uint32_t ord = (uint32_t)(sh.ckey[warp][i] & 0xffffffffull);
          const int col = warp == 3 ? (C - 1 - (int)ord) : (int)ord;
          sh.lists[warp][i] = col;
          sh.listsD[warp][i] = dsnot_metric<T>(wrow, p.sum_row, col);
        }
      }
    }
    __syncthreads();
    if (!ok || sh.fail) {
      if (tid == 0) p.fb_rows[atomicAdd(p.fb_count, 1)] = row;
      continue;
    }
    // ---- the cycle loop, one thread; ends at the row's stop cycle.  The head entry of each of the four lists sits in
    //      registers and is refilled as soon as it is consumed, so the shared-memory latency overlaps the error update ----
    if (tid == 0) {
      float err = err0;
      const float sign0 = sgnf(err0);
      int stop = 0;
      int hR = 0, tR = 0, hP = 0, tP = 0;
      int cN = sh.lists[0][0], cP = sh.lists[1][0], cH = sh.lists[2][0], cT = sh.lists[3][0];
      float dN = sh.listsD[0][0], dP = sh.listsD[1][0], dH = sh.listsD[2][0], dT = sh.listsD[3][0];
      for (int c = 1; c <= maxc; ++c) {
        int rg, pr; float rm, pm;
        if (err > 0.f) { rg = cT; rm = dT; ++tR; cT = sh.lists[3][tR & (kDsCap - 1)]; dT = sh.listsD[3][tR & (kDsCap - 1)]; }   // :654
        else { rg = cH; rm = dH; ++hR; cH = sh.lists[2][hR & (kDsCap - 1)]; dH = sh.listsD[2][hR & (kDsCap - 1)]; }
        if (err < 0.f) { pr = cP; pm = dP; ++tP; cP = sh.lists[1][tP & (kDsCap - 1)]; dP = sh.listsD[1][tP & (kDsCap - 1)]; }   // :683
        else { pr = cN; pm = dN; ++hP; cN = sh.lists[0][hP & (kDsCap - 1)]; dN = sh.listsD[0][hP & (kDsCap - 1)]; }
        const bool big = fabsf(err) > p.thr;
        bool upd = big;                                        // :717-720
        if (!p.without_same_sign) upd = big && (sign0 == sgnf(__fsub_rn(__fadd_rn(err, pm), rm)));   // :713, :722-729
        sh.walk[2 * (c - 1)] = pr;
        sh.walk[2 * (c - 1) + 1] = rg;
        if (!upd) { stop = c; break; }
        err = __fadd_rn(err, pm); err = __fsub_rn(err, rm);    // :742-751
      }
      p.row_v[row] = v;
      p.row_iv[row] = iv;
      p.row_stop[row] = stop | kDsSkipTail;
      atomicMax(p.ncycles, stop == 0 ? maxc : stop);
    }
    __syncthreads();
    for (int i = tid; i < 2 * maxc; i += kDsThreads) p.walk[(int64_t)row * 2 * maxc + i] = sh.walk[i];
  }
}

struct DsApplyParams {
  void* W; int64_t ldw; int R, C;
  const float* scaler_row;
  int prune_n, prune_m, initial_magnitude, max_cycle, ref_fixup, zero_w;
  const uint32_t* row_v; const int* row_iv; const int* row_stop; const int* walk; const int* ncycles;
  uint8_t* mask; int64_t ldm;
  const float* sq;        // sqrt(scaler_row), written by the walk entry point into the shared workspace
};

template <typename T>
__global__ void __launch_bounds__(kDsThreads)
dsnot_apply_kernel(const DsApplyParams p) {
  constexpr int V = Elem<T>::kVec;
  extern __shared__ __align__(16) uint8_t pm[];          // [C] 1 = pruned
  __shared__ int s_walk[2 * kDsCap];
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int C = p.C;
  int ncyc = *p.ncycles;
  if (ncyc > p.max_cycle) ncyc = p.max_cycle;

  for (int row = blockIdx.x; row < p.R; row += gridDim.x) {
    T* wrow = reinterpret_cast<T*>(p.W) + (int64_t)row * p.ldw;
    __syncthreads();
    for (int i = tid; i < 2 * ncyc; i += nthreads) s_walk[i] = p.walk[(int64_t)row * 2 * p.max_cycle + i];
    if (p.prune_n == 0) {
      const uint32_t v = p.row_v[row];
      const int iv = p.row_iv[row];
      for (int col = tid * V; col < C; col += nthreads * V) {
        float f[V];
        Elem<T>::unpack(ld_stream(wrow + col), f);
        uint32_t pb[V / 4] = {};
#pragma unroll
        for (int q = 0; q < V / 4; ++q) {
          float sv[4] = {1.f, 1.f, 1.f, 1.f};
          if (!p.initial_magnitude) {
            const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.sq + col) + q);
            sv[0] = s4.x; sv[1] = s4.y; sv[2] = s4.z; sv[3] = s4.w;
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float aw = fabsf(f[4 * q + e]);
            const uint32_t key = __float_as_uint(p.initial_magnitude ? aw : __fmul_rn(aw, sv[e]));
            const bool pr = key < v || (key == v && (col + 4 * q + e) <= iv);
            pb[q] |= (pr ? 1u : 0u) << (8 * e);
          }
        }
        if (V == 8) *reinterpret_cast<uint2*>(pm + col) = make_uint2(pb[0], pb[V / 4 - 1]);
        else *reinterpret_cast<uint32_t*>(pm + col) = pb[0];
      }
    } else {
      const int m = p.prune_m;
      for (int g = tid; g < C / m; g += nthreads) {
        uint32_t kk[kDsMaxM];
        for (int a = 0; a < m; ++a) {
          const float aw = fabsf(ds_to_float<T>(wrow[g * m + a]));
          kk[a] = __float_as_uint(p.initial_magnitude ? aw : __fmul_rn(aw, __fsqrt_rn(p.scaler_row[g * m + a])));
        }
        for (int a = 0; a < m; ++a) {
          int rank = 0;
          for (int b = 0; b < m; ++b) rank += (b < a) ? (kk[b] <= kk[a] ? 1 : 0) : ((b > a && kk[b] < kk[a]) ? 1 : 0);
          pm[g * m + a] = rank < p.prune_n ? 1 : 0;
        }
      }
    }
    __syncthreads();
    if (tid == 0) {
      const int stop_raw = p.row_stop[row];
      const int stop = stop_raw & ~kDsSkipTail;
      int last = ncyc;
      if ((stop_raw & kDsSkipTail) && stop > 0 && stop - 1 < last) last = stop - 1;   // later cycles change nothing (walk2)
      for (int c = 1; c <= last; ++c) {
        const int pr = s_walk[2 * (c - 1)], rg = s_walk[2 * (c - 1) + 1];
        const bool upd = stop == 0 || c < stop;
        if (p.ref_fixup && p.prune_n == 0) { pm[pr] = 0; pm[rg] = 1; }   // :731-740 net effect
        else { pm[pr] = upd ? 1 : 0; pm[rg] = upd ? 0 : 1; }             // :731-732 / :531-532
      }
    }
    __syncthreads();
    uint8_t* mrow = p.mask + (int64_t)row * p.ldm;
    for (int col = tid * V; col < C; col += nthreads * V) {
      // pm bytes are 0 / 1: the mask bytes are their complement; x * 0xff turns them into the byte masks of the weights
      uint32_t pb[2];
      if (V == 8) { const uint2 t = *reinterpret_cast<const uint2*>(pm + col); pb[0] = t.x; pb[1] = t.y; }
      else { pb[0] = *reinterpret_cast<const uint32_t*>(pm + col); pb[1] = 0; }
      if (V == 8) st_stream8(mrow + col, make_uint2(pb[0] ^ 0x01010101u, pb[1] ^ 0x01010101u));
      else st_stream4(mrow + col, pb[0] ^ 0x01010101u);
      if (p.zero_w && (pb[0] | pb[1])) {
        uint4 wv = *reinterpret_cast<const uint4*>(wrow + col);
        uint32_t* wr = reinterpret_cast<uint32_t*>(&wv);
        if (sizeof(T) == 4) {
          const uint32_t x = pb[0] * 0xffu;
          wr[0] &= ~__byte_perm(x, 0u, 0x0000); wr[1] &= ~__byte_perm(x, 0u, 0x1111);
          wr[2] &= ~__byte_perm(x, 0u, 0x2222); wr[3] &= ~__byte_perm(x, 0u, 0x3333);
        } else {
          const uint32_t x0 = pb[0] * 0xffu, x1 = pb[1] * 0xffu;
          wr[0] &= ~__byte_perm(x0, 0u, 0x1100); wr[1] &= ~__byte_perm(x0, 0u, 0x3322);
          wr[2] &= ~__byte_perm(x1, 0u, 0x1100); wr[3] &= ~__byte_perm(x1, 0u, 0x3322);
        }
        st_stream(wrow + col, wv);
      }
    }
  }
}

static size_t ds_state_bytes(int R, int C, int max_cycle) {
  return align_up((size_t)R * 4, 256) * 4 + align_up((size_t)C * 4, 256) + align_up((size_t)R * 2 * max_cycle * 4, 256);
}
size_t dsnot_refine_workspace_bytes(int R, int C, int max_cycle) {
  if (max_cycle <= 0 || max_cycle > kDsCap) max_cycle = kDsCap;
  return VLMC_WS_COUNTER_BYTES + ds_state_bytes(R, C, max_cycle);
}

struct DsState { uint32_t* row_v; int* row_iv; int* row_stop; int* fb_rows; int* fb_count; float* sq; int* walk; };
static DsState ds_carve(void* ws, int R, int C) {
  char* base = reinterpret_cast<char*>(ws) + VLMC_WS_COUNTER_BYTES;
  const size_t step = align_up((size_t)R * 4, 256);
  DsState s;
  s.row_v = reinterpret_cast<uint32_t*>(base);
  s.row_iv = reinterpret_cast<int*>(base + step);
  s.row_stop = reinterpret_cast<int*>(base + 2 * step);
  s.fb_rows = reinterpret_cast<int*>(base + 3 * step);
  s.fb_count = reinterpret_cast<int*>(ws);                  // first word of the counter area
  s.sq = reinterpret_cast<float*>(base + 4 * step);          // sqrt(scaler_row), written by the walk, read by the apply pass
  s.walk = reinterpret_cast<int*>(base + 4 * step + align_up((size_t)C * 4, 256));
  return s;
}

static int ds_common_checks(const void* W, int dtype, int R, int C, int64_t ldw, int prune_n, int prune_m, int max_cycle,
                            const void* ws, size_t ws_bytes) {
  if (!W || !ws || R < 1 || C < 1 || ldw < C) return VLMC_ERR_BAD_ARG;
  if (dtype != VLMC_F32 && dtype != VLMC_F16 && dtype != VLMC_BF16) return VLMC_ERR_BAD_ARG;
  if (max_cycle < 1 || max_cycle > kDsCap) return VLMC_ERR_UNSUPPORTED;
  if (prune_n < 0 || (prune_n > 0 && (prune_m <= prune_n || prune_m > kDsMaxM || C % prune_m != 0))) return VLMC_ERR_UNSUPPORTED;
  const int V = dtype == VLMC_F32 ? 4 : 8;
  if (C % V != 0 || ldw % V != 0 || ((uintptr_t)W & 15) != 0) return VLMC_ERR_UNSUPPORTED;
  if (C < max_cycle) return VLMC_ERR_UNSUPPORTED;     // the reference indexes out of bounds here
  if (!is_device_ptr(W) || !is_device_ptr(ws)) return VLMC_ERR_NOT_DEVICE;
  if (ws_bytes < dsnot_refine_workspace_bytes(R, C, max_cycle)) return VLMC_ERR_WORKSPACE;
  return VLMC_OK;
}

template <typename K>
static int ds_grid(K kern, size_t smem, int R, int* grid) {
  // static + dynamic shared memory together pass 48 KB long before the dynamic part alone does: always opt in
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem < 1024 ? 1024 : smem)) != cudaSuccess)
    return check_launch();
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kDsThreads, smem) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    return VLMC_ERR_UNSUPPORTED;                         // the row does not fit shared memory
  }
  *grid = kNumSMs * per_sm < R ? kNumSMs * per_sm : R;
  return VLMC_OK;
}

}  // namespace vlmc

extern "C" int vlmc_dsnot_refine_walk(const void* W, int dtype, int R, int C, int64_t ldw, const float* scaler_row,
                                      const float* sum_metric_row, const float* var, int k, int prune_n, int prune_m,
                                      float pow_of_var, int max_cycle_time, float update_threshold, int without_same_sign,
                                      int initial_magnitude, int argmin_rule, int* ncycles, void* ws, size_t ws_bytes,
                                      void* stream) {
  using namespace vlmc;
  int rc = ds_common_checks(W, dtype, R, C, ldw, prune_n, prune_m, max_cycle_time, ws, ws_bytes);
  if (rc) return rc;
  if (!scaler_row || !sum_metric_row || !var || !ncycles || k < 0) return VLMC_ERR_BAD_ARG;
  if (prune_n == 0 && C - k < max_cycle_time) return VLMC_ERR_UNSUPPORTED;
  if (!is_device_ptr(scaler_row) || !is_device_ptr(sum_metric_row) || !is_device_ptr(var) || !is_device_ptr(ncycles))
    return VLMC_ERR_NOT_DEVICE;
  cudaStream_t st = (cudaStream_t)stream;
  DsState s = ds_carve(ws, R, C);
  DsParams p;
  p.W = W; p.ldw = ldw; p.R = R; p.C = C; p.scaler_row = scaler_row; p.sum_row = sum_metric_row; p.var = var;
  p.k = prune_n ? 0 : k; p.prune_n = prune_n; p.prune_m = prune_m; p.pow_var = pow_of_var; p.max_cycle = max_cycle_time;
  p.thr = update_threshold; p.without_same_sign = without_same_sign; p.initial_magnitude = initial_magnitude;
  p.argmin_rule = argmin_rule;
  p.row_v = s.row_v; p.row_iv = s.row_iv; p.row_stop = s.row_stop; p.walk = s.walk; p.ncycles = ncycles;
  p.rows = nullptr; p.nrows = nullptr; p.fb_rows = s.fb_rows; p.fb_count = s.fb_count;
  if (cudaMemsetAsync(ncycles, 0, sizeof(int), st) != cudaSuccess) return check_launch();
  const size_t smem = (size_t)C * 8;
  int grid = 1;
  // The 5-pass kernel takes the unstructured Wanda-initialised walk; rows it cannot take (and every other configuration)
  // go through dsnot_walk_kernel.  VLMC_DSNOT_WALK_V1=1 forces the old kernel for every row (A/B runs, tests).
  const char* v1e = getenv("VLMC_DSNOT_WALK_V1");
  const bool fast = !(v1e && v1e[0] == '1') && prune_n == 0 && !initial_magnitude && k > 0 && C < 32768 &&
                    C >= 2 * max_cycle_time && C - k >= 2 * max_cycle_time && ((uintptr_t)sum_metric_row & 15) == 0;
  p.sq = s.sq;
  ds_sqrt_kernel<<<(C + 255) / 256, 256, 0, st>>>(scaler_row, s.sq, C);      // also read by vlmc_dsnot_refine_apply
  if (fast) {
    if (cudaMemsetAsync(s.fb_count, 0, sizeof(int), st) != cudaSuccess) return check_launch();
#define VLMC_DS_WALK2(TT) { rc = ds_grid(dsnot_walk2_kernel<TT>, smem, R, &grid); if (rc) return rc; \
                            dsnot_walk2_kernel<TT><<<grid, kDsThreads, smem, st>>>(p); }
    switch (dtype) {
      case VLMC_F32: VLMC_DS_WALK2(float); break;
      case VLMC_F16: VLMC_DS_WALK2(__half); break;
      default: VLMC_DS_WALK2(__nv_bfloat16); break;
    }
#undef VLMC_DS_WALK2
    if (check_launch() != VLMC_OK) return VLMC_ERR_CUDA;
    p.rows = s.fb_rows; p.nrows = s.fb_count;                // second launch: only the rows handed back (usually none)
  }
#define VLMC_DS_WALK(TT) { rc = ds_grid(dsnot_walk_kernel<TT>, smem, R, &grid); if (rc) return rc; \
                           if (fast && grid > kNumSMs) grid = kNumSMs; \
                           dsnot_walk_kernel<TT><<<grid, kDsThreads, smem, st>>>(p); }
  switch (dtype) {
    case VLMC_F32: VLMC_DS_WALK(float); break;
    case VLMC_F16: VLMC_DS_WALK(__half); break;
    default: VLMC_DS_WALK(__nv_bfloat16); break;
  }
#undef VLMC_DS_WALK
  return check_launch();
}

extern "C" int vlmc_dsnot_refine_apply(void* W, int dtype, int R, int C, int64_t ldw, const float* scaler_row,
                                       int prune_n, int prune_m, int initial_magnitude, int max_cycle_time,
                                       const int* ncycles, int ref_fixup, int zero_w, uint8_t* keep_mask, int64_t ldm,
                                       void* ws, size_t ws_bytes, void* stream) {
  using namespace vlmc;
  int rc = ds_common_checks(W, dtype, R, C, ldw, prune_n, prune_m, max_cycle_time, ws, ws_bytes);
  if (rc) return rc;
  if (!scaler_row || !ncycles || !keep_mask || ldm < C) return VLMC_ERR_BAD_ARG;
  const int V = dtype == VLMC_F32 ? 4 : 8;
  if (ldm % V != 0 || ((uintptr_t)keep_mask & 7) != 0) return VLMC_ERR_UNSUPPORTED;
  if (!is_device_ptr(scaler_row) || !is_device_ptr(ncycles) || !is_device_ptr(keep_mask)) return VLMC_ERR_NOT_DEVICE;
  cudaStream_t st = (cudaStream_t)stream;
  DsState s = ds_carve(ws, R, C);
  DsApplyParams p;
  p.W = W; p.ldw = ldw; p.R = R; p.C = C; p.scaler_row = scaler_row; p.prune_n = prune_n; p.prune_m = prune_m;
  p.initial_magnitude = initial_magnitude; p.max_cycle = max_cycle_time; p.ref_fixup = ref_fixup; p.zero_w = zero_w;
  p.row_v = s.row_v; p.row_iv = s.row_iv; p.row_stop = s.row_stop; p.walk = s.walk; p.ncycles = ncycles;
  p.mask = keep_mask; p.ldm = ldm; p.sq = s.sq;
  const size_t smem = align_up((size_t)C, 16);
  int grid = 1;
#define VLMC_DS_APPLY(TT) { rc = ds_grid(dsnot_apply_kernel<TT>, smem, R, &grid); if (rc) return rc; \
                            dsnot_apply_kernel<TT><<<grid, kDsThreads, smem, st>>>(p); }
  switch (dtype) {
    case VLMC_F32: VLMC_DS_APPLY(float); break;
    case VLMC_F16: VLMC_DS_APPLY(__half); break;
    default: VLMC_DS_APPLY(__nv_bfloat16); break;
  }
#undef VLMC_DS_APPLY
  return check_launch();
}

extern "C" int vlmc_dsnot_refine(void* W, int dtype, int R, int C, int64_t ldw, const float* scaler_row,
                                 const float* sum_metric_row, const float* var, int k, int prune_n, int prune_m,
                                 float pow_of_var, int max_cycle_time, float update_threshold, int without_same_sign,
                                 int initial_magnitude, int argmin_rule, int ref_fixup, int zero_w, uint8_t* keep_mask,
                                 int64_t ldm, int* ncycles, void* ws, size_t ws_bytes, void* stream) {
  int rc = vlmc_dsnot_refine_walk(W, dtype, R, C, ldw, scaler_row, sum_metric_row, var, k, prune_n, prune_m, pow_of_var,
                                  max_cycle_time, update_threshold, without_same_sign, initial_magnitude, argmin_rule,
                                  ncycles, ws, ws_bytes, stream);
  if (rc) return rc;
  return vlmc_dsnot_refine_apply(W, dtype, R, C, ldw, scaler_row, prune_n, prune_m, initial_magnitude, max_cycle_time,
                                 ncycles, ref_fixup, zero_w, keep_mask, ldm, ws, ws_bytes, stream);
}
