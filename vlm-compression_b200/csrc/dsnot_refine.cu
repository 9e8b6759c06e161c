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
};

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

__device__ __forceinline__ uint32_t block_sum(uint32_t v, DsShared& sh) {
  v = __reduce_add_sync(0xffffffffu, v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh.warp_tot[threadIdx.x >> 5] = v;
  __syncthreads();
  uint32_t t = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh.warp_tot[w];
  return t;
}
__device__ __forceinline__ double block_sum(double v, DsShared& sh) {
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
template <class F>
__device__ int tie_bound(F f, int n, uint32_t v, uint32_t need, DsShared& sh) {
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

  for (int row = blockIdx.x; row < p.R; row += gridDim.x) {
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

struct DsApplyParams {
  void* W; int64_t ldw; int R, C;
  const float* scaler_row;
  int prune_n, prune_m, initial_magnitude, max_cycle, ref_fixup, zero_w;
  const uint32_t* row_v; const int* row_iv; const int* row_stop; const int* walk; const int* ncycles;
  uint8_t* mask; int64_t ldm;
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
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const float aw = fabsf(f[e]);
          const uint32_t key = __float_as_uint(p.initial_magnitude ? aw : __fmul_rn(aw, __fsqrt_rn(p.scaler_row[col + e])));
          pm[col + e] = (key < v || (key == v && (col + e) <= iv)) ? 1 : 0;
        }
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
      const int stop = p.row_stop[row];
      for (int c = 1; c <= ncyc; ++c) {
        const int pr = s_walk[2 * (c - 1)], rg = s_walk[2 * (c - 1) + 1];
        const bool upd = stop == 0 || c < stop;
        if (p.ref_fixup && p.prune_n == 0) { pm[pr] = 0; pm[rg] = 1; }   // :731-740 net effect
        else { pm[pr] = upd ? 1 : 0; pm[rg] = upd ? 0 : 1; }             // :731-732 / :531-532
      }
    }
    __syncthreads();
    uint8_t* mrow = p.mask + (int64_t)row * p.ldm;
    for (int col = tid * V; col < C; col += nthreads * V) {
      uint32_t mb[V / 4] = {};
      bool any = false;
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const bool pr = pm[col + e] != 0;
        any |= pr;
        mb[e / 4] |= (pr ? 0u : 1u) << (8 * (e % 4));
      }
      if (V == 8) st_stream8(mrow + col, make_uint2(mb[0], mb[V / 4 - 1]));
      else st_stream4(mrow + col, mb[0]);
      if (p.zero_w && any) {
        uint4 wv = *reinterpret_cast<const uint4*>(wrow + col);
        uint32_t* wr = reinterpret_cast<uint32_t*>(&wv);
        if (sizeof(T) == 4) {
#pragma unroll
          for (int e = 0; e < V; ++e) if (pm[col + e]) wr[e] = 0u;
        } else {
#pragma unroll
          for (int e = 0; e < V; ++e) if (pm[col + e]) wr[e / 2] &= (e & 1) ? 0x0000ffffu : 0xffff0000u;
        }
        st_stream(wrow + col, wv);
      }
    }
  }
}

static size_t ds_state_bytes(int R, int max_cycle) {
  return align_up((size_t)R * 4, 256) * 3 + align_up((size_t)R * 2 * max_cycle * 4, 256);
}
size_t dsnot_refine_workspace_bytes(int R, int max_cycle) {
  if (max_cycle <= 0 || max_cycle > kDsCap) max_cycle = kDsCap;
  return VLMC_WS_COUNTER_BYTES + ds_state_bytes(R, max_cycle);
}

struct DsState { uint32_t* row_v; int* row_iv; int* row_stop; int* walk; };
static DsState ds_carve(void* ws, int R) {
  char* base = reinterpret_cast<char*>(ws) + VLMC_WS_COUNTER_BYTES;
  const size_t step = align_up((size_t)R * 4, 256);
  DsState s;
  s.row_v = reinterpret_cast<uint32_t*>(base);
  s.row_iv = reinterpret_cast<int*>(base + step);
  s.row_stop = reinterpret_cast<int*>(base + 2 * step);
  s.walk = reinterpret_cast<int*>(base + 3 * step);
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
  if (ws_bytes < dsnot_refine_workspace_bytes(R, max_cycle)) return VLMC_ERR_WORKSPACE;
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
  DsState s = ds_carve(ws, R);
  DsParams p;
  p.W = W; p.ldw = ldw; p.R = R; p.C = C; p.scaler_row = scaler_row; p.sum_row = sum_metric_row; p.var = var;
  p.k = prune_n ? 0 : k; p.prune_n = prune_n; p.prune_m = prune_m; p.pow_var = pow_of_var; p.max_cycle = max_cycle_time;
  p.thr = update_threshold; p.without_same_sign = without_same_sign; p.initial_magnitude = initial_magnitude;
  p.argmin_rule = argmin_rule;
  p.row_v = s.row_v; p.row_iv = s.row_iv; p.row_stop = s.row_stop; p.walk = s.walk; p.ncycles = ncycles;
  if (cudaMemsetAsync(ncycles, 0, sizeof(int), st) != cudaSuccess) return check_launch();
  const size_t smem = (size_t)C * 8;
  int grid = 1;
#define VLMC_DS_WALK(TT) { rc = ds_grid(dsnot_walk_kernel<TT>, smem, R, &grid); if (rc) return rc; \
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
  DsState s = ds_carve(ws, R);
  DsApplyParams p;
  p.W = W; p.ldw = ldw; p.R = R; p.C = C; p.scaler_row = scaler_row; p.prune_n = prune_n; p.prune_m = prune_m;
  p.initial_magnitude = initial_magnitude; p.max_cycle = max_cycle_time; p.ref_fixup = ref_fixup; p.zero_w = zero_w;
  p.row_v = s.row_v; p.row_iv = s.row_iv; p.row_stop = s.row_stop; p.walk = s.walk; p.ncycles = ncycles;
  p.mask = keep_mask; p.ldm = ldm;
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
