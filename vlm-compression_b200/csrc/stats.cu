// K1 / K2: per-input-channel calibration statistics, one streaming pass over the activations.
//
// Replaces WrappedGPT.add_batch of the reference:
//   Wanda  lavis/compression/pruners/wanda_pruner.py:66-81   (scaler_row)
//   DSnoT  lavis/compression/pruners/dsnot_pruner.py:79-101  (scaler_row, sum_metric_row, mean, var)
//
// Layout: x is [T, C] row-major (what the reference reaches with inp.reshape(-1, C); its .t() is a
// view, so the reduction there runs over a strided dimension).  Here a CTA owns a tile of
// kCX x 16 B columns (512 B for Wanda, 256 B for DSnoT) and a contiguous chunk of rows; its 256 threads cover
// 256 / kCX rows per step, 4 steps in flight per thread, fp32 partials in registers.  Partials of all row chunks go to the
// workspace; the last CTA to finish a column tile (ticket counter) combines them in fp64 in a
// fixed order, so the result is deterministic, and applies the running-average update.
//
// HBM-bound: algorithmic bytes = T*C*sizeof(x) (+ a few vectors of C floats).
#include <stdlib.h>
#include "common.cuh"

namespace vlmc {

constexpr int kStatsThreads = 256;   // = column-vector lanes (CX) x row lanes (RY) per CTA
constexpr int kUnroll = 4;

struct StatsParams {
  const void* x;
  int64_t ldx;
  int C;
  int64_t S;             // rows per segment
  int64_t nseg;
  int chunks_per_seg;
  int rows_per_chunk;
  float* part;           // [kinds][nchunks][C]
  unsigned int* tickets; // one per column tile, zero on entry, zero on exit
  // running-average state
  float* scaler_row;
  float* sum_row;
  float* mean;
  float* var;
  double n_before, b_per_seg, ntok_before;
};

// sm_100a mixed-precision FP32 ops with 16-bit operands (FHFMA / FHADD, PTX fma.rn.f32.f16 / sub.rn.f32.f16 and the .bf16
// forms): the half of a packed register is an operand selector (.H1), so a 16-bit activation enters the fp32 accumulation
// without an unpack instruction.  Same values as unpack + fmaf / fsub: the product of two 11-bit significands is exact in
// fp32, one rounding at the end.
template <typename T> struct Mix;
template <> struct Mix<__half> {
  __device__ static __forceinline__ float sq_acc(unsigned short a, float c) { float d; asm("fma.rn.f32.f16 %0, %1, %1, %2;" : "=f"(d) : "h"(a), "f"(c)); return d; }
  __device__ static __forceinline__ float sub(unsigned short a, float c) { float d; asm("sub.rn.f32.f16 %0, %1, %2;" : "=f"(d) : "h"(a), "f"(c)); return d; }
};
template <> struct Mix<__nv_bfloat16> {
  __device__ static __forceinline__ float sq_acc(unsigned short a, float c) { float d; asm("fma.rn.f32.bf16 %0, %1, %1, %2;" : "=f"(d) : "h"(a), "f"(c)); return d; }
  __device__ static __forceinline__ float sub(unsigned short a, float c) { float d; asm("sub.rn.f32.bf16 %0, %1, %2;" : "=f"(d) : "h"(a), "f"(c)); return d; }
};
template <> struct Mix<float> {   // never used (fp32 activations take the plain path); keeps the template well-formed
  __device__ static __forceinline__ float sq_acc(unsigned short, float c) { return c; }
  __device__ static __forceinline__ float sub(unsigned short, float c) { return c; }
};

// one 16-byte vector into the per-column accumulators
template <typename T, bool DSNOT>
__device__ __forceinline__ void colstats_accum(const uint4& v, float (&a0)[Elem<T>::kVec], float (&a1)[Elem<T>::kVec],
                                               const float (&x0)[Elem<T>::kVec]) {
  constexpr int V = Elem<T>::kVec;
  // Wanda's kernel keeps the unpack + fmaf form: it lives in 32 registers (8 CTAs per SM) and the FHFMA form spilled there
  if (sizeof(T) == 2 && DSNOT) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const unsigned short lo = (unsigned short)(w[q] & 0xffffu), hi = (unsigned short)(w[q] >> 16);
      constexpr int kLast = V - 1;
      const int i0 = (2 * q) & kLast, i1 = (2 * q + 1) & kLast;      // V == 8 here; the mask keeps the fp32 instantiation in bounds
      if (DSNOT) {
        const float d0 = Mix<T>::sub(lo, x0[i0]), d1 = Mix<T>::sub(hi, x0[i1]);
        a0[i0] += d0; a1[i0] = fmaf(d0, d0, a1[i0]);
        a0[i1] += d1; a1[i1] = fmaf(d1, d1, a1[i1]);
      } else {
        a1[i0] = Mix<T>::sq_acc(lo, a1[i0]);
        a1[i1] = Mix<T>::sq_acc(hi, a1[i1]);
      }
    }
  } else {
    float f[V];
    Elem<T>::unpack(v, f);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      if (DSNOT) {
        const float d = f[i] - x0[i];
        a0[i] += d;
        a1[i] = fmaf(d, d, a1[i]);
      } else {
        a1[i] = fmaf(f[i], f[i], a1[i]);
      }
    }
  }
}

// resident CTAs per SM the grid is sized for (one full wave, no tail): the Wanda variant fits 32 registers
// (8 x 256 threads = 64 warps / SM); the DSnoT variant carries 3x the per-thread state
template <bool DSNOT> struct StatsOcc {
  static constexpr int kBlocksPerSM = DSNOT ? 3 : 8;
  // 16-byte loads in flight per thread.  The DSnoT variant is bound by bytes in flight (ncu r02ab: long-scoreboard stalls
  // 16 per issue, 4.9 TB/s): its 24 registers of per-column state leave room for 4 loads at 4 CTAs per SM = 64 KB per SM in
  // flight against the 128 KB of the Wanda variant; 8 loads at 3 CTAs per SM put 96 KB in flight.
  static constexpr int kLoads = DSNOT ? 8 : 4;
};

template <typename T, bool DSNOT, int kCX>
__device__ __forceinline__ void colstats_body(const StatsParams& p, const int ct, const int64_t chunk) {
  constexpr int V = Elem<T>::kVec;
  constexpr int kRY = kStatsThreads / kCX;
  const int tx = threadIdx.x % kCX;
  const int ty = threadIdx.x / kCX;
  const int col0 = (ct * kCX + tx) * V;
  const bool col_ok = col0 < p.C;
  const int64_t nchunks = p.nseg * p.chunks_per_seg;
  const int64_t seg = chunk / p.chunks_per_seg;
  const int64_t cis = chunk % p.chunks_per_seg;
  const int64_t r0 = seg * p.S + cis * (int64_t)p.rows_per_chunk;
  int64_t r1 = r0 + p.rows_per_chunk;
  const int64_t seg_end = (seg + 1) * p.S;
  if (r1 > seg_end) r1 = seg_end;

  const T* xp = reinterpret_cast<const T*>(p.x) + col0;

  float a0[V], a1[V], x0[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { a0[i] = 0.f; a1[i] = 0.f; x0[i] = 0.f; }

  if (col_ok && r0 < r1) {
    if (DSNOT) {
      // shift by the chunk's first row: keeps sum(d^2) - sum(d)^2/n free of cancellation
      uint4 v = ld_stream(xp + r0 * p.ldx);
      Elem<T>::unpack(v, x0);
    }
    int64_t r = r0 + ty;
    constexpr int kU = StatsOcc<DSNOT>::kLoads;
    for (; r + (kU - 1) * kRY < r1; r += kU * kRY) {
      uint4 v[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) v[u] = ld_stream(xp + (r + u * kRY) * p.ldx);
#pragma unroll
      for (int u = 0; u < kU; ++u) colstats_accum<T, DSNOT>(v[u], a0, a1, x0);
    }
    for (; r < r1; r += kRY) {
      uint4 v = ld_stream(xp + r * p.ldx);
      colstats_accum<T, DSNOT>(v, a0, a1, x0);
    }
  }

  // combine the kRY row lanes of one column through shared memory
  __shared__ float red[2][kRY][kCX * V + 4];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    red[1][ty][tx * V + i] = a1[i];
    if (DSNOT) red[0][ty][tx * V + i] = a0[i];
  }
  __syncthreads();
  if (ty == 0 && col_ok) {
    float* part_sq = p.part + (size_t)chunk * p.C + col0;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float s1 = 0.f;
#pragma unroll
      for (int y = 0; y < kRY; ++y) s1 += red[1][y][tx * V + i];
      part_sq[i] = s1;
    }
    if (DSNOT) {
      float* part_s = p.part + ((size_t)nchunks + chunk) * p.C + col0;
      float* part_x0 = p.part + ((size_t)2 * nchunks + chunk) * p.C + col0;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float s0 = 0.f;
#pragma unroll
        for (int y = 0; y < kRY; ++y) s0 += red[0][y][tx * V + i];
        part_s[i] = s0;
        part_x0[i] = x0[i];
      }
    }
  }

  // ticket: the last CTA of this column tile finalises
  __shared__ unsigned int s_ticket;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(&p.tickets[ct], 1u);
  __syncthreads();
  if (s_ticket != (unsigned int)(nchunks - 1)) return;
  __threadfence();

  const double n_after = p.n_before + p.b_per_seg * (double)p.nseg;
  const float ratio = (float)(p.n_before / n_after);
  const float n_after_f = (float)n_after;
  const int tile_c0 = ct * kCX * V;
  for (int c = tile_c0 + threadIdx.x; c < tile_c0 + kCX * V && c < p.C; c += kStatsThreads) {
    if (!DSNOT) {
      // the partials of all chunks, added in chunk order; loads are issued 16 at a time so the (serial, last-CTA) tail
      // of the launch costs ~nchunks/16 L2 round trips instead of nchunks
      double tot = 0.0;
      int64_t k = 0;
      for (; k + 16 <= nchunks; k += 16) {
        float v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = __ldcg(p.part + (size_t)(k + u) * p.C + c);
#pragma unroll
        for (int u = 0; u < 16; ++u) tot += (double)v[u];
      }
      for (; k < nchunks; ++k) tot += (double)__ldcg(p.part + (size_t)k * p.C + c);
      // wanda_pruner.py:77,81 -- scaler_row *= n/(n+b); scaler_row += ||x||^2 / (n+b)
      float s = __fmul_rn(p.scaler_row[c], ratio);
      p.scaler_row[c] = __fadd_rn(s, __fdiv_rn((float)tot, n_after_f));
    } else {
      double tot_sq = 0.0, tot_s = 0.0, var_acc = 0.0, mean_acc = 0.0;
      for (int64_t sg = 0; sg < p.nseg; ++sg) {
        // Chan merge of this segment's chunks -> per-call mean and biased variance
        double n = 0.0, mu = 0.0, m2 = 0.0;
        for (int ck = 0; ck < p.chunks_per_seg; ++ck) {
          const int64_t k = sg * p.chunks_per_seg + ck;
          int64_t rr0 = ck * (int64_t)p.rows_per_chunk;
          int64_t rr1 = rr0 + p.rows_per_chunk;
          if (rr1 > p.S) rr1 = p.S;
          if (rr1 <= rr0) continue;
          const double nb = (double)(rr1 - rr0);
          const double sd2 = (double)__ldcg(p.part + (size_t)k * p.C + c);
          const double sd = (double)__ldcg(p.part + ((size_t)nchunks + k) * p.C + c);
          const double xs = (double)__ldcg(p.part + ((size_t)2 * nchunks + k) * p.C + c);
          const double mub = xs + sd / nb;
          const double m2b = sd2 - sd * sd / nb;
          tot_s += sd + nb * xs;
          tot_sq += sd2 + 2.0 * xs * sd + nb * xs * xs;
          const double nn = n + nb;
          const double delta = mub - mu;
          m2 += m2b + delta * delta * n * nb / nn;
          mu += delta * nb / nn;
          n = nn;
        }
        var_acc += m2;           // = S * biased var of the call
        mean_acc += mu * n;
      }
      // dsnot_pruner.py:89-93 -- token-weighted running means of the per-call mean / variance
      const double tok_call = (double)p.S * (double)p.nseg;
      const double tok_after = p.ntok_before + tok_call;
      if (p.ntok_before == 0.0) {
        p.var[c] = (float)(var_acc / tok_call);
        p.mean[c] = (float)(mean_acc / tok_call);
      } else {
        p.var[c] = (float)(((double)p.var[c] * p.ntok_before + var_acc) / tok_after);
        p.mean[c] = (float)(((double)p.mean[c] * p.ntok_before + mean_acc) / tok_after);
      }
      // dsnot_pruner.py:96-101
      float s = __fmul_rn(p.scaler_row[c], ratio);
      p.scaler_row[c] = __fadd_rn(s, __fdiv_rn((float)tot_sq, n_after_f));
      float m = __fmul_rn(p.sum_row[c], ratio);
      p.sum_row[c] = __fadd_rn(m, __fdiv_rn((float)tot_s, n_after_f));
    }
  }
  if (threadIdx.x == 0) p.tickets[ct] = 0u;  // leave the workspace clean
}

template <typename T, bool DSNOT, int kCX>
__global__ void __launch_bounds__(kStatsThreads, StatsOcc<DSNOT>::kBlocksPerSM)
colstats_kernel(const StatsParams p) {
  colstats_body<T, DSNOT, kCX>(p, blockIdx.x, blockIdx.y);
}

// Several accumulations in ONE launch (the linears of a block): blockIdx.z picks the item, CTAs outside an item's own
// (column tiles x chunks) rectangle exit at once.  On 8 GPUs a rank's share of a linear is a 40-100 us launch that spends
// a fifth of its time ramping up and draining; one launch for the block has one ramp and one drain.
constexpr int kStatsBatchMax = 16;
struct StatsBatch { StatsParams it[kStatsBatchMax]; int coltiles[kStatsBatchMax]; int nchunks[kStatsBatchMax]; int chunk_major; };

template <typename T, bool DSNOT, int kCX>
__global__ void __launch_bounds__(kStatsThreads, StatsOcc<DSNOT>::kBlocksPerSM)
colstats_batch_kernel(const __grid_constant__ StatsBatch b) {
  // Wanda: one resident wave, grid (column tiles, chunks, items).  DSnoT: a multi-wave grid (one chunk per call and column
  // tile), grid (column tiles, ITEMS, chunks): CTAs are dispatched x-fastest, so the linears that read the same
  // activations (q / k / v, gate / up) work on the same call at the same time and the repeats hit in L2.
  const int item = b.chunk_major ? blockIdx.y : blockIdx.z;
  const int chunk = b.chunk_major ? blockIdx.z : blockIdx.y;
  if ((int)blockIdx.x >= b.coltiles[item] || chunk >= b.nchunks[item]) return;
  colstats_body<T, DSNOT, kCX>(b.it[item], blockIdx.x, chunk);
}

struct StatsPlan {
  int coltiles, chunks_per_seg, rows_per_chunk;
  int64_t nchunks;
  size_t bytes;
};

// column lanes per CTA (x 16 B = contiguous bytes a CTA reads per row).  Measured on B200 at T = 262144 fp16 rows
// (scripts/stats_probe.py): Wanda 32 lanes (512 B x 8 rows per step) 6.52 TB/s at C = 4096 and 7.21 TB/s at C = 11008
// vs 5.93 / 7.09 with 64 lanes; DSnoT 16 lanes 4.93 / 6.16 TB/s vs 3.77 / 5.28 with 64.
template <bool DSNOT> struct StatsTile { static constexpr int kCX = DSNOT ? 16 : 32; };

// resident waves of the single-tensor Wanda launch (VLMC_STATS_WAVES, experiment switch; default below)
static int64_t k1_single_target() {
  int waves = 1;
  if (const char* e = getenv("VLMC_STATS_WAVES")) { const int v = atoi(e); if (v >= 1 && v <= 64) waves = v; }
  return (int64_t)kNumSMs * StatsOcc<false>::kBlocksPerSM * waves;
}

static StatsPlan plan_stats(int dtype, int64_t nseg, int64_t S, int C, int kinds, int blocks_per_sm = 8, int64_t target = 0) {
  StatsPlan pl;
  const int V = dtype == VLMC_F32 ? 4 : 8;
  const int kCX = kinds == 3 ? StatsTile<true>::kCX : StatsTile<false>::kCX, kRY = kStatsThreads / kCX;
  pl.coltiles = (C + kCX * V - 1) / (kCX * V);
  const int64_t target_ctas = target > 0 ? target : (int64_t)kNumSMs * blocks_per_sm;  // exactly one resident wave
  int64_t want = target_ctas / pl.coltiles;
  if (want < 1) want = 1;
  int64_t cps = want / nseg;
  if (cps < 1) cps = 1;
  const int64_t min_rows = kRY * kUnroll * 2;
  // (splitting the segments of a multi-wave DSnoT grid into 2-3 chunks to fill the last wave was measured: 0.47 -> 0.69 ms
  // at C = 4096 - the per-chunk set-up and the longer last-CTA merge cost more than the 8 % idle tail)
  int64_t max_cps = (S + min_rows - 1) / min_rows;
  if (cps > max_cps) cps = max_cps;
  if (cps < 1) cps = 1;
  pl.rows_per_chunk = (int)((S + cps - 1) / cps);
  pl.chunks_per_seg = (int)((S + pl.rows_per_chunk - 1) / pl.rows_per_chunk);
  pl.nchunks = nseg * pl.chunks_per_seg;
  pl.bytes = VLMC_WS_COUNTER_BYTES + (size_t)kinds * pl.nchunks * C * sizeof(float);
  return pl;
}

size_t stats_workspace_bytes(int dsnot, int64_t T, int C, int64_t nseg) {
  if (nseg < 1) nseg = 1;
  // dtype only changes the tile count; take the larger of the fp32 and the 16-bit plan (same occupancy as the launch)
  const int bps = dsnot ? StatsOcc<true>::kBlocksPerSM : StatsOcc<false>::kBlocksPerSM;
  const int64_t tgt = dsnot ? 0 : k1_single_target();
  size_t a = plan_stats(VLMC_F32, nseg, T / nseg, C, dsnot ? 3 : 1, bps, tgt).bytes;
  size_t b = plan_stats(VLMC_F16, nseg, T / nseg, C, dsnot ? 3 : 1, bps, tgt).bytes;
  return a > b ? a : b;
}

template <bool DSNOT>
static int launch_stats(const void* x, int dtype, int64_t nseg, int64_t S, int C, int64_t ldx,
                        float* scaler_row, float* sum_row, float* mean, float* var,
                        double n_before, double b_per_seg, double ntok_before,
                        void* ws, size_t ws_bytes, void* stream) {
  if (!x || !scaler_row || !ws || nseg < 1 || S < 1 || C < 1 || ldx < C) return VLMC_ERR_BAD_ARG;
  if (DSNOT && (!sum_row || !mean || !var)) return VLMC_ERR_BAD_ARG;
  if (dtype != VLMC_F32 && dtype != VLMC_F16 && dtype != VLMC_BF16) return VLMC_ERR_BAD_ARG;
  const int V = dtype == VLMC_F32 ? 4 : 8;
  if (C % V != 0 || ldx % V != 0 || ((uintptr_t)x & 15) != 0) return VLMC_ERR_UNSUPPORTED;
  if (!is_device_ptr(x) || !is_device_ptr(scaler_row) || !is_device_ptr(ws)) return VLMC_ERR_NOT_DEVICE;
  StatsPlan pl = plan_stats(dtype, nseg, S, C, DSNOT ? 3 : 1, StatsOcc<DSNOT>::kBlocksPerSM, DSNOT ? 0 : k1_single_target());
  if (ws_bytes < pl.bytes) return VLMC_ERR_WORKSPACE;
  if (pl.coltiles * sizeof(unsigned int) > VLMC_WS_COUNTER_BYTES) return VLMC_ERR_UNSUPPORTED;
  if (pl.nchunks > 65535) return VLMC_ERR_UNSUPPORTED;

  StatsParams p;
  p.x = x; p.ldx = ldx; p.C = C; p.S = S; p.nseg = nseg;
  p.chunks_per_seg = pl.chunks_per_seg; p.rows_per_chunk = pl.rows_per_chunk;
  p.tickets = reinterpret_cast<unsigned int*>(ws);
  p.part = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + VLMC_WS_COUNTER_BYTES);
  p.scaler_row = scaler_row; p.sum_row = sum_row; p.mean = mean; p.var = var;
  p.n_before = n_before; p.b_per_seg = b_per_seg; p.ntok_before = ntok_before;

  dim3 grid(pl.coltiles, (unsigned)pl.nchunks);
  cudaStream_t st = (cudaStream_t)stream;
  VLMC_DISPATCH_DTYPE(dtype, (colstats_kernel<scalar_t, DSNOT, StatsTile<DSNOT>::kCX><<<grid, kStatsThreads, 0, st>>>(p)));
  return check_launch();
}

}  // namespace vlmc

extern "C" int vlmc_sqnorm_accum(const void* x, int dtype, int64_t T, int C, int64_t ldx,
                                 float* scaler_row, double n_before, double b,
                                 void* ws, size_t ws_bytes, void* stream) {
  return vlmc::launch_stats<false>(x, dtype, 1, T, C, ldx, scaler_row, nullptr, nullptr, nullptr,
                                   n_before, b, 0.0, ws, ws_bytes, stream);
}

namespace vlmc {
// Waves of the batched Wanda launch.  One resident wave (the single-tensor kernel's choice) gives every tensor a thin slice
// of the GPU for the whole launch; with the grid ordered (column tile, TENSOR, row chunk) and row chunks of ~1024 rows the
// CTAs of all tensors sweep the same rows at the same time: DRAM sees one sequential stream per distinct tensor and the
// tensors that ARE the same activations (q / k / v, gate / up) hit in L2.  Measured on B200 (Vicuna block, 7 tensors, 18.66 GB
// algorithmic): 1 wave 2.44 ms, 8 waves 1.92, 16: 1.85, 32: 1.80, 64: 1.83 ms; 4 distinct tensors (12.2 GB): 1.92 -> 1.79 ms.
// (The single-tensor launch is the other way round: 6.6 TB/s at one wave, 4.8 at 32.)  VLMC_STATS_BATCH_WAVES overrides.
static int stats_batch_waves(const vlmc_stats_item* items, int count, int dtype) {
  if (const char* e = getenv("VLMC_STATS_BATCH_WAVES")) { const int v = atoi(e); if (v >= 1 && v <= 64) return v; }
  const int V = dtype == VLMC_F32 ? 4 : 8;
  double total = 0.0;
  for (int i = 0; i < count; ++i) total += (double)items[i].T * (double)items[i].C;
  if (total <= 0.0) return 1;
  const double wave = (double)kNumSMs * StatsOcc<false>::kBlocksPerSM;
  // rows per chunk of the first tensor at one wave
  const double coltiles = (double)((items[0].C + StatsTile<false>::kCX * V - 1) / (StatsTile<false>::kCX * V));
  double chunks = wave * ((double)items[0].T * (double)items[0].C / total) / coltiles;
  if (chunks < 1.0) chunks = 1.0;
  const double rows = (double)items[0].T / chunks;
  int waves = (int)(rows / 1024.0 + 0.5);
  return waves < 1 ? 1 : (waves > 64 ? 64 : waves);
}
}  // namespace vlmc

extern "C" size_t vlmc_sqnorm_accum_batch_workspace_bytes(const vlmc_stats_item* items, int count, int dtype) {
  using namespace vlmc;
  if (!items || count < 1 || count > kStatsBatchMax) return 0;
  size_t total = 0;
  const size_t waves = (size_t)stats_batch_waves(items, count, dtype);
  for (int i = 0; i < count; ++i) total += align_up(stats_workspace_bytes(0, items[i].T, items[i].C, 1) * waves, 256);
  return total;
}

extern "C" int vlmc_sqnorm_accum_batch(const vlmc_stats_item* items, int count, int dtype, void* ws, size_t ws_bytes,
                                       void* stream) {
  using namespace vlmc;
  if (!items || !ws || count < 1 || count > kStatsBatchMax) return VLMC_ERR_BAD_ARG;
  if (dtype != VLMC_F32 && dtype != VLMC_F16 && dtype != VLMC_BF16) return VLMC_ERR_BAD_ARG;
  if (!is_device_ptr(ws)) return VLMC_ERR_NOT_DEVICE;
  const int V = dtype == VLMC_F32 ? 4 : 8;
  double total_elems = 0.0;
  for (int i = 0; i < count; ++i) {
    const vlmc_stats_item& s = items[i];
    if (!s.x || !s.scaler_row || s.T < 1 || s.C < 1 || s.ldx < s.C || s.b < 0 || s.n_before < 0 || s.n_before + s.b <= 0) return VLMC_ERR_BAD_ARG;
    if (s.C % V != 0 || s.ldx % V != 0 || ((uintptr_t)s.x & 15) != 0) return VLMC_ERR_UNSUPPORTED;
    if (!is_device_ptr(s.x) || !is_device_ptr(s.scaler_row)) return VLMC_ERR_NOT_DEVICE;
    total_elems += (double)s.T * (double)s.C;
  }
  StatsBatch b;
  char* base = reinterpret_cast<char*>(ws);
  size_t used = 0;
  int max_ct = 1, max_ch = 1;
  const int waves = stats_batch_waves(items, count, dtype);
  const int64_t wave = (int64_t)kNumSMs * StatsOcc<false>::kBlocksPerSM * waves;
  for (int i = 0; i < count; ++i) {
    const vlmc_stats_item& s = items[i];
    // every item gets the share of ONE resident wave that its bytes have in the launch
    int64_t target = (int64_t)((double)wave * ((double)s.T * (double)s.C / total_elems));
    if (target < 1) target = 1;
    StatsPlan pl = plan_stats(dtype, 1, s.T, s.C, 1, StatsOcc<false>::kBlocksPerSM, target);
    if (pl.coltiles * sizeof(unsigned int) > VLMC_WS_COUNTER_BYTES || pl.nchunks > 65535) return VLMC_ERR_UNSUPPORTED;
    const size_t need = align_up(pl.bytes, 256);
    if (used + need > ws_bytes) return VLMC_ERR_WORKSPACE;
    StatsParams& p = b.it[i];
    p.x = s.x; p.ldx = s.ldx; p.C = s.C; p.S = s.T; p.nseg = 1;
    p.chunks_per_seg = pl.chunks_per_seg; p.rows_per_chunk = pl.rows_per_chunk;
    p.tickets = reinterpret_cast<unsigned int*>(base + used);
    // this item's ticket words sit where an earlier (single) launch kept partial sums: clear them
    if (cudaMemsetAsync(base + used, 0, VLMC_WS_COUNTER_BYTES, (cudaStream_t)stream) != cudaSuccess) return check_launch();
    p.part = reinterpret_cast<float*>(base + used + VLMC_WS_COUNTER_BYTES);
    p.scaler_row = s.scaler_row; p.sum_row = nullptr; p.mean = nullptr; p.var = nullptr;
    p.n_before = s.n_before; p.b_per_seg = s.b; p.ntok_before = 0.0;
    b.coltiles[i] = pl.coltiles;
    b.nchunks[i] = (int)pl.nchunks;
    max_ct = pl.coltiles > max_ct ? pl.coltiles : max_ct;
    max_ch = (int)pl.nchunks > max_ch ? (int)pl.nchunks : max_ch;
    used += need;
  }
  b.chunk_major = waves > 1 ? 1 : 0;      // several waves: linears interleaved per row chunk (see colstats_batch_kernel)
  dim3 grid(max_ct, waves > 1 ? count : max_ch, waves > 1 ? max_ch : count);
  cudaStream_t st = (cudaStream_t)stream;
  VLMC_DISPATCH_DTYPE(dtype, (colstats_batch_kernel<scalar_t, false, StatsTile<false>::kCX><<<grid, kStatsThreads, 0, st>>>(b)));
  return check_launch();
}

extern "C" size_t vlmc_dsnot_stats_batch_workspace_bytes(const vlmc_dsnot_stats_item* items, int count, int dtype) {
  using namespace vlmc;
  (void)dtype;
  if (!items || count < 1 || count > kStatsBatchMax) return 0;
  size_t total = 0;
  for (int i = 0; i < count; ++i)
    total += align_up(stats_workspace_bytes(1, items[i].nseg * items[i].S, items[i].C, items[i].nseg), 256);
  return total;
}

extern "C" int vlmc_dsnot_stats_batch(const vlmc_dsnot_stats_item* items, int count, int dtype, void* ws, size_t ws_bytes,
                                      void* stream) {
  using namespace vlmc;
  if (!items || !ws || count < 1 || count > kStatsBatchMax) return VLMC_ERR_BAD_ARG;
  if (dtype != VLMC_F32 && dtype != VLMC_F16 && dtype != VLMC_BF16) return VLMC_ERR_BAD_ARG;
  if (!is_device_ptr(ws)) return VLMC_ERR_NOT_DEVICE;
  const int V = dtype == VLMC_F32 ? 4 : 8;
  StatsBatch b;
  char* base = reinterpret_cast<char*>(ws);
  size_t used = 0;
  int max_ct = 1, max_ch = 1;
  for (int i = 0; i < count; ++i) {
    const vlmc_dsnot_stats_item& s = items[i];
    if (!s.x || !s.scaler_row || !s.sum_row || !s.mean || !s.var || s.nseg < 1 || s.S < 1 || s.C < 1 || s.ldx < s.C)
      return VLMC_ERR_BAD_ARG;
    if (s.C % V != 0 || s.ldx % V != 0 || ((uintptr_t)s.x & 15) != 0) return VLMC_ERR_UNSUPPORTED;
    if (!is_device_ptr(s.x) || !is_device_ptr(s.scaler_row)) return VLMC_ERR_NOT_DEVICE;
    // the plan of the single-tensor call: identical partial sums, identical results
    StatsPlan pl = plan_stats(dtype, s.nseg, s.S, s.C, 3, StatsOcc<true>::kBlocksPerSM);
    if (pl.coltiles * sizeof(unsigned int) > VLMC_WS_COUNTER_BYTES || pl.nchunks > 65535) return VLMC_ERR_UNSUPPORTED;
    const size_t need = align_up(pl.bytes, 256);
    if (used + need > ws_bytes) return VLMC_ERR_WORKSPACE;
    StatsParams& p = b.it[i];
    p.x = s.x; p.ldx = s.ldx; p.C = s.C; p.S = s.S; p.nseg = s.nseg;
    p.chunks_per_seg = pl.chunks_per_seg; p.rows_per_chunk = pl.rows_per_chunk;
    p.tickets = reinterpret_cast<unsigned int*>(base + used);
    if (cudaMemsetAsync(base + used, 0, VLMC_WS_COUNTER_BYTES, (cudaStream_t)stream) != cudaSuccess) return check_launch();
    p.part = reinterpret_cast<float*>(base + used + VLMC_WS_COUNTER_BYTES);
    p.scaler_row = s.scaler_row; p.sum_row = s.sum_row; p.mean = s.mean; p.var = s.var;
    p.n_before = s.n_before; p.b_per_seg = s.b_per_seg; p.ntok_before = s.ntok_before;
    b.coltiles[i] = pl.coltiles;
    b.nchunks[i] = (int)pl.nchunks;
    max_ct = pl.coltiles > max_ct ? pl.coltiles : max_ct;
    max_ch = (int)pl.nchunks > max_ch ? (int)pl.nchunks : max_ch;
    used += need;
  }
  b.chunk_major = 1;
  dim3 grid(max_ct, count, max_ch);
  cudaStream_t st = (cudaStream_t)stream;
  VLMC_DISPATCH_DTYPE(dtype, (colstats_batch_kernel<scalar_t, true, StatsTile<true>::kCX><<<grid, kStatsThreads, 0, st>>>(b)));
  return check_launch();
}

extern "C" int vlmc_dsnot_stats(const void* x, int dtype, int64_t nseg, int64_t S, int C, int64_t ldx,
                                float* scaler_row, float* sum_row, float* mean, float* var,
                                double n_before, double b_per_seg, double ntok_before,
                                void* ws, size_t ws_bytes, void* stream) {
  return vlmc::launch_stats<true>(x, dtype, nseg, S, C, ldx, scaler_row, sum_row, mean, var,
                                  n_before, b_per_seg, ntok_before, ws, ws_bytes, stream);
}
