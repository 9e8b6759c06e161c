// K3: SparseGPT Hessian accumulation  H <- H * n/(n+b) + (2/(n+b)) * X^T X  on the 5th-gen tensor cores.
//
// Replaces SparseGPT.add_batch, lavis/compression/pruners/sparsegpt_pruner.py:68-79
//   self.H *= nsamples / (nsamples + tmp); inp = sqrt(2/nsamples) * inp.float(); self.H += inp.matmul(inp.t())
// (cuBLAS SGEMM in true fp32 + a full read-modify-write of H per call in the reference).
//
// X is [T, C] row-major fp16 / bf16 (every LLM linear's input).  fp16/bf16 values are exact
// tensor-core operands and their products are exact in fp32, so ONE kind::f16 pass has SGEMM-grade accuracy
// provided the fp32 accumulation is done carefully: the tensor core accumulates a chunk of `kc` tokens in
// TMEM, the epilogue warps add the chunk results into fp32 REGISTERS with round-to-nearest adds (the
// in-TMEM accumulation rounds toward zero; short chunks keep that bias below 1e-6 relative).
//
// D[M=C, N=C] = A B with A = X^T (M-major in smem) and B = X (N-major in smem): both operands are the same
// TMA boxes of X ({64 columns, 64 tokens}, 128-byte swizzle), no transpose is ever materialised.
// SYRK: only tiles that touch the upper triangle are computed (128 x 256 tiles, bj >= bi/2); the epilogue
// writes each tile and, for tiles strictly above the diagonal band, its mirror image, so H stays exactly
// symmetric and full, as SparseGPT.fasterprune expects.
//
// Warp roles (384 threads, 1 CTA / SM, persistent over tiles):
//   warp 0      TMA producer     4-stage ring of {A 128x64, B 256x64} tiles (48 KB / stage)
//   warp 1      MMA issuer       tcgen05.mma.cta_group::1.kind::f16, M=128 N=256 K=16, fp32 accumulate in TMEM
//   warp 2      TMEM allocator   512 columns = 2 accumulator buffers (chunk n+1 overlaps the drain of chunk n)
//   warps 4-11  epilogue         tcgen05.ld -> fp32 register accumulators -> fused H update (+ mirror)
#include <cstdlib>
#include <stdlib.h>
#include "tc.cuh"
#include "gemm3x.cuh"

namespace vlmc {

constexpr int kHM = 128, kHN = 256, kHK = 64;        // CTA tile, tokens per stage
constexpr int kHStages = 4;
constexpr int kHBox = 64;                             // columns per TMA box (128 B of 16-bit)
constexpr uint32_t kBoxBytes = kHBox * kHK * 2;       // 8 KB
constexpr uint32_t kABytes = (kHM / kHBox) * kBoxBytes;   // 16 KB
constexpr uint32_t kBBytes = (kHN / kHBox) * kBoxBytes;   // 32 KB
constexpr uint32_t kStageBytes = kABytes + kBBytes;       // 48 KB
constexpr int kHThreads = 384;
constexpr int kEpiWarp0 = 4, kEpiWarps = 8;

struct HessParams {
  float* H;
  int64_t ldh;
  int C;
  int64_t T;
  int kc;            // tokens accumulated in TMEM before the RN flush into registers
  int nbi, nbj, ntiles;
  float ratio, scale;
  int beta_zero;     // n_before == 0: H is overwritten (no read)
};

// 64-bit shared-memory matrix descriptor, MN-major operand, 128-byte swizzle (cute::UMMA::SmemDescriptor):
//   [0,14) start >> 4 | [16,30) LBO >> 4 (stride between 64-element MN atoms = one TMA box)
//   [32,46) SBO >> 4 (stride between 8-token groups = 1024 B) | [46,48) version = 1 | [61,64) layout = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);
  d |= (uint64_t)(kBoxBytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void tile_from_index(int t, int nbi, int nbj, int& bi, int& bj) {
  for (bi = 0; bi < nbi; ++bi) {
    const int cnt = nbj - (bi >> 1);
    if (t < cnt) { bj = (bi >> 1) + t; return; }
    t -= cnt;
  }
  bj = nbj;  // unreachable for t < ntiles
}

struct __align__(8) HessBarriers {
  uint64_t full[kHStages], empty[kHStages], tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kHThreads, 1)
hessian_syrk_kernel(const __grid_constant__ CUtensorMap xmap, const HessParams p, const uint32_t idesc) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128-byte swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  HessBarriers* bars = reinterpret_cast<HessBarriers*>(smem + kHStages * kStageBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(&xmap) : "memory");
    for (int s = 0; s < kHStages; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&bars->tmem_full[s], 1); mbar_init(&bars->tmem_empty[s], kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(&bars->tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  const int nchunks = (int)((p.T + p.kc - 1) / p.kc);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        int bi, bj;
        tile_from_index(t, p.nbi, p.nbj, bi, bj);
        const int i0 = bi * kHM, j0 = bj * kHN;
        for (int64_t k0 = 0; k0 < p.T; k0 += kHK) {
          mbar_wait(&bars->empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          mbar_expect_tx(&bars->full[stage], kStageBytes);
#pragma unroll
          for (int q = 0; q < kHM / kHBox; ++q) tma_load_2d(sa + q * kBoxBytes, &xmap, &bars->full[stage], i0 + q * kHBox, (int)k0);
#pragma unroll
          for (int q = 0; q < kHN / kHBox; ++q) tma_load_2d(sb + q * kBoxBytes, &xmap, &bars->full[stage], j0 + q * kHBox, (int)k0);
          if (++stage == kHStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        for (int ch = 0; ch < nchunks; ++ch) {
          const int64_t kbeg = (int64_t)ch * p.kc;
          int64_t kend = kbeg + p.kc;
          if (kend > p.T) kend = p.T;
          mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * kHN;
          uint32_t accumulate = 0;
          for (int64_t k0 = kbeg; k0 < kend; k0 += kHK) {
            mbar_wait(&bars->full[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * kStageBytes);
            const uint32_t sb = sa + kABytes;
#pragma unroll
            for (int kk = 0; kk < kHK / 16; ++kk) {
              // 16 tokens = 2 swizzle atoms of 8 rows x 128 B: advance the start address by 2 KB
              const uint64_t adesc = make_desc_mn_sw128(sa + kk * 2048);
              const uint64_t bdesc = make_desc_mn_sw128(sb + kk * 2048);
              tc_mma_f16(d_tmem, adesc, bdesc, idesc, accumulate);
              accumulate = 1;
            }
            tc_commit(&bars->empty[stage]);          // frees the smem slot when these MMAs retire
            if (++stage == kHStages) { stage = 0; phase ^= 1; }
          }
          tc_commit(&bars->tmem_full[acc]);          // chunk accumulator complete
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===== epilogue: TMEM -> fp32 registers (RN adds across chunks) -> H =====
    const int ew = warp - kEpiWarp0;
    const int q = warp & 3;              // TMEM lane quarter this warp may touch
    const int half = ew >> 2;            // which 128 of the tile's 256 columns
    const int m = q * 32 + lane;         // tile row
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
      int bi, bj;
      tile_from_index(t, p.nbi, p.nbj, bi, bj);
      float sum[128];
#pragma unroll
      for (int c = 0; c < 128; ++c) sum[c] = 0.f;
      for (int ch = 0; ch < nchunks; ++ch) {
        mbar_wait(&bars->tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kHN + half * 128;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          uint32_t v[32];
          tc_ld32(taddr + c4 * 32, v);
          tc_wait_ld();
#pragma unroll
          for (int c = 0; c < 32; ++c) sum[c4 * 32 + c] += __uint_as_float(v[c]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      // fused running-average update, direct tile
      const int i = bi * kHM + m;
      const int jbase = bj * kHN + half * 128;
      if (i < p.C) {
        float* hrow = p.H + (int64_t)i * p.ldh + jbase;
#pragma unroll
        for (int c = 0; c < 128; c += 4) {
          if (jbase + c < p.C) {      // C % 4 == 0: a float4 is entirely in or out
            float4 o;
            if (p.beta_zero) {
              o.x = sum[c] * p.scale; o.y = sum[c + 1] * p.scale; o.z = sum[c + 2] * p.scale; o.w = sum[c + 3] * p.scale;
            } else {
              const float4 h = *reinterpret_cast<const float4*>(hrow + c);
              o.x = fmaf(sum[c], p.scale, h.x * p.ratio);     o.y = fmaf(sum[c + 1], p.scale, h.y * p.ratio);
              o.z = fmaf(sum[c + 2], p.scale, h.z * p.ratio); o.w = fmaf(sum[c + 3], p.scale, h.w * p.ratio);
            }
            *reinterpret_cast<float4*>(hrow + c) = o;
            sum[c] = o.x; sum[c + 1] = o.y; sum[c + 2] = o.z; sum[c + 3] = o.w;
          }
        }
        // mirror image for tiles strictly above the diagonal band (their transposes are never computed)
        if (bj > (bi >> 1)) {
#pragma unroll
          for (int c = 0; c < 128; ++c) {
            const int j = jbase + c;
            if (j < p.C) p.H[(int64_t)j * p.ldh + i] = sum[c];   // lanes = consecutive i: coalesced
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- CTA-pair variant -----------------------------------------------------------------------------------
// The single-CTA kernel above is bound by operand traffic, not by the tensor pipe: every 128 x 256 x 64 MMA step
// needs 48 KB of X from L2 (87 flop / B; 14 TB/s of L2 -> SM traffic at the measured rate).  Here two CTAs of a
// cluster (a TPC pair) share one 256 x 256 tile: each loads its own 128 rows of A and only HALF of B (128 of the 256
// columns), `tcgen05.mma.cta_group::2` (M = 256) reads the other half from the partner's shared memory.  32 KB per
// CTA and step for the same MMA work: 131 flop / B.
//   full[s]        lives in the LEADER (rank 0): one arrival (the leader's expect_tx for both CTAs' bytes), both CTAs' TMA bytes land on it
//   empty[s]       one per CTA, released by the leader's tcgen05.commit multicast to both CTAs
//   tmem_full[a]   one per CTA (each CTA's epilogue drains its own 128 accumulator rows), commit multicast
//   tmem_empty[a]  leader only: 2 x 8 epilogue warps arrive (the partner's remotely)
constexpr int kH2Stages = 7;
constexpr uint32_t kH2ABytes = (128 / kHBox) * kBoxBytes;     // 16 KB: this CTA's 128 rows of the tile
constexpr uint32_t kH2BBytes = (128 / kHBox) * kBoxBytes;     // 16 KB: this CTA's half of the tile's 256 columns
constexpr uint32_t kH2StageBytes = kH2ABytes + kH2BBytes;     // 32 KB
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;                   // clears the CTA-rank bit of a shared::cluster address

struct __align__(8) Hess2Barriers {
  uint64_t full[kH2Stages], empty[kH2Stages], tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive (+ expected bytes) on a barrier of the LEADER CTA, from either CTA of the pair
__device__ __forceinline__ void mbar_expect_tx_leader(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;"
               :: "r"(smem_u32(bar) & kPeerMask), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(smem_u32(bar) & kPeerMask) : "memory");
}
// TMA load into THIS CTA's shared memory whose bytes are accounted on the leader's barrier
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :: "r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               :: "r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// tile pair t -> (bi2, bj) over the upper block triangle of 256 x 256 blocks, bj >= bi2, row by row.
// (A banded order - 8 block rows, column by column, so that concurrent tiles share more columns of X - measured 9 % slower.)
__device__ __forceinline__ void tile2_from_index(int t, int nb2, int& bi2, int& bj) {
  for (bi2 = 0; bi2 < nb2; ++bi2) {
    const int cnt = nb2 - bi2;
    if (t < cnt) { bj = bi2 + t; return; }
    t -= cnt;
  }
  bj = nb2;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kHThreads, 1)
hessian_syrk2_kernel(const __grid_constant__ CUtensorMap xmap, const HessParams p, const uint32_t idesc) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Hess2Barriers* bars = reinterpret_cast<Hess2Barriers*>(smem + kH2Stages * kH2StageBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(&xmap) : "memory");
    for (int s = 0; s < kH2Stages; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&bars->tmem_full[s], 1); mbar_init(&bars->tmem_empty[s], 2 * kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(&bars->tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // both CTAs' barriers are initialised before any remote arrive / TMA
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const int nchunks = (int)((p.T + p.kc - 1) / p.kc);

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own 128 A rows + own 128 B columns =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = cluster_id; t < p.ntiles; t += nclusters) {
        int bi2, bj;
        tile2_from_index(t, p.nbi, bi2, bj);
        const int i0 = bi2 * 256 + (int)rank * 128, j0 = bj * 256 + (int)rank * 128;
        for (int64_t k0 = 0; k0 < p.T; k0 += kHK) {
          mbar_wait(&bars->empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kH2StageBytes;
          uint8_t* sb = sa + kH2ABytes;
          // the leader alone announces the bytes of BOTH CTAs (one local arrive; the partner's loads only complete_tx)
          if (rank == 0) mbar_expect_tx(&bars->full[stage], 2 * kH2StageBytes);
#pragma unroll
          for (int q = 0; q < 128 / kHBox; ++q) tma_load_2d_pair(sa + q * kBoxBytes, &xmap, &bars->full[stage], i0 + q * kHBox, (int)k0);
#pragma unroll
          for (int q = 0; q < 128 / kHBox; ++q) tma_load_2d_pair(sb + q * kBoxBytes, &xmap, &bars->full[stage], j0 + q * kHBox, (int)k0);
          if (++stage == kH2Stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread of the LEADER CTA drives both tensor cores =====
    if (lane == 0 && rank == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int t = cluster_id; t < p.ntiles; t += nclusters) {
        for (int ch = 0; ch < nchunks; ++ch) {
          const int64_t kbeg = (int64_t)ch * p.kc;
          int64_t kend = kbeg + p.kc;
          if (kend > p.T) kend = p.T;
          mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * kHN;
          uint32_t accumulate = 0;
          for (int64_t k0 = kbeg; k0 < kend; k0 += kHK) {
            mbar_wait(&bars->full[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * kH2StageBytes);
            const uint32_t sb = sa + kH2ABytes;
#pragma unroll
            for (int kk = 0; kk < kHK / 16; ++kk) {
              tc_mma_f16_pair(d_tmem, make_desc_mn_sw128(sa + kk * 2048), make_desc_mn_sw128(sb + kk * 2048), idesc, accumulate);
              accumulate = 1;
            }
            tc_commit_pair(&bars->empty[stage]);       // frees this smem slot in BOTH CTAs when the MMAs retire
            if (++stage == kH2Stages) { stage = 0; phase ^= 1; }
          }
          tc_commit_pair(&bars->tmem_full[acc]);       // chunk accumulator complete, in both CTAs
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===== epilogue (both CTAs): own 128 accumulator rows x 256 columns =====
    const int ew = warp - kEpiWarp0;
    const int q = warp & 3;
    const int half = ew >> 2;
    const int m = q * 32 + lane;
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = cluster_id; t < p.ntiles; t += nclusters) {
      int bi2, bj;
      tile2_from_index(t, p.nbi, bi2, bj);
      float sum[128];
#pragma unroll
      for (int c = 0; c < 128; ++c) sum[c] = 0.f;
      for (int ch = 0; ch < nchunks; ++ch) {
        mbar_wait(&bars->tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kHN + half * 128;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          uint32_t v[32];
          tc_ld32(taddr + c4 * 32, v);
          tc_wait_ld();
#pragma unroll
          for (int c = 0; c < 32; ++c) sum[c4 * 32 + c] += __uint_as_float(v[c]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&bars->tmem_empty[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      const int i = bi2 * 256 + (int)rank * 128 + m;
      const int jbase = bj * 256 + half * 128;
      if (i < p.C) {
        float* hrow = p.H + (int64_t)i * p.ldh + jbase;
#pragma unroll
        for (int c = 0; c < 128; c += 4) {
          if (jbase + c < p.C) {
            float4 o;
            if (p.beta_zero) {
              o.x = sum[c] * p.scale; o.y = sum[c + 1] * p.scale; o.z = sum[c + 2] * p.scale; o.w = sum[c + 3] * p.scale;
            } else {
              const float4 h = *reinterpret_cast<const float4*>(hrow + c);
              o.x = fmaf(sum[c], p.scale, h.x * p.ratio);     o.y = fmaf(sum[c + 1], p.scale, h.y * p.ratio);
              o.z = fmaf(sum[c + 2], p.scale, h.z * p.ratio); o.w = fmaf(sum[c + 3], p.scale, h.w * p.ratio);
            }
            *reinterpret_cast<float4*>(hrow + c) = o;
            sum[c] = o.x; sum[c + 1] = o.y; sum[c + 2] = o.z; sum[c + 3] = o.w;
          }
        }
        if (bj > bi2) {                // off-diagonal block: write the transpose (diagonal blocks are computed in full)
#pragma unroll
          for (int c = 0; c < 128; ++c) {
            const int j = jbase + c;
            if (j < p.C) p.H[(int64_t)j * p.ldh + i] = sum[c];
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // the partner may still be reading this CTA's shared memory / signalling its barriers
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
  }
}

constexpr size_t kHess2Smem = kH2Stages * kH2StageBytes + sizeof(Hess2Barriers) + 1024;

// ---- host side -----------------------------------------------------------------------------------------
constexpr size_t kHessSmem = kHStages * kStageBytes + sizeof(HessBarriers) + 1024;

}  // namespace vlmc

extern "C" int vlmc_hessian_accum(const void* x, int dtype, int64_t T, int C, int64_t ldx,
                                  float* H, int64_t ldh, double n_before, double b, int kc, int64_t slab_tokens,
                                  void* stream) {
  using namespace vlmc;
  if (!x || !H || T < 1 || C < 1 || ldx < C || ldh < C || b < 0 || n_before < 0 || n_before + b <= 0) return VLMC_ERR_BAD_ARG;   // b == 0 with n_before == N: plain += (2/N) X^T X (token-sharded accumulation)
  if (dtype == VLMC_F32) {
    // fp32 activations (EVA-ViT qkv / fc1 inputs): fp32 values are not exact tensor-core operands, so the
    // contraction runs as the 3xTF32 split GEMM  H = ratio * H + scale * X^T X  (gemm3x.cu, A and B both = X, [K,M] / [K,N]).
    if (C % 4 != 0 || ldx % 4 != 0 || ldh % 4 != 0 || ((uintptr_t)x & 15) != 0 || ((uintptr_t)H & 15) != 0)
      return VLMC_ERR_UNSUPPORTED;
    if (T > 0x7fffffff) return VLMC_ERR_UNSUPPORTED;
    if (!is_device_ptr(x) || !is_device_ptr(H)) return VLMC_ERR_NOT_DEVICE;
    const double n_after32 = n_before + b;
    const float* xf = reinterpret_cast<const float*>(x);
    return gemm3x(false, C, C, (int)T, (float)(2.0 / n_after32), xf, ldx, xf, ldx, (float)(n_before / n_after32), H, ldh, 0,
                  kc > 0 ? (kc < 256 ? kc : 256) : 128, (cudaStream_t)stream, 0, true);
  }
  if (dtype != VLMC_F16 && dtype != VLMC_BF16) return VLMC_ERR_BAD_ARG;
  if (C % 8 != 0 || ldx % 8 != 0 || ldh % 4 != 0 || ((uintptr_t)x & 15) != 0 || ((uintptr_t)H & 15) != 0)
    return VLMC_ERR_UNSUPPORTED;
  if (T > 0x7fffffff) return VLMC_ERR_UNSUPPORTED;
  if (!is_device_ptr(x) || !is_device_ptr(H)) return VLMC_ERR_NOT_DEVICE;
  if (kc <= 0) kc = 512;
  kc = (kc + kHK - 1) / kHK * kHK;

  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return VLMC_ERR_CUDA;

  HessParams p;
  p.H = H; p.ldh = ldh; p.C = C; p.kc = kc;
  p.nbi = (C + kHM - 1) / kHM;
  p.nbj = (C + kHN - 1) / kHN;
  p.ntiles = 0;
  for (int bi = 0; bi < p.nbi; ++bi) p.ntiles += p.nbj - (bi >> 1);
  const double n_after = n_before + b;
  p.scale = (float)(2.0 / n_after);

  // instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, A and B MN-major, N=256, M=128
  const uint32_t fmt = dtype == VLMC_F16 ? 0u : 1u;
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) |
                         ((uint32_t)(kHN >> 3) << 17) | ((uint32_t)(kHM >> 4) << 24);

  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(hessian_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHessSmem) != cudaSuccess)
      return check_launch();
    attr_set = true;
  }
  // CTA-pair kernel (256 x 256 tiles on a 2-CTA cluster) by default; VLMC_HESS_2CTA=0 selects the single-CTA kernel
  // (kept for A/B measurements: scripts/hessian_pair_probe.py)
  static const bool use_pair = [] { const char* e = getenv("VLMC_HESS_2CTA"); return !e || atoi(e) != 0; }();
  const uint32_t idesc2 = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) |
                          ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
  if (use_pair) {
    static bool attr2_set = false;
    if (!attr2_set) {
      if (cudaFuncSetAttribute(hessian_syrk2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHess2Smem) != cudaSuccess)
        return check_launch();
      attr2_set = true;
    }
    const int nb2 = (C + 255) / 256;
    p.nbi = nb2; p.nbj = nb2;
    p.ntiles = nb2 * (nb2 + 1) / 2;
  }
  // VLMC_HESS_MAX_SMS (read per call): SMs this accumulation may occupy.  A caller that runs latency-bound chains on other
  // streams at the same time (the overlapped SparseGPT block schedule) leaves them a few SMs this way: the persistent grid
  // would otherwise hold every SM until the launch ends.
  int sm_cap = kNumSMs;
  if (const char* e = getenv("VLMC_HESS_MAX_SMS")) { const int v = atoi(e); if (v >= 2 && v < kNumSMs) sm_cap = v; }
  const int grid = use_pair ? 2 * (p.ntiles < sm_cap / 2 ? p.ntiles : sm_cap / 2) : (p.ntiles < sm_cap ? p.ntiles : sm_cap);

  // Long calibration sets are processed in slabs of tokens, all tiles per slab: the CTAs of a wave then read the
  // same slab of X at about the same time and the operand re-reads (each column block feeds ~C/256 tiles) hit in
  // L2 instead of HBM.  The extra cost is one read-modify-write of H per slab.
  int64_t slab = slab_tokens > 0 ? slab_tokens : (C >= 8192 ? 65536 : ((int64_t)1 << 30));
  slab = (slab + kc - 1) / kc * kc;
  for (int64_t t0 = 0; t0 < T; t0 += slab) {
    const int64_t tn = (T - t0 < slab) ? (T - t0) : slab;
    CUtensorMap map;
    const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)tn};
    const cuuint64_t strides[1] = {(cuuint64_t)ldx * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kHBox, (cuuint32_t)kHK};
    const cuuint32_t estr[2] = {1, 1};
    const char* base = reinterpret_cast<const char*>(x) + (size_t)t0 * ldx * 2;
    CUresult cr = encode(&map, dtype == VLMC_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                         2, const_cast<char*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { g_last_cuda_error = (int)cr; return VLMC_ERR_CUDA; }
    p.T = tn;
    if (t0 == 0) {
      p.ratio = (float)(n_before / n_after);
      p.beta_zero = n_before == 0.0 ? 1 : 0;
    } else {
      p.ratio = 1.0f;
      p.beta_zero = 0;
    }
    if (use_pair) hessian_syrk2_kernel<<<grid, kHThreads, kHess2Smem, (cudaStream_t)stream>>>(map, p, idesc2);
    else hessian_syrk_kernel<<<grid, kHThreads, kHessSmem, (cudaStream_t)stream>>>(map, p, idesc);
    int rc = check_launch();
    if (rc) return rc;
  }
  return VLMC_OK;
}
