// fp32-accurate GEMM on tcgen05 (3xTF32 split): the tensor-core building block of the SparseGPT factorisation
// (K10: panel solves, symmetric trailing updates, triangular-inverse merges) and of the lazy OBS update (K13).
//
// Replaces the fp32 cuBLAS / cuSOLVER GEMMs behind lavis/compression/pruners/sparsegpt_pruner.py:114-157 and :210
//   W[:, i2:] -= Err1.matmul(Hinv[i1:i2, i2:])
// which torch runs in true fp32 (allow_tf32 = False).  A single TF32 pass keeps 11 mantissa bits (8.9e-6 max-norm
// error); the split a = hi + lo with hi = tf32(a), lo = a - hi and the three products
// lo*hi + hi*lo + hi*hi recover fp32-grade accuracy (4.5e-8 measured in the survey's emulation).
//
// C[M,N] = beta * C + alpha * A[M,K] * op(B); A, B, C are plain fp32 row-major arrays in HBM.  The split is done
// ON CHIP: TMA brings raw fp32 k-slices (128-byte swizzle) into shared memory, two converter warps rewrite each
// slice in place as `hi` and write `lo` next to it (element-wise, so the swizzle is irrelevant to them), then one
// thread issues 3 tcgen05.mma.kind::tf32 per 8-wide k-step into a TMEM accumulator.  Operands therefore cross
// L2 -> SM once, un-split.
//
// Warp roles (384 threads, 1 CTA / SM, persistent over output tiles):
//   warp 0        TMA producer    ring of {A 128x32, B TNx32} fp32 slices
//   warp 1        TMEM allocator + MMA issuer (one thread)
//   warps 2-3     converters      raw -> (hi, lo), fence.proxy.async, arrive
//   warps 4-11    epilogue        tcgen05.ld -> alpha * acc + beta * C -> global (coalesced through a staging buffer)
// Long K: the in-TMEM accumulation rounds toward zero, so K is cut into chunks of `kc` (default 128) whose partial
// sums are added round-to-nearest in fp32 registers by the epilogue warps (double-buffered accumulators), exactly
// like the Hessian kernel (hessian.cu).
#include "tc.cuh"
#include "gemm3x.cuh"

namespace vlmc {

constexpr int kG3M = 128;                 // tile rows (UMMA M)
constexpr int kG3K = 32;                  // k-slice per stage: 32 fp32 = one 128-byte swizzle row
constexpr int kG3Threads = 12 * 32;
constexpr int kG3CvtWarp0 = 2, kG3CvtWarps = 2, kG3EpiWarp0 = 4, kG3EpiWarps = 8;
constexpr uint32_t kG3ABytes = kG3M * kG3K * 4;    // 16 KB

template <int TN> struct G3Cfg {
  static constexpr uint32_t kBBytes = TN * kG3K * 4;               // 32 KB (TN = 256) / 16 KB (TN = 128)
  static constexpr uint32_t kStageBytes = 2 * (kG3ABytes + kBBytes);
  static constexpr int kStages = TN == 256 ? 2 : 3;
};

constexpr int kG3MaxStages = 3;
struct __align__(8) G3Barriers {
  uint64_t full[kG3MaxStages], conv[kG3MaxStages], empty[kG3MaxStages], tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
};

struct G3Params {
  float* C;
  int64_t ldc;
  int M, N, K;
  float alpha, beta;
  int tri, kc;
  int kmode;         // 0 dense; 1: B[k][j] == 0 for k < j (lower-triangular [K,N]); 2: A[i][k] == 0 for k > i
  int nbi, nbj, ntiles;
};

// K-major operand, 128-byte swizzle: rows are 128 B apart, 8-row groups 1024 B apart (SBO); LBO is unused
__device__ __forceinline__ uint64_t g3_desc_k(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// MN-major fp32 operand: the only legal layout is SWIZZLE_128B_BASE32B (32-byte chunks swizzled inside 128-byte rows,
// TMA's CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): 32 consecutive n per 128-byte row, atoms of 4 k-rows = 512 B
// (SBO between k groups of 4), LBO = stride between 32-wide n atoms = one TMA box of 32 k-rows
__device__ __forceinline__ uint64_t g3_desc_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);
  d |= (uint64_t)((kG3K * 128) >> 4) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}

template <int TN>
__device__ __forceinline__ void g3_tile(int t, const G3Params& p, int& bi, int& bj) {
  if (!p.tri) { bi = t / p.nbj; bj = t - bi * p.nbj; return; }
  for (bi = 0; bi < p.nbi; ++bi) {
    int cnt = (bi * kG3M + kG3M - 1) / TN + 1;
    if (cnt > p.nbj) cnt = p.nbj;
    if (t < cnt) { bj = t; return; }
    t -= cnt;
  }
  bj = 0;  // unreachable for t < ntiles
}

// k-slices [sb, se) of tile (bi, bj) that can be non-zero (structural zeros of a triangular operand are skipped)
template <int TN>
__device__ __forceinline__ void g3_krange(const G3Params& p, int bi, int bj, int nslices, int& sb, int& se) {
  sb = 0; se = nslices;
  if (p.kmode == 1) {
    sb = (bj * TN) / kG3K;
  } else if (p.kmode == 2) {
    const int e = (bi * kG3M + kG3M + kG3K - 1) / kG3K;
    if (e < se) se = e;
  }
}

__device__ __forceinline__ void g3_split(const uint4& raw, uint4& hi, uint4& lo) {
  const uint32_t r[4] = {raw.x, raw.y, raw.z, raw.w};
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h[e]) : "f"(__uint_as_float(r[e])));
    l[e] = __float_as_uint(__fsub_rn(__uint_as_float(r[e]), __uint_as_float(h[e])));
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

template <int TN, bool B_MN, bool CHUNKED, bool A_MN>
__global__ void __launch_bounds__(kG3Threads, 1)
gemm3x_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap, const G3Params p,
              const uint32_t idesc) {
  using Cfg = G3Cfg<TN>;
  constexpr int kStages = Cfg::kStages;
  constexpr uint32_t kBBytes = Cfg::kBBytes, kStageBytes = Cfg::kStageBytes;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  G3Barriers* bars = reinterpret_cast<G3Barriers*>(smem + kStages * kStageBytes + kG3EpiWarps * 4096);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(&amap) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(&bmap) : "memory");
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->conv[s], kG3CvtWarps * 32);
      mbar_init(&bars->empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) { mbar_init(&bars->tmem_full[s], 1); mbar_init(&bars->tmem_empty[s], kG3EpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(&bars->tmem_base)), "r"((uint32_t)(2 * TN)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  const int nslices = (p.K + kG3K - 1) / kG3K;
  const int slices_per_chunk = CHUNKED ? p.kc / kG3K : nslices;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        int bi, bj;
        g3_tile<TN>(t, p, bi, bj);
        const int m0 = bi * kG3M, n0 = bj * TN;
        int sbeg, send;
        g3_krange<TN>(p, bi, bj, nslices, sbeg, send);
        for (int ks = sbeg; ks < send; ++ks) {
          const int k0 = ks * kG3K;
          mbar_wait(&bars->empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + 2 * kG3ABytes;
          mbar_expect_tx(&bars->full[stage], kG3ABytes + kBBytes);
          if (A_MN) {
#pragma unroll
            for (int q = 0; q < kG3M / 32; ++q) tma_load_2d(sa + q * (kG3K * 128), &amap, &bars->full[stage], m0 + q * 32, k0);
          } else {
            tma_load_2d(sa, &amap, &bars->full[stage], k0, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int q = 0; q < TN / 32; ++q) tma_load_2d(sb + q * (kG3K * 128), &bmap, &bars->full[stage], n0 + q * 32, k0);
          } else {
            tma_load_2d(sb, &bmap, &bars->full[stage], k0, n0);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        int bi, bj, tb, te;
        g3_tile<TN>(t, p, bi, bj);
        g3_krange<TN>(p, bi, bj, nslices, tb, te);
        for (int sbeg = tb; sbeg < te; sbeg += slices_per_chunk) {
          int send = sbeg + slices_per_chunk;
          if (send > te) send = te;
          mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * TN;
          uint32_t accumulate = 0;
          for (int ks = sbeg; ks < send; ++ks) {
            mbar_wait(&bars->conv[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * kStageBytes);
            const uint32_t sb = sa + 2 * kG3ABytes;
#pragma unroll
            for (int kk = 0; kk < kG3K / 8; ++kk) {
              const uint64_t a_hi = A_MN ? g3_desc_mn(sa + kk * 1024) : g3_desc_k(sa + kk * 32);
              const uint64_t a_lo = A_MN ? g3_desc_mn(sa + kG3ABytes + kk * 1024) : g3_desc_k(sa + kG3ABytes + kk * 32);
              const uint64_t b_hi = B_MN ? g3_desc_mn(sb + kk * 1024) : g3_desc_k(sb + kk * 32);
              const uint64_t b_lo = B_MN ? g3_desc_mn(sb + kBBytes + kk * 1024) : g3_desc_k(sb + kBBytes + kk * 32);
              tc_mma_tf32(d_tmem, a_lo, b_hi, idesc, accumulate);    // small terms first
              tc_mma_tf32(d_tmem, a_hi, b_lo, idesc, 1);
              tc_mma_tf32(d_tmem, a_hi, b_hi, idesc, 1);
              accumulate = 1;
            }
            tc_commit(&bars->empty[stage]);          // frees the smem slot when these MMAs retire
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          tc_commit(&bars->tmem_full[acc]);          // chunk accumulator complete
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else if (warp >= kG3CvtWarp0 && warp < kG3CvtWarp0 + kG3CvtWarps) {
    // ===== converters: raw fp32 slice -> hi (in place) and lo =====
    const int ct = threadIdx.x - kG3CvtWarp0 * 32;
    int stage = 0; uint32_t phase = 0;
    for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
      int bi, bj, tb, te;
      g3_tile<TN>(t, p, bi, bj);
      g3_krange<TN>(p, bi, bj, nslices, tb, te);
      for (int ks = tb; ks < te; ++ks) {
        mbar_wait(&bars->full[stage], phase);
        uint8_t* sa = smem + stage * kStageBytes;
        uint8_t* sb = sa + 2 * kG3ABytes;
        // A and B slices are adjacent pairs {raw/hi, lo}: batches of 8 independent 16-byte vectors per thread
        auto convert = [&](uint8_t* base, uint32_t bytes) {
          constexpr int kCT = kG3CvtWarps * 32, kBatch = 8;
          for (int i0 = ct; i0 < (int)(bytes / 16); i0 += kCT * kBatch) {
            uint4 raw[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) raw[u] = *reinterpret_cast<const uint4*>(base + (i0 + u * kCT) * 16);
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
              uint4 hi, lo;
              g3_split(raw[u], hi, lo);
              *reinterpret_cast<uint4*>(base + (i0 + u * kCT) * 16) = hi;
              *reinterpret_cast<uint4*>(base + bytes + (i0 + u * kCT) * 16) = lo;
            }
          }
        };
        convert(sa, kG3ABytes);
        convert(sb, kBBytes);
        fence_proxy_async_smem();
        mbar_arrive(&bars->conv[stage]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= kG3EpiWarp0) {
    // ===== epilogue =====
    // tcgen05.ld hands every thread one accumulator ROW (32 consecutive columns per load).  Writing rows straight to
    // global memory would touch 32 different cache lines per instruction, so each 32 x 32 block goes through a
    // swizzled 4 KB staging buffer and leaves as 4 rows x 128 contiguous bytes per instruction.
    constexpr int kCols = TN / 2;           // columns per epilogue warp
    constexpr int kBlocks = kCols / 32;
    const int ew = warp - kG3EpiWarp0;
    const int q = warp & 3;                 // TMEM lane quarter this warp may touch
    const int half = ew >> 2;
    float* stg = reinterpret_cast<float*>(smem + kStages * kStageBytes) + ew * 1024;
    const int rsub = lane >> 3, jsub = lane & 7;
    int acc = 0; uint32_t acc_phase = 0;

    // C values of block `blk` of the tile at (row0, col0): 8 float4 per lane, rows 4*i + rsub, columns 4*jsub
    auto load_c = [&](float4 (&h)[8], int row0, int col0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int gi = row0 + 4 * i + rsub, gj = col0 + 4 * jsub;
        h[i] = (p.beta != 0.f && gi < p.M && gj < p.N)
                   ? *reinterpret_cast<const float4*>(p.C + (int64_t)gi * p.ldc + gj) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    // v = this lane's row of the block -> C (alpha * v + beta * h)
    auto store_block = [&](const float (&v)[32], const float4 (&h)[8], int row0, int col0) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = 4 * i + rsub;
        const float4 a = *reinterpret_cast<const float4*>(stg + r * 32 + ((jsub ^ (r & 7)) << 2));
        const int gi = row0 + r, gj = col0 + 4 * jsub;
        if (gi < p.M && gj < p.N) {         // N % 4 == 0: a float4 is entirely in or out
          float4 o;
          o.x = fmaf(p.alpha, a.x, p.beta * h[i].x); o.y = fmaf(p.alpha, a.y, p.beta * h[i].y);
          o.z = fmaf(p.alpha, a.z, p.beta * h[i].z); o.w = fmaf(p.alpha, a.w, p.beta * h[i].w);
          *reinterpret_cast<float4*>(p.C + (int64_t)gi * p.ldc + gj) = o;
        }
      }
      __syncwarp();
    };

    for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
      int bi, bj;
      g3_tile<TN>(t, p, bi, bj);
      const int row0 = bi * kG3M + q * 32;
      const int col0 = bj * TN + half * kCols;
      if (CHUNKED) {
        float sum[kCols];
#pragma unroll
        for (int c = 0; c < kCols; ++c) sum[c] = 0.f;
        int tb, te;
        g3_krange<TN>(p, bi, bj, nslices, tb, te);
        const int nchunks = (te - tb + slices_per_chunk - 1) / slices_per_chunk;
        for (int ch = 0; ch < nchunks; ++ch) {
          mbar_wait(&bars->tmem_full[acc], acc_phase);
          tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * TN + half * kCols;
#pragma unroll
          for (int c4 = 0; c4 < kBlocks; ++c4) {
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
#pragma unroll
        for (int c4 = 0; c4 < kBlocks; ++c4) {
          float4 h[8];
          load_c(h, row0, col0 + c4 * 32);
          float v[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] = sum[c4 * 32 + c];
          store_block(v, h, row0, col0 + c4 * 32);
        }
      } else {
        float4 h[2][8];
        load_c(h[0], row0, col0);           // in flight while the tensor core is still working on this tile
        mbar_wait(&bars->tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * TN + half * kCols;
#pragma unroll
        for (int c4 = 0; c4 < kBlocks; ++c4) {
          if (c4 + 1 < kBlocks) load_c(h[(c4 + 1) & 1], row0, col0 + (c4 + 1) * 32);
          uint32_t u[32];
          tc_ld32(taddr + c4 * 32, u);
          tc_wait_ld();
          float v[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] = __uint_as_float(u[c]);
          store_block(v, h[c4 & 1], row0, col0 + c4 * 32);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)(2 * TN)) : "memory");
  }
}

template <int TN> constexpr size_t g3_smem() {
  return G3Cfg<TN>::kStages * G3Cfg<TN>::kStageBytes + kG3EpiWarps * 4096 + sizeof(G3Barriers) + 1024;
}

template <int TN, bool B_MN, bool CHUNKED, bool A_MN = false>
static int g3_launch(const CUtensorMap& amap, const CUtensorMap& bmap, G3Params p, cudaStream_t st, int max_ctas = 0) {
  auto kern = gemm3x_kernel<TN, B_MN, CHUNKED, A_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g3_smem<TN>()) != cudaSuccess)
      return check_launch();
    attr_set = true;
  }
  p.nbi = (p.M + kG3M - 1) / kG3M;
  p.nbj = (p.N + TN - 1) / TN;
  if (p.tri) {
    p.ntiles = 0;
    for (int bi = 0; bi < p.nbi; ++bi) {
      int cnt = (bi * kG3M + kG3M - 1) / TN + 1;
      p.ntiles += cnt < p.nbj ? cnt : p.nbj;
    }
  } else {
    p.ntiles = p.nbi * p.nbj;
  }
  // fp32 accumulate, A and B TF32, A K-major, B K-major or MN-major, N = TN, M = 128 (cute::UMMA::InstrDescriptor)
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                         ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(kG3M >> 4) << 24);
  const int cap = (max_ctas > 0 && max_ctas < kNumSMs) ? max_ctas : kNumSMs;
  const int grid = p.ntiles < cap ? p.ntiles : cap;
  kern<<<grid, kG3Threads, g3_smem<TN>(), st>>>(amap, bmap, p, idesc);
  return check_launch();
}

static int g3_map(CUtensorMap* map, const float* base, uint64_t inner, uint64_t outer, int64_t ld, uint32_t box_inner,
                  uint32_t box_outer, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return VLMC_ERR_CUDA;
  const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {box_inner, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  CUresult cr = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { g_last_cuda_error = (int)cr; return VLMC_ERR_CUDA; }
  return VLMC_OK;
}

int gemm3x(bool b_nk, int M, int N, int K, float alpha, const float* A, int64_t lda, const float* B, int64_t ldb,
           float beta, float* C, int64_t ldc, int tri, int kc, cudaStream_t st, int kmode, bool a_km, int max_ctas) {
  if (M <= 0 || N <= 0) return VLMC_OK;
  if (K <= 0 || !A || !B || !C) return VLMC_ERR_BAD_ARG;
  // K is the contiguous dimension of a K-major operand only: [K,M] / [K,N] operands take any K
  if (a_km && (b_nk || kmode != 0 || (M & 3))) return VLMC_ERR_UNSUPPORTED;
  if ((!(a_km && !b_nk) && (K & 3)) || (N & 3) || (lda & 3) || (ldb & 3) || (ldc & 3) || ((uintptr_t)A & 15) || ((uintptr_t)B & 15) ||
      ((uintptr_t)C & 15))
    return VLMC_ERR_UNSUPPORTED;
  if (kc <= 0) kc = 128;
  if (kmode < 0 || kmode > 2 || (kmode == 1 && b_nk)) return VLMC_ERR_BAD_ARG;
  kc = (kc + kG3K - 1) / kG3K * kG3K;
  const bool chunked = K > kc;
  const int TN = (!chunked && N > 128 && !a_km) ? 256 : 128;

  CUtensorMap amap, bmap;
  int rc = a_km ? g3_map(&amap, A, (uint64_t)M, (uint64_t)K, lda, 32, kG3K, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
                : g3_map(&amap, A, (uint64_t)K, (uint64_t)M, lda, kG3K, kG3M, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  if (b_nk) rc = g3_map(&bmap, B, (uint64_t)K, (uint64_t)N, ldb, kG3K, (uint32_t)TN, CU_TENSOR_MAP_SWIZZLE_128B);
  else rc = g3_map(&bmap, B, (uint64_t)N, (uint64_t)K, ldb, 32, kG3K, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc) return rc;

  G3Params p;
  p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K; p.alpha = alpha; p.beta = beta; p.tri = tri; p.kc = kc; p.kmode = kmode;
  p.nbi = p.nbj = p.ntiles = 0;
  if (a_km) {   // A given as [K,M] (A^T B products: the fp32-activation Hessian X^T X)
    if (chunked) return g3_launch<128, true, true, true>(amap, bmap, p, st);
    return g3_launch<128, true, false, true>(amap, bmap, p, st);
  }
  if (chunked) return b_nk ? g3_launch<128, false, true>(amap, bmap, p, st) : g3_launch<128, true, true>(amap, bmap, p, st);
  if (TN == 256) return b_nk ? g3_launch<256, false, false>(amap, bmap, p, st, max_ctas) : g3_launch<256, true, false>(amap, bmap, p, st, max_ctas);
  return b_nk ? g3_launch<128, false, false>(amap, bmap, p, st, max_ctas) : g3_launch<128, true, false>(amap, bmap, p, st, max_ctas);
}

}  // namespace vlmc

extern "C" int vlmc_gemm_tf32x3(int b_nk, int M, int N, int K, float alpha, const float* A, int64_t lda,
                                const float* B, int64_t ldb, float beta, float* C, int64_t ldc, int tri, int kc,
                                void* stream) {
  using namespace vlmc;
  if (!A || !B || !C || M < 1 || N < 1 || K < 1 || lda < K || ldc < N || ldb < (b_nk ? K : N)) return VLMC_ERR_BAD_ARG;
  if (!is_device_ptr(A) || !is_device_ptr(B) || !is_device_ptr(C)) return VLMC_ERR_NOT_DEVICE;
  return gemm3x(b_nk != 0, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, tri, kc, (cudaStream_t)stream, 0, false);
}
