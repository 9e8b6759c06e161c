// K23: the SparseLoRA masked training forward as ONE kernel: y = x W_eff^T (+ bias) with the effective weight built ON CHIP
// ("fused masked-weight GEMM").
//
// Replaces lavis/peft/src/peft/tuners/lora.py:359-382 (Linear.forward, r > 0, not merged)
//   sparse:      F.linear(x, (W + (B @ A).to(W.dtype) * scaling) * mask, bias)        (:364-369)
//   not sparse:  F.linear(x,  W * mask + (B @ A).to(W.dtype) * scaling,  bias)        (:370-375)
// which materialises a dense [R, C] weight every step.  K15 + a library GEMM (lora_forward.cu) still writes and re-reads that
// weight (5 + 2 B / weight of HBM traffic per step); here no effective weight ever reaches HBM: TMA brings raw W (16-bit),
// the keep mask (bytes) and the matching slice of lora_A into shared memory, converter warps rewrite the W slice in place
// as W_eff (rank-r dot in fp32 registers, k ascending, the reference's three roundings - the same lt_effective() as K15, so
// the operand the tensor core sees is bit-identical to K15's output), and one thread issues tcgen05.mma.kind::f16 against
// the token tiles with the accumulators in TMEM.
//
// Work unit = 512 tokens x 128 output features: 4 accumulators of 128 x 128 fp32 = all 512 TMEM columns, so a converted
// 128 x 64 weight slice feeds FOUR 128-token MMAs (conversion ~700 issue cycles per slice against 1024 tensor cycles).
//
// Warp roles (448 threads, 1 CTA / SM, persistent over units, feature tiles fastest so that a wave shares token blocks):
//   warp 0        TMA producer    W ring (3 x {W 128x64, mask 128x64 B, lora_A RKx64 fp32}) and x ring (2 x 4 x {128x64})
//   warp 1        TMEM allocator + MMA issuer (one thread)
//   warps 2-9     converters      W slice -> W_eff slice in place, fence.proxy.async, arrive
//   warps 10-13   epilogue        tcgen05.ld -> (+ bias) -> round to the layer dtype -> global (64 contiguous bytes per thread)
#include "tc.cuh"
#include "lora_tile.cuh"

namespace vlmc {

constexpr int kLLM = 128;                 // tokens per MMA (UMMA M)
constexpr int kLLN = 128;                 // output features per unit (UMMA N)
constexpr int kLLMT = 4;                  // token tiles per unit
constexpr int kLLK = 64;                  // k-slice: 64 16-bit elements = one 128-byte swizzle row
constexpr int kLLCvtWarp0 = 2, kLLCvtWarps = 8, kLLEpiWarp0 = 10, kLLEpiWarps = 4;
constexpr int kLLCvtThreads = kLLCvtWarps * 32;
constexpr int kLLThreads = (kLLEpiWarp0 + kLLEpiWarps) * 32;      // 448
constexpr int kLLXStages = 2, kLLWStages = 3;
constexpr uint32_t kLLXTile = kLLM * 128;            // 16 KB
constexpr uint32_t kLLXStage = kLLMT * kLLXTile;     // 64 KB
constexpr uint32_t kLLWTile = kLLN * 128;            // 16 KB
constexpr uint32_t kLLMaskTile = kLLN * kLLK;        // 8 KB

struct __align__(8) LLBarriers {
  uint64_t full_x[kLLXStages], empty_x[kLLXStages];
  uint64_t full_w[kLLWStages], conv_w[kLLWStages], empty_w[kLLWStages];
  uint64_t tmem_full, tmem_empty;
  uint32_t tmem_base;
};

template <int RK> struct LLCfg {
  static constexpr uint32_t kABytes = RK * kLLK * 4;                          // 2 KB (RK = 8) / 4 KB (RK = 16)
  static constexpr uint32_t kWStage = kLLWTile + kLLMaskTile + kABytes;       // 26 / 28 KB, a multiple of 1024
  static constexpr uint32_t kSB = kLLN * RK * 4;                              // lora_B rows of the unit
  static constexpr size_t kSmem = 1024 + kLLXStages * kLLXStage + kLLWStages * kWStage + kSB + sizeof(LLBarriers);
};

struct LLParams {
  void* y;
  int64_t ldy;
  const void* bias;
  const float* B;
  int T, R, C, rank, sparse;
  float scaling;
  int nbt, nbn, nunits;
};

// K-major 16-bit operand, 128-byte swizzle: rows 128 B apart, 8-row groups 1024 B apart (SBO); LBO unused
__device__ __forceinline__ uint64_t ll_desc_k(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <typename T> __device__ __forceinline__ uint32_t ll_pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t ll_pack2<__half>(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t ll_pack2<__nv_bfloat16>(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <typename T> __device__ __forceinline__ float ll_to_float(T v);
template <> __device__ __forceinline__ float ll_to_float<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float ll_to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T, int RK>
__global__ void __launch_bounds__(kLLThreads, 1)
lora_linear_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap wmap,
                   const __grid_constant__ CUtensorMap mmap, const __grid_constant__ CUtensorMap amap, const LLParams p,
                   const uint32_t idesc) {
  using Cfg = LLCfg<RK>;
  constexpr uint32_t kWStage = Cfg::kWStage;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* xring = smem;
  uint8_t* wring = xring + kLLXStages * kLLXStage;
  float* sB = reinterpret_cast<float*>(wring + kLLWStages * kWStage);
  LLBarriers* bars = reinterpret_cast<LLBarriers*>(reinterpret_cast<uint8_t*>(sB) + Cfg::kSB);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(&xmap) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(&wmap) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(&mmap) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(&amap) : "memory");
    for (int s = 0; s < kLLXStages; ++s) { mbar_init(&bars->full_x[s], 1); mbar_init(&bars->empty_x[s], 1); }
    for (int s = 0; s < kLLWStages; ++s) {
      mbar_init(&bars->full_w[s], 1);
      mbar_init(&bars->conv_w[s], kLLCvtThreads);
      mbar_init(&bars->empty_w[s], 1);
    }
    mbar_init(&bars->tmem_full, 1);
    mbar_init(&bars->tmem_empty, kLLEpiWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(&bars->tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  const int nslices = (p.C + kLLK - 1) / kLLK;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int xs = 0, ws = 0; uint32_t xphase = 0, wphase = 0;
      for (int u = blockIdx.x; u < p.nunits; u += gridDim.x) {
        const int bt = u / p.nbn, bn = u - bt * p.nbn;
        const int t0 = bt * (kLLMT * kLLM), n0 = bn * kLLN;
        int mt_cnt = (p.T - t0 + kLLM - 1) / kLLM;
        if (mt_cnt > kLLMT) mt_cnt = kLLMT;
        for (int ks = 0; ks < nslices; ++ks) {
          const int k0 = ks * kLLK;
          mbar_wait(&bars->empty_w[ws], wphase ^ 1);
          uint8_t* sw = wring + ws * kWStage;
          mbar_expect_tx(&bars->full_w[ws], kWStage);
          tma_load_2d(sw, &wmap, &bars->full_w[ws], k0, n0);
          tma_load_2d(sw + kLLWTile, &mmap, &bars->full_w[ws], k0, n0);
          tma_load_2d(sw + kLLWTile + kLLMaskTile, &amap, &bars->full_w[ws], k0, 0);
          if (++ws == kLLWStages) { ws = 0; wphase ^= 1; }

          mbar_wait(&bars->empty_x[xs], xphase ^ 1);
          uint8_t* sx = xring + xs * kLLXStage;
          mbar_expect_tx(&bars->full_x[xs], (uint32_t)mt_cnt * kLLXTile);
          for (int mt = 0; mt < mt_cnt; ++mt) tma_load_2d(sx + mt * kLLXTile, &xmap, &bars->full_x[xs], k0, t0 + mt * kLLM);
          if (++xs == kLLXStages) { xs = 0; xphase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      int xs = 0, ws = 0; uint32_t xphase = 0, wphase = 0, acc_phase = 0;
      for (int u = blockIdx.x; u < p.nunits; u += gridDim.x) {
        const int bt = u / p.nbn;
        const int t0 = bt * (kLLMT * kLLM);
        int mt_cnt = (p.T - t0 + kLLM - 1) / kLLM;
        if (mt_cnt > kLLMT) mt_cnt = kLLMT;
        mbar_wait(&bars->tmem_empty, acc_phase ^ 1);          // the epilogue has drained the previous unit
        tc_fence_after();
        for (int ks = 0; ks < nslices; ++ks) {
          mbar_wait(&bars->conv_w[ws], wphase);
          mbar_wait(&bars->full_x[xs], xphase);
          tc_fence_after();
          const uint32_t sx = smem_u32(xring + xs * kLLXStage);
          const uint32_t sw = smem_u32(wring + ws * kWStage);
#pragma unroll
          for (int kk = 0; kk < kLLK / 16; ++kk) {
            const uint64_t bdesc = ll_desc_k(sw + kk * 32);
            for (int mt = 0; mt < mt_cnt; ++mt)
              tc_mma_f16(tmem_base + mt * kLLN, ll_desc_k(sx + mt * kLLXTile + kk * 32), bdesc, idesc, (ks > 0 || kk > 0) ? 1u : 0u);
          }
          tc_commit(&bars->empty_w[ws]);          // frees both slots when these MMAs retire
          tc_commit(&bars->empty_x[xs]);
          if (++ws == kLLWStages) { ws = 0; wphase ^= 1; }
          if (++xs == kLLXStages) { xs = 0; xphase ^= 1; }
        }
        tc_commit(&bars->tmem_full);
        acc_phase ^= 1;
      }
    }
  } else if (warp < kLLEpiWarp0) {
    // ===== converters: W slice -> W_eff slice, in place =====
    // thread -> 16-byte chunk j (8 columns) of rows rb, rb + 32, rb + 64, rb + 96: a quarter-warp covers one 128-byte row
    // (conflict-free under the swizzle), the lora_A values of the 8 columns are held in registers for the four rows
    constexpr int V = 8;
    const int ct = threadIdx.x - kLLCvtWarp0 * 32;
    const int j = ct & 7, rb = ct >> 3;
    const uint32_t chunk_off = (uint32_t)((j ^ (rb & 7)) << 4);
    int ws = 0; uint32_t wphase = 0;
    for (int u = blockIdx.x; u < p.nunits; u += gridDim.x) {
      const int bt = u / p.nbn, bn = u - bt * p.nbn;
      const int n0 = bn * kLLN;
      asm volatile("bar.sync 1, %0;" :: "n"(kLLCvtThreads) : "memory");       // everyone is done with the previous unit's rows
      for (int idx = ct; idx < kLLN * RK; idx += kLLCvtThreads) {
        const int row = idx / RK, k = idx - row * RK;
        sB[idx] = (n0 + row < p.R && k < p.rank) ? __ldg(p.B + (int64_t)(n0 + row) * p.rank + k) : 0.f;
      }
      asm volatile("bar.sync 1, %0;" :: "n"(kLLCvtThreads) : "memory");
      for (int ks = 0; ks < nslices; ++ks) {
        mbar_wait(&bars->full_w[ws], wphase);
        uint8_t* sw = wring + ws * kWStage;
        const uint8_t* sm = sw + kLLWTile;
        const float* sa = reinterpret_cast<const float*>(sm + kLLMaskTile);
        uint4 wv[4];
        uint2 mv[4];
        float2 acc[4][V / 2];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = rb + 32 * i;
          wv[i] = *reinterpret_cast<const uint4*>(sw + row * 128 + chunk_off);
          mv[i] = *reinterpret_cast<const uint2*>(sm + row * kLLK + j * 8);
#pragma unroll
          for (int e = 0; e < V / 2; ++e) acc[i][e] = make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int kh = 0; kh < RK / 8; ++kh) {
          float2 a[8][V / 2];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 lo = *reinterpret_cast<const float4*>(sa + (kh * 8 + k) * kLLK + j * 8);
            const float4 hi = *reinterpret_cast<const float4*>(sa + (kh * 8 + k) * kLLK + j * 8 + 4);
            a[k][0] = make_float2(lo.x, lo.y); a[k][1] = make_float2(lo.z, lo.w);
            a[k][2] = make_float2(hi.x, hi.y); a[k][3] = make_float2(hi.z, hi.w);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float* brow = sB + (rb + 32 * i) * RK + kh * 8;
            const float4 b0 = *reinterpret_cast<const float4*>(brow), b1 = *reinterpret_cast<const float4*>(brow + 4);
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float2 b2 = make_float2(b[k], b[k]);
#pragma unroll
              for (int e = 0; e < V / 2; ++e) acc[i][e] = __ffma2_rn(b2, a[k][e], acc[i][e]);      // k ascending, like K15
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float f[V], ac[V];
          Elem<T>::unpack(wv[i], f);
#pragma unroll
          for (int e = 0; e < V / 2; ++e) { ac[2 * e] = acc[i][e].x; ac[2 * e + 1] = acc[i][e].y; }
          const uint32_t mb[2] = {mv[i].x, mv[i].y};
          lt_effective<T>(f, ac, mb, p.scaling, p.sparse);
          *reinterpret_cast<uint4*>(sw + (rb + 32 * i) * 128 + chunk_off) = Elem<T>::pack(f);
        }
        fence_proxy_async_smem();
        mbar_arrive(&bars->conv_w[ws]);
        if (++ws == kLLWStages) { ws = 0; wphase ^= 1; }
      }
    }
  } else {
    // ===== epilogue =====
    const int q = warp & 3;                 // TMEM lane quarter this warp may touch
    uint32_t acc_phase = 0;
    T* y = reinterpret_cast<T*>(p.y);
    const T* bias = reinterpret_cast<const T*>(p.bias);
    for (int u = blockIdx.x; u < p.nunits; u += gridDim.x) {
      const int bt = u / p.nbn, bn = u - bt * p.nbn;
      const int t0 = bt * (kLLMT * kLLM), n0 = bn * kLLN;
      int mt_cnt = (p.T - t0 + kLLM - 1) / kLLM;
      if (mt_cnt > kLLMT) mt_cnt = kLLMT;
      mbar_wait(&bars->tmem_full, acc_phase);
      tc_fence_after();
      for (int mt = 0; mt < mt_cnt; ++mt) {
        const int gi = t0 + mt * kLLM + q * 32 + lane;
#pragma unroll
        for (int c4 = 0; c4 < kLLN / 32; ++c4) {
          uint32_t v[32];
          tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + mt * kLLN + c4 * 32, v);
          tc_wait_ld();
          const int gj = n0 + c4 * 32;
          if (gi < p.T) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (gj + 8 * g < p.R) {       // R % 8 == 0: a 16-byte vector is entirely in or out
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(v[8 * g + e]);
                if (bias) {
                  const uint4 bv = __ldg(reinterpret_cast<const uint4*>(bias + gj + 8 * g));
                  float bf[8];
                  Elem<T>::unpack(bv, bf);
#pragma unroll
                  for (int e = 0; e < 8; ++e) o[e] += bf[e];
                }
                uint4 ov;
                ov.x = ll_pack2<T>(o[0], o[1]); ov.y = ll_pack2<T>(o[2], o[3]);
                ov.z = ll_pack2<T>(o[4], o[5]); ov.w = ll_pack2<T>(o[6], o[7]);
                *reinterpret_cast<uint4*>(y + (int64_t)gi * p.ldy + gj + 8 * g) = ov;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tmem_empty);
      acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
  }
}

static int ll_map(CUtensorMap* map, CUtensorMapDataType type, int elem_bytes, const void* base, uint64_t inner, uint64_t outer,
                  int64_t ld, uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return VLMC_ERR_CUDA;
  const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * elem_bytes};
  const cuuint32_t box[2] = {box_inner, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  CUresult cr = encode(map, type, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { g_last_cuda_error = (int)cr; return VLMC_ERR_CUDA; }
  return VLMC_OK;
}

template <typename T, int RK>
static int ll_launch(const CUtensorMap& xmap, const CUtensorMap& wmap, const CUtensorMap& mmap, const CUtensorMap& amap,
                     const LLParams& p, uint32_t idesc, cudaStream_t st) {
  auto kern = lora_linear_kernel<T, RK>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LLCfg<RK>::kSmem) != cudaSuccess)
      return check_launch();
    attr_set = true;
  }
  const int grid = p.nunits < kNumSMs ? p.nunits : kNumSMs;
  kern<<<grid, kLLThreads, LLCfg<RK>::kSmem, st>>>(xmap, wmap, mmap, amap, p, idesc);
  return check_launch();
}

}  // namespace vlmc

extern "C" int vlmc_sparselora_linear_forward(const void* x, int dtype, int64_t T, int C, int64_t ldx, const void* W, int R,
                                              int64_t ldw, const float* A, const float* B, int rank, float scaling,
                                              const uint8_t* keep_mask, int64_t ldm, int sparse, const void* bias, void* y,
                                              int64_t ldy, void* stream) {
  using namespace vlmc;
  if (!x || !W || !A || !B || !keep_mask || !y || T < 0 || R < 1 || C < 1 || rank < 1 || ldx < C || ldw < C || ldm < C || ldy < R)
    return VLMC_ERR_BAD_ARG;
  if (dtype != VLMC_F16 && dtype != VLMC_BF16) return dtype == VLMC_F32 ? VLMC_ERR_UNSUPPORTED : VLMC_ERR_BAD_ARG;
  if (T == 0) return VLMC_OK;
  if (rank > 16 || T > 0x7fffffff) return VLMC_ERR_UNSUPPORTED;
  // TMA: 16-byte aligned bases and row pitches; the epilogue writes 16-byte vectors
  if (C % 8 != 0 || R % 8 != 0 || ldx % 8 != 0 || ldw % 8 != 0 || ldm % 16 != 0 || ldy % 8 != 0 || ((uintptr_t)x & 15) != 0 ||
      ((uintptr_t)W & 15) != 0 || ((uintptr_t)keep_mask & 15) != 0 || ((uintptr_t)A & 15) != 0 || ((uintptr_t)y & 15) != 0 ||
      (bias && ((uintptr_t)bias & 15) != 0))
    return VLMC_ERR_UNSUPPORTED;
  if (!is_device_ptr(x) || !is_device_ptr(W) || !is_device_ptr(A) || !is_device_ptr(B) || !is_device_ptr(keep_mask) ||
      !is_device_ptr(y) || (bias && !is_device_ptr(bias)))
    return VLMC_ERR_NOT_DEVICE;

  const int RK = rank <= 8 ? 8 : 16;
  const CUtensorMapDataType dt = dtype == VLMC_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap xmap, wmap, mmap, amap;
  int rc = ll_map(&xmap, dt, 2, x, (uint64_t)C, (uint64_t)T, ldx, kLLK, kLLM, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = ll_map(&wmap, dt, 2, W, (uint64_t)C, (uint64_t)R, ldw, kLLK, kLLN, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = ll_map(&mmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, keep_mask, (uint64_t)C, (uint64_t)R, ldm, kLLK, kLLN, CU_TENSOR_MAP_SWIZZLE_NONE);
  if (rc) return rc;
  rc = ll_map(&amap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, A, (uint64_t)C, (uint64_t)rank, C, kLLK, (uint32_t)RK, CU_TENSOR_MAP_SWIZZLE_NONE);
  if (rc) return rc;

  LLParams p;
  p.y = y; p.ldy = ldy; p.bias = bias; p.B = B; p.T = (int)T; p.R = R; p.C = C; p.rank = rank; p.sparse = sparse ? 1 : 0;
  p.scaling = scaling;
  p.nbt = (int)((T + kLLMT * kLLM - 1) / (kLLMT * kLLM));
  p.nbn = (R + kLLN - 1) / kLLN;
  const int64_t nunits = (int64_t)p.nbt * p.nbn;
  if (nunits > 0x7fffffff) return VLMC_ERR_UNSUPPORTED;
  p.nunits = (int)nunits;
  // instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, A and B K-major, N = 128, M = 128
  const uint32_t fmt = dtype == VLMC_F16 ? 0u : 1u;
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(kLLN >> 3) << 17) | ((uint32_t)(kLLM >> 4) << 24);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VLMC_F16) return RK == 8 ? ll_launch<__half, 8>(xmap, wmap, mmap, amap, p, idesc, st)
                                        : ll_launch<__half, 16>(xmap, wmap, mmap, amap, p, idesc, st);
  return RK == 8 ? ll_launch<__nv_bfloat16, 8>(xmap, wmap, mmap, amap, p, idesc, st)
                 : ll_launch<__nv_bfloat16, 16>(xmap, wmap, mmap, amap, p, idesc, st);
}
