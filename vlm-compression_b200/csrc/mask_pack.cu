// Bit-packed mask exchange for the row-sharded multi-GPU plan (new: the reference runs replicas).
//
// After a rank has selected the masks of ITS output rows (wanda_pruner.py:323-341 on a row shard), every other rank
// needs the same rows of `module.mask` (:339) and of the pruned weight (:341).  The weights are replicated inputs, so
// only the mask has to travel: 1 bit per weight instead of the 3 bytes of an fp16 weight + a bool (24x less NVLink
// traffic).  vlmc_mask_pack turns mask bytes into bits; vlmc_mask_apply_packed expands received bits into mask
// bytes and zeroes the pruned weights of the local replica in the same pass.
// Bit e of byte j of a row is column 8 j + e (numpy.packbits(bitorder="little")).
#include "common.cuh"

namespace vlmc {

__device__ __forceinline__ void mask_pack_body(const uint8_t* __restrict__ keep, int64_t ldm, int R, int C,
                                               uint8_t* __restrict__ bits, int64_t ldb) {
  // one thread: 16 mask bytes -> 2 bytes of bits
  const int per_row = C >> 4;
  const int64_t n = (int64_t)R * per_row;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(idx / per_row), j = (int)(idx % per_row);
    const uint4 v = ld_stream(keep + (int64_t)row * ldm + j * 16);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t out = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      // bytes are 0 / 1: gather bit 0 of each byte of w[q] into 4 bits
      const uint32_t m = w[q] & 0x01010101u;
      const uint32_t nib = (m | (m >> 7) | (m >> 14) | (m >> 21)) & 0xfu;
      out |= nib << (4 * q);
    }
    *reinterpret_cast<uint16_t*>(bits + (int64_t)row * ldb + j * 2) = (uint16_t)out;
  }
}

__global__ void __launch_bounds__(256)
mask_pack_kernel(const uint8_t* __restrict__ keep, int64_t ldm, int R, int C, uint8_t* __restrict__ bits, int64_t ldb) {
  mask_pack_body(keep, ldm, R, C, bits, ldb);
}

// up to kMpBatchMax matrices per launch: grid (blocks, items) - the row-sharded block plan packs / expands the masks of
// all linears of a block around ONE all-gather, and seven 10-40 us launches each way were most of that phase
constexpr int kMpBatchMax = 16;
struct MpPackBatch { vlmc_pack_item it[kMpBatchMax]; };
struct MpApplyBatch { vlmc_apply_item it[kMpBatchMax]; };
__global__ void __launch_bounds__(256) mask_pack_batch_kernel(const __grid_constant__ MpPackBatch b) {
  const vlmc_pack_item& p = b.it[blockIdx.y];
  mask_pack_body(p.keep_mask, p.ldm, p.R, p.C, p.bits, p.ldb);
}

template <typename T>
__device__ __forceinline__ void mask_apply_packed_body(T* __restrict__ W, int64_t ldw, int R, int C,
                                                       const uint8_t* __restrict__ bits, int64_t ldb, int rows_per_seg,
                                                       int64_t seg_stride, uint8_t* __restrict__ keep, int64_t ldm, int zero_w) {
  // one thread: 2 bytes of bits -> 16 mask bytes, 16 weights
  constexpr int V = Elem<T>::kVec;            // weights per 16-byte vector
  constexpr int NV = 16 / V;
  const int per_row = C >> 4;
  const int64_t n = (int64_t)R * per_row;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(idx / per_row), j = (int)(idx % per_row);
    const int seg = row / rows_per_seg;
    const uint32_t b = *reinterpret_cast<const uint16_t*>(bits + seg * seg_stride + (int64_t)(row - seg * rows_per_seg) * ldb + j * 2);
    if (keep) {
      uint32_t m[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t nib = (b >> (4 * q)) & 0xfu;
        m[q] = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
      }
      st_stream(keep + (int64_t)row * ldm + j * 16, make_uint4(m[0], m[1], m[2], m[3]));
    }
    if (zero_w && b != 0xffffu) {
      T* wp = W + (int64_t)row * ldw + j * 16;
#pragma unroll
      for (int q = 0; q < NV; ++q) {
        const uint32_t sub = (b >> (q * V)) & ((1u << V) - 1u);
        if (sub == (1u << V) - 1u) continue;
        uint4 wv = *reinterpret_cast<const uint4*>(wp + q * V);
        uint32_t* wr = reinterpret_cast<uint32_t*>(&wv);
        if (sizeof(T) == 4) {
#pragma unroll
          for (int e = 0; e < V; ++e) if (!((sub >> e) & 1u)) wr[e] = 0u;
        } else {
#pragma unroll
          for (int e = 0; e < V; ++e) if (!((sub >> e) & 1u)) wr[e / 2] &= (e & 1) ? 0x0000ffffu : 0xffff0000u;
        }
        st_stream(wp + q * V, wv);
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
mask_apply_packed_kernel(T* __restrict__ W, int64_t ldw, int R, int C, const uint8_t* __restrict__ bits, int64_t ldb,
                         int rows_per_seg, int64_t seg_stride, uint8_t* __restrict__ keep, int64_t ldm, int zero_w) {
  mask_apply_packed_body<T>(W, ldw, R, C, bits, ldb, rows_per_seg, seg_stride, keep, ldm, zero_w);
}
template <typename T>
__global__ void __launch_bounds__(256) mask_apply_packed_batch_kernel(const __grid_constant__ MpApplyBatch b, int zero_w) {
  const vlmc_apply_item& p = b.it[blockIdx.y];
  mask_apply_packed_body<T>(reinterpret_cast<T*>(p.W), p.ldw, p.R, p.C, p.bits, p.ldb, p.rows_per_seg, p.seg_stride,
                            p.keep_mask, p.ldm, zero_w);
}

static int pack_item_check(const uint8_t* keep_mask, int R, int C, int64_t ldm, const uint8_t* bits, int64_t ldb) {
  if (!keep_mask || !bits || R < 1 || C < 1 || ldm < C || ldb < C / 8) return VLMC_ERR_BAD_ARG;
  if ((C & 15) || (ldm & 15) || (ldb & 1) || ((uintptr_t)keep_mask & 15) || ((uintptr_t)bits & 1)) return VLMC_ERR_UNSUPPORTED;
  if (!is_device_ptr(keep_mask) || !is_device_ptr(bits)) return VLMC_ERR_NOT_DEVICE;
  return VLMC_OK;
}

static int apply_item_check(void* W, int dtype, int R, int C, int64_t ldw, const uint8_t* bits, int64_t ldb, int& rows_per_seg,
                            int64_t& seg_stride, uint8_t* keep_mask, int64_t ldm, int zero_w) {
  if (!bits || R < 1 || C < 1 || ldb < C / 8 || (!W && zero_w) || (!keep_mask && !zero_w)) return VLMC_ERR_BAD_ARG;
  if (zero_w && ldw < C) return VLMC_ERR_BAD_ARG;
  if (keep_mask && ldm < C) return VLMC_ERR_BAD_ARG;
  if (dtype != VLMC_F32 && dtype != VLMC_F16 && dtype != VLMC_BF16) return VLMC_ERR_BAD_ARG;
  const int V = dtype == VLMC_F32 ? 4 : 8;
  if (rows_per_seg <= 0) { rows_per_seg = R; seg_stride = 0; }
  if ((C & 15) || (ldb & 1) || (seg_stride & 1) || ((uintptr_t)bits & 1)) return VLMC_ERR_UNSUPPORTED;
  if (zero_w && ((ldw % V) || ((uintptr_t)W & 15))) return VLMC_ERR_UNSUPPORTED;
  if (keep_mask && ((ldm & 15) || ((uintptr_t)keep_mask & 15))) return VLMC_ERR_UNSUPPORTED;
  if (!is_device_ptr(bits) || (W && !is_device_ptr(W)) || (keep_mask && !is_device_ptr(keep_mask))) return VLMC_ERR_NOT_DEVICE;
  return VLMC_OK;
}

}  // namespace vlmc

extern "C" int vlmc_mask_pack_batch(const vlmc_pack_item* items, int count, void* stream) {
  using namespace vlmc;
  if (!items || count < 1 || count > kMpBatchMax) return VLMC_ERR_BAD_ARG;
  MpPackBatch b;
  int64_t nmax = 0;
  for (int i = 0; i < count; ++i) {
    const vlmc_pack_item& p = items[i];
    int rc = pack_item_check(p.keep_mask, p.R, p.C, p.ldm, p.bits, p.ldb);
    if (rc) return rc;
    b.it[i] = p;
    const int64_t n = (int64_t)p.R * (p.C >> 4);
    nmax = n > nmax ? n : nmax;
  }
  int gx = (int)((nmax + 255) / 256);
  const int cap = (kNumSMs * 8 + count - 1) / count;
  if (gx > cap) gx = cap;
  mask_pack_batch_kernel<<<dim3(gx, count), 256, 0, (cudaStream_t)stream>>>(b);
  return check_launch();
}

extern "C" int vlmc_mask_apply_packed_batch(const vlmc_apply_item* items, int count, int dtype, int zero_w, void* stream) {
  using namespace vlmc;
  if (!items || count < 1 || count > kMpBatchMax) return VLMC_ERR_BAD_ARG;
  MpApplyBatch b;
  int64_t nmax = 0;
  for (int i = 0; i < count; ++i) {
    vlmc_apply_item p = items[i];
    int rc = apply_item_check(p.W, dtype, p.R, p.C, p.ldw, p.bits, p.ldb, p.rows_per_seg, p.seg_stride, p.keep_mask, p.ldm, zero_w);
    if (rc) return rc;
    b.it[i] = p;
    const int64_t n = (int64_t)p.R * (p.C >> 4);
    nmax = n > nmax ? n : nmax;
  }
  int gx = (int)((nmax + 255) / 256);
  const int cap = (kNumSMs * 8 + count - 1) / count;
  if (gx > cap) gx = cap;
  cudaStream_t st = (cudaStream_t)stream;
  VLMC_DISPATCH_DTYPE(dtype, (mask_apply_packed_batch_kernel<scalar_t><<<dim3(gx, count), 256, 0, st>>>(b, zero_w)));
  return check_launch();
}

extern "C" int vlmc_mask_pack(const uint8_t* keep_mask, int R, int C, int64_t ldm, uint8_t* bits, int64_t ldb,
                              void* stream) {
  using namespace vlmc;
  if (!keep_mask || !bits || R < 1 || C < 1 || ldm < C || ldb < C / 8) return VLMC_ERR_BAD_ARG;
  if ((C & 15) || (ldm & 15) || (ldb & 1) || ((uintptr_t)keep_mask & 15) || ((uintptr_t)bits & 1)) return VLMC_ERR_UNSUPPORTED;
  if (!is_device_ptr(keep_mask) || !is_device_ptr(bits)) return VLMC_ERR_NOT_DEVICE;
  const int64_t n = (int64_t)R * (C >> 4);
  int grid = (int)((n + 255) / 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  mask_pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(keep_mask, ldm, R, C, bits, ldb);
  return check_launch();
}

extern "C" int vlmc_mask_apply_packed(void* W, int dtype, int R, int C, int64_t ldw, const uint8_t* bits, int64_t ldb,
                                      int rows_per_seg, int64_t seg_stride, uint8_t* keep_mask, int64_t ldm, int zero_w,
                                      void* stream) {
  using namespace vlmc;
  if (!bits || R < 1 || C < 1 || ldb < C / 8 || (!W && zero_w) || (!keep_mask && !zero_w)) return VLMC_ERR_BAD_ARG;
  if (zero_w && ldw < C) return VLMC_ERR_BAD_ARG;
  if (keep_mask && ldm < C) return VLMC_ERR_BAD_ARG;
  if (dtype != VLMC_F32 && dtype != VLMC_F16 && dtype != VLMC_BF16) return VLMC_ERR_BAD_ARG;
  const int V = dtype == VLMC_F32 ? 4 : 8;
  if (rows_per_seg <= 0) { rows_per_seg = R; seg_stride = 0; }
  if ((C & 15) || (ldb & 1) || (seg_stride & 1) || ((uintptr_t)bits & 1)) return VLMC_ERR_UNSUPPORTED;
  if (zero_w && ((ldw % V) || ((uintptr_t)W & 15))) return VLMC_ERR_UNSUPPORTED;
  if (keep_mask && ((ldm & 15) || ((uintptr_t)keep_mask & 15))) return VLMC_ERR_UNSUPPORTED;
  if (!is_device_ptr(bits) || (W && !is_device_ptr(W)) || (keep_mask && !is_device_ptr(keep_mask))) return VLMC_ERR_NOT_DEVICE;
  const int64_t n = (int64_t)R * (C >> 4);
  int grid = (int)((n + 255) / 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  cudaStream_t st = (cudaStream_t)stream;
  VLMC_DISPATCH_DTYPE(dtype, (mask_apply_packed_kernel<scalar_t><<<grid, 256, 0, st>>>(
                                 reinterpret_cast<scalar_t*>(W), ldw, R, C, bits, ldb, rows_per_seg, seg_stride, keep_mask,
                                 ldm, zero_w)));
  return check_launch();
}
