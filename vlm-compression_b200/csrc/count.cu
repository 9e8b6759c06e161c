// K17: how much of the model is left - the number of non-zero elements of every parameter, many tensors per launch.
//
// Replaces evaluate_old.py:331-334
//   distilled_total_size = sum((param != 0).float().sum() for param in model.parameters())
// (per parameter a bool temporary, a float temporary and a reduction: 3 launches and ~9 B / element; ~1000 parameters).
// Here a launch takes up to 64 tensors, walks them as one list of 16 K-element units with a grid-stride loop, and adds
// integer counts (exact, order independent): 2 B / element for 16-bit weights.  `x != 0` as torch evaluates it:
// -0.0 is zero, NaN is not.
#include "common.cuh"

namespace vlmc {

constexpr int kCntMax = 64;
constexpr int kCntThreads = 256;
constexpr int64_t kCntUnit = 16384;          // elements per work unit

struct CountBatch {
  const void* ptr[kCntMax];
  int64_t numel[kCntMax];
  int64_t unit_begin[kCntMax + 1];
  int count;
  int esize;                                  // 2 or 4 bytes per element
};

template <int ES>
__device__ __forceinline__ uint32_t count_word(uint32_t w) {
  if (ES == 4) return (w & 0x7fffffffu) != 0u ? 1u : 0u;
  return ((w & 0x7fffu) != 0u ? 1u : 0u) + ((w & 0x7fff0000u) != 0u ? 1u : 0u);
}

template <int ES>
__global__ void __launch_bounds__(kCntThreads)
count_nonzero_kernel(const __grid_constant__ CountBatch b, unsigned long long* __restrict__ out) {
  __shared__ uint32_t red[kCntThreads / 32];
  const int64_t total = b.unit_begin[b.count];
  for (int64_t unit = blockIdx.x; unit < total; unit += gridDim.x) {
    int p = 0;
    while (unit >= b.unit_begin[p + 1]) ++p;
    const int64_t e0 = (unit - b.unit_begin[p]) * kCntUnit;
    const int64_t e1 = e0 + kCntUnit < b.numel[p] ? e0 + kCntUnit : b.numel[p];
    const char* base = reinterpret_cast<const char*>(b.ptr[p]);
    uint32_t cnt = 0;
    // 16-byte vectors where the address allows, single elements at the ragged ends
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(base) + (uintptr_t)e0 * ES;
    const uintptr_t a1 = reinterpret_cast<uintptr_t>(base) + (uintptr_t)e1 * ES;
    uintptr_t v0 = (a0 + 15) & ~(uintptr_t)15, v1 = a1 & ~(uintptr_t)15;
    if (v0 > v1) { v0 = a1; v1 = a1; }
    for (uintptr_t a = a0 + (uintptr_t)threadIdx.x * ES; a < v0; a += (uintptr_t)kCntThreads * ES) {
      if (ES == 4) cnt += count_word<4>(*reinterpret_cast<const uint32_t*>(a));
      else cnt += (*reinterpret_cast<const uint16_t*>(a) & 0x7fffu) != 0 ? 1u : 0u;
    }
    for (uintptr_t a = v0 + (uintptr_t)threadIdx.x * 16; a < v1; a += (uintptr_t)kCntThreads * 16) {
      const uint4 v = ld_stream(reinterpret_cast<const void*>(a));
      cnt += count_word<ES>(v.x) + count_word<ES>(v.y) + count_word<ES>(v.z) + count_word<ES>(v.w);
    }
    for (uintptr_t a = v1 + (uintptr_t)threadIdx.x * ES; a < a1; a += (uintptr_t)kCntThreads * ES) {
      if (ES == 4) cnt += count_word<4>(*reinterpret_cast<const uint32_t*>(a));
      else cnt += (*reinterpret_cast<const uint16_t*>(a) & 0x7fffu) != 0 ? 1u : 0u;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t s = 0;
#pragma unroll
      for (int w = 0; w < kCntThreads / 32; ++w) s += red[w];
      if (s) atomicAdd(out + p, (unsigned long long)s);
    }
    __syncthreads();
  }
}

}  // namespace vlmc

extern "C" int vlmc_count_nonzero_batch(const vlmc_tensor_item* items, int count, int dtype, unsigned long long* out,
                                        void* stream) {
  using namespace vlmc;
  if (!items || !out || count < 1 || count > kCntMax) return VLMC_ERR_BAD_ARG;
  if (dtype != VLMC_F32 && dtype != VLMC_F16 && dtype != VLMC_BF16) return VLMC_ERR_BAD_ARG;
  if (!is_device_ptr(out)) return VLMC_ERR_NOT_DEVICE;
  CountBatch b;
  b.count = count;
  b.esize = elem_size(dtype);
  b.unit_begin[0] = 0;
  for (int i = 0; i < count; ++i) {
    if (items[i].numel < 0 || (items[i].numel > 0 && !items[i].ptr)) return VLMC_ERR_BAD_ARG;
    if (items[i].numel > 0 && !is_device_ptr(items[i].ptr)) return VLMC_ERR_NOT_DEVICE;
    if (((uintptr_t)items[i].ptr & (uintptr_t)(b.esize - 1)) != 0) return VLMC_ERR_UNSUPPORTED;
    b.ptr[i] = items[i].ptr;
    b.numel[i] = items[i].numel;
    b.unit_begin[i + 1] = b.unit_begin[i] + (items[i].numel + kCntUnit - 1) / kCntUnit;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(out, 0, (size_t)count * sizeof(unsigned long long), st) != cudaSuccess) return check_launch();
  const int64_t total = b.unit_begin[count];
  if (total == 0) return VLMC_OK;
  int grid = kNumSMs * 8;
  if (grid > total) grid = (int)total;
  if (b.esize == 4) count_nonzero_kernel<4><<<grid, kCntThreads, 0, st>>>(b, out);
  else count_nonzero_kernel<2><<<grid, kCntThreads, 0, st>>>(b, out);
  return check_launch();
}
