// Shared helpers for the vlmc kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/vlmc.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "vlmc kernels are written for sm_100a (B200) only"
#endif

namespace vlmc {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

extern thread_local int g_last_cuda_error;

inline int check_launch() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    g_last_cuda_error = (int)e;
    return VLMC_ERR_CUDA;
  }
  return VLMC_OK;
}

// Rejects host pointers: there is no CPU path behind this ABI.
inline bool is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- element traits: 16-byte vectors of the activation / weight dtype -------------------
template <typename T> struct Elem;
template <> struct Elem<float> {
  static constexpr int kVec = 4;  // elements per 16 B
  __device__ static __forceinline__ void unpack(const uint4& v, float* f) {
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y);
    f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
  }
  __device__ static __forceinline__ uint4 pack(const float* f) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]),
                      __float_as_uint(f[2]), __float_as_uint(f[3]));
  }
};
template <> struct Elem<__half> {
  static constexpr int kVec = 8;
  __device__ static __forceinline__ void unpack(const uint4& v, float* f) {
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __half22float2(h[i]);
      f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
  }
  __device__ static __forceinline__ uint4 pack(const float* f) {
    uint4 v;
    __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    return v;
  }
};
template <> struct Elem<__nv_bfloat16> {
  static constexpr int kVec = 8;
  __device__ static __forceinline__ void unpack(const uint4& v, float* f) {
    // bf16 -> fp32 is a 16-bit shift
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static __forceinline__ uint4 pack(const float* f) {
    uint4 v;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return v;
  }
};

// streaming 16-byte load that does not pollute L1 (data is touched once)
__device__ __forceinline__ uint4 ld_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream8(void* p, const uint2& v) {
  asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};"
               :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void st_stream4(void* p, uint32_t v) {
  asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

#define VLMC_DISPATCH_DTYPE(dtype, ...)                                   \
  switch (dtype) {                                                        \
    case VLMC_F32: { using scalar_t = float; __VA_ARGS__; break; }        \
    case VLMC_F16: { using scalar_t = __half; __VA_ARGS__; break; }       \
    case VLMC_BF16: { using scalar_t = __nv_bfloat16; __VA_ARGS__; break; } \
    default: return VLMC_ERR_BAD_ARG;                                     \
  }

inline int elem_size(int dtype) { return dtype == VLMC_F32 ? 4 : 2; }

}  // namespace vlmc
