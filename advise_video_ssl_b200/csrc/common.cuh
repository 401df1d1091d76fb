// Shared helpers for the sm_100a kernels behind include/avssl_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "avssl_b200.h"

namespace avssl {

void set_error(const char* fmt, ...);

#define AVSSL_REQUIRE(cond, code, ...)      \
  do {                                      \
    if (!(cond)) {                          \
      ::avssl::set_error(__VA_ARGS__);      \
      return (code);                        \
    }                                       \
  } while (0)

#define AVSSL_CUDA_OK(expr)                                                         \
  do {                                                                              \
    cudaError_t e__ = (expr);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      ::avssl::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),  \
                         __FILE__, __LINE__);                                       \
      return AVSSL_ERR_CUDA;                                                        \
    }                                                                               \
  } while (0)

#define AVSSL_LAUNCH_OK(name)                                                        \
  do {                                                                               \
    cudaError_t e__ = cudaGetLastError();                                            \
    if (e__ != cudaSuccess) {                                                        \
      ::avssl::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__)); \
      return AVSSL_ERR_CUDA;                                                         \
    }                                                                                \
  } while (0)

int sm_count();  // cached per process; <0 on error

// cudaFuncSetAttribute() applies to the CURRENT device only: kernels that opt in to large dynamic
// shared memory keep a bitmask of the device ordinals already configured (per kernel instantiation).
// Returns true the first time it is called for the current device.
inline bool first_use_on_device(unsigned long long& mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return true;
  const unsigned long long bit = 1ull << (dev & 63);
  if (mask & bit) return false;
  mask |= bit;
  return true;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Sum of squares of one row by ONE warp, fixed order (lane l takes columns l, l+32, ...): every kernel
// that normalises the same row obtains the same bits.
__device__ __forceinline__ float row_sumsq(const float* __restrict__ x, int D, int lane) {
  float ss = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float v = x[c];
    ss = fmaf(v, v, ss);
  }
  return warp_sum(ss);
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32). `scratch` >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (lane < nw) ? scratch[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

// Streaming 128-bit global accesses (data touched once per step).
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

}  // namespace avssl
