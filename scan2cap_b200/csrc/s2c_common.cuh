// Shared helpers of libs2c.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/s2c.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libs2c is written for sm_100a (B200) only"
#endif

namespace s2c {

// thread-local error string behind s2c_last_error()
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define S2C_REQUIRE(cond, ...)               \
  do {                                       \
    if (!(cond)) {                           \
      ::s2c::set_error(__VA_ARGS__);         \
      return S2C_ERR_INVALID_ARGUMENT;       \
    }                                        \
  } while (0)

#define S2C_CHECK_LAUNCH(what)                                   \
  do {                                                           \
    cudaError_t e__ = cudaGetLastError();                        \
    if (e__ != cudaSuccess) return ::s2c::cuda_fail(e__, what);  \
  } while (0)

#define S2C_CUDA(call, what)                                     \
  do {                                                           \
    cudaError_t e__ = (call);                                    \
    if (e__ != cudaSuccess) return ::s2c::cuda_fail(e__, what);  \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// The one distance expression every reference search kernel compiles to (SASS of
// ball_query_gpu.cu:30-31, sampling_gpu.cu:103-104, interpolate_gpu.cu:33 built by nvcc 12.9):
//   (ax-bx)^2 + (ay-by)^2 + (az-bz)^2  ->  FMUL(dy,dy); FFMA(dx,dx,.); FFMA(dz,dz,.)
__device__ __forceinline__ float sqdist3(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

__device__ __forceinline__ float sqnorm3(float x, float y, float z) {
  return __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
}

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// streaming (read-once / write-once) accesses: keep them out of L1
__device__ __forceinline__ float ld_stream(const float *p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream(float *p, float v) {
  asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void st_stream4(float *p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

}  // namespace s2c
