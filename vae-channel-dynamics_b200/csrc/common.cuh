// common.cuh — shared helpers for libvcd_b200 (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vcd.h"

typedef __nv_bfloat16 bf16;

void vcd_set_error(const char* fmt, ...);

#define VCD_CHECK_ARG(cond, ...)       \
  do {                                 \
    if (!(cond)) {                     \
      vcd_set_error(__VA_ARGS__);      \
      return -1;                       \
    }                                  \
  } while (0)

#define VCD_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t e__ = (expr);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      vcd_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
      return -2;                                                                    \
    }                                                                               \
  } while (0)

#define VCD_LAUNCH_CHECK()                                                               \
  do {                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) {                                                            \
      vcd_set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__));  \
      return -3;                                                                         \
    }                                                                                    \
  } while (0)

static inline cudaStream_t as_stream(vcd_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
int vcd_num_sms();

// ---- device helpers ---------------------------------------------------------------
struct alignas(16) bf16x8 {
  __nv_bfloat162 v[4];
};

__device__ __forceinline__ void unpack8(const bf16x8& p, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(p.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
  bf16x8 p;
#pragma unroll
  for (int i = 0; i < 4; ++i) p.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return p;
}
__device__ __forceinline__ bf16x8 ld8(const bf16* ptr) { return *reinterpret_cast<const bf16x8*>(ptr); }
__device__ __forceinline__ void st8(bf16* ptr, const bf16x8& v) { *reinterpret_cast<bf16x8*>(ptr) = v; }

__device__ __forceinline__ float load_param(const void* p, int dtype, int i) {
  return dtype == VCD_F32 ? reinterpret_cast<const float*>(p)[i]
                          : __bfloat162float(reinterpret_cast<const bf16*>(p)[i]);
}
__device__ __forceinline__ void store_param(void* p, int dtype, int64_t i, float v) {
  if (dtype == VCD_F32) reinterpret_cast<float*>(p)[i] = v;
  else reinterpret_cast<bf16*>(p)[i] = __float2bfloat16_rn(v);
}
__device__ __forceinline__ float silu_f(float y) { return y / (1.f + __expf(-y)); }
__device__ __forceinline__ float silu_grad_f(float y) {
  float s = 1.f / (1.f + __expf(-y));
  return s * (1.f + y * (1.f - s));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// float atomic max for non-negative values (bit pattern order == value order)
__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) {
  atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
}
