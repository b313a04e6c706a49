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
// eight bf16 values moved as ONE 128-bit access (LDG.E.128 / STG.E.128; a struct of bfloat162 is split
// into four 32-bit accesses by nvcc, which quarters the bytes per L2 request)
struct alignas(16) bf16x8 {
  uint4 u;
};

__device__ __forceinline__ float2 bf2_to_f2(uint32_t w) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&w);
  return __bfloat1622float2(h);
}
__device__ __forceinline__ uint32_t f2_to_bf2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void unpack8(const bf16x8& p, float* f) {
  float2 t;
  t = bf2_to_f2(p.u.x); f[0] = t.x; f[1] = t.y;
  t = bf2_to_f2(p.u.y); f[2] = t.x; f[3] = t.y;
  t = bf2_to_f2(p.u.z); f[4] = t.x; f[5] = t.y;
  t = bf2_to_f2(p.u.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
  bf16x8 p;
  p.u = make_uint4(f2_to_bf2(f[0], f[1]), f2_to_bf2(f[2], f[3]), f2_to_bf2(f[4], f[5]), f2_to_bf2(f[6], f[7]));
  return p;
}
__device__ __forceinline__ bf16x8 ld8(const bf16* ptr) {
  bf16x8 p;
  p.u = *reinterpret_cast<const uint4*>(ptr);
  return p;
}
__device__ __forceinline__ void st8(bf16* ptr, const bf16x8& v) { *reinterpret_cast<uint4*>(ptr) = v.u; }

__device__ __forceinline__ float load_param(const void* p, int dtype, int i) {
  return dtype == VCD_F32 ? reinterpret_cast<const float*>(p)[i]
                          : __bfloat162float(reinterpret_cast<const bf16*>(p)[i]);
}
__device__ __forceinline__ void store_param(void* p, int dtype, int64_t i, float v) {
  if (dtype == VCD_F32) reinterpret_cast<float*>(p)[i] = v;
  else reinterpret_cast<bf16*>(p)[i] = __float2bfloat16_rn(v);
}
// sigmoid through one MUFU op: sigmoid(y) = 0.5 * tanh(0.5 y) + 0.5  (tanh.approx: ~2^-11 rel. error, far inside
// the bf16 output precision; the precise form costs an EX2, an RCP and a multi-instruction division)
__device__ __forceinline__ float sigmoid_fast(float y) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * y));
  return fmaf(t, 0.5f, 0.5f);
}
__device__ __forceinline__ float silu_f(float y) { return y * sigmoid_fast(y); }
__device__ __forceinline__ float silu_grad_f(float y) {
  float s = sigmoid_fast(y);
  return s * fmaf(y, 1.f - s, 1.f);
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
