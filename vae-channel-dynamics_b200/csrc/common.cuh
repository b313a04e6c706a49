// common.cuh — shared helpers for libvcd_b200 (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <utility>

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
// per-DEVICE one-time flags (cudaFuncSetAttribute applies to the current device's context only): slot in [0, 16)
bool* vcd_device_once(int slot);

// ---- programmatic dependent launch ---------------------------------------------------------------------------------
// vcd_launch(..., overlap_prev = true) launches with cudaLaunchAttributeProgrammaticStreamSerialization: the grid may start
// while the kernel enqueued before it in the same stream is still running — as soon as every CTA of that kernel has executed
// vcd_pdl_trigger() (griddepcontrol.launch_dependents) or exited — and its CTAs then take the SMs the earlier kernel's tail
// leaves idle.  The launched kernel never executes griddepcontrol.wait, so it must not read anything the earlier kernel
// writes nor write anything it reads; everything enqueued before THAT kernel is complete, and any later plain launch waits
// for both.  Used for weight-gradient GEMMs behind the data-gradient GEMM of the same layer (VCD_WGRAD_OVERLAP_PREV).
template <typename... P, typename... A>
static inline cudaError_t vcd_launch(void (*kernel)(P...), int grid, int block, size_t smem, cudaStream_t st,
                                     bool overlap_prev, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = overlap_prev ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<A>(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void vcd_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

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

// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 / FMUL2) -------------------------------------------------------
// Two fp32 lanes per instruction on the FMA pipe: the streaming kernels (GroupNorm, statistics, optimizer) spend their
// issue slots on per-element fp32 math, and the packed forms halve that.  A pair lives in a 64-bit register
// (lo = first element, hi = second).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f32x2 dup2(float v) { return pk2(v, v); }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 neg2(f32x2 a) { return a ^ 0x8000000080000000ull; }
__device__ __forceinline__ f32x2 abs2(f32x2 a) { return a & 0x7fffffff7fffffffull; }
__device__ __forceinline__ f32x2 max2(f32x2 a, f32x2 b) {
  float al, ah, bl, bh;
  upk2(a, al, ah);
  upk2(b, bl, bh);
  return pk2(fmaxf(al, bl), fmaxf(ah, bh));
}
__device__ __forceinline__ f32x2 tanh2(f32x2 a) {
  float lo, hi;
  upk2(a, lo, hi);
  asm("tanh.approx.f32 %0, %0;" : "+f"(lo));
  asm("tanh.approx.f32 %0, %0;" : "+f"(hi));
  return pk2(lo, hi);
}
// one 32-bit word holding two bf16 (first element in the low half) <-> an fp32 pair
__device__ __forceinline__ f32x2 bf2_to_f32x2(uint32_t w) {
  f32x2 r;
  uint32_t lo = w << 16, hi = w & 0xffff0000u;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ uint32_t f32x2_to_bf2(f32x2 v) {
  float lo, hi;
  upk2(v, lo, hi);
  return f2_to_bf2(lo, hi);
}
__device__ __forceinline__ void unpack8x(const bf16x8& p, f32x2* f) {
  f[0] = bf2_to_f32x2(p.u.x);
  f[1] = bf2_to_f32x2(p.u.y);
  f[2] = bf2_to_f32x2(p.u.z);
  f[3] = bf2_to_f32x2(p.u.w);
}
__device__ __forceinline__ bf16x8 pack8x(const f32x2* f) {
  bf16x8 p;
  p.u = make_uint4(f32x2_to_bf2(f[0]), f32x2_to_bf2(f[1]), f32x2_to_bf2(f[2]), f32x2_to_bf2(f[3]));
  return p;
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
