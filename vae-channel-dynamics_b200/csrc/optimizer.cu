// optimizer.cu — fused multi-tensor gradient-norm + clip + AdamW.
//
// Replaces, for reference src/train.py:184-187,301-302: torch.nn.utils.clip_grad_norm_(params, max_norm) followed by
// torch.optim.AdamW.step() ([upstream] torch foreach implementation: ~25 multi_tensor_apply launches, ~1.7 GB of HBM
// traffic in ~10 passes over 84 M parameters) with TWO launches: one pass over the gradients (sum of squares, fp64)
// and one pass that applies the clip coefficient, the decoupled weight decay and the Adam update:
//     g   = grad * min(1, max_norm / (||grad||_2 + 1e-6))
//     p   = p * (1 - lr * wd)
//     m   = m + (1 - b1) * (g - m)              (torch lerp)
//     v   = b2 * v + (1 - b2) * g * g
//     p   = p - (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// Parameters and gradients fp32 or bf16 (as train.py:150-154 loads them), moments always fp32.
// HBM-bound: bf16 parameters move 2 (g) + 2+2 (p) + 8+8 (m, v) = 22 bytes per parameter and step.
#include "common.cuh"

namespace {

constexpr int kChunk = 8192;   // elements per block: 256 threads x 8 elements x 4 iterations
constexpr int kThreads = 256;

__device__ __forceinline__ void load8(const void* base, int dt, int64_t i, int64_t n, float* f) {
  // 8 consecutive elements starting at i (i % 8 == 0); vector access when the address allows it, else scalar
  if (dt == VCD_BF16) {
    const bf16* p = reinterpret_cast<const bf16*>(base) + i;
    if (i + 8 <= n && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
      unpack8(ld8(p), f);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (i + j < n) ? __bfloat162float(p[j]) : 0.f;
    }
  } else {
    const float* p = reinterpret_cast<const float*>(base) + i;
    if (i + 8 <= n && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
      const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
      f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (i + j < n) ? p[j] : 0.f;
    }
  }
}
__device__ __forceinline__ void store8(void* base, int dt, int64_t i, int64_t n, const float* f) {
  if (dt == VCD_BF16) {
    bf16* p = reinterpret_cast<bf16*>(base) + i;
    if (i + 8 <= n && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
      st8(p, pack8(f));
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (i + j < n) p[j] = __float2bfloat16_rn(f[j]);
    }
  } else {
    float* p = reinterpret_cast<float*>(base) + i;
    if (i + 8 <= n && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
      *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (i + j < n) p[j] = f[j];
    }
  }
}

// chunk c of the launch covers elements [chunk_off[c], chunk_off[c] + kChunk) of tensor chunk_tensor[c]
__global__ void __launch_bounds__(kThreads) multi_sqnorm_kernel(const void* const* __restrict__ grads,
                                                               const int64_t* __restrict__ numels,
                                                               const int32_t* __restrict__ dtypes,
                                                               const int32_t* __restrict__ chunk_tensor,
                                                               const int64_t* __restrict__ chunk_off,
                                                               double* __restrict__ out) {
  const int t = chunk_tensor[blockIdx.x];
  const int64_t n = numels[t], base = chunk_off[blockIdx.x];
  const void* g = grads[t];
  const int dt = dtypes[t];
  float acc = 0.f;
  if (g != nullptr) {
    for (int64_t i = base + (int64_t)threadIdx.x * 8; i < min(base + kChunk, n); i += (int64_t)kThreads * 8) {
      float f[8];
      load8(g, dt, i, n, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(f[j], f[j], acc);
    }
  }
  double d = warp_sum_d((double)acc);
  __shared__ double red[kThreads / 32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) s += red[i];
    if (s != 0.0) atomicAdd(out, s);
  }
}

__global__ void __launch_bounds__(kThreads) clip_adamw_kernel(void* const* __restrict__ params,
                                                             const void* const* __restrict__ grads,
                                                             float* const* __restrict__ exp_avg,
                                                             float* const* __restrict__ exp_avg_sq,
                                                             const int64_t* __restrict__ numels,
                                                             const int32_t* __restrict__ dtypes,
                                                             const int32_t* __restrict__ chunk_tensor,
                                                             const int64_t* __restrict__ chunk_off,
                                                             const double* __restrict__ grad_sqnorm, float max_norm,
                                                             float decay, float b1, float b2, float w1, float w2,
                                                             float eps, double lr, double beta1, double beta2,
                                                             const int32_t* __restrict__ steps, int64_t step) {
  const int t = chunk_tensor[blockIdx.x];
  const void* g = grads[t];
  if (g == nullptr) return;                    // parameter without a gradient this step: untouched (torch skips it)
  const int64_t n = numels[t], base = chunk_off[blockIdx.x];
  void* p = params[t];
  float* m = exp_avg[t];
  float* v = exp_avg_sq[t];
  const int dt = dtypes[t];
  float coef = 1.f;
  if (grad_sqnorm != nullptr && max_norm > 0.f) {
    const float total = (float)sqrt(*grad_sqnorm);
    coef = fminf(max_norm / (total + 1e-6f), 1.f);     // torch.nn.utils.clip_grad_norm_
  }
  // bias correction for THIS tensor's own step count (torch keeps state['step'] per parameter: a parameter that had no
  // gradient in some step lags behind), in fp64 like torch's python scalars
  const double ts = (double)(steps != nullptr ? (int64_t)steps[t] : step);
  const float step_size = (float)(lr / (1.0 - pow(beta1, ts)));
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow(beta2, ts)));
  for (int64_t i = base + (int64_t)threadIdx.x * 8; i < min(base + kChunk, n); i += (int64_t)kThreads * 8) {
    float fp[8], fg[8], fm[8], fv[8];
    load8(p, dt, i, n, fp);
    load8(g, dt, i, n, fg);
    load8(m, VCD_F32, i, n, fm);
    load8(v, VCD_F32, i, n, fv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gg = fg[j] * coef;
      const float pp = fp[j] * decay;
      fm[j] = fmaf(w1, gg - fm[j], fm[j]);
      fv[j] = fmaf(w2 * gg, gg, fv[j] * b2);
      const float denom = fmaf(sqrtf(fv[j]), inv_sqrt_bc2, eps);
      fp[j] = pp - step_size * (fm[j] / denom);
    }
    store8(p, dt, i, n, fp);
    store8(m, VCD_F32, i, n, fm);
    store8(v, VCD_F32, i, n, fv);
  }
}

}  // namespace

extern "C" int vcd_optim_chunk_elems(void) { return kChunk; }

extern "C" int vcd_multi_sqnorm(const void* const* grads, const int64_t* numels, const int32_t* dtypes,
                                const int32_t* chunk_tensor, const int64_t* chunk_off, int n_chunks, double* out_sqnorm,
                                vcd_stream_t stream) {
  VCD_CHECK_ARG(n_chunks > 0, "vcd_multi_sqnorm: no chunks");
  VCD_CUDA(cudaMemsetAsync(out_sqnorm, 0, sizeof(double), as_stream(stream)));
  multi_sqnorm_kernel<<<n_chunks, kThreads, 0, as_stream(stream)>>>(grads, numels, dtypes, chunk_tensor, chunk_off, out_sqnorm);
  VCD_LAUNCH_CHECK();
  return 0;
}

extern "C" int vcd_clip_adamw_step(void* const* params, const void* const* grads, float* const* exp_avg,
                                   float* const* exp_avg_sq, const int64_t* numels, const int32_t* dtypes,
                                   const int32_t* chunk_tensor, const int64_t* chunk_off, int n_chunks,
                                   const double* grad_sqnorm, double max_norm, double lr, double beta1, double beta2,
                                   double eps, double weight_decay, const int32_t* steps, int64_t step,
                                   vcd_stream_t stream) {
  VCD_CHECK_ARG(n_chunks > 0 && (steps != nullptr || step >= 1), "vcd_clip_adamw_step: bad arguments");
  clip_adamw_kernel<<<n_chunks, kThreads, 0, as_stream(stream)>>>(
      params, grads, exp_avg, exp_avg_sq, numels, dtypes, chunk_tensor, chunk_off, grad_sqnorm, (float)max_norm,
      (float)(1.0 - lr * weight_decay), (float)beta1, (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps,
      lr, beta1, beta2, steps, step);
  VCD_LAUNCH_CHECK();
  return 0;
}
