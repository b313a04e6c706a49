// tracker.cu — device side of src/tracking (monitor.py, deadneuron.py), src/classification
// (classifier.py) and src/intervention (nudger.py).
#include "common.cuh"

namespace {

// ---------------------------------------------------------------- stand-alone per-channel statistics
// channels_last: x is [N][HW][C] (bf16 or fp32).  Each block owns a pixel range, thread t owns channels
// t, t+blockDim, ...: loads are coalesced across channels.
template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ldf<bf16>(const bf16* p) { return __bfloat162float(*p); }

template <typename T>
__global__ void __launch_bounds__(256) chan_stats_cl_kernel(const T* __restrict__ x, float* __restrict__ cs,
                                                            float near_zero, int64_t pixels, int C, int64_t ppb) {
  int64_t p0 = blockIdx.x * ppb, p1 = min(p0 + ppb, pixels);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f, q = 0.f, sa = 0.f, mx = 0.f, nz = 0.f;
    for (int64_t p = p0; p < p1; ++p) {
      float v = ldf<T>(x + p * C + c);
      float a = fabsf(v);
      s += v; q += v * v; sa += a; mx = fmaxf(mx, a); nz += (a < near_zero) ? 1.f : 0.f;
    }
    atomicAdd(&cs[0 * C + c], s);
    atomicAdd(&cs[1 * C + c], q);
    atomicAdd(&cs[2 * C + c], sa);
    atomic_max_nonneg(&cs[3 * C + c], mx);
    atomicAdd(&cs[4 * C + c], nz);
  }
}
// channels_first: x is [N][C][HW]; one block per (chunk, c, n), contiguous reads along HW.
template <typename T>
__global__ void __launch_bounds__(256) chan_stats_cf_kernel(const T* __restrict__ x, float* __restrict__ cs,
                                                            float near_zero, int64_t HW, int C) {
  const int c = blockIdx.y, n = blockIdx.z;
  const T* xb = x + ((int64_t)n * C + c) * HW;
  float s = 0.f, q = 0.f, sa = 0.f, mx = 0.f, nz = 0.f;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < HW; p += (int64_t)gridDim.x * blockDim.x) {
    float v = ldf<T>(xb + p);
    float a = fabsf(v);
    s += v; q += v * v; sa += a; mx = fmaxf(mx, a); nz += (a < near_zero) ? 1.f : 0.f;
  }
  s = warp_sum(s); q = warp_sum(q); sa = warp_sum(sa); mx = warp_max(mx); nz = warp_sum(nz);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&cs[0 * C + c], s);
    atomicAdd(&cs[1 * C + c], q);
    atomicAdd(&cs[2 * C + c], sa);
    atomic_max_nonneg(&cs[3 * C + c], mx);
    atomicAdd(&cs[4 * C + c], nz);
  }
}

// ---------------------------------------------------------------- per-forward finalisation (monitor.py:64-75,101)
__global__ void __launch_bounds__(256) stats_finalize_kernel(float* __restrict__ cs, float* __restrict__ run,
                                                             double* __restrict__ scal, double n_per_channel, int C) {
  __shared__ double red[2][8];
  double ts = 0.0, tq = 0.0;
  const float inv = (float)(1.0 / n_per_channel);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = cs[0 * C + c], q = cs[1 * C + c], sa = cs[2 * C + c], mx = cs[3 * C + c], nz = cs[4 * C + c];
    ts += (double)s;
    tq += (double)q;
    float mean = s * inv;
    run[0 * C + c] += sa * inv;
    run[1 * C + c] += mean;
    run[2 * C + c] += fmaxf(q * inv - mean * mean, 0.f);
    run[3 * C + c] = fmaxf(run[3 * C + c], mx);
    run[4 * C + c] += nz * inv;
    cs[0 * C + c] = cs[1 * C + c] = cs[2 * C + c] = cs[3 * C + c] = cs[4 * C + c] = 0.f;
  }
  ts = warp_sum_d(ts);
  tq = warp_sum_d(tq);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = ts; red[1][threadIdx.x >> 5] = tq; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0, q = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { s += red[0][i]; q += red[1][i]; }
    double n = n_per_channel * (double)C;
    double mean = s / n;
    double var = n > 1.0 ? (q - s * s / n) / (n - 1.0) : 0.0;  // torch.std: unbiased (monitor.py:74-75)
    scal[0] += mean;
    scal[1] += sqrt(var > 0.0 ? var : 0.0);
    scal[2] += 1.0;
  }
}

// ---------------------------------------------------------------- classifier.py:135  (strict <, fp32)
__global__ void classify_kernel(const float* __restrict__ v, float thr, uint8_t* __restrict__ mask,
                                int32_t* __restrict__ count, int C) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  bool hit = (c < C) && (v[c] < thr);
  if (c < C) mask[c] = hit ? 1 : 0;
  unsigned b = __ballot_sync(0xffffffffu, hit);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(count, __popc(b));
}

// ---------------------------------------------------------------- nudger.py:127-143 / 162-168
__global__ void nudge_kernel(void* __restrict__ gamma, int dt, int C, const int64_t* __restrict__ idx, int n_idx,
                             double factor, double cap, int mode, int32_t* __restrict__ applied) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_idx) return;
  int64_t k = idx[i];
  if (k < 0 || k >= C) return;  // nudger.py:129 bounds check
  double v = dt == VCD_F32 ? (double)((float*)gamma)[k] : (double)__bfloat162float(((bf16*)gamma)[k]);
  double nv = mode == 1 ? 1.0 : fmin(v * factor, cap);
  float f = __double2float_rn(nv);
  if (dt == VCD_F32) ((float*)gamma)[k] = f;
  else ((bf16*)gamma)[k] = __float2bfloat16_rn(f);
  atomicAdd(applied, 1);
}

// ---------------------------------------------------------------- deadneuron.py:78-115 (multi-tensor)
__device__ __forceinline__ float ld_any(const void* p, int dt, int64_t i) {
  return dt == VCD_F32 ? ((const float*)p)[i] : __bfloat162float(((const bf16*)p)[i]);
}
__global__ void __launch_bounds__(256) dead_sum_kernel(const void* const* __restrict__ tensors,
                                                       const int64_t* __restrict__ numels,
                                                       const int32_t* __restrict__ dtypes, double* __restrict__ sums) {
  const int t = blockIdx.y;
  const void* p = tensors[t];
  const int64_t n = numels[t];
  const int dt = dtypes[t];
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    acc += (double)fabsf(ld_any(p, dt, i));
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0 && acc != 0.0) atomicAdd(&sums[t], acc);
}
__device__ __forceinline__ float round_to(double v, int dt) {
  float f = __double2float_rn(v);
  return dt == VCD_F32 ? f : __bfloat162float(__float2bfloat16_rn(f));
}
// dead_type: 0 threshold, 1 percent_of_mean, 2 both
__global__ void __launch_bounds__(256) dead_count_kernel(const void* const* __restrict__ tensors,
                                                         const int64_t* __restrict__ numels,
                                                         const int32_t* __restrict__ dtypes,
                                                         const double* __restrict__ sums, double threshold,
                                                         double mean_pct, int dead_type,
                                                         unsigned long long* __restrict__ counts) {
  const int t = blockIdx.y;
  const void* p = tensors[t];
  const int64_t n = numels[t];
  const int dt = dtypes[t];
  if (n == 0) return;
  // torch compares a tensor with a python scalar in the tensor's dtype
  const float thr = round_to(threshold, dt);
  const float mean_abs = round_to(sums[t] / (double)n, dt);  // param_abs.mean().item()
  const bool tiny = fabs((double)mean_abs) < 1e-9;
  const float athr = tiny ? round_to(1e-9, dt) : round_to(mean_pct * (double)mean_abs, dt);
  unsigned cnt = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float a = fabsf(ld_any(p, dt, i));
    bool fixed = a < thr, adaptive = a < athr;
    bool dead = dead_type == 0 ? fixed : (dead_type == 1 ? adaptive : (fixed && adaptive));
    cnt += dead ? 1u : 0u;
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&counts[t], (unsigned long long)cnt);
}

}  // namespace

extern "C" int vcd_chan_stats(const void* x, int x_dtype, float* chan_stats, float near_zero, int N, int HW, int C,
                              int channels_last, vcd_stream_t stream) {
  cudaStream_t st = as_stream(stream);
  if (channels_last) {
    int64_t pixels = (int64_t)N * HW;
    int64_t ppb = 64;
    unsigned grid = (unsigned)ceil_div64(pixels, ppb);
    if (x_dtype == VCD_F32) chan_stats_cl_kernel<float><<<grid, 256, 0, st>>>((const float*)x, chan_stats, near_zero, pixels, C, ppb);
    else chan_stats_cl_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, chan_stats, near_zero, pixels, C, ppb);
  } else {
    unsigned gx = (unsigned)ceil_div64(HW, 256 * 8);
    if (gx > 64) gx = 64;
    dim3 grid(gx, C, N);
    if (x_dtype == VCD_F32) chan_stats_cf_kernel<float><<<grid, 256, 0, st>>>((const float*)x, chan_stats, near_zero, HW, C);
    else chan_stats_cf_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, chan_stats, near_zero, HW, C);
  }
  VCD_LAUNCH_CHECK();
  return 0;
}

extern "C" int vcd_stats_finalize(float* chan_stats, float* run, double* scal, int64_t n_per_channel, int C,
                                  vcd_stream_t stream) {
  VCD_CHECK_ARG(n_per_channel > 0, "stats_finalize: empty forward");
  stats_finalize_kernel<<<1, 256, 0, as_stream(stream)>>>(chan_stats, run, scal, (double)n_per_channel, C);
  VCD_LAUNCH_CHECK();
  return 0;
}

extern "C" int vcd_classify_mask(const float* mean_abs, float threshold, uint8_t* mask, int32_t* count, int C,
                                 vcd_stream_t stream) {
  VCD_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t), as_stream(stream)));
  if (C > 0) {
    classify_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(mean_abs, threshold, mask, count, C);
    VCD_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int vcd_nudge_gamma(void* gamma, int dtype, int C, const int64_t* idx, int n_idx, double factor, double cap,
                               int mode, int32_t* applied_count, vcd_stream_t stream) {
  VCD_CUDA(cudaMemsetAsync(applied_count, 0, sizeof(int32_t), as_stream(stream)));
  if (n_idx > 0) {
    nudge_kernel<<<(n_idx + 127) / 128, 128, 0, as_stream(stream)>>>(gamma, dtype, C, idx, n_idx, factor, cap, mode,
                                                                     applied_count);
    VCD_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int vcd_dead_weight_count(const void* const* tensors, const int64_t* numels, const int32_t* dtypes, int T,
                                     double threshold, double mean_percentage, int dead_type, double* sum_abs_ws,
                                     int64_t* counts, vcd_stream_t stream) {
  if (T <= 0) return 0;
  cudaStream_t st = as_stream(stream);
  VCD_CUDA(cudaMemsetAsync(sum_abs_ws, 0, sizeof(double) * T, st));
  VCD_CUDA(cudaMemsetAsync(counts, 0, sizeof(int64_t) * T, st));
  dim3 grid(32, T);
  dead_sum_kernel<<<grid, 256, 0, st>>>(tensors, numels, dtypes, sum_abs_ws);
  VCD_LAUNCH_CHECK();
  dead_count_kernel<<<grid, 256, 0, st>>>(tensors, numels, dtypes, sum_abs_ws, threshold, mean_percentage, dead_type,
                                          (unsigned long long*)counts);
  VCD_LAUNCH_CHECK();
  return 0;
}
