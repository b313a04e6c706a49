// elementwise.cu — layout, activation, attention-softmax, latent-distribution and loss kernels.
// All HBM-bound: coalesced 16-byte access where the layout allows, warp-shuffle reductions.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------- weight packing
// OIHW (fp32|bf16) -> w_fprop [tap][Cout][Cin], w_dgrad [tap][Cin][Cout] (bf16), bias -> fp32
__global__ void pack_weight_kernel(const void* __restrict__ w, int dt, int Cout, int Cin, int taps,
                                   bf16* __restrict__ wf, bf16* __restrict__ wd) {
  int64_t total = (int64_t)Cout * Cin * taps;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    // i indexes the fprop pack: ((tap*Cout)+co)*Cin + ci  (coalesced writes of wf)
    int ci = (int)(i % Cin);
    int64_t r = i / Cin;
    int co = (int)(r % Cout);
    int tap = (int)(r / Cout);
    float v = load_param(w, dt, ((int64_t)co * Cin + ci) * taps + tap);
    bf16 b = __float2bfloat16_rn(v);
    wf[i] = b;
    if (wd) wd[((int64_t)tap * Cin + ci) * Cout + co] = b;
  }
}
__global__ void bias_to_f32_kernel(const void* __restrict__ b, int dt, int n, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = load_param(b, dt, i);
}

// ---------------------------------------------------------------- parity planes / upsample
// x [N][H][W][C] -> xp [N][2][2][H/2][W/2][C]; vectors of 8 channels
__global__ void space_planes_kernel(const bf16* __restrict__ x, bf16* __restrict__ xp, int N, int H, int W, int C,
                                    int to_planes) {
  const int V = C / 8;
  const int H2 = H / 2, W2 = W / 2;
  int64_t total = (int64_t)N * H * W * V;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int v = (int)(i % V);
    int64_t r = i / V;
    int w = (int)(r % W);
    r /= W;
    int h = (int)(r % H);
    int n = (int)(r / H);
    int64_t src = (((int64_t)n * H + h) * W + w) * C + v * 8;
    int64_t dst = (((((int64_t)n * 2 + (h & 1)) * 2 + (w & 1)) * H2 + (h >> 1)) * W2 + (w >> 1)) * C + v * 8;
    if (to_planes) st8(xp + dst, ld8(x + src));
    else st8(const_cast<bf16*>(x) + src, ld8(xp + dst));
  }
}
// y [N][2H][2W][C] = x [N][H][W][C] nearest
__global__ void upsample2x_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int N, int H, int W, int C) {
  const int V = C / 8;
  int64_t total = (int64_t)N * H * W * V;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int v = (int)(i % V);
    int64_t r = i / V;
    int w = (int)(r % W);
    r /= W;
    int h = (int)(r % H);
    int n = (int)(r / H);
    bf16x8 val = ld8(x + (((int64_t)n * H + h) * W + w) * C + v * 8);
    int64_t o = (((int64_t)n * 2 * H + 2 * h) * 2 * W + 2 * w) * C + v * 8;
    st8(y + o, val);
    st8(y + o + C, val);
    st8(y + o + (int64_t)2 * W * C, val);
    st8(y + o + (int64_t)2 * W * C + C, val);
  }
}
__global__ void upsample2x_bwd_kernel(const bf16* __restrict__ dy, bf16* __restrict__ dx, int N, int H, int W, int C) {
  const int V = C / 8;
  int64_t total = (int64_t)N * H * W * V;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int v = (int)(i % V);
    int64_t r = i / V;
    int w = (int)(r % W);
    r /= W;
    int h = (int)(r % H);
    int n = (int)(r / H);
    int64_t o = (((int64_t)n * 2 * H + 2 * h) * 2 * W + 2 * w) * C + v * 8;
    float a[8], b[8], c[8], d[8];
    unpack8(ld8(dy + o), a);
    unpack8(ld8(dy + o + C), b);
    unpack8(ld8(dy + o + (int64_t)2 * W * C), c);
    unpack8(ld8(dy + o + (int64_t)2 * W * C + C), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = (a[j] + b[j]) + (c[j] + d[j]);
    st8(dx + (((int64_t)n * H + h) * W + w) * C + v * 8, pack8(a));
  }
}

// ---------------------------------------------------------------- NCHW <-> NHWC (small C: 3, 4, 8)
__global__ void nchw_to_nhwc_kernel(const void* __restrict__ x, int dt, bf16* __restrict__ y, int N, int C, int64_t HW) {
  int64_t total = (int64_t)N * HW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t n = i / HW, p = i % HW;
    for (int c = 0; c < C; ++c) {
      int64_t s = (n * C + c) * HW + p;
      float v = dt == VCD_F32 ? ((const float*)x)[s] : __bfloat162float(((const bf16*)x)[s]);
      y[i * C + c] = __float2bfloat16_rn(v);
    }
  }
}
__global__ void nhwc_to_nchw_kernel(const bf16* __restrict__ x, void* __restrict__ y, int dt, int N, int C, int64_t HW) {
  int64_t total = (int64_t)N * HW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t n = i / HW, p = i % HW;
    for (int c = 0; c < C; ++c) {
      float v = __bfloat162float(x[i * C + c]);
      int64_t d = (n * C + c) * HW + p;
      if (dt == VCD_F32) ((float*)y)[d] = v;
      else ((bf16*)y)[d] = __float2bfloat16_rn(v);
    }
  }
}

__global__ void add_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ o, int64_t n8) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float x[8], y[8];
    unpack8(ld8(a + i * 8), x);
    unpack8(ld8(b + i * 8), y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] += y[j];
    st8(o + i * 8, pack8(x));
  }
}
__global__ void silu_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, bf16* __restrict__ o, int64_t n8) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float f[8], g[8];
    unpack8(ld8(x + i * 8), f);
    if (dy) {
      unpack8(ld8(dy + i * 8), g);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = g[j] * silu_grad_f(f[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = silu_f(f[j]);
    }
    st8(o + i * 8, pack8(f));
  }
}

// ---------------------------------------------------------------- softmax over rows of S (bf16), one warp per row
// cols % 8 == 0.  fwd: p = softmax(s) (s already scaled by 1/sqrt(d) in the GEMM epilogue).
// bwd: ds = scale * p * (dp - sum(dp*p)).
__global__ void __launch_bounds__(256) softmax_fwd_kernel(const bf16* __restrict__ s, bf16* __restrict__ p, int64_t rows,
                                                          int cols) {
  int64_t row = blockIdx.x * (int64_t)(blockDim.x / 32) + threadIdx.x / 32;
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const bf16* sr = s + row * cols;
  bf16* pr = p + row * cols;
  float mx = -INFINITY;
  for (int c = lane * 8; c < cols; c += 256) {
    float f[8];
    unpack8(ld8(sr + c), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) mx = fmaxf(mx, f[j]);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int c = lane * 8; c < cols; c += 256) {
    float f[8];
    unpack8(ld8(sr + c), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) sum += __expf(f[j] - mx);
  }
  sum = warp_sum(sum);
  float inv = 1.f / sum;
  for (int c = lane * 8; c < cols; c += 256) {
    float f[8];
    unpack8(ld8(sr + c), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = __expf(f[j] - mx) * inv;
    st8(pr + c, pack8(f));
  }
}
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const bf16* __restrict__ p, const bf16* __restrict__ dp,
                                                          bf16* __restrict__ ds, float scale, int64_t rows, int cols) {
  int64_t row = blockIdx.x * (int64_t)(blockDim.x / 32) + threadIdx.x / 32;
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const bf16* pr = p + row * cols;
  const bf16* dr = dp + row * cols;
  float dot = 0.f;
  for (int c = lane * 8; c < cols; c += 256) {
    float a[8], b[8];
    unpack8(ld8(pr + c), a);
    unpack8(ld8(dr + c), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) dot += a[j] * b[j];
  }
  dot = warp_sum(dot);
  for (int c = lane * 8; c < cols; c += 256) {
    float a[8], b[8];
    unpack8(ld8(pr + c), a);
    unpack8(ld8(dr + c), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = scale * a[j] * (b[j] - dot);
    st8(ds + row * cols + c, pack8(a));
  }
}

// y[b][c][r] = x[b][r][c], 32x32 smem tiles
__global__ void transpose_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int rows, int cols) {
  __shared__ bf16 t[32][33];
  const bf16* xb = x + (int64_t)blockIdx.z * rows * cols;
  bf16* yb = y + (int64_t)blockIdx.z * rows * cols;
  int c = blockIdx.x * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += 8) {
    int r = blockIdx.y * 32 + j;
    if (r < rows && c < cols) t[j][threadIdx.x] = xb[(int64_t)r * cols + c];
  }
  __syncthreads();
  int r2 = blockIdx.y * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += 8) {
    int c2 = blockIdx.x * 32 + j;
    if (r2 < rows && c2 < cols) yb[(int64_t)c2 * rows + r2] = t[threadIdx.x][j];
  }
}

// ---------------------------------------------------------------- latent distribution (L = 4)
// moments NHWC [N][hw][8] bf16 -> z NHWC [N][hw][4]; kl[n] = 0.5*sum(mean^2 + var - 1 - logvar)
__global__ void __launch_bounds__(256) gauss_fwd_kernel(const bf16* __restrict__ mom, const float* __restrict__ noise,
                                                        bf16* __restrict__ z, float* __restrict__ mean_out,
                                                        float* __restrict__ logvar_out, float* __restrict__ kl, int hw) {
  const int n = blockIdx.y;
  float acc = 0.f;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x) {
    float f[8];
    unpack8(ld8(mom + ((int64_t)n * hw + p) * 8), f);
    float zz[4];
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      float mu = f[l];
      float lv = fminf(fmaxf(f[4 + l], -30.f), 20.f);
      float e = noise ? noise[((int64_t)n * 4 + l) * hw + p] : 0.f;
      zz[l] = mu + expf(0.5f * lv) * e;
      acc += mu * mu + expf(lv) - 1.f - lv;
      if (mean_out) mean_out[((int64_t)n * 4 + l) * hw + p] = mu;
      if (logvar_out) logvar_out[((int64_t)n * 4 + l) * hw + p] = lv;
    }
    __nv_bfloat162* zo = reinterpret_cast<__nv_bfloat162*>(z + ((int64_t)n * hw + p) * 4);
    zo[0] = __floats2bfloat162_rn(zz[0], zz[1]);
    zo[1] = __floats2bfloat162_rn(zz[2], zz[3]);
  }
  acc = warp_sum(acc);
  __shared__ float ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += ws[i];
    atomicAdd(&kl[n], 0.5f * t);
  }
}
__global__ void __launch_bounds__(256) gauss_bwd_kernel(const bf16* __restrict__ mom, const float* __restrict__ noise,
                                                        const bf16* __restrict__ dz, const float* __restrict__ dkl,
                                                        bf16* __restrict__ dmom, int hw) {
  const int n = blockIdx.y;
  const float gk = dkl ? dkl[n] : 0.f;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x) {
    float f[8], o[8];
    unpack8(ld8(mom + ((int64_t)n * hw + p) * 8), f);
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    if (dz) {
      const __nv_bfloat162* zi = reinterpret_cast<const __nv_bfloat162*>(dz + ((int64_t)n * hw + p) * 4);
      float2 a = __bfloat1622float2(zi[0]), b = __bfloat1622float2(zi[1]);
      g[0] = a.x; g[1] = a.y; g[2] = b.x; g[3] = b.y;
    }
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      float mu = f[l], raw = f[4 + l];
      bool inside = (raw >= -30.f) && (raw <= 20.f);
      float lv = fminf(fmaxf(raw, -30.f), 20.f);
      float e = noise ? noise[((int64_t)n * 4 + l) * hw + p] : 0.f;
      o[l] = g[l] + gk * mu;
      o[4 + l] = inside ? (g[l] * 0.5f * expf(0.5f * lv) * e + gk * 0.5f * (expf(lv) - 1.f)) : 0.f;
    }
    st8(dmom + ((int64_t)n * hw + p) * 8, pack8(o));
  }
}

// ---------------------------------------------------------------- MSE forward + gradient
__global__ void __launch_bounds__(256) mse_kernel(const bf16* __restrict__ rec, const float* __restrict__ x,
                                                  double* __restrict__ loss, bf16* __restrict__ drec, float gscale, int N,
                                                  int C, int64_t HW) {
  float acc = 0.f;
  int64_t total = (int64_t)N * HW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t n = i / HW, p = i % HW;
    for (int c = 0; c < C; ++c) {
      float d = __bfloat162float(rec[i * C + c]) - x[(n * C + c) * HW + p];
      acc += d * d;
      if (drec) drec[i * C + c] = __float2bfloat16_rn(2.f * d * gscale);
    }
  }
  acc = warp_sum(acc);
  __shared__ float ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += (double)ws[i];
    atomicAdd(loss, t);
  }
}

int ew_grid(int64_t work, int threads) {
  int64_t b = ceil_div64(work, threads);
  int64_t cap = (int64_t)vcd_num_sms() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" int vcd_pack_conv_weight(const void* w, const void* bias, int dtype, int Cout, int Cin, int KH, int KW,
                                    void* w_fprop, void* w_dgrad, float* bias_f32, vcd_stream_t stream) {
  VCD_CHECK_ARG(w && w_fprop, "vcd_pack_conv_weight: null pointer");
  int64_t total = (int64_t)Cout * Cin * KH * KW;
  pack_weight_kernel<<<ew_grid(total, 256), 256, 0, as_stream(stream)>>>(w, dtype, Cout, Cin, KH * KW, (bf16*)w_fprop,
                                                                         (bf16*)w_dgrad);
  VCD_LAUNCH_CHECK();
  if (bias && bias_f32) {
    bias_to_f32_kernel<<<(Cout + 127) / 128, 128, 0, as_stream(stream)>>>(bias, dtype, Cout, bias_f32);
    VCD_LAUNCH_CHECK();
  }
  return 0;
}

// ---------------------------------------------------------------- all weight packs of the model in ONE launch
// One block = a 32 (Cout) x 64 (Cin) tile of one layer, all taps: the OIHW rows of the tile (64 x taps contiguous elements
// per output channel) are read with 16-byte loads into shared memory (fp32), then written as the fprop pack
// [tap][Cout][Cin] and the dgrad pack [tap][Cin][Cout] with ONE 16-byte store per thread and tap (eight consecutive input /
// output channels gathered from the tile; 128-byte / 64-byte contiguous segments per row) and, for mode 1, as the 16
// pre-summed phase taps of the Upsample2D form (conv_dispatch.cu).  HBM-bound: 2 (bf16) or 4 (fp32) bytes read + 4 bytes
// written per weight = 0.5 GB for the 83.6 M parameters of the VAE.  Partial tiles (the small-channel layers) and
// misaligned tensors take the element-wise path below.  Replaces ~140 per-layer launches of pack_weight_kernel /
// bias_to_f32_kernel per forward.
namespace {
constexpr int kPackCo = 32, kPackCi = 64;
constexpr int kPackRow = kPackCi * 9 + 1;   // floats per tile row (odd: conflict-free column walks)
constexpr int kPackSmem = kPackCo * kPackRow * (int)sizeof(float);
__device__ __forceinline__ int up_group_pack(int a, int k) { return a == 0 ? (k == 0 ? 0 : 1) : (k <= 1 ? 0 : 1); }
// value of output tap q at (co, ci) of the tile: the source tap itself, or the pre-summed phase tap of an Upsample2D conv
template <int MODE>
__device__ __forceinline__ float pack_value(const float* __restrict__ trow, int ci, int taps, int q) {
  if (MODE == 0) return trow[ci * taps + q];
  const int dw = q & 1, dh = (q >> 1) & 1, b = (q >> 2) & 1, a = (q >> 3) & 1;
  float v = 0.f;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int kw = 0; kw < 3; ++kw)
      if (up_group_pack(a, kh) == dh && up_group_pack(b, kw) == dw) v += trow[ci * 9 + kh * 3 + kw];
  return v;
}
template <int MODE>
__device__ __forceinline__ void pack_full_tile(const float* __restrict__ tile, const vcd_pack_desc& d, int co0, int ci0) {
  bf16* wf = reinterpret_cast<bf16*>(d.wf);
  bf16* wd = reinterpret_cast<bf16*>(d.wd);
  const int taps = d.taps, otaps = MODE == 1 ? 16 : taps;
  {   // fprop pack: thread = (output channel, group of 8 input channels)
    const int ci8 = threadIdx.x & 7, co = threadIdx.x >> 3;
    const float* trow = tile + co * kPackRow;
    for (int q = 0; q < otaps; ++q) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = pack_value<MODE>(trow, ci8 * 8 + k, taps, q);
      st8(wf + ((int64_t)q * d.cout + co0 + co) * d.cin + ci0 + ci8 * 8, pack8(v));
    }
  }
  if (wd == nullptr) return;
  {   // dgrad pack: thread = (input channel, group of 8 output channels)
    const int co8 = threadIdx.x & 3, ci = threadIdx.x >> 2;
    for (int q = 0; q < otaps; ++q) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = pack_value<MODE>(tile + (co8 * 8 + k) * kPackRow, ci, taps, q);
      st8(wd + ((int64_t)q * d.cin + ci0 + ci) * d.cout + co0 + co8 * 8, pack8(v));
    }
  }
}
__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

__global__ void __launch_bounds__(256) multi_pack_kernel(const vcd_pack_desc* __restrict__ descs,
                                                        const int32_t* __restrict__ tile_layer,
                                                        const int32_t* __restrict__ tile_co,
                                                        const int32_t* __restrict__ tile_ci) {
  extern __shared__ float tile[];   // [kPackCo][kPackRow]
  const vcd_pack_desc d = descs[tile_layer[blockIdx.x]];
  const int co0 = tile_co[blockIdx.x], ci0 = tile_ci[blockIdx.x];
  const int nco = min(kPackCo, d.cout - co0), nci = min(kPackCi, d.cin - ci0);
  const int taps = d.taps;             // taps of the SOURCE tensor (9 for mode 1)
  const int row = nci * taps;          // contiguous source elements per output channel of this tile
  const bool full = nco == kPackCo && nci == kPackCi && (taps == 9 || taps == 1) && (d.cin & 7) == 0 && (d.cout & 7) == 0 &&
                    aligned16(d.w) && aligned16(d.wf) && aligned16(d.wd) && (d.mode == 0 || taps == 9);
  if (full && d.dtype == VCD_F32) {
    const int nv = row >> 2;
    for (int i = threadIdx.x; i < kPackCo * nv; i += blockDim.x) {
      const int co = i / nv, e = (i - co * nv) << 2;
      const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(d.w) +
                                                        ((int64_t)(co0 + co) * d.cin + ci0) * taps + e);
      float* t = tile + co * kPackRow + e;
      t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
    }
  } else if (full) {
    const int nv = row >> 3;
    for (int i = threadIdx.x; i < kPackCo * nv; i += blockDim.x) {
      const int co = i / nv, e = (i - co * nv) << 3;
      unpack8(ld8(reinterpret_cast<const bf16*>(d.w) + ((int64_t)(co0 + co) * d.cin + ci0) * taps + e),
              tile + co * kPackRow + e);
    }
  } else {
    for (int i = threadIdx.x; i < nco * row; i += blockDim.x) {
      const int co = i / row, e = i - co * row;
      tile[co * kPackRow + e] = load_param(d.w, d.dtype, ((int64_t)(co0 + co) * d.cin + ci0) * taps + e);
    }
  }
  if (ci0 == 0 && d.bias != nullptr && d.bias_f32 != nullptr)
    for (int i = threadIdx.x; i < nco; i += blockDim.x) d.bias_f32[co0 + i] = load_param(d.bias, d.dtype, co0 + i);
  __syncthreads();
  if (full) {
    if (d.mode == 1) pack_full_tile<1>(tile, d, co0, ci0);
    else pack_full_tile<0>(tile, d, co0, ci0);
    return;
  }
  // element-wise path: partial tiles (small-channel layers), unusual tap counts, misaligned tensors
  bf16* wf = reinterpret_cast<bf16*>(d.wf);
  bf16* wd = reinterpret_cast<bf16*>(d.wd);
  const int otaps = d.mode == 1 ? 16 : taps;
  for (int i = threadIdx.x; i < otaps * nco * nci; i += blockDim.x) {   // fprop pack: consecutive threads along ci
    const int ci = i % nci, r = i / nci, co = r % nco, q = r / nco;
    const float v = d.mode == 1 ? pack_value<1>(tile + co * kPackRow, ci, 9, q) : pack_value<0>(tile + co * kPackRow, ci, taps, q);
    wf[((int64_t)q * d.cout + co0 + co) * d.cin + ci0 + ci] = __float2bfloat16_rn(v);
  }
  if (wd == nullptr) return;
  for (int i = threadIdx.x; i < otaps * nco * nci; i += blockDim.x) {   // dgrad pack: consecutive threads along co
    const int co = i % nco, r = i / nco, ci = r % nci, q = r / nci;
    const float v = d.mode == 1 ? pack_value<1>(tile + co * kPackRow, ci, 9, q) : pack_value<0>(tile + co * kPackRow, ci, taps, q);
    wd[((int64_t)q * d.cin + ci0 + ci) * d.cout + co0 + co] = __float2bfloat16_rn(v);
  }
}
}  // namespace

extern "C" int vcd_pack_tile_co(void) { return kPackCo; }
extern "C" int vcd_pack_tile_ci(void) { return kPackCi; }
extern "C" int vcd_multi_pack_weights(const vcd_pack_desc* descs, const int32_t* tile_layer, const int32_t* tile_co,
                                      const int32_t* tile_ci, int n_tiles, vcd_stream_t stream) {
  VCD_CHECK_ARG(descs && tile_layer && tile_co && tile_ci && n_tiles > 0, "vcd_multi_pack_weights: bad arguments");
  bool& attr = *vcd_device_once(8);
  if (!attr) {
    VCD_CUDA(cudaFuncSetAttribute(multi_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPackSmem));
    attr = true;
  }
  multi_pack_kernel<<<n_tiles, 256, kPackSmem, as_stream(stream)>>>(descs, tile_layer, tile_co, tile_ci);
  VCD_LAUNCH_CHECK();
  return 0;
}

extern "C" int vcd_space_to_planes(const void* x, void* xp, int N, int H, int W, int C, vcd_stream_t stream) {
  VCD_CHECK_ARG(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "space_to_planes: need C%%8==0 and even H,W");
  space_planes_kernel<<<ew_grid((int64_t)N * H * W * (C / 8), 256), 256, 0, as_stream(stream)>>>(
      (const bf16*)x, (bf16*)xp, N, H, W, C, 1);
  VCD_LAUNCH_CHECK();
  return 0;
}
extern "C" int vcd_planes_to_space(const void* xp, void* x, int N, int H, int W, int C, vcd_stream_t stream) {
  VCD_CHECK_ARG(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "planes_to_space: need C%%8==0 and even H,W");
  space_planes_kernel<<<ew_grid((int64_t)N * H * W * (C / 8), 256), 256, 0, as_stream(stream)>>>(
      (const bf16*)x, (bf16*)const_cast<void*>(xp), N, H, W, C, 0);
  VCD_LAUNCH_CHECK();
  return 0;
}
extern "C" int vcd_upsample2x_fwd(const void* x, void* y, int N, int H, int W, int C, vcd_stream_t stream) {
  VCD_CHECK_ARG(C % 8 == 0, "upsample2x: need C%%8==0");
  upsample2x_kernel<<<ew_grid((int64_t)N * H * W * (C / 8), 256), 256, 0, as_stream(stream)>>>((const bf16*)x, (bf16*)y,
                                                                                                N, H, W, C);
  VCD_LAUNCH_CHECK();
  return 0;
}
extern "C" int vcd_upsample2x_bwd(const void* dy, void* dx, int N, int H, int W, int C, vcd_stream_t stream) {
  VCD_CHECK_ARG(C % 8 == 0, "upsample2x_bwd: need C%%8==0");
  upsample2x_bwd_kernel<<<ew_grid((int64_t)N * H * W * (C / 8), 256), 256, 0, as_stream(stream)>>>(
      (const bf16*)dy, (bf16*)dx, N, H, W, C);
  VCD_LAUNCH_CHECK();
  return 0;
}
extern "C" int vcd_nchw_to_nhwc(const void* x, int x_dtype, void* y_bf16, int N, int C, int H, int W,
                                vcd_stream_t stream) {
  nchw_to_nhwc_kernel<<<ew_grid((int64_t)N * H * W, 256), 256, 0, as_stream(stream)>>>(x, x_dtype, (bf16*)y_bf16, N, C,
                                                                                        (int64_t)H * W);
  VCD_LAUNCH_CHECK();
  return 0;
}
extern "C" int vcd_nhwc_to_nchw(const void* x_bf16, void* y, int y_dtype, int N, int C, int H, int W,
                                vcd_stream_t stream) {
  nhwc_to_nchw_kernel<<<ew_grid((int64_t)N * H * W, 256), 256, 0, as_stream(stream)>>>((const bf16*)x_bf16, y, y_dtype,
                                                                                        N, C, (int64_t)H * W);
  VCD_LAUNCH_CHECK();
  return 0;
}
extern "C" int vcd_add(const void* a, const void* b, void* out, int64_t n, vcd_stream_t stream) {
  VCD_CHECK_ARG(n % 8 == 0, "vcd_add: n %% 8 != 0");
  add_kernel<<<ew_grid(n / 8, 256), 256, 0, as_stream(stream)>>>((const bf16*)a, (const bf16*)b, (bf16*)out, n / 8);
  VCD_LAUNCH_CHECK();
  return 0;
}
extern "C" int vcd_silu_fwd(const void* x, void* y, int64_t n, vcd_stream_t stream) {
  VCD_CHECK_ARG(n % 8 == 0, "vcd_silu_fwd: n %% 8 != 0");
  silu_kernel<<<ew_grid(n / 8, 256), 256, 0, as_stream(stream)>>>((const bf16*)x, nullptr, (bf16*)y, n / 8);
  VCD_LAUNCH_CHECK();
  return 0;
}
extern "C" int vcd_silu_bwd(const void* x, const void* dy, void* dx, int64_t n, vcd_stream_t stream) {
  VCD_CHECK_ARG(n % 8 == 0, "vcd_silu_bwd: n %% 8 != 0");
  silu_kernel<<<ew_grid(n / 8, 256), 256, 0, as_stream(stream)>>>((const bf16*)x, (const bf16*)dy, (bf16*)dx, n / 8);
  VCD_LAUNCH_CHECK();
  return 0;
}
extern "C" int vcd_softmax_fwd(const void* s, void* p, int64_t rows, int cols, vcd_stream_t stream) {
  VCD_CHECK_ARG(cols % 8 == 0, "softmax: cols %% 8 != 0");
  softmax_fwd_kernel<<<(unsigned)ceil_div64(rows, 8), 256, 0, as_stream(stream)>>>((const bf16*)s, (bf16*)p, rows, cols);
  VCD_LAUNCH_CHECK();
  return 0;
}
extern "C" int vcd_softmax_bwd(const void* p, const void* dp, void* ds, float scale, int64_t rows, int cols,
                               vcd_stream_t stream) {
  VCD_CHECK_ARG(cols % 8 == 0, "softmax: cols %% 8 != 0");
  softmax_bwd_kernel<<<(unsigned)ceil_div64(rows, 8), 256, 0, as_stream(stream)>>>((const bf16*)p, (const bf16*)dp,
                                                                                    (bf16*)ds, scale, rows, cols);
  VCD_LAUNCH_CHECK();
  return 0;
}
extern "C" int vcd_transpose_bf16(const void* x, void* y, int batch, int rows, int cols, vcd_stream_t stream) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, batch);
  transpose_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>((const bf16*)x, (bf16*)y, rows, cols);
  VCD_LAUNCH_CHECK();
  return 0;
}
extern "C" int vcd_gauss_sample_kl_fwd(const void* moments, const float* eps_noise, void* z, float* mean_out,
                                       float* logvar_out, float* kl_per_sample, int N, int hw, int L,
                                       vcd_stream_t stream) {
  VCD_CHECK_ARG(L == 4, "gauss_sample_kl: latent_channels must be 4 (SDXL-VAE), got %d", L);
  VCD_CUDA(cudaMemsetAsync(kl_per_sample, 0, sizeof(float) * N, as_stream(stream)));
  dim3 grid((unsigned)((hw + 255) / 256 > 64 ? 64 : (hw + 255) / 256), N);
  gauss_fwd_kernel<<<grid, 256, 0, as_stream(stream)>>>((const bf16*)moments, eps_noise, (bf16*)z, mean_out, logvar_out,
                                                        kl_per_sample, hw);
  VCD_LAUNCH_CHECK();
  return 0;
}
extern "C" int vcd_gauss_sample_kl_bwd(const void* moments, const float* eps_noise, const void* dz, const float* dkl,
                                       void* dmoments, int N, int hw, int L, vcd_stream_t stream) {
  VCD_CHECK_ARG(L == 4, "gauss_sample_kl: latent_channels must be 4 (SDXL-VAE), got %d", L);
  dim3 grid((unsigned)((hw + 255) / 256 > 64 ? 64 : (hw + 255) / 256), N);
  gauss_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>((const bf16*)moments, eps_noise, (const bf16*)dz, dkl,
                                                        (bf16*)dmoments, hw);
  VCD_LAUNCH_CHECK();
  return 0;
}
extern "C" int vcd_mse_fwd_bwd(const void* rec, const float* x_nchw, double* loss_sum, void* drec, float grad_scale,
                               int N, int C, int H, int W, vcd_stream_t stream) {
  VCD_CUDA(cudaMemsetAsync(loss_sum, 0, sizeof(double), as_stream(stream)));
  mse_kernel<<<ew_grid((int64_t)N * H * W, 256), 256, 0, as_stream(stream)>>>((const bf16*)rec, x_nchw, loss_sum,
                                                                              (bf16*)drec, grad_scale, N, C,
                                                                              (int64_t)H * W);
  VCD_LAUNCH_CHECK();
  return 0;
}
