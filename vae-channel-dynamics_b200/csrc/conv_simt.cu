// conv_simt.cu — CUDA-core direct convolutions.
//
// Production use: the six small-channel layers of the SDXL VAE (encoder.conv_in 3->128,
// encoder.conv_out 512->8, quant_conv 8->8, post_quant_conv 4->4, decoder.conv_in 4->512,
// decoder.conv_out 128->3; 0.11 % of the FLOPs, HBM-bound) whose channel counts cannot fill a
// tcgen05 tile.  Also the independent on-device cross-check for the tcgen05 path (VCD_IMPL_SIMT).
//
// One generalised form serves fprop and dgrad:
//   out[n,h,w,o] = sum_t sum_i in[n, hi(t), wi(t), i] * wt[t][o][i]
//   fprop : hi = h*stride + dh[t],            wt = w_fprop [tap][Cout][Cin]
//   dgrad : q = h + dh[t]; hi = q/stride if q % stride == 0 (else skipped), wt = w_dgrad [tap][Cin][Cout]
#include "common.cuh"
#include "conv_dispatch.h"

namespace {

struct Taps {
  int n;
  int dh[9], dw[9];
};

// ---------------------------------------------------------------- G: one thread per (pixel, out channel)
__global__ void __launch_bounds__(256) conv_general_kernel(const bf16* __restrict__ in, const bf16* __restrict__ wt,
                                                           const float* __restrict__ bias,
                                                           const bf16* __restrict__ residual, bf16* __restrict__ out,
                                                           int N, int Hin, int Win, int Ci, int Hout, int Wout, int Co,
                                                           int stride, int transposed, Taps taps) {
  int64_t total = (int64_t)N * Hout * Wout * Co;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int o = (int)(idx % Co);
    int64_t px = idx / Co;
    int w = (int)(px % Wout);
    int64_t r = px / Wout;
    int h = (int)(r % Hout);
    int n = (int)(r / Hout);
    float acc = bias ? bias[o] : 0.f;
    for (int t = 0; t < taps.n; ++t) {
      int hi, wi;
      if (transposed) {
        int qh = h + taps.dh[t], qw = w + taps.dw[t];
        if (qh < 0 || qw < 0 || (qh % stride) || (qw % stride)) continue;
        hi = qh / stride;
        wi = qw / stride;
      } else {
        hi = h * stride + taps.dh[t];
        wi = w * stride + taps.dw[t];
      }
      if (hi < 0 || hi >= Hin || wi < 0 || wi >= Win) continue;
      const bf16* ip = in + (((int64_t)n * Hin + hi) * Win + wi) * Ci;
      const bf16* wp = wt + ((int64_t)t * Co + o) * Ci;
      for (int i = 0; i < Ci; ++i) acc = fmaf(__bfloat162float(ip[i]), __bfloat162float(wp[i]), acc);
    }
    if (residual) acc += __bfloat162float(residual[idx]);
    out[idx] = __float2bfloat16_rn(acc);
  }
}

// ---------------------------------------------------------------- S: one thread per pixel, <= 8 out channels,
// Ci % 8 == 0, weights staged in shared memory (broadcast reads), 16-byte input vectors.
__global__ void __launch_bounds__(128) conv_small_out_kernel(const bf16* __restrict__ in, const bf16* __restrict__ wt,
                                                             const float* __restrict__ bias, bf16* __restrict__ out, int N,
                                                             int Hin, int Win, int Ci, int Hout, int Wout, int Co,
                                                             int stride, int transposed, Taps taps) {
  extern __shared__ __align__(16) bf16 sw[];  // [tap][Co][Ci]
  const int wn = taps.n * Co * Ci;
  for (int i = threadIdx.x; i < wn; i += blockDim.x) sw[i] = wt[i];
  __syncthreads();
  int64_t total = (int64_t)N * Hout * Wout;
  for (int64_t px = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; px < total; px += (int64_t)gridDim.x * blockDim.x) {
    int w = (int)(px % Wout);
    int64_t r = px / Wout;
    int h = (int)(r % Hout);
    int n = (int)(r / Hout);
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = (bias && o < Co) ? bias[o] : 0.f;
    for (int t = 0; t < taps.n; ++t) {
      int hi, wi;
      if (transposed) {
        int qh = h + taps.dh[t], qw = w + taps.dw[t];
        if (qh < 0 || qw < 0 || (qh % stride) || (qw % stride)) continue;
        hi = qh / stride;
        wi = qw / stride;
      } else {
        hi = h * stride + taps.dh[t];
        wi = w * stride + taps.dw[t];
      }
      if (hi < 0 || hi >= Hin || wi < 0 || wi >= Win) continue;
      const bf16* ip = in + (((int64_t)n * Hin + hi) * Win + wi) * Ci;
      const bf16* wp = sw + (int64_t)t * Co * Ci;
      for (int i = 0; i < Ci; i += 8) {
        float f[8];
        unpack8(ld8(ip + i), f);
#pragma unroll
        for (int o = 0; o < 8; ++o) {
          if (o < Co) {
            float g[8];
            unpack8(ld8(wp + o * Ci + i), g);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[o] = fmaf(f[j], g[j], acc[o]);
          }
        }
      }
    }
    for (int o = 0; o < Co; ++o) out[px * Co + o] = __float2bfloat16_rn(acc[o]);
  }
}

// ---------------------------------------------------------------- weight gradient
// ws fp32 [tap][Co][Ci]; general: one thread per output element, block = pixel chunk (atomics).
__global__ void __launch_bounds__(256) wgrad_general_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                                            float* __restrict__ ws, int N, int H, int W, int Ci, int Ho,
                                                            int Wo, int Co, int stride, Taps taps, int64_t chunk) {
  const int64_t pixels = (int64_t)N * Ho * Wo;
  const int64_t p0 = blockIdx.y * chunk, p1 = min(p0 + chunk, pixels);
  const int64_t outs = (int64_t)taps.n * Co * Ci;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < outs; e += (int64_t)gridDim.x * blockDim.x) {
    int ci = (int)(e % Ci);
    int64_t r = e / Ci;
    int co = (int)(r % Co);
    int t = (int)(r / Co);
    float acc = 0.f;
    for (int64_t p = p0; p < p1; ++p) {
      int w = (int)(p % Wo);
      int64_t rr = p / Wo;
      int h = (int)(rr % Ho);
      int n = (int)(rr / Ho);
      int hi = h * stride + taps.dh[t], wi = w * stride + taps.dw[t];
      if (hi < 0 || hi >= H || wi < 0 || wi >= W) continue;
      acc = fmaf(__bfloat162float(dy[p * Co + co]), __bfloat162float(x[(((int64_t)n * H + hi) * W + wi) * Ci + ci]), acc);
    }
    atomicAdd(&ws[e], acc);
  }
}

// small Cin (<= 8): thread = co; registers acc[tap][ci]
template <int SMALL>
__global__ void __launch_bounds__(128) wgrad_small_in_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                                             float* __restrict__ ws, int N, int H, int W, int Ci, int Ho,
                                                             int Wo, int Co, int stride, Taps taps, int64_t chunk) {
  const int co = blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t pixels = (int64_t)N * Ho * Wo;
  const int64_t p0 = blockIdx.y * chunk, p1 = min(p0 + chunk, pixels);
  float acc[9][SMALL];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < SMALL; ++i) acc[t][i] = 0.f;
  if (co < Co) {
    for (int64_t p = p0; p < p1; ++p) {
      int w = (int)(p % Wo);
      int64_t rr = p / Wo;
      int h = (int)(rr % Ho);
      int n = (int)(rr / Ho);
      float g = __bfloat162float(dy[p * Co + co]);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        if (t < taps.n) {
          int hi = h * stride + taps.dh[t], wi = w * stride + taps.dw[t];
          if (hi >= 0 && hi < H && wi >= 0 && wi < W) {
            const bf16* xp = x + (((int64_t)n * H + hi) * W + wi) * Ci;
#pragma unroll
            for (int i = 0; i < SMALL; ++i)
              if (i < Ci) acc[t][i] = fmaf(g, __bfloat162float(xp[i]), acc[t][i]);
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t)
      if (t < taps.n)
#pragma unroll
        for (int i = 0; i < SMALL; ++i)
          if (i < Ci) atomicAdd(&ws[((int64_t)t * Co + co) * Ci + i], acc[t][i]);
  }
}

// small Cout (<= 8): thread = ci; registers acc[tap][co]
template <int SMALL>
__global__ void __launch_bounds__(128) wgrad_small_out_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                                              float* __restrict__ ws, int N, int H, int W, int Ci, int Ho,
                                                              int Wo, int Co, int stride, Taps taps, int64_t chunk) {
  const int ci = blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t pixels = (int64_t)N * Ho * Wo;
  const int64_t p0 = blockIdx.y * chunk, p1 = min(p0 + chunk, pixels);
  float acc[9][SMALL];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int o = 0; o < SMALL; ++o) acc[t][o] = 0.f;
  if (ci < Ci) {
    for (int64_t p = p0; p < p1; ++p) {
      int w = (int)(p % Wo);
      int64_t rr = p / Wo;
      int h = (int)(rr % Ho);
      int n = (int)(rr / Ho);
      float g[SMALL];
#pragma unroll
      for (int o = 0; o < SMALL; ++o) g[o] = (o < Co) ? __bfloat162float(dy[p * Co + o]) : 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        if (t < taps.n) {
          int hi = h * stride + taps.dh[t], wi = w * stride + taps.dw[t];
          if (hi >= 0 && hi < H && wi >= 0 && wi < W) {
            float xv = __bfloat162float(x[(((int64_t)n * H + hi) * W + wi) * Ci + ci]);
#pragma unroll
            for (int o = 0; o < SMALL; ++o) acc[t][o] = fmaf(g[o], xv, acc[t][o]);
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t)
      if (t < taps.n)
#pragma unroll
        for (int o = 0; o < SMALL; ++o)
          if (o < Co) atomicAdd(&ws[((int64_t)t * Co + o) * Ci + ci], acc[t][o]);
  }
}

// db[c] = sum over pixels dy[p][c]
// C % 8 == 0: thread (pl, v) owns 8 channels (one 16-byte load per pixel) and walks pixels pl, pl+PL, ... of the block's
// range; partials meet in shared memory, one atomicAdd per (block, channel).
__global__ void __launch_bounds__(256) colsum_vec_kernel(const bf16* __restrict__ dy, float* __restrict__ out, int64_t pixels,
                                                         int C, int64_t chunk) {
  extern __shared__ float red[];  // [8][PL][V]
  const int V = C >> 3, PL = 256 / V;
  const int v = threadIdx.x % V, pl = threadIdx.x / V;
  const int64_t p0 = blockIdx.x * chunk, p1 = min(p0 + chunk, pixels);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const bf16* base = dy + v * 8;
  int64_t p = p0 + pl;
  for (; p + 3 * (int64_t)PL < p1; p += 4 * (int64_t)PL) {
    bf16x8 t[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) t[u] = ld8(base + (p + (int64_t)u * PL) * C);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float f[8];
      unpack8(t[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
  }
  for (; p < p1; p += PL) {
    float f[8];
    unpack8(ld8(base + p * C), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += f[j];
  }
  if (pl < PL) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red[(j * PL + pl) * V + v] = acc[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    const int cv = c >> 3, j = c & 7;
    float t = 0.f;
    for (int q = 0; q < PL; ++q) t += red[(j * PL + q) * V + cv];
    atomicAdd(&out[c], t);
  }
}
// any C (the 3 / 4 / 8-channel tensors): one pixel per thread and iteration, warp-shuffle sum per channel
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ dy, float* __restrict__ out, int64_t pixels,
                                                     int C, int64_t chunk) {
  const int64_t p0 = blockIdx.x * chunk, p1 = min(p0 + chunk, pixels);
  for (int c0 = 0; c0 < C; c0 += 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int64_t p = p0 + threadIdx.x; p < p1; p += 256) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c0 + j < C) acc[j] += __bfloat162float(dy[p * C + c0 + j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float t = warp_sum(acc[j]);
      if ((threadIdx.x & 31) == 0 && c0 + j < C) atomicAdd(&out[c0 + j], t);
    }
  }
}

// ws [tap][Co][Ci] fp32 (+ bias) -> OIHW dw / db in the parameter dtype.  One block = one output channel x 64 input
// channels x all taps, transposed through shared memory: both the [tap][Co][Ci] reads and the [Co][Ci][tap] writes are
// contiguous runs (the round-1 kernel read with a stride of Co*Ci floats per thread).
constexpr int kFinCi = 64;
__global__ void __launch_bounds__(128) wgrad_finalize_kernel(const float* __restrict__ ws, const float* __restrict__ bias_src,
                                                             void* __restrict__ dw, void* __restrict__ db, int dt, int Co,
                                                             int Ci, int taps) {
  __shared__ float tile[16][kFinCi + 1];
  const int tiles_ci = (Ci + kFinCi - 1) / kFinCi;
  const int co = blockIdx.x / tiles_ci, ci0 = (blockIdx.x % tiles_ci) * kFinCi;
  const int nci = min(kFinCi, Ci - ci0);
  for (int i = threadIdx.x; i < taps * nci; i += blockDim.x) {
    const int t = i / nci, c = i - t * nci;
    tile[t][c] = ws[((int64_t)t * Co + co) * Ci + ci0 + c];
  }
  __syncthreads();
  const int64_t obase = ((int64_t)co * Ci + ci0) * taps;
  for (int j = threadIdx.x; j < nci * taps; j += blockDim.x) store_param(dw, dt, obase + j, tile[j % taps][j / taps]);
  if (db && ci0 == 0 && threadIdx.x == 0) store_param(db, dt, co, bias_src[co]);
}

Taps make_taps(int KH, int KW, int pad_t, int pad_l, bool dgrad) {
  Taps t;
  t.n = KH * KW;
  for (int kh = 0; kh < KH; ++kh)
    for (int kw = 0; kw < KW; ++kw) {
      t.dh[kh * KW + kw] = dgrad ? (pad_t - kh) : (kh - pad_t);
      t.dw[kh * KW + kw] = dgrad ? (pad_l - kw) : (kw - pad_l);
    }
  return t;
}

int grid_for(int64_t work, int threads, int waves = 8) {
  int64_t b = ceil_div64(work, threads);
  int64_t cap = (int64_t)vcd_num_sms() * waves;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

int simt_conv_fprop(const void* x, const void* wf, const float* bias, const void* residual, void* y, int N, int H, int W,
                    int Cin, int Cout, int KH, int KW, int stride, int pad_t, int pad_l, int Ho, int Wo,
                    cudaStream_t st) {
  VCD_CHECK_ARG(KH * KW <= 9, "conv: kernel larger than 3x3 unsupported");
  Taps taps = make_taps(KH, KW, pad_t, pad_l, false);
  if (Cout <= 8 && Cin % 8 == 0 && !residual && (size_t)taps.n * Cout * Cin * 2 <= 96 * 1024) {
    size_t smem = (size_t)taps.n * Cout * Cin * sizeof(bf16);
    VCD_CUDA(cudaFuncSetAttribute(conv_small_out_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    conv_small_out_kernel<<<grid_for((int64_t)N * Ho * Wo, 128), 128, smem, st>>>(
        (const bf16*)x, (const bf16*)wf, bias, (bf16*)y, N, H, W, Cin, Ho, Wo, Cout, stride, 0, taps);
  } else {
    conv_general_kernel<<<grid_for((int64_t)N * Ho * Wo * Cout, 256, 32), 256, 0, st>>>(
        (const bf16*)x, (const bf16*)wf, bias, (const bf16*)residual, (bf16*)y, N, H, W, Cin, Ho, Wo, Cout, stride, 0, taps);
  }
  VCD_LAUNCH_CHECK();
  return 0;
}

int simt_conv_dgrad(const void* dy, const void* wd, void* dx, int N, int H, int W, int Cin, int Cout, int KH, int KW,
                    int stride, int pad_t, int pad_l, int Ho, int Wo, cudaStream_t st) {
  VCD_CHECK_ARG(KH * KW <= 9, "conv: kernel larger than 3x3 unsupported");
  VCD_CHECK_ARG(wd != nullptr, "conv dgrad (SIMT) needs the w_dgrad pack");
  Taps taps = make_taps(KH, KW, pad_t, pad_l, true);
  // roles: in = dy [N][Ho][Wo][Cout], out = dx [N][H][W][Cin], wt = w_dgrad [tap][Cin][Cout]
  if (Cin <= 8 && Cout % 8 == 0 && (size_t)taps.n * Cout * Cin * 2 <= 96 * 1024) {
    size_t smem = (size_t)taps.n * Cout * Cin * sizeof(bf16);
    VCD_CUDA(cudaFuncSetAttribute(conv_small_out_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    conv_small_out_kernel<<<grid_for((int64_t)N * H * W, 128), 128, smem, st>>>(
        (const bf16*)dy, (const bf16*)wd, nullptr, (bf16*)dx, N, Ho, Wo, Cout, H, W, Cin, stride, 1, taps);
  } else {
    conv_general_kernel<<<grid_for((int64_t)N * H * W * Cin, 256, 32), 256, 0, st>>>(
        (const bf16*)dy, (const bf16*)wd, nullptr, nullptr, (bf16*)dx, N, Ho, Wo, Cout, H, W, Cin, stride, 1, taps);
  }
  VCD_LAUNCH_CHECK();
  return 0;
}

// ws must already be zeroed: [tap][Cout][Cin] then [Cout]
int simt_conv_wgrad(const void* x, const void* dy, float* ws, int N, int H, int W, int Cin, int Cout, int KH, int KW,
                    int stride, int pad_t, int pad_l, int Ho, int Wo, cudaStream_t st) {
  VCD_CHECK_ARG(KH * KW <= 9, "conv: kernel larger than 3x3 unsupported");
  Taps taps = make_taps(KH, KW, pad_t, pad_l, false);
  const int64_t pixels = (int64_t)N * Ho * Wo;
  const int sms = vcd_num_sms();
  if (Cin <= 8 && Cout >= 32) {
    int bx = (Cout + 127) / 128;
    int64_t chunks = (2 * sms + bx - 1) / bx;
    int64_t chunk = ceil_div64(pixels, chunks);
    dim3 grid(bx, (unsigned)ceil_div64(pixels, chunk));
    wgrad_small_in_kernel<8><<<grid, 128, 0, st>>>((const bf16*)x, (const bf16*)dy, ws, N, H, W, Cin, Ho, Wo, Cout, stride,
                                                  taps, chunk);
  } else if (Cout <= 8 && Cin >= 32) {
    int bx = (Cin + 127) / 128;
    int64_t chunks = (2 * sms + bx - 1) / bx;
    int64_t chunk = ceil_div64(pixels, chunks);
    dim3 grid(bx, (unsigned)ceil_div64(pixels, chunk));
    wgrad_small_out_kernel<8><<<grid, 128, 0, st>>>((const bf16*)x, (const bf16*)dy, ws, N, H, W, Cin, Ho, Wo, Cout,
                                                   stride, taps, chunk);
  } else {
    int64_t outs = (int64_t)taps.n * Cout * Cin;
    int bx = (int)ceil_div64(outs, 256);
    if (bx > 4096) bx = 4096;
    int64_t chunks = ceil_div64(4 * sms, bx);
    if (chunks < 1) chunks = 1;
    int64_t chunk = ceil_div64(pixels, chunks);
    if (chunk < 64) chunk = 64;
    dim3 grid(bx, (unsigned)ceil_div64(pixels, chunk));
    wgrad_general_kernel<<<grid, 256, 0, st>>>((const bf16*)x, (const bf16*)dy, ws, N, H, W, Cin, Ho, Wo, Cout, stride,
                                               taps, chunk);
  }
  VCD_LAUNCH_CHECK();
  return 0;
}

int conv_bias_grad(const void* dy, float* out, int64_t pixels, int C, cudaStream_t st) {
  int64_t chunks = (int64_t)vcd_num_sms() * 4;
  int64_t chunk = ceil_div64(pixels, chunks);
  if (chunk < 32) chunk = 32;
  if (C % 8 == 0 && 256 % (C / 8) == 0 && C / 8 <= 256)
    colsum_vec_kernel<<<(unsigned)ceil_div64(pixels, chunk), 256, 8 * 256 * sizeof(float), st>>>((const bf16*)dy, out, pixels,
                                                                                                  C, chunk);
  else
    colsum_kernel<<<(unsigned)ceil_div64(pixels, chunk), 256, 0, st>>>((const bf16*)dy, out, pixels, C, chunk);
  VCD_LAUNCH_CHECK();
  return 0;
}

// bias_src: fp32 [Cout] bias gradient (the column sums a previous kernel already produced), or NULL = ws + taps*Cout*Cin
int conv_wgrad_finalize(const float* ws, const float* bias_src, void* dw, void* db, int dtype, int Cout, int Cin, int taps,
                        cudaStream_t st) {
  int64_t total = (int64_t)Cout * Cin * taps;
  if (taps > 16) {
    vcd_set_error("conv_wgrad_finalize: at most 16 taps (got %d)", taps);
    return -1;
  }
  const int64_t blocks = (int64_t)Cout * ((Cin + kFinCi - 1) / kFinCi);
  wgrad_finalize_kernel<<<(unsigned)blocks, 128, 0, st>>>(ws, bias_src ? bias_src : ws + total, dw, db, dtype, Cout, Cin, taps);
  VCD_LAUNCH_CHECK();
  return 0;
}
