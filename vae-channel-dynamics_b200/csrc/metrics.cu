// metrics.cu — evaluation metrics and input preprocessing on the device.
//
// Replaces, for reference src/evaluate.py:163-176,238-249,281-288: torchmetrics' PeakSignalNoiseRatio(data_range) and
// StructuralSimilarityIndexMeasure(gaussian_kernel=True, sigma, kernel_size) — [upstream] torchmetrics computes SSIM as
// a depth-wise Gaussian convolution of the reflect-padded images and then crops the padded border, i.e. the Gaussian
// window slides over VALID positions only; per-image mean over (C, H-k+1, W-k+1), then mean over images — and
// src/data_utils.py:13-30 (Resize(bilinear) -> CenterCrop -> ToTensor -> Normalize(0.5, 0.5)) for uint8 batches.
// HBM-bound one-pass kernels: one read of both images for SSIM (+ the squared error of PSNR in the same pass).
#include "common.cuh"

namespace {

constexpr int kMaxK = 15;
constexpr int TW = 32, TH = 16;  // output tile (valid positions) per block

struct GaussK {
  float w[kMaxK];
};

// One block: TH x TW valid-window outputs of one (image, channel) plane.  Stage the (TH+K-1) x (TW+K-1) inputs in shared
// memory, horizontal pass for the five moments (x, y, xx, yy, xy), vertical pass, SSIM map value, block reduction.
// The block also owns the squared error of its TH x TW top-left aligned input pixels (edge blocks own the K-1 border).
__global__ void __launch_bounds__(256) ssim_psnr_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                        int C, int H, int W, int K, GaussK g, float c1, float c2,
                                                        double* __restrict__ ssim_sum, double* __restrict__ sse,
                                                        double inv_valid) {
  const int Ho = H - K + 1, Wo = W - K + 1;
  const int plane = blockIdx.z;                 // n * C + c
  const int ox0 = blockIdx.x * TW, oy0 = blockIdx.y * TH;
  const int IW = TW + kMaxK - 1, IH = TH + kMaxK - 1;
  __shared__ float sx[IH][IW + 1], sy[IH][IW + 1];
  __shared__ float hm[5][IH][TW + 1];
  const float* px = pred + (int64_t)plane * H * W;
  const float* py = target + (int64_t)plane * H * W;
  const int iw = TW + K - 1, ih = TH + K - 1;
  const bool last_x = (ox0 + TW >= Wo), last_y = (oy0 + TH >= Ho);
  double my_sse = 0.0;
  for (int i = threadIdx.x; i < ih * iw; i += blockDim.x) {
    const int r = i / iw, c = i - r * iw;
    const int y = oy0 + r, x = ox0 + c;
    float a = 0.f, b = 0.f;
    if (y < H && x < W) {
      a = px[(int64_t)y * W + x];
      b = py[(int64_t)y * W + x];
      // ownership of input pixels for the squared error: the tile's own TH x TW region, extended to the image edge
      // by the last tile in each direction
      const bool own_x = (c < TW) || last_x, own_y = (r < TH) || last_y;
      if (own_x && own_y) {
        const double d = (double)a - (double)b;
        my_sse += d * d;
      }
    }
    sx[r][c] = a;
    sy[r][c] = b;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ih * TW; i += blockDim.x) {
    const int r = i / TW, c = i - r * TW;
    float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f, m4 = 0.f;
    for (int k = 0; k < K; ++k) {
      const float a = sx[r][c + k], b = sy[r][c + k], w = g.w[k];
      m0 = fmaf(w, a, m0);
      m1 = fmaf(w, b, m1);
      m2 = fmaf(w, a * a, m2);
      m3 = fmaf(w, b * b, m3);
      m4 = fmaf(w, a * b, m4);
    }
    hm[0][r][c] = m0; hm[1][r][c] = m1; hm[2][r][c] = m2; hm[3][r][c] = m3; hm[4][r][c] = m4;
  }
  __syncthreads();
  double my_ssim = 0.0;
  for (int i = threadIdx.x; i < TH * TW; i += blockDim.x) {
    const int r = i / TW, c = i - r * TW;
    if (oy0 + r >= Ho || ox0 + c >= Wo) continue;
    float mx = 0.f, my = 0.f, xx = 0.f, yy = 0.f, xy = 0.f;
    for (int k = 0; k < K; ++k) {
      const float w = g.w[k];
      mx = fmaf(w, hm[0][r + k][c], mx);
      my = fmaf(w, hm[1][r + k][c], my);
      xx = fmaf(w, hm[2][r + k][c], xx);
      yy = fmaf(w, hm[3][r + k][c], yy);
      xy = fmaf(w, hm[4][r + k][c], xy);
    }
    const float vx = xx - mx * mx, vy = yy - my * my, cxy = xy - mx * my;
    const float num = (2.f * mx * my + c1) * (2.f * cxy + c2);
    const float den = (mx * mx + my * my + c1) * (vx + vy + c2);
    my_ssim += (double)(num / den);
  }
  my_ssim = warp_sum_d(my_ssim);
  my_sse = warp_sum_d(my_sse);
  __shared__ double red[2][8];
  const int wid = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { red[0][wid] = my_ssim; red[1][wid] = my_sse; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += red[0][i]; b += red[1][i]; }
    atomicAdd(ssim_sum, a * inv_valid);   // sum over images of the per-image mean SSIM
    if (b != 0.0) atomicAdd(sse, b);
  }
}

// data_utils.py:13-30 for a decoded uint8 batch: bilinear resize (shorter side -> R, half-pixel centres, no antialias),
// centre crop R x R, /255, (x - 0.5) / 0.5.  in: [N][H][W][3] uint8, out: [N][3][R][R] fp32.
__global__ void preprocess_u8_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, int N, int H, int W, int R,
                                     int RH, int RW, float sy, float sx, int oy, int ox) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)N * R * R;
  if (i >= total) return;
  const int x = (int)(i % R), y = (int)((i / R) % R), n = (int)(i / ((int64_t)R * R));
  // position in the resized (RH x RW) image, then in the source
  float fy = ((float)(y + oy) + 0.5f) * sy - 0.5f, fx = ((float)(x + ox) + 0.5f) * sx - 0.5f;
  fy = fmaxf(fy, 0.f);
  fx = fmaxf(fx, 0.f);
  int y0 = (int)fy, x0 = (int)fx;
  y0 = min(y0, H - 1);
  x0 = min(x0, W - 1);
  const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
  const float ly = fy - (float)y0, lx = fx - (float)x0;
  const uint8_t* b = in + (int64_t)n * H * W * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v00 = b[((int64_t)y0 * W + x0) * 3 + c], v01 = b[((int64_t)y0 * W + x1) * 3 + c];
    const float v10 = b[((int64_t)y1 * W + x0) * 3 + c], v11 = b[((int64_t)y1 * W + x1) * 3 + c];
    const float top = v00 + (v01 - v00) * lx, bot = v10 + (v11 - v10) * lx;
    const float v = top + (bot - top) * ly;
    out[(((int64_t)n * 3 + c) * R + y) * R + x] = v * (2.f / 255.f) - 1.f;
  }
}

}  // namespace

extern "C" int vcd_ssim_psnr_update(const float* pred, const float* target, int N, int C, int H, int W,
                                    float data_range, int kernel_size, float sigma, double* ssim_sum, double* sse,
                                    vcd_stream_t stream) {
  VCD_CHECK_ARG(kernel_size >= 1 && kernel_size <= kMaxK && (kernel_size & 1), "SSIM kernel_size must be odd and <= %d", kMaxK);
  VCD_CHECK_ARG(H >= kernel_size && W >= kernel_size, "SSIM needs images of at least kernel_size pixels per side");
  VCD_CHECK_ARG(N > 0 && C > 0 && sigma > 0.f, "bad SSIM arguments");
  GaussK g;
  double s = 0.0;
  for (int k = 0; k < kMaxK; ++k) g.w[k] = 0.f;
  for (int k = 0; k < kernel_size; ++k) {   // [upstream] torchmetrics _gaussian: exp(-(d/sigma)^2 / 2), normalised
    const double d = (double)k - (double)(kernel_size - 1) / 2.0;
    const double w = exp(-0.5 * (d / (double)sigma) * (d / (double)sigma));
    g.w[k] = (float)w;
    s += w;
  }
  for (int k = 0; k < kernel_size; ++k) g.w[k] = (float)((double)g.w[k] / s);
  const int Ho = H - kernel_size + 1, Wo = W - kernel_size + 1;
  dim3 grid((Wo + TW - 1) / TW, (Ho + TH - 1) / TH, N * C);
  const float c1 = (0.01f * data_range) * (0.01f * data_range), c2 = (0.03f * data_range) * (0.03f * data_range);
  ssim_psnr_kernel<<<grid, 256, 0, as_stream(stream)>>>(pred, target, C, H, W, kernel_size, g, c1, c2, ssim_sum, sse,
                                                        1.0 / ((double)C * Ho * Wo));
  VCD_LAUNCH_CHECK();
  return 0;
}

extern "C" int vcd_preprocess_u8(const uint8_t* images, float* out, int N, int H, int W, int R, vcd_stream_t stream) {
  VCD_CHECK_ARG(N > 0 && H > 0 && W > 0 && R > 0, "bad preprocess arguments");
  // torchvision Resize(int): shorter side -> R, the other side int(R * long / short)
  int RH, RW;
  if (H <= W) { RH = R; RW = (int)((int64_t)R * W / H); } else { RW = R; RH = (int)((int64_t)R * H / W); }
  const float sy = (float)H / (float)RH, sx = (float)W / (float)RW;
  // CenterCrop: top = round((RH - R) / 2), left = round((RW - R) / 2)  (python round: half to even)
  auto pyround = [](double v) { double r = nearbyint(v); return (int)r; };
  const int oy = pyround((RH - R) / 2.0), ox = pyround((RW - R) / 2.0);
  const int64_t total = (int64_t)N * R * R;
  preprocess_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(images, out, N, H, W, R, RH, RW, sy,
                                                                                       sx, oy, ox);
  VCD_LAUNCH_CHECK();
  return 0;
}
