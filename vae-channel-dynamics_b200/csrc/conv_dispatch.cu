// conv_dispatch.cu — C-ABI convolution / GEMM entry points: shape checks, TMA tensor maps, tile
// decomposition and tap tables for the tcgen05 kernel (umma_gemm.cu), SIMT path for small channels.
#include <stdarg.h>
#include <string.h>

#include "conv_dispatch.h"
#include "umma_gemm.cuh"

// ---------------------------------------------------------------- error / device info (C ABI)
static thread_local char g_err[512] = "";
void vcd_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* vcd_last_error(void) { return g_err; }
extern "C" int vcd_version(void) { return 100; }
int vcd_num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  return sms;
}

namespace {

int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// split PIX pixels into a (w, h, n) box that tiles the (W, H, N) pixel space
void choose_tile(int PIX, int W, int H, int N, UmmaParams& p) {
  int tw = W >= PIX ? PIX : pow2_ceil(W);
  if (tw > PIX) tw = PIX;
  int rem = PIX / tw;
  int th = H >= rem ? rem : pow2_ceil(H);
  if (th > rem) th = rem;
  int tn = rem / th;
  p.W = W; p.H = H; p.Nimg = N;
  p.tile_w = tw; p.tile_h = th; p.tile_n = tn;
  p.tiles_w = (W + tw - 1) / tw;
  p.tiles_h = (H + th - 1) / th;
  p.tiles_n = (N + tn - 1) / tn;
}

constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO 1024 B, version 1, SWIZZLE_128B
uint32_t make_idesc(int n, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
void set_form0_desc(UmmaParams& p, int block_n) {
  p.a_desc_hi = p.b_desc_hi = kDescHi;
  p.a_lbo = p.b_lbo = 1;
  p.a_kstep = p.b_kstep = 32 >> 4;  // 16 bf16 along K inside the 128-byte swizzle row
  p.idesc = make_idesc(block_n, 0, 0);
}
void set_form1_desc(UmmaParams& p, int block_n) {
  p.a_desc_hi = p.b_desc_hi = kDescHi;
  p.a_lbo = p.b_lbo = 8192 >> 4;     // next 64-channel box
  p.a_kstep = p.b_kstep = 2048 >> 4;  // 16 pixel rows of 128 B
  p.idesc = make_idesc(block_n, 1, 1);
}
int pick_block_n(int n) { return (n % 256 == 0) ? 256 : 128; }

// taps of a conv seen from the OUTPUT pixel grid, reading the (possibly parity-plane) input
void fill_fprop_taps(UmmaParams& p, int KH, int KW, int stride, int pad_t, int pad_l, int rows_per_tap) {
  p.ntaps = KH * KW;
  for (int kh = 0; kh < KH; ++kh)
    for (int kw = 0; kw < KW; ++kw) {
      int t = kh * KW + kw;
      int eh = kh - pad_t, ew = kw - pad_l;
      if (stride == 1) {
        p.tap_dh[t] = eh; p.tap_dw[t] = ew; p.tap_plane[t] = 0;
      } else {
        int ph = eh & 1, pw = ew & 1;
        p.tap_dh[t] = (eh - ph) / 2; p.tap_dw[t] = (ew - pw) / 2; p.tap_plane[t] = ph * 2 + pw;
      }
      p.tap_brow[t] = t * rows_per_tap;
    }
}

bool umma_shape_ok(int Cin, int Cout, int KH, int KW, int stride) {
  if (Cin % 128 != 0 || Cout % 128 != 0) return false;
  if (!((KH == 3 && KW == 3) || (KH == 1 && KW == 1))) return false;
  if (stride != 1 && stride != 2) return false;
  return true;
}

int umma_fprop(const void* x, const void* wf, const float* bias, const void* residual, void* y, int N, int H, int W,
               int Cin, int Cout, int KH, int KW, int stride, int pad_t, int pad_l, int Ho, int Wo, int x_planes,
               cudaStream_t st) {
  UmmaParams p;
  memset(&p, 0, sizeof(p));
  p.form = 0;
  VCD_CHECK_ARG(stride == 1 || x_planes, "tcgen05 stride-2 conv needs the parity-plane input (vcd_space_to_planes)");
  VCD_CHECK_ARG(stride == 2 || (Ho == H && Wo == W), "tcgen05 conv: stride-1 convs must preserve the spatial size");
  choose_tile(128, Wo, Ho, N, p);
  const int bn = pick_block_n(Cout);
  fill_fprop_taps(p, KH, KW, stride, pad_t, pad_l, Cout);
  p.n_tiles = Cout / bn;
  p.kc_per_tap = Cin / 64;
  p.b_batch_rows = 0;
  p.out = (bf16*)y; p.residual = (const bf16*)residual; p.bias = bias; p.alpha = 1.f;
  p.out_sn = (long long)Ho * Wo * Cout; p.out_sh = (long long)Wo * Cout; p.out_sw = Cout;
  p.Nout = Cout;
  set_form0_desc(p, bn);
  p.total_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles;
  CUtensorMap mA, mB;
  int rc;
  if (stride == 1) rc = make_act_map(&mA, x, Cin, W, H, 1, N, 64, p.tile_w, p.tile_h, p.tile_n);
  else rc = make_act_map(&mA, x, Cin, W / 2, H / 2, 4, N, 64, p.tile_w, p.tile_h, p.tile_n);
  if (rc) return rc;
  if ((rc = make_act_map(&mB, wf, Cin, KH * KW * Cout, 1, 1, 1, 64, bn, 1, 1))) return rc;
  return umma_launch(mA, mB, p, bn, st);
}

int umma_dgrad(const void* dy, const void* wd, void* dx, int N, int H, int W, int Cin, int Cout, int KH, int KW,
               int stride, int pad_t, int pad_l, int Ho, int Wo, int dx_planes, cudaStream_t st) {
  VCD_CHECK_ARG(wd != nullptr, "tcgen05 dgrad needs the w_dgrad pack");
  const int bn = pick_block_n(Cin);
  CUtensorMap mA, mB;
  int rc;
  if (stride == 1) {
    VCD_CHECK_ARG(Ho == H && Wo == W, "tcgen05 dgrad: stride-1 convs must preserve the spatial size");
    UmmaParams p;
    memset(&p, 0, sizeof(p));
    p.form = 0;
    choose_tile(128, W, H, N, p);
    p.ntaps = KH * KW;
    for (int kh = 0; kh < KH; ++kh)
      for (int kw = 0; kw < KW; ++kw) {
        int t = kh * KW + kw;
        p.tap_dh[t] = pad_t - kh; p.tap_dw[t] = pad_l - kw; p.tap_plane[t] = 0;
        p.tap_brow[t] = t * Cin;
      }
    p.n_tiles = Cin / bn; p.kc_per_tap = Cout / 64;
    p.out = (bf16*)dx; p.alpha = 1.f;
    p.out_sn = (long long)H * W * Cin; p.out_sh = (long long)W * Cin; p.out_sw = Cin;
    p.Nout = Cin;
    set_form0_desc(p, bn);
    p.total_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles;
    if ((rc = make_act_map(&mA, dy, Cout, Wo, Ho, 1, N, 64, p.tile_w, p.tile_h, p.tile_n))) return rc;
    if ((rc = make_act_map(&mB, wd, Cout, KH * KW * Cin, 1, 1, 1, 64, bn, 1, 1))) return rc;
    return umma_launch(mA, mB, p, bn, st);
  }
  // stride 2: one launch per parity plane of dx (plane (ph,pw) receives the taps with matching parity)
  VCD_CHECK_ARG(dx_planes, "tcgen05 stride-2 dgrad writes the parity-plane layout (dx_planes = 1)");
  const int H2 = H / 2, W2 = W / 2;
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      UmmaParams p;
      memset(&p, 0, sizeof(p));
      p.form = 0;
      choose_tile(128, W2, H2, N, p);
      int nt = 0;
      for (int kh = 0; kh < KH; ++kh)
        for (int kw = 0; kw < KW; ++kw) {
          int eh = kh - pad_t, ew = kw - pad_l;
          if ((eh & 1) != ph || (ew & 1) != pw) continue;
          p.tap_dh[nt] = -(eh - ph) / 2; p.tap_dw[nt] = -(ew - pw) / 2; p.tap_plane[nt] = 0;
          p.tap_brow[nt] = (kh * KW + kw) * Cin;
          ++nt;
        }
      const long long plane_elems = (long long)H2 * W2 * Cin;
      bf16* outp = (bf16*)dx + (ph * 2 + pw) * plane_elems;
      if (nt == 0) {  // no tap reaches this plane: its gradient is zero
        for (int n = 0; n < N; ++n)
          VCD_CUDA(cudaMemsetAsync(outp + n * 4 * plane_elems, 0, plane_elems * sizeof(bf16), st));
        continue;
      }
      p.ntaps = nt;
      p.n_tiles = Cin / bn; p.kc_per_tap = Cout / 64;
      p.out = outp; p.alpha = 1.f;
      p.out_sn = 4 * plane_elems; p.out_sh = (long long)W2 * Cin; p.out_sw = Cin;
      p.Nout = Cin;
      set_form0_desc(p, bn);
      p.total_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles;
      if ((rc = make_act_map(&mA, dy, Cout, Wo, Ho, 1, N, 64, p.tile_w, p.tile_h, p.tile_n))) return rc;
      if ((rc = make_act_map(&mB, wd, Cout, KH * KW * Cin, 1, 1, 1, 64, bn, 1, 1))) return rc;
      if ((rc = umma_launch(mA, mB, p, bn, st))) return rc;
    }
  return 0;
}

// ws: zeroed fp32 [tap][Cout][Cin]
int umma_wgrad(const void* x, const void* dy, float* ws, int N, int H, int W, int Cin, int Cout, int KH, int KW,
               int stride, int pad_t, int pad_l, int Ho, int Wo, int x_planes, cudaStream_t st) {
  UmmaParams p;
  memset(&p, 0, sizeof(p));
  p.form = 1;
  VCD_CHECK_ARG(stride == 1 || x_planes, "tcgen05 stride-2 wgrad needs the parity-plane input");
  choose_tile(64, Wo, Ho, N, p);
  const int bn = pick_block_n(Cin);
  fill_fprop_taps(p, KH, KW, stride, pad_t, pad_l, 0);
  p.n_tiles = Cin / bn;
  p.m_tiles = Cout / 128;
  p.Mout = Cout; p.Nout = Cin;
  p.acc = ws;
  p.batches = 1;
  p.k_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int base_tiles = p.ntaps * p.m_tiles * p.n_tiles;
  int splits = (2 * vcd_num_sms() + base_tiles - 1) / base_tiles;
  if (splits > p.k_tiles) splits = p.k_tiles;
  if (splits < 1) splits = 1;
  p.k_per_split = (p.k_tiles + splits - 1) / splits;
  p.splits = (p.k_tiles + p.k_per_split - 1) / p.k_per_split;
  set_form1_desc(p, bn);
  p.total_tiles = base_tiles * p.splits;
  CUtensorMap mA, mB;
  int rc;
  if ((rc = make_act_map(&mA, dy, Cout, Wo, Ho, 1, N, 64, p.tile_w, p.tile_h, p.tile_n))) return rc;
  if (stride == 1) rc = make_act_map(&mB, x, Cin, W, H, 1, N, 64, p.tile_w, p.tile_h, p.tile_n);
  else rc = make_act_map(&mB, x, Cin, W / 2, H / 2, 4, N, 64, p.tile_w, p.tile_h, p.tile_n);
  if (rc) return rc;
  return umma_launch(mA, mB, p, bn, st);
}

__global__ void convert_f32_kernel(const float* __restrict__ in, void* __restrict__ out, int dt, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    store_param(out, dt, i, in[i]);
}

}  // namespace

extern "C" int vcd_conv_umma_supported(int Cin, int Cout, int KH, int KW, int stride) {
  return umma_shape_ok(Cin, Cout, KH, KW, stride) ? 1 : 0;
}

extern "C" int vcd_conv2d_fprop(const void* x, const void* w_fprop, const float* bias, const void* residual, void* y,
                                int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad_t, int pad_l,
                                int Ho, int Wo, int x_planes, int impl, vcd_stream_t stream) {
  VCD_CHECK_ARG(x && w_fprop && y, "conv fprop: null pointer");
  bool ok = umma_shape_ok(Cin, Cout, KH, KW, stride);
  if (impl == VCD_IMPL_UMMA) VCD_CHECK_ARG(ok, "conv fprop: shape (Cin=%d,Cout=%d,k=%d,s=%d) not supported by the tcgen05 path", Cin, Cout, KH, stride);
  if (ok && impl != VCD_IMPL_SIMT)
    return umma_fprop(x, w_fprop, bias, residual, y, N, H, W, Cin, Cout, KH, KW, stride, pad_t, pad_l, Ho, Wo, x_planes,
                      as_stream(stream));
  VCD_CHECK_ARG(!x_planes, "conv fprop (SIMT): parity-plane input not supported");
  return simt_conv_fprop(x, w_fprop, bias, residual, y, N, H, W, Cin, Cout, KH, KW, stride, pad_t, pad_l, Ho, Wo,
                         as_stream(stream));
}

extern "C" int vcd_conv2d_dgrad(const void* dy, const void* w_fprop, const void* w_dgrad, void* dx, int N, int H, int W,
                                int Cin, int Cout, int KH, int KW, int stride, int pad_t, int pad_l, int Ho, int Wo,
                                int dx_planes, int impl, vcd_stream_t stream) {
  (void)w_fprop;
  VCD_CHECK_ARG(dy && dx, "conv dgrad: null pointer");
  bool ok = umma_shape_ok(Cin, Cout, KH, KW, stride);
  if (impl == VCD_IMPL_UMMA) VCD_CHECK_ARG(ok, "conv dgrad: shape not supported by the tcgen05 path");
  if (ok && impl != VCD_IMPL_SIMT)
    return umma_dgrad(dy, w_dgrad, dx, N, H, W, Cin, Cout, KH, KW, stride, pad_t, pad_l, Ho, Wo, dx_planes,
                      as_stream(stream));
  VCD_CHECK_ARG(!dx_planes, "conv dgrad (SIMT): parity-plane output not supported");
  return simt_conv_dgrad(dy, w_dgrad, dx, N, H, W, Cin, Cout, KH, KW, stride, pad_t, pad_l, Ho, Wo, as_stream(stream));
}

extern "C" int64_t vcd_conv2d_wgrad_ws_bytes(int Cin, int Cout, int KH, int KW) {
  return ((int64_t)KH * KW * Cout * Cin + Cout) * (int64_t)sizeof(float);
}

extern "C" int vcd_conv2d_wgrad(const void* x, const void* dy, void* dw, void* db, int dtype, void* ws, int N, int H,
                                int W, int Cin, int Cout, int KH, int KW, int stride, int pad_t, int pad_l, int Ho, int Wo,
                                int x_planes, int impl, vcd_stream_t stream) {
  VCD_CHECK_ARG(x && dy && dw && ws, "conv wgrad: null pointer");
  cudaStream_t st = as_stream(stream);
  VCD_CUDA(cudaMemsetAsync(ws, 0, (size_t)vcd_conv2d_wgrad_ws_bytes(Cin, Cout, KH, KW), st));
  float* wsf = (float*)ws;
  bool ok = umma_shape_ok(Cin, Cout, KH, KW, stride);
  if (impl == VCD_IMPL_UMMA) VCD_CHECK_ARG(ok, "conv wgrad: shape not supported by the tcgen05 path");
  int rc;
  if (ok && impl != VCD_IMPL_SIMT) {
    rc = umma_wgrad(x, dy, wsf, N, H, W, Cin, Cout, KH, KW, stride, pad_t, pad_l, Ho, Wo, x_planes, st);
  } else {
    VCD_CHECK_ARG(!x_planes, "conv wgrad (SIMT): parity-plane input not supported");
    rc = simt_conv_wgrad(x, dy, wsf, N, H, W, Cin, Cout, KH, KW, stride, pad_t, pad_l, Ho, Wo, st);
  }
  if (rc) return rc;
  if (db && (rc = conv_bias_grad(dy, wsf + (int64_t)KH * KW * Cout * Cin, (int64_t)N * Ho * Wo, Cout, st))) return rc;
  return conv_wgrad_finalize(wsf, dw, db, dtype, Cout, Cin, KH * KW, st);
}

// D[b][m][n] = alpha * sum_k A[b][m][k] B[(b)][n][k] (+bias[n]) (+residual[b][m][n])
extern "C" int vcd_gemm_nt(const void* A, const void* B, const float* bias, const void* residual, void* D, int batch,
                           int M, int Nn, int K, int b_batched, float alpha, vcd_stream_t stream) {
  VCD_CHECK_ARG(A && B && D, "gemm_nt: null pointer");
  VCD_CHECK_ARG(K % 64 == 0 && Nn % 8 == 0, "gemm_nt: need K %% 64 == 0 and N %% 8 == 0 (K=%d N=%d)", K, Nn);
  UmmaParams p;
  memset(&p, 0, sizeof(p));
  p.form = 0;
  p.W = M; p.H = 1; p.Nimg = batch;
  p.tile_w = 128; p.tile_h = 1; p.tile_n = 1;
  p.tiles_w = (M + 127) / 128; p.tiles_h = 1; p.tiles_n = batch;
  p.ntaps = 1;
  const int bn = pick_block_n(Nn);
  p.n_tiles = (Nn + bn - 1) / bn;
  p.kc_per_tap = K / 64;
  p.b_batch_rows = b_batched ? Nn : 0;
  p.out = (bf16*)D; p.residual = (const bf16*)residual; p.bias = bias; p.alpha = alpha;
  p.out_sn = (long long)M * Nn; p.out_sh = 0; p.out_sw = Nn;
  p.Nout = Nn;
  set_form0_desc(p, bn);
  p.total_tiles = p.tiles_w * p.tiles_n * p.n_tiles;
  CUtensorMap mA, mB;
  int rc;
  if ((rc = make_act_map(&mA, A, K, M, 1, 1, batch, 64, 128, 1, 1))) return rc;
  if ((rc = make_act_map(&mB, B, K, b_batched ? batch * Nn : Nn, 1, 1, 1, 64, bn, 1, 1))) return rc;
  return umma_launch(mA, mB, p, bn, as_stream(stream));
}

// D[b][m][n] = sum_k A[b][k][m] B[b][k][n]; reduce_batch sums over b as well (Linear wgrad)
extern "C" int vcd_gemm_tn(const void* A, const void* B, void* D, int d_dtype, void* ws_f32, int batch, int M, int Nn,
                           int K, int reduce_batch, vcd_stream_t stream) {
  VCD_CHECK_ARG(A && B && D && ws_f32, "gemm_tn: null pointer");
  VCD_CHECK_ARG(M % 8 == 0 && Nn % 8 == 0, "gemm_tn: need M %% 8 == 0 and N %% 8 == 0");
  cudaStream_t st = as_stream(stream);
  UmmaParams p;
  memset(&p, 0, sizeof(p));
  p.form = 1;
  p.W = K; p.H = 1; p.Nimg = batch;
  p.tile_w = 64; p.tile_h = 1; p.tile_n = 1;
  p.tiles_w = (K + 63) / 64; p.tiles_h = 1; p.tiles_n = batch;
  p.ntaps = 1;
  const int bn = pick_block_n(Nn);
  p.n_tiles = (Nn + bn - 1) / bn;
  p.m_tiles = (M + 127) / 128;
  p.Mout = M; p.Nout = Nn;
  p.acc = (float*)ws_f32;
  p.batches = reduce_batch ? 1 : batch;
  p.k_tiles = reduce_batch ? p.tiles_w * batch : p.tiles_w;
  const int base_tiles = p.batches * p.m_tiles * p.n_tiles;
  int splits = (vcd_num_sms() + base_tiles - 1) / base_tiles;
  if (splits > p.k_tiles) splits = p.k_tiles;
  if (splits < 1) splits = 1;
  p.k_per_split = (p.k_tiles + splits - 1) / splits;
  p.splits = (p.k_tiles + p.k_per_split - 1) / p.k_per_split;
  set_form1_desc(p, bn);
  p.total_tiles = base_tiles * p.splits;
  const int64_t out_elems = (int64_t)p.batches * M * Nn;
  VCD_CUDA(cudaMemsetAsync(ws_f32, 0, out_elems * sizeof(float), st));
  CUtensorMap mA, mB;
  int rc;
  if ((rc = make_act_map(&mA, A, M, K, 1, 1, batch, 64, 64, 1, 1))) return rc;
  if ((rc = make_act_map(&mB, B, Nn, K, 1, 1, batch, 64, 64, 1, 1))) return rc;
  if ((rc = umma_launch(mA, mB, p, bn, st))) return rc;
  int64_t blocks = ceil_div64(out_elems, 256);
  if (blocks > vcd_num_sms() * 8) blocks = vcd_num_sms() * 8;
  convert_f32_kernel<<<(unsigned)blocks, 256, 0, st>>>((const float*)ws_f32, D, d_dtype, out_elems);
  VCD_LAUNCH_CHECK();
  return 0;
}
