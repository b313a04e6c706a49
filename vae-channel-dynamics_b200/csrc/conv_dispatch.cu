// conv_dispatch.cu — C-ABI convolution / GEMM entry points: shape checks, TMA tensor maps, tile
// decomposition and tap tables for the tcgen05 kernel (umma_gemm.cu), SIMT path for small channels.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "conv_dispatch.h"
#include "umma_gemm.cuh"
#include "umma_pair.cuh"

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
bool* vcd_device_once(int slot) {
  static bool flags[64][16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  return &flags[dev & 63][slot & 15];
}
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
// split-K factor: the largest split that keeps tiles <= 2 full waves of the persistent grid (a partial third
// wave costs a whole wave), but never more splits than K steps
int pick_splits(int base_tiles, int k_tiles) {
  const int target = 2 * vcd_num_sms();
  int splits = base_tiles >= target ? 1 : target / base_tiles;
  if (splits > k_tiles) splits = k_tiles;
  return splits < 1 ? 1 : splits;
}

// ---------------------------------------------------------------- CTA-pair kernel (umma_pair.cu) front end
// BLOCK_N for the pair kernel: 256 when Nout allows it, unless 128 fills the last wave of 74 clusters visibly better
int pair_block_n(int pairs, int Nout) {
  // N = 128 tiles are bound by the shared-memory operand bandwidth (A 4 KB + B 2 KB per 64-cycle MMA), so the wider
  // tile wins whenever the channel count allows it, even with a ragged last wave
  (void)pairs;
  return (Nout % 256 != 0) ? 128 : 256;
}
// attach the fused GroupNorm-sum epilogue when the output shape allows it (zeroes the sums buffer)
int attach_gn(PairParams& p, GnEpilogue* gn, int n_images, int Nout, int rows_per_img, cudaStream_t st) {
  if (!gn || !gn->sums || gn->groups <= 0 || Nout % gn->groups != 0) return 0;
  const int D = Nout / gn->groups;
  const int logD = D == 4 ? 2 : D == 8 ? 3 : D == 16 ? 4 : -1;
  if (logD < 0 || Nout % 32 != 0) return 0;
  if (p.mode == 1 && (rows_per_img <= 0 || rows_per_img % 128 != 0)) return 0;
  p.gn_sums = gn->sums; p.gn_G = gn->groups; p.gn_logD = logD; p.gn_rows_per_img = rows_per_img;
  if (!gn->accumulate && !gn->prezeroed)
    VCD_CUDA(cudaMemsetAsync(gn->sums, 0, sizeof(double) * 2 * n_images * gn->groups, st));
  gn->fused = true;
  return 0;
}
// implicit-GEMM convolution with halo reuse: returns 1 when launched, 0 when the shape is not eligible, < 0 on error
int try_pair_halo(const void* act, int C, int Wa, int Ha, int P, int N, int Wt, int Ht, const PairTap* taps, int ntaps,
                  const void* wpack, int wrows, const float* bias, const void* residual, void* out, long long sn,
                  long long sh, long long sw, int Nout, cudaStream_t st, GnEpilogue* gn = nullptr,
                  const GnBwdPrologue* gnb = nullptr, int es = 1) {
  if (!pair_enabled() || C % 64 != 0) return 0;
  PairParams p;
  memset(&p, 0, sizeof(p));
  int box_h = 0;
  if (!pair_setup_halo(p, Wt, Ht, N, taps, ntaps, 128, &box_h)) return 0;
  const int bn = pair_block_n(p.pairs, Nout);
  pair_setup_halo(p, Wt, Ht, N, taps, ntaps, bn, &box_h);
  p.n_tiles = (Nout + bn - 1) / bn;
  p.kc = C / 64;
  p.out = (bf16*)out; p.residual = (const bf16*)residual; p.bias = bias; p.alpha = 1.f;
  p.out_sn = sn; p.out_sh = sh; p.out_sw = sw;
  p.Nout = Nout;
  CUtensorMap mA, mB;
  int rc;
  // 128 -> 128 channel 3x3 layers (N = 128 tile, K = 1152): the item's MMAs last ~4.6k cycles and the GroupNorm-sum epilogue
  // (32 adds + 32 FMAs + a 15-shuffle transpose-reduce + fp64 atomics per 32-column chunk) does not fit behind them: measured
  // on B200, 128->128 @512^2 B=8: 0.410 ms without the sums, 0.559 ms with them, against 0.101 ms for the stand-alone
  // vcd_gn_stats pass over the output (5.3 TB/s).  For these layers the caller's fallback (vcd_gn_stats after the GEMM) wins;
  // VCD_GN_FUSE_128=1 restores the fused form (A/B measurement).
  static int fuse128 = -1;
  if (fuse128 < 0) { const char* e = getenv("VCD_GN_FUSE_128"); fuse128 = (e && e[0] == '1') ? 1 : 0; }
  const bool short_item = bn == 128 && ntaps * C <= 1152;
  if ((!short_item || fuse128) && (rc = attach_gn(p, gn, N, Nout, 0, st))) return rc;
  if (gnb) {
    VCD_CHECK_ARG(Nout <= 512 && Nout % 32 == 0, "fused GroupNorm backward: channels must be a multiple of 32, <= 512");
    p.gnb_x = (const bf16*)gnb->x; p.gnb_ab = gnb->ab; p.gnb_dsdb = gnb->dsdb; p.gnb_act = gnb->act;
  }
  p.a_es = es;
  // 128 -> 128 channel 3x3 layers: the CTA's whole weight operand (ntaps * kc boxes of 8 KB) fits beside two A boxes, so it
  // is loaded once per kernel instead of once per 256-pixel item (VCD_BRES=0 disables, for A/B measurement)
  {
    static int bres = -1;
    if (bres < 0) { const char* e = getenv("VCD_BRES"); bres = (e && e[0] == '0') ? 0 : 1; }
    p.b_resident = (bres && bn == 128 && p.n_tiles == 1 && ntaps * p.kc * 64 * 128 <= 147456 && p.pairs >= 4 * 74) ? 1 : 0;
  }
  if ((rc = make_act_map(&mA, act, C, Wa, Ha, P, N, 64, p.box_w, box_h, 1, es))) return rc;
  if ((rc = make_act_map(&mB, wpack, C, wrows, 1, 1, 1, 64, bn / 2, 1, 1))) return rc;
  if ((rc = pair_launch(mA, mB, p, bn, st))) return rc;
  return 1;
}
// plain GEMM rows x Nout (1x1 convolutions, Linear, attention products): D[b][m][n] = alpha * A[b][m][:] . B[(b)][n][:]
int pair_rows_gemm(const void* A, const void* B, const float* bias, const void* residual, void* D, int batch, int M, int Nn,
                   int K, int b_batched, float alpha, cudaStream_t st, GnEpilogue* gn = nullptr, int gn_images = 0,
                   int gn_rows_per_img = 0) {
  PairParams p;
  memset(&p, 0, sizeof(p));
  const int tiles = (M + 127) / 128;
  const int pairs = b_batched ? ((tiles + 1) / 2) * batch : (tiles * batch + 1) / 2;
  const int bn = pair_block_n(pairs, Nn);
  pair_setup_rows(p, M, batch, b_batched ? 1 : 0, 0, bn);
  p.n_tiles = (Nn + bn - 1) / bn;
  p.kc = K / 64;
  p.b_batch_rows = b_batched ? Nn : 0;
  p.out = (bf16*)D; p.residual = (const bf16*)residual; p.bias = bias; p.alpha = alpha;
  p.out_sn = (long long)M * Nn; p.out_sh = 0; p.out_sw = Nn;
  p.Nout = Nn;
  CUtensorMap mA, mB;
  int rc;
  if ((rc = attach_gn(p, gn, gn_images, Nn, gn_rows_per_img, st))) return rc;
  if ((rc = make_act_map(&mA, A, K, M, 1, 1, batch, 64, 128, 1, 1))) return rc;
  if ((rc = make_act_map(&mB, B, K, b_batched ? batch * Nn : Nn, 1, 1, 1, 64, bn / 2, 1, 1))) return rc;
  return pair_launch(mA, mB, p, bn, st);
}

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
               cudaStream_t st, GnEpilogue* gn = nullptr) {
  UmmaParams p;
  memset(&p, 0, sizeof(p));
  p.form = 0;
  VCD_CHECK_ARG(stride == 2 || (Ho == H && Wo == W), "tcgen05 conv: stride-1 convs must preserve the spatial size");
  // stride 2: parity planes are either a physical layout (x_planes, vcd_space_to_planes) or, by default, read in place
  // from the NHWC tensor through an element-strided TMA map (make_act_map es = 2)
  const int es = (stride == 2 && !x_planes) ? 2 : 1;
  choose_tile(128, Wo, Ho, N, p);
  const int bn = pick_block_n(Cout);
  fill_fprop_taps(p, KH, KW, stride, pad_t, pad_l, Cout);
  if (pair_enabled()) {
    int rc;
    if (KH * KW == 1 && stride == 1) {  // 1x1 shortcut: a plain GEMM over all pixels
      VCD_CHECK_ARG((long long)N * H * W < (1ll << 31), "1x1 conv: too many pixels");
      return pair_rows_gemm(x, wf, bias, residual, y, 1, N * H * W, Cout, Cin, 0, 1.f, st, gn, N, H * W);
    }
    PairTap taps[16];
    for (int t = 0; t < p.ntaps; ++t) taps[t] = PairTap{p.tap_plane[t], p.tap_dh[t], p.tap_dw[t], p.tap_brow[t]};
    rc = stride == 1 ? try_pair_halo(x, Cin, W, H, 1, N, Wo, Ho, taps, p.ntaps, wf, KH * KW * Cout, bias, residual, y,
                                     (long long)Ho * Wo * Cout, (long long)Wo * Cout, Cout, Cout, st, gn)
                     : es == 2
                           ? try_pair_halo(x, Cin, W, H, 1, N, Wo, Ho, taps, p.ntaps, wf, KH * KW * Cout, bias, residual, y,
                                           (long long)Ho * Wo * Cout, (long long)Wo * Cout, Cout, Cout, st, gn, nullptr, 2)
                           : try_pair_halo(x, Cin, W / 2, H / 2, 4, N, Wo, Ho, taps, p.ntaps, wf, KH * KW * Cout, bias,
                                           residual, y, (long long)Ho * Wo * Cout, (long long)Wo * Cout, Cout, Cout, st, gn);
    if (rc != 0) return rc < 0 ? rc : 0;
  }
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
  p.a_es = es;
  if (stride == 1 || es == 2) rc = make_act_map(&mA, x, Cin, W, H, 1, N, 64, p.tile_w, p.tile_h, p.tile_n, es);
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
    if (pair_enabled()) {
      if (KH * KW == 1) {
        VCD_CHECK_ARG((long long)N * H * W < (1ll << 31), "1x1 conv: too many pixels");
        return pair_rows_gemm(dy, wd, nullptr, nullptr, dx, 1, N * H * W, Cin, Cout, 0, 1.f, st);
      }
      PairTap taps[16];
      for (int t = 0; t < p.ntaps; ++t) taps[t] = PairTap{0, p.tap_dh[t], p.tap_dw[t], p.tap_brow[t]};
      rc = try_pair_halo(dy, Cout, Wo, Ho, 1, N, W, H, taps, p.ntaps, wd, KH * KW * Cin, nullptr, nullptr, dx,
                         (long long)H * W * Cin, (long long)W * Cin, Cin, Cin, st);
      if (rc != 0) return rc < 0 ? rc : 0;
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
  // stride 2: one launch per parity plane of dx (plane (ph,pw) receives the taps with matching parity); the plane is
  // either a slab of the parity-plane layout (dx_planes) or a strided view of the NHWC tensor written in place
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
      bf16* outp = dx_planes ? (bf16*)dx + (ph * 2 + pw) * plane_elems : (bf16*)dx + ((long long)ph * W + pw) * Cin;
      const long long o_sn = 4 * plane_elems;
      const long long o_sh = dx_planes ? (long long)W2 * Cin : 2ll * W * Cin;
      const long long o_sw = dx_planes ? Cin : 2ll * Cin;
      if (nt == 0) {  // no tap reaches this plane: its gradient is zero
        VCD_CHECK_ARG(dx_planes, "stride-2 dgrad: a parity plane without taps needs the plane layout");
        for (int n = 0; n < N; ++n)
          VCD_CUDA(cudaMemsetAsync(outp + n * 4 * plane_elems, 0, plane_elems * sizeof(bf16), st));
        continue;
      }
      p.ntaps = nt;
      {
        PairTap taps[16];
        for (int t = 0; t < nt; ++t) taps[t] = PairTap{0, p.tap_dh[t], p.tap_dw[t], p.tap_brow[t]};
        rc = try_pair_halo(dy, Cout, Wo, Ho, 1, N, W2, H2, taps, nt, wd, KH * KW * Cin, nullptr, nullptr, outp, o_sn, o_sh,
                           o_sw, Cin, st);
        if (rc < 0) return rc;
        if (rc == 1) continue;
      }
      p.n_tiles = Cin / bn; p.kc_per_tap = Cout / 64;
      p.out = outp; p.alpha = 1.f;
      p.out_sn = o_sn; p.out_sh = o_sh; p.out_sw = o_sw;
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
               int stride, int pad_t, int pad_l, int Ho, int Wo, int x_planes, cudaStream_t st, bool overlap_prev = false) {
  UmmaParams p;
  memset(&p, 0, sizeof(p));
  p.form = 1;
  const int es = (stride == 2 && !x_planes) ? 2 : 1;  // stride 2 without the plane layout: element-strided TMA map
  choose_tile(64, Wo, Ho, N, p);
  const int bn = pick_block_n(Cin);
  fill_fprop_taps(p, KH, KW, stride, pad_t, pad_l, 0);
  if (pair_enabled() && Cout % 256 == 0) {  // CTA-pair kernel: M = 256 output channels per cluster
    PairWgradParams q;
    memset(&q, 0, sizeof(q));
    q.W = p.W; q.H = p.H; q.Nimg = p.Nimg;
    q.tile_w = p.tile_w; q.tile_h = p.tile_h; q.tile_n = p.tile_n;
    q.tiles_w = p.tiles_w; q.tiles_h = p.tiles_h; q.tiles_n = p.tiles_n;
    q.ntaps = p.ntaps;
    for (int t = 0; t < p.ntaps; ++t) {
      q.tap_dw[t] = p.tap_dw[t]; q.tap_dh[t] = p.tap_dh[t]; q.tap_plane[t] = p.tap_plane[t]; q.tap_plane_a[t] = 0;
    }
    q.m_pairs = Cout / 256;
    q.n_tiles = Cin / bn;
    q.k_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    const int base = q.ntaps * q.m_pairs * q.n_tiles, target = vcd_num_sms();  // two waves of 74 clusters
    int splits = base >= target ? 1 : target / base;
    if (splits > q.k_tiles) splits = q.k_tiles;
    if (splits < 1) splits = 1;
    q.k_per_split = (q.k_tiles + splits - 1) / splits;
    q.splits = (q.k_tiles + q.k_per_split - 1) / q.k_per_split;
    q.acc = ws; q.Mout = Cout; q.Nout = Cin;
    CUtensorMap mA, mB;
    int rc;
    q.b_es = es;
    if ((rc = make_act_map(&mA, dy, Cout, Wo, Ho, 1, N, 64, p.tile_w, p.tile_h, p.tile_n))) return rc;
    if (stride == 1 || es == 2) rc = make_act_map(&mB, x, Cin, W, H, 1, N, 64, p.tile_w, p.tile_h, p.tile_n, es);
    else rc = make_act_map(&mB, x, Cin, W / 2, H / 2, 4, N, 64, p.tile_w, p.tile_h, p.tile_n);
    if (rc) return rc;
    return pair_wgrad_launch(mA, mB, q, bn, st, overlap_prev);
  }
  // Cin = 128: one 256-column tile = two taps (umma_gemm.cuh tap_pairs)
  const bool pairs = Cin == 128 && p.ntaps > 1;
  const int bn_eff = pairs ? 256 : bn;
  p.tap_pairs = pairs ? 1 : 0;
  p.tap_items = pairs ? (p.ntaps + 1) / 2 : p.ntaps;
  p.n_tiles = pairs ? 1 : Cin / bn;
  p.m_tiles = Cout / 128;
  p.Mout = Cout; p.Nout = Cin;
  p.acc = ws;
  p.batches = 1;
  p.k_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int base_tiles = p.tap_items * p.m_tiles * p.n_tiles;
  int splits = pick_splits(base_tiles, p.k_tiles);
  p.k_per_split = (p.k_tiles + splits - 1) / splits;
  p.splits = (p.k_tiles + p.k_per_split - 1) / p.k_per_split;
  set_form1_desc(p, bn_eff);
  p.total_tiles = base_tiles * p.splits;
  CUtensorMap mA, mB;
  int rc;
  p.b_es = es;
  if ((rc = make_act_map(&mA, dy, Cout, Wo, Ho, 1, N, 64, p.tile_w, p.tile_h, p.tile_n))) return rc;
  if (stride == 1 || es == 2) rc = make_act_map(&mB, x, Cin, W, H, 1, N, 64, p.tile_w, p.tile_h, p.tile_n, es);
  else rc = make_act_map(&mB, x, Cin, W / 2, H / 2, 4, N, 64, p.tile_w, p.tile_h, p.tile_n);
  if (rc) return rc;
  return umma_launch(mA, mB, p, bn_eff, st, overlap_prev);
}


// ---------------------------------------------------------------- small-channel layers on the GEMM kernel
// patch[px][col], col = t*S + s  ->  src[n, h + dh[t], w + dw[t], s]   (zero outside the image / beyond taps*S)
struct PatchTaps { int n; int dh[9], dw[9]; };
// One block = 4 image rows x 64 consecutive pixels.  The six source rows (h0-1 .. h0+4, 66 pixels each, zero outside the
// image) are staged in shared memory with coalesced loads, every patch column's source offset comes from a small table
// (no per-element division), and each thread writes whole 16-byte vectors: the kernel runs at the rate the patch can be
// WRITTEN (round 1's per-element gather issued eight scattered 2-byte loads and eight divisions per vector: 0.27 ms per
// 512^2 patch, four of them per step; one image row per block kept only 8 KB of stores in flight per block: 0.12-0.24 ms).
constexpr int kI2cPx = 64, kI2cRows = 4;
__global__ void __launch_bounds__(256) im2col_small_kernel(const bf16* __restrict__ src, bf16* __restrict__ patch, int N,
                                                           int H, int W, int S, int Kp, PatchTaps taps) {
  extern __shared__ unsigned char i2c_smem[];
  bf16* rows = reinterpret_cast<bf16*>(i2c_smem);                                   // [kI2cRows + 2][kI2cPx + 2][S]
  short* off = reinterpret_cast<short*>(rows + (kI2cRows + 2) * (kI2cPx + 2) * S);  // [Kp]: offset relative to the pixel, -1 = zero
  const int segs = (W + kI2cPx - 1) / kI2cPx, hblocks = (H + kI2cRows - 1) / kI2cRows;
  const int seg = blockIdx.x % segs;
  const int t0 = blockIdx.x / segs;
  const int h0 = (t0 % hblocks) * kI2cRows;
  const int64_t n = t0 / hblocks;
  const int w0 = seg * kI2cPx;
  const int RW = (kI2cPx + 2) * S;
  for (int i = threadIdx.x; i < (kI2cRows + 2) * RW; i += blockDim.x) {
    const int r = i / RW, e = i - r * RW;
    const int wi = w0 - 1 + e / S, hi = h0 - 1 + r;
    bf16 v = __float2bfloat16_rn(0.f);
    if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = src[((n * H + hi) * W + wi) * S + (e % S)];
    rows[i] = v;
  }
  for (int col = threadIdx.x; col < Kp; col += blockDim.x) {
    const int t = col / S, c = col - t * S;
    off[col] = t < taps.n ? (short)(((taps.dh[t] + 1) * (kI2cPx + 2) + (taps.dw[t] + 1)) * S + c) : (short)-1;
  }
  __syncthreads();
  const int V = Kp / 8;
  const int npx = min(kI2cPx, W - w0), nrows = min(kI2cRows, H - h0);
  for (int i = threadIdx.x; i < nrows * npx * V; i += blockDim.x) {
    const int v = i % V, q = i / V;
    const int px = q % npx, r = q / npx;
    const bf16* base = rows + (r * (kI2cPx + 2) + px) * S;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int o = off[v * 8 + j];
      f[j] = o >= 0 ? __bfloat162float(base[o]) : 0.f;
    }
    st8(patch + (((n * H + h0 + r) * W + w0 + px) * (int64_t)Kp) + v * 8, pack8(f));
  }
}
static int im2col_small_launch(const bf16* src, bf16* patch, int N, int H, int W, int S, int Kp, const PatchTaps& taps,
                               cudaStream_t st) {
  for (int t = 0; t < taps.n; ++t)
    VCD_CHECK_ARG(taps.dh[t] >= -1 && taps.dh[t] <= 1 && taps.dw[t] >= -1 && taps.dw[t] <= 1, "im2col: taps beyond 3x3");
  const int segs = (W + kI2cPx - 1) / kI2cPx, hblocks = (H + kI2cRows - 1) / kI2cRows;
  const int64_t blocks = (int64_t)N * hblocks * segs;
  VCD_CHECK_ARG(blocks < (1ll << 31), "im2col: too many rows");
  const size_t smem = (size_t)(kI2cRows + 2) * (kI2cPx + 2) * S * sizeof(bf16) + (size_t)Kp * sizeof(short);
  im2col_small_kernel<<<(unsigned)blocks, 256, smem, st>>>(src, patch, N, H, W, S, Kp, taps);
  VCD_LAUNCH_CHECK();
  return 0;
}
// wk[o][t*S + s] = pack[t][o][s]   (pack = w_fprop [tap][Cout][Cin] or w_dgrad [tap][Cin][Cout])
__global__ void repack_small_kernel(const bf16* __restrict__ pack, bf16* __restrict__ wk, int taps, int O, int S, int Kp) {
  const int total = O * Kp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int col = i % Kp, o = i / Kp;
    const int t = col / S, c = col - t * S;
    wk[i] = t < taps ? pack[((int64_t)t * O + o) * S + c] : __float2bfloat16_rn(0.f);
  }
}
// D [big][Np] (col = t*S + s) -> ws [tap][Cout][Cin]; big_is_cout: big = co, s = ci; else big = ci, s = co
__global__ void wgrad_small_reorder_kernel(const float* __restrict__ D, float* __restrict__ ws, int taps, int Cout, int Cin,
                                           int Np, int big_is_cout) {
  const int total = taps * Cout * Cin;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ci = i % Cin;
    const int r = i / Cin;
    const int co = r % Cout, t = r / Cout;
    ws[i] = big_is_cout ? D[(int64_t)co * Np + t * Cin + ci] : D[(int64_t)ci * Np + t * Cout + co];
  }
}
int round_up(int v, int m) { return (v + m - 1) / m * m; }
int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }

int gemm_nt_impl(const void* A, const void* B, const float* bias, const void* residual, void* D, int batch, int M, int Nn,
                 int K, int b_batched, float alpha, cudaStream_t st) {
  if (pair_enabled() && Nn >= 128) return pair_rows_gemm(A, B, bias, residual, D, batch, M, Nn, K, b_batched, alpha, st);
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
  return umma_launch(mA, mB, p, bn, st);
}

// acc (fp32 [batches][M][Nn]) = sum_k A[b][k][m] B[b][k][n]; acc is zeroed here
int gemm_tn_impl(const void* A, const void* B, float* acc, int batch, int M, int Nn, int64_t K, int reduce_batch,
                 cudaStream_t st) {
  UmmaParams p;
  memset(&p, 0, sizeof(p));
  p.form = 1;
  p.W = (int)K; p.H = 1; p.Nimg = batch;
  p.tile_w = 64; p.tile_h = 1; p.tile_n = 1;
  p.tiles_w = (int)((K + 63) / 64); p.tiles_h = 1; p.tiles_n = batch;
  p.ntaps = 1;
  const int bn = pick_block_n(Nn);
  p.n_tiles = (Nn + bn - 1) / bn;
  p.m_tiles = (M + 127) / 128;
  p.Mout = M; p.Nout = Nn;
  p.acc = acc;
  p.batches = reduce_batch ? 1 : batch;
  p.k_tiles = reduce_batch ? p.tiles_w * batch : p.tiles_w;
  p.tap_items = p.ntaps;
  const int base_tiles = p.batches * p.m_tiles * p.n_tiles;
  int splits = pick_splits(base_tiles, p.k_tiles);
  p.k_per_split = (p.k_tiles + splits - 1) / splits;
  p.splits = (p.k_tiles + p.k_per_split - 1) / p.k_per_split;
  set_form1_desc(p, bn);
  p.total_tiles = base_tiles * p.splits;
  VCD_CUDA(cudaMemsetAsync(acc, 0, (size_t)p.batches * M * Nn * sizeof(float), st));
  CUtensorMap mA, mB;
  int rc;
  if ((rc = make_act_map(&mA, A, M, (int)K, 1, 1, batch, 64, 64, 1, 1))) return rc;
  if ((rc = make_act_map(&mB, B, Nn, (int)K, 1, 1, batch, 64, 64, 1, 1))) return rc;
  return umma_launch(mA, mB, p, bn, st);
}

// small-channel classification (stride 1 only)
bool narrow_out_ok(int Kc, int Nout, int stride) { return stride == 1 && Kc % 64 == 0 && Nout < 128 && Nout >= 1; }
bool patch_ok(int small, int big, int stride, int taps) { return stride == 1 && small <= 8 && big >= 32 && big % 8 == 0 && taps <= 9; }

PatchTaps patch_taps(int KH, int KW, int pad_t, int pad_l, int sign) {
  PatchTaps t;
  t.n = KH * KW;
  for (int kh = 0; kh < KH; ++kh)
    for (int kw = 0; kw < KW; ++kw) {
      t.dh[kh * KW + kw] = sign * (kh - pad_t);
      t.dw[kh * KW + kw] = sign * (kw - pad_l);
    }
  return t;
}
int ew_blocks(int64_t work) {
  int64_t b = (work + 255) / 256, cap = (int64_t)vcd_num_sms() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// out[px][O] = bias + sum_{t,s} src[px + shift_t][s] * pack[t][O][s]  via  im2col patch + GEMM (K = taps*S padded to 64)
int patch_gemm(const void* src, const void* pack, const float* bias, void* out, void* ws, int N, int H, int W, int S, int O,
               int KH, int KW, int pad_t, int pad_l, int sign, cudaStream_t st) {
  VCD_CHECK_ARG(ws != nullptr, "small-channel conv needs a workspace (vcd_conv2d_*_ws_bytes)");
  const int taps = KH * KW, Kp = round_up(taps * S, 64);
  const int64_t px = (int64_t)N * H * W;
  bf16* patch = (bf16*)ws;
  bf16* wk = (bf16*)((char*)ws + align256(px * Kp * 2));
  {
    int rc_i = im2col_small_launch((const bf16*)src, patch, N, H, W, S, Kp, patch_taps(KH, KW, pad_t, pad_l, sign), st);
    if (rc_i) return rc_i;
  }
  repack_small_kernel<<<ew_blocks((int64_t)O * Kp), 256, 0, st>>>((const bf16*)pack, wk, taps, O, S, Kp);
  VCD_LAUNCH_CHECK();
  VCD_CHECK_ARG(px < (1ll << 31), "small-channel conv: too many pixels");
  return gemm_nt_impl(patch, wk, bias, nullptr, out, 1, (int)px, O, Kp, 0, 1.f, st);
}
int64_t patch_gemm_ws(int64_t px, int S, int O, int taps) {
  const int Kp = round_up(taps * S, 64);
  return align256(px * Kp * 2) + align256((int64_t)O * Kp * 2);
}

// narrow-N implicit GEMM: fprop with Cout < 128 (or dgrad with Cin < 128); weights pack [tap][Nout][Kc] used as is
int narrow_conv(const void* in, const void* pack, const float* bias, void* out, int N, int H, int W, int Kc, int Nout,
                int KH, int KW, int pad_t, int pad_l, int sign, cudaStream_t st) {
  UmmaParams p;
  memset(&p, 0, sizeof(p));
  p.form = 0;
  choose_tile(128, W, H, N, p);
  p.ntaps = KH * KW;
  for (int kh = 0; kh < KH; ++kh)
    for (int kw = 0; kw < KW; ++kw) {
      int t = kh * KW + kw;
      p.tap_dh[t] = sign * (kh - pad_t); p.tap_dw[t] = sign * (kw - pad_l); p.tap_plane[t] = 0;
      p.tap_brow[t] = t * Nout;
    }
  {
    PairTap taps[16];
    for (int t = 0; t < p.ntaps; ++t) taps[t] = PairTap{0, p.tap_dh[t], p.tap_dw[t], p.tap_brow[t]};
    int rc = try_pair_halo(in, Kc, W, H, 1, N, W, H, taps, p.ntaps, pack, KH * KW * Nout, bias, nullptr, out,
                           (long long)H * W * Nout, (long long)W * Nout, Nout, Nout, st);
    if (rc != 0) return rc < 0 ? rc : 0;
  }
  p.n_tiles = 1; p.kc_per_tap = Kc / 64;
  p.out = (bf16*)out; p.bias = bias; p.alpha = 1.f;
  p.out_sn = (long long)H * W * Nout; p.out_sh = (long long)W * Nout; p.out_sw = Nout;
  p.Nout = Nout;
  set_form0_desc(p, 128);
  p.total_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  CUtensorMap mA, mB;
  int rc;
  if ((rc = make_act_map(&mA, in, Kc, W, H, 1, N, 64, p.tile_w, p.tile_h, p.tile_n))) return rc;
  if ((rc = make_act_map(&mB, pack, Kc, KH * KW * Nout, 1, 1, 1, 64, 128, 1, 1))) return rc;
  return umma_launch(mA, mB, p, 128, st);
}

__global__ void convert_f32_from_param_kernel(const void* __restrict__ b, int dt, int n, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = load_param(b, dt, i);
}

__global__ void convert_f32_kernel(const float* __restrict__ in, void* __restrict__ out, int dt, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    store_param(out, dt, i, in[i]);
}

}  // namespace

// ---------------------------------------------------------------- Upsample2D (nearest x2) + 3x3 conv, fused
// [upstream] F.interpolate(x, 2, "nearest") followed by conv3x3(pad 1) equals four 2x2 convolutions on the
// LOW-resolution tensor, one per output parity (a, b), with pre-summed weights:
//   y[2h+a, 2w+b] = sum_{dh,dw in {0,1}} x[h + dh-1+a, w + dw-1+b] * Wsum[a][b][dh][dw]
//   Wsum[a][b][dh][dw] = sum_{kh in R(a,dh)} sum_{kw in R(b,dw)} W[kh][kw],  R(0,0)={0} R(0,1)={1,2} R(1,0)={0,1} R(1,1)={2}
// 16 tap-products per low-res pixel instead of 36: 2.25x fewer FLOPs, and the upsampled tensor never exists.
namespace {
__device__ __forceinline__ int up_group(int a, int k) { return a == 0 ? (k == 0 ? 0 : 1) : (k <= 1 ? 0 : 1); }

__global__ void pack_upconv_kernel(const void* __restrict__ w, int dt, int Cout, int Cin, bf16* __restrict__ wf,
                                   bf16* __restrict__ wd) {
  const int64_t total = (int64_t)16 * Cout * Cin;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    int64_t r = i / Cin;
    const int co = (int)(r % Cout);
    const int q = (int)(r / Cout);
    const int dw = q & 1, dh = (q >> 1) & 1, b = (q >> 2) & 1, a = (q >> 3) & 1;
    float acc = 0.f;
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw)
        if (up_group(a, kh) == dh && up_group(b, kw) == dw) acc += load_param(w, dt, (((int64_t)co * Cin + ci) * 3 + kh) * 3 + kw);
    const bf16 v = __float2bfloat16_rn(acc);
    wf[i] = v;
    wd[((int64_t)q * Cin + ci) * Cout + co] = v;
  }
}
// acc16 fp32 [16][Cout][Cin] -> ws [9][Cout][Cin]: dW[kh][kw] = sum_{a,b} acc16[a][b][group(a,kh)][group(b,kw)]
__global__ void upconv_wgrad_combine_kernel(const float* __restrict__ acc16, float* __restrict__ ws, int Cout, int Cin) {
  const int64_t per = (int64_t)Cout * Cin, total = 9 * per;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i / per);
    const int64_t e = i % per;
    const int kh = t / 3, kw = t % 3;
    float v = 0.f;
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) v += acc16[(int64_t)(((a * 2 + b) * 2 + up_group(a, kh)) * 2 + up_group(b, kw)) * per + e];
    ws[i] = v;
  }
}
}  // namespace

extern "C" int vcd_pack_upconv_weight(const void* w, const void* bias, int dtype, int Cout, int Cin, void* wf16,
                                      void* wd16, float* bias_f32, vcd_stream_t stream) {
  VCD_CHECK_ARG(w && wf16 && wd16, "pack_upconv_weight: null pointer");
  cudaStream_t st = as_stream(stream);
  pack_upconv_kernel<<<ew_blocks((int64_t)16 * Cout * Cin), 256, 0, st>>>(w, dtype, Cout, Cin, (bf16*)wf16, (bf16*)wd16);
  VCD_LAUNCH_CHECK();
  if (bias && bias_f32) {
    convert_f32_from_param_kernel<<<(Cout + 127) / 128, 128, 0, st>>>(bias, dtype, Cout, bias_f32);
    VCD_LAUNCH_CHECK();
  }
  return 0;
}

// x [N][H][W][Cin] -> y [N][2H][2W][Cout]
extern "C" int vcd_upconv2d_fprop(const void* x, const void* wf16, const float* bias, void* y, int N, int H, int W,
                                  int Cin, int Cout, double* gn_sums, int gn_groups, vcd_stream_t stream) {
  VCD_CHECK_ARG(x && wf16 && y, "upconv fprop: null pointer");
  VCD_CHECK_ARG(Cin % 128 == 0 && Cout % 128 == 0, "upconv: channels must be multiples of 128 (Cin=%d Cout=%d)", Cin, Cout);
  const int bn = pick_block_n(Cout);
  CUtensorMap mA, mB;
  int rc;
  // GroupNorm sums of y accumulate over the four phase launches (the first one zeroes the buffer)
  bool all_fused = gn_sums != nullptr;
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      UmmaParams p;
      memset(&p, 0, sizeof(p));
      p.form = 0;
      choose_tile(128, W, H, N, p);
      p.ntaps = 4;
      for (int dh = 0; dh < 2; ++dh)
        for (int dw = 0; dw < 2; ++dw) {
          const int t = dh * 2 + dw;
          p.tap_dh[t] = dh - 1 + a; p.tap_dw[t] = dw - 1 + b; p.tap_plane[t] = 0;
          p.tap_brow[t] = (((a * 2 + b) * 2 + dh) * 2 + dw) * Cout;
        }
      {
        PairTap taps[4];
        for (int t = 0; t < 4; ++t) taps[t] = PairTap{0, p.tap_dh[t], p.tap_dw[t], p.tap_brow[t]};
        GnEpilogue gph{gn_sums, gn_groups, false, (a | b) != 0};
        rc = try_pair_halo(x, Cin, W, H, 1, N, W, H, taps, 4, wf16, 16 * Cout, bias, nullptr,
                           (bf16*)y + ((long long)a * 2 * W + b) * Cout, 4ll * H * W * Cout, 4ll * W * Cout, 2ll * Cout,
                           Cout, as_stream(stream), (gn_sums && all_fused) ? &gph : nullptr);
        if (rc < 0) return rc;
        if (rc == 1) { all_fused = all_fused && gph.fused; continue; }
        all_fused = false;
      }
      p.n_tiles = Cout / bn; p.kc_per_tap = Cin / 64;
      p.out = (bf16*)y + ((long long)a * 2 * W + b) * Cout; p.bias = bias; p.alpha = 1.f;
      p.out_sn = 4ll * H * W * Cout; p.out_sh = 4ll * W * Cout; p.out_sw = 2ll * Cout;
      p.Nout = Cout;
      set_form0_desc(p, bn);
      p.total_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles;
      if ((rc = make_act_map(&mA, x, Cin, W, H, 1, N, 64, p.tile_w, p.tile_h, p.tile_n))) return rc;
      if ((rc = make_act_map(&mB, wf16, Cin, 16 * Cout, 1, 1, 1, 64, bn, 1, 1))) return rc;
      if ((rc = umma_launch(mA, mB, p, bn, as_stream(stream)))) return rc;
    }
  if (gn_sums && !all_fused) return vcd_gn_stats(y, gn_sums, nullptr, 0.f, N, 4 * H * W, Cout, gn_groups, stream);
  return 0;
}

// dy [N][2H][2W][Cout] (its four parity planes are read in place through an element-strided TMA map) -> dx [N][H][W][Cin]
extern "C" int vcd_upconv2d_dgrad(const void* dy_planes, const void* wd16, void* dx, int N, int H, int W, int Cin,
                                  int Cout, vcd_stream_t stream) {
  VCD_CHECK_ARG(dy_planes && wd16 && dx, "upconv dgrad: null pointer");
  VCD_CHECK_ARG(Cin % 128 == 0 && Cout % 128 == 0, "upconv: channels must be multiples of 128");
  const int bn = pick_block_n(Cin);
  UmmaParams p;
  memset(&p, 0, sizeof(p));
  p.form = 0;
  choose_tile(128, W, H, N, p);
  p.ntaps = 16;
  for (int q = 0; q < 16; ++q) {
    const int dw = q & 1, dh = (q >> 1) & 1, b = (q >> 2) & 1, a = (q >> 3) & 1;
    p.tap_dh[q] = -(dh - 1 + a); p.tap_dw[q] = -(dw - 1 + b); p.tap_plane[q] = a * 2 + b;
    p.tap_brow[q] = q * Cin;
  }
  {
    PairTap taps[16];
    for (int t = 0; t < 16; ++t) taps[t] = PairTap{p.tap_plane[t], p.tap_dh[t], p.tap_dw[t], p.tap_brow[t]};
    int prc = try_pair_halo(dy_planes, Cout, 2 * W, 2 * H, 1, N, W, H, taps, 16, wd16, 16 * Cin, nullptr, nullptr, dx,
                            (long long)H * W * Cin, (long long)W * Cin, Cin, Cin, as_stream(stream), nullptr, nullptr, 2);
    if (prc != 0) return prc < 0 ? prc : 0;
  }
  p.n_tiles = Cin / bn; p.kc_per_tap = Cout / 64;
  p.out = (bf16*)dx; p.alpha = 1.f;
  p.out_sn = (long long)H * W * Cin; p.out_sh = (long long)W * Cin; p.out_sw = Cin;
  p.Nout = Cin;
  set_form0_desc(p, bn);
  p.total_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles;
  CUtensorMap mA, mB;
  int rc;
  p.a_es = 2;
  if ((rc = make_act_map(&mA, dy_planes, Cout, 2 * W, 2 * H, 1, N, 64, p.tile_w, p.tile_h, p.tile_n, 2))) return rc;
  if ((rc = make_act_map(&mB, wd16, Cout, 16 * Cin, 1, 1, 1, 64, bn, 1, 1))) return rc;
  return umma_launch(mA, mB, p, bn, as_stream(stream));
}

extern "C" int64_t vcd_upconv2d_wgrad_ws_bytes(int Cin, int Cout) {
  return align256(((int64_t)9 * Cout * Cin + Cout) * 4) + align256((int64_t)16 * Cout * Cin * 4);
}

extern "C" int vcd_upconv2d_wgrad(const void* x, const void* dy_planes, void* dw, void* db, const float* db_colsum,
                                  int dtype, void* ws, int N, int H, int W, int Cin, int Cout, vcd_stream_t stream) {
  VCD_CHECK_ARG(x && dy_planes && dw && ws, "upconv wgrad: null pointer");
  VCD_CHECK_ARG(Cin % 128 == 0 && Cout % 128 == 0, "upconv: channels must be multiples of 128");
  cudaStream_t st = as_stream(stream);
  const int64_t main_elems = (int64_t)9 * Cout * Cin;
  float* wsf = (float*)ws;
  float* acc16 = (float*)((char*)ws + align256((main_elems + Cout) * 4));
  VCD_CUDA(cudaMemsetAsync(wsf + main_elems, 0, Cout * sizeof(float), st));
  VCD_CUDA(cudaMemsetAsync(acc16, 0, (size_t)16 * Cout * Cin * sizeof(float), st));
  UmmaParams p;
  memset(&p, 0, sizeof(p));
  p.form = 1;
  choose_tile(64, W, H, N, p);
  const int bn = pick_block_n(Cin);
  p.ntaps = 16;
  for (int q = 0; q < 16; ++q) {
    const int dw_ = q & 1, dh = (q >> 1) & 1, b = (q >> 2) & 1, a = (q >> 3) & 1;
    p.tap_dh[q] = dh - 1 + a; p.tap_dw[q] = dw_ - 1 + b; p.tap_plane[q] = 0;
    p.tap_plane_a[q] = a * 2 + b;
  }
  if (pair_enabled() && Cout % 256 == 0) {  // CTA-pair kernel: M = 256 output channels per cluster
    PairWgradParams q;
    memset(&q, 0, sizeof(q));
    q.W = p.W; q.H = p.H; q.Nimg = p.Nimg;
    q.tile_w = p.tile_w; q.tile_h = p.tile_h; q.tile_n = p.tile_n;
    q.tiles_w = p.tiles_w; q.tiles_h = p.tiles_h; q.tiles_n = p.tiles_n;
    q.ntaps = 16;
    for (int t = 0; t < 16; ++t) {
      q.tap_dw[t] = p.tap_dw[t]; q.tap_dh[t] = p.tap_dh[t]; q.tap_plane[t] = 0; q.tap_plane_a[t] = p.tap_plane_a[t];
    }
    q.m_pairs = Cout / 256;
    q.n_tiles = Cin / bn;
    q.k_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    const int base = q.ntaps * q.m_pairs * q.n_tiles, target = vcd_num_sms();
    int splits = base >= target ? 1 : target / base;
    if (splits > q.k_tiles) splits = q.k_tiles;
    if (splits < 1) splits = 1;
    q.k_per_split = (q.k_tiles + splits - 1) / splits;
    q.splits = (q.k_tiles + q.k_per_split - 1) / q.k_per_split;
    q.acc = acc16; q.Mout = Cout; q.Nout = Cin;
    CUtensorMap mA, mB;
    int rc;
    q.a_es = 2;
    if ((rc = make_act_map(&mA, dy_planes, Cout, 2 * W, 2 * H, 1, N, 64, p.tile_w, p.tile_h, p.tile_n, 2))) return rc;
    if ((rc = make_act_map(&mB, x, Cin, W, H, 1, N, 64, p.tile_w, p.tile_h, p.tile_n))) return rc;
    if ((rc = pair_wgrad_launch(mA, mB, q, bn, st))) return rc;
  } else {
  p.n_tiles = Cin / bn; p.m_tiles = Cout / 128;
  p.Mout = Cout; p.Nout = Cin;
  p.acc = acc16;
  p.batches = 1;
  p.k_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  p.tap_items = p.ntaps;
  const int base_tiles = p.ntaps * p.m_tiles * p.n_tiles;
  int splits = pick_splits(base_tiles, p.k_tiles);
  p.k_per_split = (p.k_tiles + splits - 1) / splits;
  p.splits = (p.k_tiles + p.k_per_split - 1) / p.k_per_split;
  set_form1_desc(p, bn);
  p.total_tiles = base_tiles * p.splits;
  CUtensorMap mA, mB;
  int rc;
  p.a_es = 2;
  if ((rc = make_act_map(&mA, dy_planes, Cout, 2 * W, 2 * H, 1, N, 64, p.tile_w, p.tile_h, p.tile_n, 2))) return rc;
  if ((rc = make_act_map(&mB, x, Cin, W, H, 1, N, 64, p.tile_w, p.tile_h, p.tile_n))) return rc;
  if ((rc = umma_launch(mA, mB, p, bn, st))) return rc;
  }
  int rc;
  upconv_wgrad_combine_kernel<<<ew_blocks(main_elems), 256, 0, st>>>(acc16, wsf, Cout, Cin);
  VCD_LAUNCH_CHECK();
  if (db && !db_colsum && (rc = conv_bias_grad(dy_planes, wsf + main_elems, (int64_t)N * 4 * H * W, Cout, st))) return rc;
  return conv_wgrad_finalize(wsf, db ? db_colsum : nullptr, dw, db, dtype, Cout, Cin, 9, st);
}

extern "C" int vcd_conv_umma_supported(int Cin, int Cout, int KH, int KW, int stride) {
  return umma_shape_ok(Cin, Cout, KH, KW, stride) ? 1 : 0;
}

// ---- which path serves a layer (AUTO): 2 = tcgen05 implicit GEMM, 3 = narrow-N implicit GEMM,
//      4 = im2col patch + GEMM, 1 = SIMT
static int fprop_path(int Cin, int Cout, int KH, int KW, int stride) {
  if (umma_shape_ok(Cin, Cout, KH, KW, stride)) return 2;
  if (narrow_out_ok(Cin, Cout, stride) && KH * KW <= 9) return 3;
  if (patch_ok(Cin, Cout, stride, KH * KW)) return 4;
  return 1;
}
static int dgrad_path(int Cin, int Cout, int KH, int KW, int stride) {
  if (umma_shape_ok(Cin, Cout, KH, KW, stride)) return 2;
  if (narrow_out_ok(Cout, Cin, stride) && KH * KW <= 9) return 3;
  if (patch_ok(Cout, Cin, stride, KH * KW)) return 4;
  return 1;
}
static int wgrad_path(int Cin, int Cout, int KH, int KW, int stride) {
  if (umma_shape_ok(Cin, Cout, KH, KW, stride)) return 2;
  if (patch_ok(Cin, Cout, stride, KH * KW) || patch_ok(Cout, Cin, stride, KH * KW)) return 4;
  return 1;
}

extern "C" int64_t vcd_conv2d_fprop_ws_bytes(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride) {
  if (fprop_path(Cin, Cout, KH, KW, stride) != 4 || small_in_conv_ok(Cin, Cout, KH, KW, stride)) return 0;
  return patch_gemm_ws((int64_t)N * H * W, Cin, Cout, KH * KW);
}
extern "C" int64_t vcd_conv2d_dgrad_ws_bytes(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride) {
  if (dgrad_path(Cin, Cout, KH, KW, stride) != 4 || small_in_conv_ok(Cout, Cin, KH, KW, stride)) return 0;
  return patch_gemm_ws((int64_t)N * H * W, Cout, Cin, KH * KW);
}

static int conv_fprop_impl(const void* x, const void* w_fprop, const float* bias, const void* residual, void* y, void* ws,
                           int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad_t, int pad_l, int Ho,
                           int Wo, int x_planes, int impl, GnEpilogue* gn, cudaStream_t st);

extern "C" int vcd_conv2d_fprop(const void* x, const void* w_fprop, const float* bias, const void* residual, void* y,
                                void* ws, int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad_t,
                                int pad_l, int Ho, int Wo, int x_planes, int impl, double* gn_sums, int gn_groups,
                                vcd_stream_t stream) {
  VCD_CHECK_ARG(x && w_fprop && y, "conv fprop: null pointer");
  cudaStream_t st = as_stream(stream);
  const bool prezeroed = (impl & VCD_ACC_PREZEROED) != 0;
  impl &= 0xff;
  if (gn_sums) {  // GroupNorm sums of y: fused into the GEMM epilogue when the pair kernel serves the layer
    GnEpilogue gn{gn_sums, gn_groups, false, false, prezeroed};
    const int rc = conv_fprop_impl(x, w_fprop, bias, residual, y, ws, N, H, W, Cin, Cout, KH, KW, stride, pad_t, pad_l, Ho,
                                   Wo, x_planes, impl, &gn, st);
    if (rc || gn.fused) return rc;
    return gn_stats_launch(y, gn_sums, nullptr, 0.f, N, Ho * Wo, Cout, gn_groups, st, prezeroed);
  }
  return conv_fprop_impl(x, w_fprop, bias, residual, y, ws, N, H, W, Cin, Cout, KH, KW, stride, pad_t, pad_l, Ho, Wo,
                         x_planes, impl, nullptr, st);
}

static int conv_fprop_impl(const void* x, const void* w_fprop, const float* bias, const void* residual, void* y, void* ws,
                           int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad_t, int pad_l, int Ho,
                           int Wo, int x_planes, int impl, GnEpilogue* gn, cudaStream_t st) {
  const int path = impl == VCD_IMPL_SIMT ? 1 : fprop_path(Cin, Cout, KH, KW, stride);
  if (impl == VCD_IMPL_UMMA)
    VCD_CHECK_ARG(path != 1, "conv fprop: shape (Cin=%d,Cout=%d,k=%d,s=%d) has no tcgen05 path", Cin, Cout, KH, stride);
  if (path == 2)
    return umma_fprop(x, w_fprop, bias, residual, y, N, H, W, Cin, Cout, KH, KW, stride, pad_t, pad_l, Ho, Wo, x_planes, st,
                      gn);
  VCD_CHECK_ARG(!x_planes, "conv fprop: parity-plane input only on the tcgen05 stride-2 path");
  if (path == 3 && !residual)
    return narrow_conv(x, w_fprop, bias, y, N, H, W, Cin, Cout, KH, KW, pad_t, pad_l, +1, st);
  if (path == 4 && !residual && small_in_conv_ok(Cin, Cout, KH, KW, stride))     // conv_in: patch only in shared memory
    return small_in_conv_launch(x, w_fprop, bias, y, N, H, W, Cin, Cout, KH, KW, pad_t, pad_l, +1, st);
  if (path == 4 && !residual)
    return patch_gemm(x, w_fprop, bias, y, ws, N, H, W, Cin, Cout, KH, KW, pad_t, pad_l, +1, st);
  return simt_conv_fprop(x, w_fprop, bias, residual, y, N, H, W, Cin, Cout, KH, KW, stride, pad_t, pad_l, Ho, Wo, st);
}

extern "C" int vcd_conv2d_dgrad(const void* dy, const void* w_fprop, const void* w_dgrad, void* dx, void* ws, int N,
                                int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad_t, int pad_l, int Ho,
                                int Wo, int dx_planes, int impl, vcd_stream_t stream) {
  (void)w_fprop;
  VCD_CHECK_ARG(dy && dx, "conv dgrad: null pointer");
  cudaStream_t st = as_stream(stream);
  impl &= 0xff;
  const int path = impl == VCD_IMPL_SIMT ? 1 : dgrad_path(Cin, Cout, KH, KW, stride);
  if (impl == VCD_IMPL_UMMA) VCD_CHECK_ARG(path != 1, "conv dgrad: shape has no tcgen05 path");
  if (path == 2)
    return umma_dgrad(dy, w_dgrad, dx, N, H, W, Cin, Cout, KH, KW, stride, pad_t, pad_l, Ho, Wo, dx_planes, st);
  VCD_CHECK_ARG(!dx_planes, "conv dgrad: parity-plane output only on the tcgen05 stride-2 path");
  VCD_CHECK_ARG(w_dgrad != nullptr, "conv dgrad needs the w_dgrad pack");
  // dx[q][ci] = sum_t sum_co dy[q - (k - pad)][co] * w_dgrad[t][ci][co]
  if (path == 3) return narrow_conv(dy, w_dgrad, nullptr, dx, N, H, W, Cout, Cin, KH, KW, pad_t, pad_l, -1, st);
  if (path == 4 && small_in_conv_ok(Cout, Cin, KH, KW, stride))                  // conv_out's data gradient
    return small_in_conv_launch(dy, w_dgrad, nullptr, dx, N, H, W, Cout, Cin, KH, KW, pad_t, pad_l, -1, st);
  if (path == 4) return patch_gemm(dy, w_dgrad, nullptr, dx, ws, N, H, W, Cout, Cin, KH, KW, pad_t, pad_l, -1, st);
  return simt_conv_dgrad(dy, w_dgrad, dx, N, H, W, Cin, Cout, KH, KW, stride, pad_t, pad_l, Ho, Wo, st);
}

// ---- dgrad with the fused GroupNorm backward prologue (see include/vcd.h)
extern "C" int vcd_conv2d_dgrad_gn_supported(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride) {
  (void)N;
  // Cin, Cout >= 256 only: with 128 output channels per item column (N = 128) or a short reduction (K = 9 * Cout =
  // 1152) the GEMM item is too short to hide the ~900-instruction fused epilogue — measured on B200: 128->128 @512^2
  // +348 us on the dgrad against 210 us for the stand-alone vcd_gn_bwd_reduce; 256->128 +700 us against 410 us
  // (tools/prof_conv2.py)
  static int min_c = -1;   // experiment knob: VCD_GNB_MIN_C=128 also fuses the 128-channel layers
  if (min_c < 0) { const char* e = getenv("VCD_GNB_MIN_C"); min_c = e ? atoi(e) : 256; }
  return (pair_enabled() && umma_shape_ok(Cin, Cout, KH, KW, stride) && KH == 3 && KW == 3 && stride == 1 && W >= 8 &&
          H >= 16 && Cin <= 512 && Cin >= min_c && Cout >= min_c && Cin % 32 == 0) ? 1 : 0;
}
extern "C" int vcd_conv2d_dgrad_gn(const void* dy, const void* w_dgrad, void* g_out, int N, int H, int W, int Cin, int Cout,
                                   int KH, int KW, int pad_t, int pad_l, const void* gn_x, const double* gn_sums,
                                   const void* gn_gamma, const void* gn_beta, int param_dtype, int gn_groups, float gn_eps,
                                   int gn_act, float* gn_dsdb, float* gn_ab_ws, vcd_stream_t stream) {
  VCD_CHECK_ARG(dy && w_dgrad && g_out && gn_x && gn_sums && gn_gamma && gn_beta && gn_dsdb && gn_ab_ws,
                "conv dgrad+GN: null pointer");
  VCD_CHECK_ARG(vcd_conv2d_dgrad_gn_supported(N, H, W, Cin, Cout, KH, KW, 1), "conv dgrad+GN: shape not supported");
  cudaStream_t st = as_stream(stream);
  int rc;
  // the kernel that writes (a, b) also zeroes gn_dsdb (same [N][Cin][2] shape): no memset between it and the GEMM
  if ((rc = gn_make_ab(gn_sums, gn_gamma, gn_beta, param_dtype, gn_eps, N, H * W, Cin, gn_groups, gn_ab_ws, gn_dsdb, st)))
    return rc;
  GnBwdPrologue gnb{gn_x, gn_ab_ws, gn_dsdb, gn_act};
  PairTap taps[9];
  for (int kh = 0; kh < KH; ++kh)
    for (int kw = 0; kw < KW; ++kw) taps[kh * KW + kw] = PairTap{0, pad_t - kh, pad_l - kw, (kh * KW + kw) * Cin};
  rc = try_pair_halo(dy, Cout, W, H, 1, N, W, H, taps, KH * KW, w_dgrad, KH * KW * Cin, nullptr, nullptr, g_out,
                     (long long)H * W * Cin, (long long)W * Cin, Cin, Cin, st, nullptr, &gnb);
  if (rc < 0) return rc;
  VCD_CHECK_ARG(rc == 1, "conv dgrad+GN: the pair kernel did not take the shape");
  return 0;
}

// workspace: fp32 [tap][Cout][Cin] + [Cout]  (+ small-channel path: bf16 patch [px][Np] + fp32 D [big][Np])
extern "C" int64_t vcd_conv2d_wgrad_ws_bytes(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride) {
  int64_t base = align256(((int64_t)KH * KW * Cout * Cin + Cout) * (int64_t)sizeof(float));
  if (wgrad_path(Cin, Cout, KH, KW, stride) == 4) {
    const int small = Cin <= 8 ? Cin : Cout, big = Cin <= 8 ? Cout : Cin;
    const int Np = round_up(KH * KW * small, 8);
    base += align256((int64_t)N * H * W * Np * 2) + align256((int64_t)big * Np * 4);
  }
  return base;
}

extern "C" int vcd_conv2d_wgrad(const void* x, const void* dy, void* dw, void* db, const float* db_colsum, int dtype,
                                void* ws, int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad_t,
                                int pad_l, int Ho, int Wo, int x_planes, int impl, vcd_stream_t stream) {
  VCD_CHECK_ARG(x && dy && dw && ws, "conv wgrad: null pointer");
  cudaStream_t st = as_stream(stream);
  const bool overlap_prev = (impl & VCD_WGRAD_OVERLAP_PREV) != 0;   // ws zeroed by vcd_conv2d_wgrad_prepare
  const bool prezeroed = overlap_prev || (impl & VCD_ACC_PREZEROED) != 0;
  impl &= 0xff;
  const int taps = KH * KW;
  const int64_t main_elems = (int64_t)taps * Cout * Cin;
  if (!prezeroed) VCD_CUDA(cudaMemsetAsync(ws, 0, (size_t)(main_elems + Cout) * sizeof(float), st));
  float* wsf = (float*)ws;
  const int path = impl == VCD_IMPL_SIMT ? 1 : wgrad_path(Cin, Cout, KH, KW, stride);
  if (impl == VCD_IMPL_UMMA) VCD_CHECK_ARG(path != 1, "conv wgrad: shape has no tcgen05 path");
  int rc;
  if (path == 2) {
    rc = umma_wgrad(x, dy, wsf, N, H, W, Cin, Cout, KH, KW, stride, pad_t, pad_l, Ho, Wo, x_planes, st, overlap_prev);
  } else if (path == 4) {
    VCD_CHECK_ARG(!x_planes, "conv wgrad: parity-plane input only on the tcgen05 stride-2 path");
    const bool small_in = Cin <= 8 && patch_ok(Cin, Cout, stride, taps);
    const int small = small_in ? Cin : Cout, big = small_in ? Cout : Cin;
    const int Np = round_up(taps * small, 8);
    const int64_t px = (int64_t)N * H * W;
    VCD_CHECK_ARG(px < (1ll << 31), "conv wgrad: too many pixels");
    char* base = (char*)ws + align256((main_elems + Cout) * (int64_t)sizeof(float));
    bf16* patch = (bf16*)base;
    float* D = (float*)(base + align256(px * Np * 2));
    // small_in : patch = im2col(x, +shift), big tensor = dy   -> D[co][t*Cin+ci]
    // small_out: patch = im2col(dy, -shift), big tensor = x   -> D[ci][t*Cout+co]
    if ((rc = im2col_small_launch((const bf16*)(small_in ? x : dy), patch, N, H, W, small, Np,
                                  patch_taps(KH, KW, pad_t, pad_l, small_in ? +1 : -1), st)))
      return rc;
    if ((rc = gemm_tn_impl(small_in ? dy : x, patch, D, 1, big, Np, px, 1, st))) return rc;
    wgrad_small_reorder_kernel<<<ew_blocks(main_elems), 256, 0, st>>>(D, wsf, taps, Cout, Cin, Np, small_in ? 1 : 0);
    VCD_LAUNCH_CHECK();
    rc = 0;
  } else {
    VCD_CHECK_ARG(!x_planes, "conv wgrad (SIMT): parity-plane input not supported");
    rc = simt_conv_wgrad(x, dy, wsf, N, H, W, Cin, Cout, KH, KW, stride, pad_t, pad_l, Ho, Wo, st);
  }
  if (rc) return rc;
  if (db && !db_colsum && (rc = conv_bias_grad(dy, wsf + main_elems, (int64_t)N * Ho * Wo, Cout, st))) return rc;
  return conv_wgrad_finalize(wsf, db ? db_colsum : nullptr, dw, db, dtype, Cout, Cin, taps, st);
}

extern "C" int vcd_conv2d_wgrad_prepare(void* ws, int Cin, int Cout, int KH, int KW, vcd_stream_t stream) {
  VCD_CHECK_ARG(ws != nullptr, "conv wgrad prepare: null workspace");
  VCD_CUDA(cudaMemsetAsync(ws, 0, (size_t)((int64_t)KH * KW * Cout * Cin + Cout) * sizeof(float), as_stream(stream)));
  return 0;
}

// D[b][m][n] = alpha * sum_k A[b][m][k] B[(b)][n][k] (+bias[n]) (+residual[b][m][n])
extern "C" int vcd_gemm_nt(const void* A, const void* B, const float* bias, const void* residual, void* D, int batch,
                           int M, int Nn, int K, int b_batched, float alpha, vcd_stream_t stream) {
  VCD_CHECK_ARG(A && B && D, "gemm_nt: null pointer");
  VCD_CHECK_ARG(K % 64 == 0 && Nn % 8 == 0, "gemm_nt: need K %% 64 == 0 and N %% 8 == 0 (K=%d N=%d)", K, Nn);
  return gemm_nt_impl(A, B, bias, residual, D, batch, M, Nn, K, b_batched, alpha, as_stream(stream));
}

// D[b][m][n] = sum_k A[b][k][m] B[b][k][n]; reduce_batch sums over b as well (Linear wgrad)
extern "C" int vcd_gemm_tn(const void* A, const void* B, void* D, int d_dtype, void* ws_f32, int batch, int M, int Nn,
                           int K, int reduce_batch, vcd_stream_t stream) {
  VCD_CHECK_ARG(A && B && D && ws_f32, "gemm_tn: null pointer");
  VCD_CHECK_ARG(M % 8 == 0 && Nn % 8 == 0, "gemm_tn: need M %% 8 == 0 and N %% 8 == 0");
  cudaStream_t st = as_stream(stream);
  int rc;
  if ((rc = gemm_tn_impl(A, B, (float*)ws_f32, batch, M, Nn, K, reduce_batch, st))) return rc;
  const int64_t out_elems = (int64_t)(reduce_batch ? 1 : batch) * M * Nn;
  int64_t blocks = ceil_div64(out_elems, 256);
  if (blocks > vcd_num_sms() * 8) blocks = vcd_num_sms() * 8;
  convert_f32_kernel<<<(unsigned)blocks, 256, 0, st>>>((const float*)ws_f32, D, d_dtype, out_elems);
  VCD_LAUNCH_CHECK();
  return 0;
}
