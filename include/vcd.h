/* vcd.h — C ABI of libvcd_b200.so (sm_100a only).
 *
 * Drop-in boundary for the hot path of olegroshka/vae-channel-dynamics: the SDXL
 * AutoencoderKL training step + per-channel activation tracker + dead-channel
 * classifier + GroupNorm-gamma nudge.  The reference has no FFI of its own: its
 * interface is four Python modules.  Each entry point below
 * names the reference call site (file:line under /root/reference) whose arithmetic
 * it replaces; the Python host layer (vae-channel-dynamics_b200/src/...) keeps the
 * reference's class/method names and reaches these symbols through ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host
 *   - activations are bf16, NHWC contiguous ([N][H][W][C]); statistics, GroupNorm
 *     sums, loss accumulators and weight-gradient accumulators are fp32/fp64
 *   - dtype codes: 0 = fp32, 1 = bf16
 *   - no entry point allocates, synchronises the device, or touches a stream other
 *     than `stream`; workspaces are caller-provided
 *   - return 0 on success; otherwise a negative code and vcd_last_error() holds text
 */
#ifndef VCD_H_
#define VCD_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* vcd_stream_t; /* cudaStream_t */

#define VCD_F32 0
#define VCD_BF16 1

/* conv implementation selector.  AUTO: tcgen05 implicit GEMM for 128-multiple channels; small-channel layers
 * (3->128, 128->3, 4->512, 512->8) also run on the tcgen05 kernel, as a narrow-N implicit GEMM or an im2col
 * patch + GEMM with a caller-provided workspace (vcd_conv2d_*_ws_bytes); SIMT only for the 1x1 quant convs */
#define VCD_IMPL_AUTO 0
#define VCD_IMPL_SIMT 1 /* CUDA-core direct convolution (small-channel layers, cross-check) */
#define VCD_IMPL_UMMA 2 /* tcgen05/TMEM/TMA implicit GEMM; error if the shape is unsupported */
/* flag OR-ed into the `impl` argument of vcd_conv2d_wgrad: the caller has already zeroed the workspace with
 * vcd_conv2d_wgrad_prepare() and enqueued, after it and directly before this call, a kernel that neither writes x / dy
 * nor reads dw / db / ws (the data-gradient GEMM of the same layer): the weight-gradient GEMM is then launched as a
 * programmatic dependent launch and its CTAs fill the SMs which that kernel's last wave leaves idle. */
#define VCD_WGRAD_OVERLAP_PREV 0x100
/* flag: the accumulator buffers this call would zero before its kernels add into them are ALREADY zero (the caller carved
 * them out of a zero-filled arena), so the call enqueues no memset — a training step otherwise issues ~245 of them, each a
 * 2 us device activity plus one more launch boundary.  OR-ed into `impl` of vcd_conv2d_fprop (gn_sums) and vcd_conv2d_wgrad
 * (ws), and into `act_silu` of vcd_gn_bwd_reduce (dsdb) and vcd_gn_bwd_apply (dx_colsum). */
#define VCD_ACC_PREZEROED 0x200

const char* vcd_last_error(void);
int vcd_version(void);
/* 1 if (Cin, Cout) can run on the tcgen05 implicit-GEMM path */
int vcd_conv_umma_supported(int Cin, int Cout, int KH, int KW, int stride);

/* ---- weights -------------------------------------------------------------------
 * [upstream diffusers Conv2d/Linear parameters, OIHW] -> GEMM operand packs, redone
 * after every optimizer step (train.py:302).
 *   w_fprop : bf16 [KH*KW][Cout][Cin]   (K-major B operand of fprop)
 *   w_dgrad : bf16 [KH*KW][Cin][Cout]   (K-major B operand of dgrad)   (may be NULL)
 *   bias_f32: fp32 [Cout]                                              (may be NULL) */
int vcd_pack_conv_weight(const void* w, const void* bias, int dtype, int Cout, int Cin, int KH, int KW,
                         void* w_fprop, void* w_dgrad, float* bias_f32, vcd_stream_t stream);

/* All operand packs of a model in ONE launch (the per-layer entry point above costs ~140 launches per forward).
 * descs (device array): one vcd_pack_desc per layer — mode 0: w OIHW [cout][cin][taps] -> wf [taps][cout][cin],
 * wd [taps][cin][cout] (wd may be NULL); mode 1 (Upsample2D conv, taps = 9): the 16 pre-summed phase taps of
 * vcd_pack_upconv_weight; bias -> bias_f32 when both are given.  Work list (device int32 arrays of length n_tiles): tile t
 * covers output channels [tile_co[t], +vcd_pack_tile_co()) x input channels [tile_ci[t], +vcd_pack_tile_ci()) of layer
 * tile_layer[t]. */
typedef struct {
  const void* w;
  const void* bias;
  void* wf;
  void* wd;
  float* bias_f32;
  int32_t dtype, cout, cin, taps, mode, pad_;
} vcd_pack_desc;
int vcd_pack_tile_co(void);
int vcd_pack_tile_ci(void);
int vcd_multi_pack_weights(const vcd_pack_desc* descs, const int32_t* tile_layer, const int32_t* tile_co,
                           const int32_t* tile_ci, int n_tiles, vcd_stream_t stream);

/* ---- convolution (AutoencoderKL.encode/decode conv layers; sdxl_vae_wrapper.py:60,71;
 *      backward reached from train.py:299) ------------------------------------------
 * y[n,ho,wo,co] = bias[co] + sum x[n, ho*stride - pad_t + kh, wo*stride - pad_l + kw, ci] * w[co,ci,kh,kw]
 *                 (+ residual[n,ho,wo,co])
 * Downsample2D's asymmetric (0,1,0,1) pad is pad_t = pad_l = 0 with Ho = H/2 (zero fill
 * past the bottom/right edge).  Stride-2 layers read the NHWC tensor in place (x_planes = 0: the tcgen05 path uses
 * element-strided TMA maps); x_planes != 0 means x is the parity-plane layout [N][2][2][H/2][W/2][C] written by
 * vcd_space_to_planes (kept for callers that already hold that layout). */
int64_t vcd_conv2d_fprop_ws_bytes(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride);
/* gn_sums (fp64 [N][gn_groups][2], may be NULL): on return holds sum and sum of squares of y per (image, group) —
 * the statistics the GroupNorm that consumes y needs (vcd_gn_apply_fwd), produced by the GEMM epilogue when the
 * tcgen05 pair kernel serves the layer (no extra pass over y), otherwise by vcd_gn_stats inside the call. */
int vcd_conv2d_fprop(const void* x, const void* w_fprop, const float* bias, const void* residual, void* y, void* ws,
                     int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad_t, int pad_l,
                     int Ho, int Wo, int x_planes, int impl, double* gn_sums, int gn_groups, vcd_stream_t stream);
/* dx = conv_transpose(dy).  dx_planes != 0 (stride 2 only): dx is written in parity-plane layout. */
int64_t vcd_conv2d_dgrad_ws_bytes(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride);
int vcd_conv2d_dgrad(const void* dy, const void* w_fprop, const void* w_dgrad, void* dx, void* ws,
                     int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad_t, int pad_l,
                     int Ho, int Wo, int dx_planes, int impl, vcd_stream_t stream);
/* dgrad of a 3x3 stride-1 conv whose INPUT was act(GroupNorm(gn_x)) (ResnetBlock2D norm1->conv1, norm2->conv2), fused
 * with the first half of that GroupNorm's backward: instead of dL/d(input) the call stores
 *     g = dL/d(input) * SiLU'(a*gn_x + b)        (gn_act = 0: g = dL/d(input))
 * in g_out, and accumulates gn_dsdb[n][c] = (sum_p g*gn_x, sum_p g) — exactly what vcd_gn_bwd_reduce computes — in the
 * GEMM epilogue.  The GroupNorm backward then is vcd_gn_bwd_apply(x, g, ..., act_silu = 0) + vcd_gn_param_grad: one
 * pass over the tensors instead of two.  gn_ab_ws: fp32 [N][Cin][2] scratch.  Only shapes for which
 * vcd_conv2d_dgrad_gn_supported() returns 1. */
int vcd_conv2d_dgrad_gn_supported(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride);
int vcd_conv2d_dgrad_gn(const void* dy, const void* w_dgrad, void* g_out, int N, int H, int W, int Cin, int Cout,
                        int KH, int KW, int pad_t, int pad_l, const void* gn_x, const double* gn_sums,
                        const void* gn_gamma, const void* gn_beta, int param_dtype, int gn_groups, float gn_eps,
                        int gn_act, float* gn_dsdb, float* gn_ab_ws, vcd_stream_t stream);
/* dw (OIHW, `dtype`) and db ([Cout], `dtype`, may be NULL).  ws: workspace of vcd_conv2d_wgrad_ws_bytes()
 * bytes (zeroed by the call unless VCD_WGRAD_OVERLAP_PREV or VCD_ACC_PREZEROED is set).  db_colsum (fp32 [Cout], may be NULL): column sums of dy already produced by
 * the kernel that wrote dy (vcd_gn_bwd_apply), which saves the bias-gradient pass over dy. */
int64_t vcd_conv2d_wgrad_ws_bytes(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride);
int vcd_conv2d_wgrad(const void* x, const void* dy, void* dw, void* db, const float* db_colsum, int dtype, void* ws,
                     int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad_t, int pad_l,
                     int Ho, int Wo, int x_planes, int impl, vcd_stream_t stream);
/* zeroes the accumulators inside ws ahead of a vcd_conv2d_wgrad(..., impl | VCD_WGRAD_OVERLAP_PREV, ...) call */
int vcd_conv2d_wgrad_prepare(void* ws, int Cin, int Cout, int KH, int KW, vcd_stream_t stream);

/* ---- Upsample2D fused: nearest x2 + conv3x3(pad 1) as four 2x2 phase convolutions on the low-resolution
 * tensor with pre-summed weights ([upstream] Upsample2D in decoder.up_blocks.{0,1,2}.upsamplers.0).
 * 16 tap products per low-res pixel instead of 36; the upsampled tensor is never materialised.
 *   wf16: bf16 [16][Cout][Cin], wd16: bf16 [16][Cin][Cout]  (index ((a*2+b)*2+dh)*2+dw)
 *   x [N][H][W][Cin] -> y [N][2H][2W][Cout];  the backward entry points take dy [N][2H][2W][Cout] as stored: its four
 *   parity planes are read in place through element-strided TMA maps (no space-to-planes copy) */
int vcd_pack_upconv_weight(const void* w, const void* bias, int dtype, int Cout, int Cin, void* wf16, void* wd16,
                           float* bias_f32, vcd_stream_t stream);
int vcd_upconv2d_fprop(const void* x, const void* wf16, const float* bias, void* y, int N, int H, int W, int Cin,
                       int Cout, double* gn_sums, int gn_groups, vcd_stream_t stream);
int vcd_upconv2d_dgrad(const void* dy_planes, const void* wd16, void* dx, int N, int H, int W, int Cin, int Cout,
                       vcd_stream_t stream);
int64_t vcd_upconv2d_wgrad_ws_bytes(int Cin, int Cout);
int vcd_upconv2d_wgrad(const void* x, const void* dy_planes, void* dw, void* db, const float* db_colsum, int dtype,
                       void* ws, int N, int H, int W, int Cin, int Cout, vcd_stream_t stream);

/* Number of launches of the CTA-pair (tcgen05.mma.cta_group::2, halo-reuse) kernel since the library was loaded:
 * lets tests and the bench assert that the pair path, not the single-CTA kernel, served a layer. */
int64_t vcd_pair_kernel_launches(void);
/* Route the GEMM-path layers through the CTA-pair kernels (default, 1) or the single-CTA tcgen05 kernel (0): two
 * independent device implementations of the same arithmetic, compared against each other at full layer sizes by
 * tests/test_fullsize_gpu.py.  Returns the previous setting.  Env VCD_PAIR=0 sets the initial value. */
int vcd_set_pair_kernels(int enabled);

/* NHWC [N][H][W][C] <-> parity planes [N][2][2][H/2][W/2][C] (stride-2 convs), H and W even */
int vcd_space_to_planes(const void* x, void* xp, int N, int H, int W, int C, vcd_stream_t stream);
int vcd_planes_to_space(const void* xp, void* x, int N, int H, int W, int C, vcd_stream_t stream);
/* Upsample2D nearest x2 ([upstream] F.interpolate) and its adjoint (2x2 sum) */
int vcd_upsample2x_fwd(const void* x, void* y, int N, int H, int W, int C, vcd_stream_t stream);
int vcd_upsample2x_bwd(const void* dy, void* dx, int N, int H, int W, int C, vcd_stream_t stream);
/* loader tensor (data_utils.py:28-29: fp32 NCHW in [-1,1]) -> bf16 NHWC, and back */
int vcd_nchw_to_nhwc(const void* x, int x_dtype, void* y_bf16, int N, int C, int H, int W, vcd_stream_t stream);
int vcd_nhwc_to_nchw(const void* x_bf16, void* y, int y_dtype, int N, int C, int H, int W, vcd_stream_t stream);
int vcd_add(const void* a, const void* b, void* out, int64_t n, vcd_stream_t stream); /* bf16 */

/* ---- GroupNorm(32, C, eps) [+ SiLU] with fused per-channel statistics -------------
 * ([upstream] torch.nn.GroupNorm + F.silu inside ResnetBlock2D / Attention / conv_norm_out;
 *  statistics: src/tracking/monitor.py:64-75)
 * Pass 1  vcd_gn_stats      : sums[n][g] = {sum x, sum x^2} (fp64), optional per-channel
 *                             statistics of the GroupNorm INPUT  (capture_point "input")
 * Pass 2  vcd_gn_apply_fwd  : y = gamma*(x-mean)*rstd+beta ; optional statistics of y
 *                             (capture_point "output", before SiLU) ; out = silu(y) if act ;
 *                             optional statistics of the INPUT x in the same pass (chan_stats_in:
 *                             capture_point "input" of this GroupNorm / "output" of the layer that
 *                             produced x, e.g. vae.encoder.conv_in) when the group sums came from the
 *                             producing GEMM's epilogue and pass 1 does not run
 * chan stats layout (fp32 [5][C]): sum x, sum x^2, sum |x|, max |x|, count(|x| < near_zero)
 * HW = H*W; x is [N][HW][C]. */
int vcd_gn_stats(const void* x, double* sums, float* chan_stats_in, float near_zero,
                 int N, int HW, int C, int G, vcd_stream_t stream);
int vcd_gn_apply_fwd(const void* x, const double* sums, const void* gamma, const void* beta, int param_dtype,
                     void* out, float* chan_stats_in, float* chan_stats_out, float near_zero, float eps, int act_silu,
                     int N, int HW, int C, int G, vcd_stream_t stream);
/* backward: dsdb[n][c] = {sum g*x, sum g} with g = dout * silu'(y) ; then dx; then dgamma/dbeta.
 * act_silu: 1 = SiLU follows the GroupNorm; VCD_ACC_PREZEROED may be OR-ed in (dsdb / dx_colsum already zero). */
int vcd_gn_bwd_reduce(const void* x, const void* dout, const double* sums, const void* gamma, const void* beta,
                      int param_dtype, float* dsdb, float eps, int act_silu,
                      int N, int HW, int C, int G, vcd_stream_t stream);
/* dx = GroupNorm/SiLU backward (+ dres, the gradient arriving over the block's skip connection, may be NULL);
 * dx_colsum (fp32 [C], may be NULL) receives the per-channel sums of dx = the bias gradient of the conv that
 * produced x */
/* dgamma / dbeta ([C], parameter dtype, may both be NULL): when given, the kernel's first block also writes the
 * parameter gradients (the arithmetic of vcd_gn_param_grad), saving that launch. */
int vcd_gn_bwd_apply(const void* x, const void* dout, const double* sums, const void* gamma, const void* beta,
                     int param_dtype, const float* dsdb, void* dx, const void* dres, float* dx_colsum,
                     void* dgamma, void* dbeta, float eps, int act_silu, int N, int HW, int C, int G,
                     vcd_stream_t stream);
int vcd_gn_param_grad(const double* sums, const float* dsdb, void* dgamma, void* dbeta, int param_dtype,
                      float eps, int N, int HW, int C, int G, vcd_stream_t stream);
/* stand-alone SiLU (only used when a foreign forward hook needs the pre-activation tensor) */
int vcd_silu_fwd(const void* x, void* y, int64_t n, vcd_stream_t stream);
int vcd_silu_bwd(const void* x, const void* dy, void* dx, int64_t n, vcd_stream_t stream);

/* ---- attention (mid_block.attentions.0; [upstream] AttnProcessor2_0, 1 head, d = C) ---
 * GEMMs go through vcd_gemm_*; these are the row-softmax passes over S[N][T][T] (bf16). */
int vcd_softmax_fwd(const void* s, void* p, int64_t rows, int cols, vcd_stream_t stream);
int vcd_softmax_bwd(const void* p, const void* dp, void* ds, float scale, int64_t rows, int cols,
                    vcd_stream_t stream);
int vcd_transpose_bf16(const void* x, void* y, int batch, int rows, int cols, vcd_stream_t stream);
/* D[b][m][n] = alpha * sum_k A[b][m][k] * B[b][n][k] (+ bias[n]) (+ residual[b][m][n]); bf16 in/out,
 * K-major operands (tcgen05); b_batch_stride_rows = 0 shares B across the batch (Linear). */
int vcd_gemm_nt(const void* A, const void* B, const float* bias, const void* residual, void* D,
                int batch, int M, int Nn, int K, int b_batched, float alpha, vcd_stream_t stream);
/* D[b][m][n] = sum_k A[b][k][m] * B[b][k][n]  (both operands reduction-major; fp32 result via ws) */
int vcd_gemm_tn(const void* A, const void* B, void* D, int d_dtype, void* ws_f32,
                int batch, int M, int Nn, int K, int reduce_batch, vcd_stream_t stream);

/* ---- latent distribution + losses -------------------------------------------------
 * [upstream] DiagonalGaussianDistribution (sdxl_vae_wrapper.py:60-66) and train.py:289-291.
 * moments: bf16 NHWC [N][hw][2*L]; eps_noise: fp32 NCHW-order [N][L][hw] (torch.randn order) or NULL (mode()).
 * z: bf16 NHWC [N][hw][L]; kl_per_sample: fp32 [N] (zeroed by the call). */
int vcd_gauss_sample_kl_fwd(const void* moments, const float* eps_noise, void* z, float* mean_out,
                            float* logvar_out, float* kl_per_sample, int N, int hw, int L, vcd_stream_t stream);
/* dmoments from dz (bf16 NHWC, may be NULL) and dkl (fp32 [N], may be NULL) */
int vcd_gauss_sample_kl_bwd(const void* moments, const float* eps_noise, const void* dz, const float* dkl,
                            void* dmoments, int N, int hw, int L, vcd_stream_t stream);
/* mse: loss_sum (fp64, zeroed by the call) = sum (rec - x)^2 ; drec = 2*(rec-x)*grad_scale (bf16 NHWC).
 * rec: bf16 NHWC; x: fp32 NCHW (the loader tensor, train.py:289). */
int vcd_mse_fwd_bwd(const void* rec, const float* x_nchw, double* loss_sum, void* drec, float grad_scale,
                    int N, int C, int H, int W, vcd_stream_t stream);

/* ---- tracker / classifier / nudger / dead-weight scan -------------------------------
 * vcd_chan_stats        : stand-alone per-channel statistics of any [N][HW][C] tensor
 *                         (monitor.py:64-67 for non-GroupNorm targets such as encoder.conv_in)
 * vcd_stats_finalize    : per-forward normalisation + running accumulation (monitor.py:101,
 *                         176-186): run[0][c] += sumabs/Nf, run[1][c] += mean_c, run[2][c] += var_c,
 *                         run[3][c] = max(.., maxabs), run[4][c] += nearzero/Nf;
 *                         scal[0] += mean_activation, scal[1] += std_activation (unbiased), scal[2] += 1
 *                         (forward count); then chan_stats is zeroed for the next forward
 * vcd_classify_mask     : mask[c] = mean_abs[c] < thr (strict, fp32; classifier.py:135), count
 * vcd_nudge_gamma       : gamma[idx] = min(gamma*factor, cap) in fp64, RN to dtype (nudger.py:127-143);
 *                         mode 1 = reset to 1.0 (nudger.py:162-168); out-of-range indices skipped;
 *                         applied_count (int32, device) receives the number applied
 * vcd_dead_weight_count : for T tensors, count |w| < thr and/or |w| < pct*mean|w| (deadneuron.py:78-115) */
int vcd_chan_stats(const void* x, int x_dtype, float* chan_stats, float near_zero, int N, int HW, int C,
                   int channels_last, vcd_stream_t stream);
int vcd_stats_finalize(float* chan_stats, float* run, double* scal, int64_t n_per_channel, int C,
                       vcd_stream_t stream);
int vcd_classify_mask(const float* mean_abs, float threshold, uint8_t* mask, int32_t* count, int C,
                      vcd_stream_t stream);
int vcd_nudge_gamma(void* gamma, int dtype, int C, const int64_t* idx, int n_idx, double factor, double cap,
                    int mode, int32_t* applied_count, vcd_stream_t stream);
int vcd_dead_weight_count(const void* const* tensors, const int64_t* numels, const int32_t* dtypes, int T,
                          double threshold, double mean_percentage, int dead_type, double* sum_abs_ws,
                          int64_t* counts, vcd_stream_t stream);

/* ---- fused multi-tensor gradient norm + clip + AdamW (train.py:184-187,301-302) ----------------
 * Tables (device arrays of length T): params / grads (fp32 or bf16 per dtypes[t]; grads[t] may be NULL = parameter
 * skipped this step), exp_avg / exp_avg_sq (fp32), numels.  Work is cut into chunks of vcd_optim_chunk_elems()
 * elements: chunk c covers elements [chunk_off[c], chunk_off[c] + chunk) of tensor chunk_tensor[c] (device arrays of
 * length n_chunks, built once by the caller).
 * vcd_multi_sqnorm    : *out_sqnorm (fp64, zeroed by the call) = sum over all tensors of sum g^2
 * vcd_clip_adamw_step : g *= min(1, max_norm / (sqrt(*grad_sqnorm) + 1e-6)) (skipped when grad_sqnorm is NULL or
 *                       max_norm <= 0; torch.nn.utils.clip_grad_norm_), then torch.optim.AdamW's update with decoupled
 *                       weight decay and bias correction for step number `step` (1-based), or — torch keeps one step
 *                       counter per parameter — for steps[t] when `steps` (device int32 [T]) is given. */
int vcd_optim_chunk_elems(void);
int vcd_multi_sqnorm(const void* const* grads, const int64_t* numels, const int32_t* dtypes,
                     const int32_t* chunk_tensor, const int64_t* chunk_off, int n_chunks, double* out_sqnorm,
                     vcd_stream_t stream);
int vcd_clip_adamw_step(void* const* params, const void* const* grads, float* const* exp_avg,
                        float* const* exp_avg_sq, const int64_t* numels, const int32_t* dtypes,
                        const int32_t* chunk_tensor, const int64_t* chunk_off, int n_chunks,
                        const double* grad_sqnorm, double max_norm, double lr, double beta1, double beta2,
                        double eps, double weight_decay, const int32_t* steps, int64_t step, vcd_stream_t stream);

/* ---- evaluation metrics and input preprocessing (evaluate.py:163-176,238-249; data_utils.py:13-30)
 * vcd_ssim_psnr_update : pred / target fp32 NCHW in [0, data_range].  ACCUMULATES (caller zeroes once):
 *                        *ssim_sum += sum over the N images of the per-image mean SSIM ([upstream] torchmetrics
 *                        StructuralSimilarityIndexMeasure: Gaussian window kernel_size x kernel_size, sigma, valid
 *                        positions, c1 = (0.01 R)^2, c2 = (0.03 R)^2); *sse += sum (pred - target)^2
 *                        (PeakSignalNoiseRatio: 10 log10(R^2 / (sse / elements))).
 * vcd_preprocess_u8    : [N][H][W][3] uint8 -> fp32 [N][3][R][R] in [-1, 1]: Resize(shorter side -> R, bilinear) ->
 *                        CenterCrop(R) -> ToTensor -> Normalize(0.5, 0.5). */
int vcd_ssim_psnr_update(const float* pred, const float* target, int N, int C, int H, int W, float data_range,
                         int kernel_size, float sigma, double* ssim_sum, double* sse, vcd_stream_t stream);
int vcd_preprocess_u8(const uint8_t* images, float* out, int N, int H, int W, int R, vcd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VCD_H_ */
