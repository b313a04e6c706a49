// conv_dispatch.h — internal (non-ABI) entry points shared between conv_simt.cu and conv_dispatch.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

int simt_conv_fprop(const void* x, const void* wf, const float* bias, const void* residual, void* y, int N, int H, int W,
                    int Cin, int Cout, int KH, int KW, int stride, int pad_t, int pad_l, int Ho, int Wo, cudaStream_t st);
int simt_conv_dgrad(const void* dy, const void* wd, void* dx, int N, int H, int W, int Cin, int Cout, int KH, int KW,
                    int stride, int pad_t, int pad_l, int Ho, int Wo, cudaStream_t st);
int simt_conv_wgrad(const void* x, const void* dy, float* ws, int N, int H, int W, int Cin, int Cout, int KH, int KW,
                    int stride, int pad_t, int pad_l, int Ho, int Wo, cudaStream_t st);
int conv_bias_grad(const void* dy, float* out, int64_t pixels, int C, cudaStream_t st);
int conv_wgrad_finalize(const float* ws, const float* bias_src, void* dw, void* db, int dtype, int Cout, int Cin, int taps,
                        cudaStream_t st);
// ab[n][c] = (a, b) of the affine GroupNorm y = a x + b; zero_dsdb (may be NULL): a [N][C][2] fp32 buffer zeroed in passing
int gn_make_ab(const double* sums, const void* gamma, const void* beta, int pdt, float eps, int N, int HW, int C, int G,
               float* ab, float* zero_dsdb, cudaStream_t st);
// vcd_gn_stats with the zeroing of `sums` optional (prezeroed: the caller guarantees it is zero)
int gn_stats_launch(const void* x, double* sums, float* chan_stats_in, float near_zero, int N, int HW, int C, int G,
                    cudaStream_t st, bool prezeroed);
// conv_small.cu: S <= 8 source channels -> 128 output channels with the im2col patch held in shared memory only
bool small_in_conv_ok(int S, int O, int KH, int KW, int stride);
int small_in_conv_launch(const void* src, const void* pack, const float* bias, void* out, int N, int H, int W, int S, int O,
                         int KH, int KW, int pad_t, int pad_l, int sign, cudaStream_t st);
