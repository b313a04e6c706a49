// umma_pair.cuh — host-visible description of the CTA-pair (cta_group::2) implicit-GEMM kernel (umma_pair.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

// D[pixel][n] = alpha * sum_{tap} sum_{c} A[pixel + shift(tap)][c] * B[tap_brow + n][c]  (+bias, +residual)
//
// Two CTAs of a cluster (an SM pair) execute ONE tcgen05.mma.cta_group::2 stream: M = 256 (128 pixels per CTA),
// N = BLOCK_N; each CTA stages its own 128-pixel A operand and HALF of the B operand (BLOCK_N/2 weight rows), so
// the weight traffic from L2 and the shared-memory reads per FLOP are half those of a single-CTA tile.
//
// HALO mode (3x3 / 2x2-phase / stride-2 convolutions): a CTA's pixel tile is 8 (w) x 16 (h) pixels of one image.
// For every 64-channel K chunk ONE TMA box of (8+halo_w) x (16+halo_h) pixels is loaded and ALL taps that read the
// same tensor plane are served from it: the A descriptor of a tap starts (dh*box_w + dw) 128-byte rows into the
// box and strides box_w*128 bytes between its 8-row groups (the 128B swizzle is a function of the absolute
// shared-memory address, so a row-shifted start is legal — tools/probe_umma.py).  The activation is read from L2
// ~1.4x instead of 9x per 3x3 convolution.
// ROWS mode (1x1 convolutions, Linear, Q.K^T, P.V): the pixel space is a flat list of rows, tile = 128 rows, one tap.
struct PairParams {
  int mode;              // 0 = HALO, 1 = ROWS
  int W, H, Nimg;        // pixel space; ROWS: W = rows per batch entry, H = 1, Nimg = batch entries
  int tiles_w, tiles_h;  // tiles per image
  int pix_tiles;         // tiles_w * tiles_h * Nimg
  int pair_in_image;     // 1: both tiles of a pair lie in the same image (per-image B operand)
  int pairs;             // pair-tiles
  int n_tiles;           // ceil(Nout / BLOCK_N)
  int total_items;       // pairs * n_tiles
  int a_es;              // TMA element stride of the A map (1 | 2, see make_act_map)
  int box_w;             // A box width in pixels (HALO: 8 + halo_w, ROWS: 128)
  uint32_t a_box_bytes;  // bytes of one A box
  int ngroups;           // A boxes per K chunk (one per tensor plane the taps touch)
  int g_plane[4], g_dh[4], g_dw[4];  // plane and origin offset of each group's box relative to the tile origin
  int g_tap0[5];         // taps [g_tap0[g], g_tap0[g+1]) belong to group g
  int ntaps;
  int tap_aoff[16];      // byte offset of the tap's first row inside its group's box
  int tap_brow[16];      // B row offset of the tap
  int kc;                // 64-channel K chunks per tap
  int b_batch_rows;      // B row offset per image (batched GEMM); 0 = shared weights
  int b_resident;        // 1: the whole B operand of this CTA (ntaps * kc boxes, <= 144 KB) is loaded ONCE per kernel and
                         //    stays in shared memory; only the A halo boxes stream (128 -> 128 channel 3x3 convs)
  bf16* out;
  const bf16* residual;
  const float* bias;
  float alpha;
  long long out_sn, out_sh, out_sw;  // element strides of (n, h, w) in out / residual
  int Nout;
  uint32_t idesc;
  // fused GroupNorm sums of the OUTPUT tensor (the next layer's GroupNorm): sum and sum of squares of the bf16-rounded
  // outputs per (image, group) added to gn_sums[n][G][2]; every tile must lie inside one image
  double* gn_sums;
  int gn_G, gn_logD;       // groups, log2(channels per group) in {2, 3, 4}
  int gn_rows_per_img;     // ROWS mode: rows of one image (a multiple of 128)
  // fused GroupNorm backward prologue (dgrad of the conv that consumed act(GroupNorm(x))): the epilogue multiplies the
  // data gradient by SiLU'(a x + b), stores that pre-activation gradient g, and accumulates sum_p g*x and sum_p g per
  // (image, channel) into gnb_dsdb[n][c][2] (vcd_gn_bwd_reduce's result) — the GroupNorm backward then needs one pass
  const bf16* gnb_x;       // GroupNorm input, same layout / strides as out
  const float* gnb_ab;     // [N][C][2]: a = gamma * rstd, b = beta - mean * a
  float* gnb_dsdb;         // [N][C][2] fp32, zeroed by the launcher
  int gnb_act;             // 1: SiLU follows the GroupNorm
  int side_prefetch;   // 1 (default): L2-prefetch the next item's epilogue side input (VCD_PAIR_PREFETCH=0 disables)
  int release_arrive;  // A/B knob (VCD_PAIR_RELEASE=1): epilogue hands accumulator stages back with a RELEASING arrive
  int dbg;  // profiling aid (VCD_PAIR_DBG): 1 = epilogue only hand-shakes, 2 = MMA warp issues no MMAs, 4 = no stores
};

// request / result of the fused GroupNorm sums (host side)
struct GnEpilogue {
  double* sums;  // [N][groups][2], zeroed by the launcher when it fuses
  int groups;
  bool fused;    // set by the launcher: the kernel produced the sums
  bool accumulate = false;  // add to sums already started by an earlier launch (phase convolutions): no zeroing
  bool prezeroed = false;   // the caller guarantees sums is zero (VCD_ACC_PREZEROED): no zeroing
};

// fused GroupNorm backward prologue of a dgrad launch (host side)
struct GnBwdPrologue {
  const void* x;
  const float* ab;
  float* dsdb;   // zeroed by the kernel that writes ab (gn_make_ab)
  int act;
};

struct PairTap {
  int plane, dh, dw, brow;
};

// ---- CTA-pair weight-gradient kernel: Dacc[tap][co][ci] += sum_pixels dy[pixel][co] * x[pixel + shift(tap)][ci]
// M = 256 output channels (128 per CTA), N = BLOCK_N input channels (each CTA stages half of them), K = pixels in
// steps of 64 (both operands MN-major 128B-swizzle boxes of 64 channels x 64 pixels), split over clusters (split-K),
// fp32 red.global.add epilogue.  Per MMA each CTA reads 4 KB of A and 4 KB of B from shared memory (N = 256) instead of
// 4 + 8 KB in the single-CTA kernel, which lifts the shared-memory operand bandwidth cap from 67 % to 100 % of the
// tensor rate.
struct PairWgradParams {
  int W, H, Nimg;
  int tile_w, tile_h, tile_n;     // the 64-pixel K-step box
  int tiles_w, tiles_h, tiles_n;
  int ntaps;
  int tap_dw[16], tap_dh[16], tap_plane[16], tap_plane_a[16];
  int a_es, b_es;                 // TMA element strides of the A (dy) and B (x) maps
  int m_pairs;                    // Cout / 256
  int n_tiles;                    // Cin / BLOCK_N
  int splits, k_per_split, k_tiles;
  float* acc;                     // [tap][Mout][Nout] fp32
  int Mout, Nout;
  uint32_t idesc;
  int total_items;
};
// overlap_prev: programmatic dependent launch behind the preceding kernel of the stream (common.cuh vcd_launch)
int pair_wgrad_launch(const CUtensorMap& mapA, const CUtensorMap& mapB, PairWgradParams& p, int block_n, cudaStream_t st,
                      bool overlap_prev = false);

// fills mode / tiling / groups / taps / descriptors of a HALO launch; returns false when the shape is not eligible
bool pair_setup_halo(PairParams& p, int W, int H, int N, const PairTap* taps, int ntaps, int block_n, int* box_h_out);
void pair_setup_rows(PairParams& p, int rows, int batch, int pair_in_image, int brow, int block_n);
int pair_launch(const CUtensorMap& mapA, const CUtensorMap& mapB, PairParams& p, int block_n, cudaStream_t st);
bool pair_enabled();
