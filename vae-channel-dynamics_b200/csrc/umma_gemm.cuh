// umma_gemm.cuh — host-visible description of the tcgen05 implicit-GEMM kernel (umma_gemm.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

// One persistent warp-specialised kernel serves two operand arrangements.
//
// FORM 0 ("pixel rows", K-major operands) — conv fprop / dgrad, 1x1 shortcuts, Linear, Q.K^T, P.V^T:
//     D[pixel][n] = alpha * sum_{tap} sum_{c} A[pixel + shift(tap)][c] * B[tap*b_tap_rows + n][c]  (+bias, +residual)
//   A: 5-D TMA map (C, W, H, plane, N) over a bf16 NHWC activation; one box = 64 channels x 128 pixels,
//      landing in shared memory as the canonical 128B-swizzled K-major UMMA tile (128 rows x 128 B).
//      Conv padding = TMA out-of-bounds zero fill; a tap is a coordinate shift.
//   B: 5-D map (K, rows, 1, 1, 1) over packed weights [tap][Nout][K]; one box = 64 x BLOCK_N rows.
// FORM 1 ("pixel reduction", MN-major operands) — conv/Linear wgrad, dV = P^T dO, dK = dS^T Q:
//     Dacc[batch][tap][m][n] += sum_{pixel} A[pixel][m] * B[pixel + shift(tap)][n]      (fp32, red.global.add)
//   A, B: 5-D activation maps; one box = 64 channels x 64 pixels = an MN-major 128B-swizzle column block.
//   Split over pixel tiles across CTAs (split-K).
struct UmmaParams {
  int form;
  // pixel space and its tiling (form 0: M tiles of 128 pixels; form 1: K steps of 64 pixels)
  int W, H, Nimg;
  int tile_w, tile_h, tile_n;
  int tiles_w, tiles_h, tiles_n;
  // taps
  int ntaps;
  int tap_dw[16], tap_dh[16], tap_plane[16];
  int tap_plane_a[16];  // form 1: plane of the A (dy) operand per tap (upsample-conv wgrad), else 0
  // form 0
  int n_tiles;       // Nout / BLOCK_N
  int kc_per_tap;    // K / 64
  int tap_brow[16];  // B row offset per tap
  int b_batch_rows;  // B row offset per image (batched GEMM); 0 = shared weights
  bf16* out;
  const bf16* residual;
  const float* bias;
  float alpha;
  long long out_sn, out_sh, out_sw;  // element strides of (n, h, w) in out / residual
  int Nout;
  // form 1
  float* acc;   // [batch][tap][Mout][Nout] fp32
  int Mout;
  int m_tiles;  // Mout / 128
  int splits, k_per_split, k_tiles;  // pixel tiles per batch entry, split over CTAs
  int batches;  // 1 for conv wgrad (all images reduced), Nimg for attention (per-image result)
  // form 1, Nout = 128 only: one BLOCK_N = 256 tile covers TWO taps x 128 input channels (columns 0-127: tap 2g,
  // 128-255: tap 2g+1), so each MMA reads 4 KB of A for 256 columns instead of 128 (shared-memory operand bound)
  int a_es, b_es;  // TMA element stride (1 | 2) of the A / B activation map: coordinates (w*es + pw, h*es + ph)
  int tap_pairs;
  int tap_items;  // tiles along the tap axis: ntaps, or ceil(ntaps / 2) with tap_pairs
  // descriptor constants
  uint32_t a_desc_hi, b_desc_hi;   // upper 32 bits of the shared-memory matrix descriptors
  uint32_t a_lbo, b_lbo;           // leading byte offsets >> 4
  uint32_t a_kstep, b_kstep;       // descriptor start-address advance (>>4) per K=16 MMA
  uint32_t idesc;
  int total_tiles;
  int release_arrive;  // A/B knob (VCD_GEMM_RELEASE=1): hand accumulator stages back with a releasing arrive
};

int umma_launch(const CUtensorMap& mapA, const CUtensorMap& mapB, const UmmaParams& p, int block_n, cudaStream_t st,
                bool overlap_prev = false);

// activation map: dims (C, W, H, P, N); strides derive from a dense [N][P][H][W][C] bf16 tensor.
// es = 2: the box samples every second pixel along W and H (TMA element strides), so a coordinate (2*w + pw, 2*h + ph)
// addresses pixel (w, h) of parity plane (ph, pw) of a plain NHWC tensor — stride-2 convolutions and the phase form of
// Upsample2D read their operands in place, without a space-to-planes copy.  box_w / box_h count LOADED pixels.
int make_act_map(CUtensorMap* m, const void* base, int C, int W, int H, int P, int N, int box_c, int box_w, int box_h,
                 int box_n, int es = 1);
