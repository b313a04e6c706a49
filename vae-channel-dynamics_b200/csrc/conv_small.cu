// conv_small.cu — the 3-channel ends of the VAE (encoder.conv_in fprop: 3 -> 128, decoder.conv_out dgrad: 3 -> 128) as ONE
// HBM-bound kernel: out[px][O] = bias[O] + sum_{t,s} src[px + shift_t][s] * pack[t][O][s]   with S <= 8 source channels.
//
// [upstream] diffusers Conv2d(3, 128, 3, padding=1) / the data gradient of Conv2d(128, 3, 3, padding=1), reached from
// sdxl_vae_wrapper.py:60,71 and train.py:299.  These layers hold 0.3 % of the FLOPs (K = 27) but write / read the largest
// tensors of the network ([8,512,512,128] = 537 MB): they are bound by how fast that tensor moves, not by arithmetic.
// Round 2's first form materialised an im2col patch in HBM (268 MB written + read again per pass) and ran a K = 64 GEMM on
// the tcgen05 kernel whose per-item epilogue hand-shake capped it at ~1.6 TB/s of output: 0.40 ms per pass against
// 0.08 ms for the output write alone.  Here the patch lives only in shared memory:
//   block (256 threads, persistent) : weights wk[O = 128][Kp] staged once; then per tile of 128 consecutive pixels of one
//   image row: 3 halo rows -> shared memory, patch [128][Kp] built there from an offset table, 8 warps x (16 pixels x 128
//   outputs) with mma.sync.m16n8k16 (bf16 in, fp32 accumulate; the arithmetic is ~2 % of the kernel), bias, bf16, and the
//   warp's 16 x 256-byte rows leave through a shared-memory slab as 16-byte coalesced stores.
// Algorithmic bytes per pixel: 2*S read + 2*O written (262 B for 3 -> 128).
#include <stdlib.h>

#include "common.cuh"
#include "conv_dispatch.h"

namespace {

constexpr int kThreads = 256;
constexpr int kPx = 128;              // pixels per tile (one image row segment)
constexpr int kO = 128;               // output channels (one chunk)
constexpr int kMaxKp = 64;            // taps * S padded to a multiple of 16
constexpr int kSlabStride = kO + 8;   // bf16 elements per slab row (272 B: 16-byte aligned, conflict-free fragment writes)

struct SmallTaps {
  int n;
  int dh[9], dw[9];
};

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(kThreads) small_in_conv_kernel(const bf16* __restrict__ src, const bf16* __restrict__ pack,
                                                                const float* __restrict__ bias, bf16* __restrict__ out,
                                                                int N, int H, int W, int S, int O_total, int Kp,
                                                                SmallTaps taps, int total_tiles, int late_bias) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int ks = Kp + 8;                                       // row stride (elements) of wk and patch: conflict-free
  bf16* wk = reinterpret_cast<bf16*>(smem);                    // [kO][ks]
  bf16* patch = wk + kO * ks;                                  // [kPx][ks]
  bf16* slab = patch + kPx * ks;                               // [8 warps][16][kSlabStride]
  bf16* rows = slab + 8 * 16 * kSlabStride;                    // [3][kPx + 2][S]
  float* sbias = reinterpret_cast<float*>(rows + 3 * (kPx + 2) * 8);   // [kO]
  short* off = reinterpret_cast<short*>(sbias + kO);           // [Kp]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  const int KS = taps.n * S;

  // weights: wk[o][t*S + s] = pack[t][o][s], zero beyond taps*S; bias; column -> halo offset table
  for (int i = threadIdx.x; i < kO * Kp; i += kThreads) {
    const int o = i / Kp, col = i - o * Kp;
    const int t = col / S, s = col - t * S;
    wk[o * ks + col] = col < KS ? pack[((int64_t)t * O_total + o) * S + s] : __float2bfloat16_rn(0.f);
  }
  for (int i = threadIdx.x; i < kO; i += kThreads) sbias[i] = bias ? bias[i] : 0.f;
  for (int col = threadIdx.x; col < Kp; col += kThreads) {
    const int t = col / S, c = col - t * S;
    off[col] = col < KS ? (short)(((taps.dh[t] + 1) * (kPx + 2) + (taps.dw[t] + 1)) * S + c) : (short)-1;
  }
  const int segs = (W + kPx - 1) / kPx;
  const int RW = (kPx + 2) * S;
  const int V = Kp >> 3;
  bf16* myslab = slab + warp * 16 * kSlabStride;

  // halo rows of a tile -> registers (issued one tile ahead: their latency hides behind the patch / MMA / store phases of
  // the current tile; with two resident blocks per SM nothing else would cover it)
  constexpr int kPre = (3 * (kPx + 2) * 8 + kThreads - 1) / kThreads;
  bf16 pre[kPre];
  auto prefetch = [&](int tile) {
    const int seg = tile % segs;
    const int64_t row = tile / segs;
    const int h = (int)(row % H);
    const int64_t n = row / H;
    const int w0 = seg * kPx;
#pragma unroll
    for (int j = 0; j < kPre; ++j) {
      const int i = threadIdx.x + j * kThreads;
      bf16 v = __float2bfloat16_rn(0.f);
      if (i < 3 * RW) {
        const int r = i / RW, e = i - r * RW;
        const int wi = w0 - 1 + e / S, hi = h - 1 + r;
        if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = src[((n * H + hi) * W + wi) * S + (e % S)];
      }
      pre[j] = v;
    }
  };
  if ((int)blockIdx.x < total_tiles) prefetch(blockIdx.x);

  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int seg = tile % segs;
    const int64_t row = tile / segs;            // n * H + h
    const int w0 = seg * kPx;
    const int npx = min(kPx, W - w0);
    __syncthreads();                            // previous tile's patch / rows fully consumed (and wk / off written)
#pragma unroll
    for (int j = 0; j < kPre; ++j) {
      const int i = threadIdx.x + j * kThreads;
      if (i < 3 * RW) rows[i] = pre[j];
    }
    __syncthreads();
    if (tile + (int)gridDim.x < total_tiles) prefetch(tile + gridDim.x);
    for (int i = threadIdx.x; i < kPx * V; i += kThreads) {   // patch[px][8v .. 8v+7]
      const int v = i % V, px = i / V;
      const bf16* base = rows + px * S;
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int o = off[v * 8 + j];
        f[j] = (o >= 0 && px < npx) ? __bfloat162float(base[o]) : 0.f;
      }
      *reinterpret_cast<uint4*>(patch + px * ks + v * 8) = pack8(f).u;
    }
    __syncthreads();
    if (warp * 16 < npx) {
      float acc[16][4];
#pragma unroll
      for (int nt = 0; nt < 16; ++nt) {
        const float b0 = late_bias ? 0.f : sbias[nt * 8 + 2 * tg], b1 = late_bias ? 0.f : sbias[nt * 8 + 2 * tg + 1];
        acc[nt][0] = b0; acc[nt][1] = b1; acc[nt][2] = b0; acc[nt][3] = b1;
      }
      const bf16* arow = patch + (warp * 16 + g) * ks;
      for (int k0 = 0; k0 < Kp; k0 += 16) {
        uint32_t a[4];
        a[0] = *reinterpret_cast<const uint32_t*>(arow + k0 + 2 * tg);
        a[1] = *reinterpret_cast<const uint32_t*>(arow + 8 * ks + k0 + 2 * tg);
        a[2] = *reinterpret_cast<const uint32_t*>(arow + k0 + 8 + 2 * tg);
        a[3] = *reinterpret_cast<const uint32_t*>(arow + 8 * ks + k0 + 8 + 2 * tg);
#pragma unroll
        for (int nt = 0; nt < 16; ++nt) {
          const bf16* brow = wk + (nt * 8 + g) * ks + k0 + 2 * tg;
          mma_bf16_16816(acc[nt], a, *reinterpret_cast<const uint32_t*>(brow), *reinterpret_cast<const uint32_t*>(brow + 8));
        }
      }
      if (late_bias) {   // default: bias added AFTER the products, as the GEMM epilogue of the im2col path does — seeding the
#pragma unroll           // accumulator with it (VCD_SMALL_CONV_LATE_BIAS=0) rounds in another order and flips 1 output in 2 M
        for (int nt = 0; nt < 16; ++nt) {
          const float b0 = sbias[nt * 8 + 2 * tg], b1 = sbias[nt * 8 + 2 * tg + 1];
          acc[nt][0] += b0; acc[nt][1] += b1; acc[nt][2] += b0; acc[nt][3] += b1;
        }
      }
      // fragment -> slab (row g / g+8, columns nt*8 + 2tg, +1), then 16 x 256-byte rows as 16-byte stores
#pragma unroll
      for (int nt = 0; nt < 16; ++nt) {
        *reinterpret_cast<uint32_t*>(myslab + g * kSlabStride + nt * 8 + 2 * tg) = f2_to_bf2(acc[nt][0], acc[nt][1]);
        *reinterpret_cast<uint32_t*>(myslab + (g + 8) * kSlabStride + nt * 8 + 2 * tg) = f2_to_bf2(acc[nt][2], acc[nt][3]);
      }
      __syncwarp();
      bf16* obase = out + ((row * W + w0 + warp * 16) * (int64_t)O_total);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = i * 2 + (lane >> 4), c8 = lane & 15;
        if (warp * 16 + r < npx)
          *reinterpret_cast<uint4*>(obase + (int64_t)r * O_total + c8 * 8) =
              *reinterpret_cast<const uint4*>(myslab + r * kSlabStride + c8 * 8);
      }
      __syncwarp();
    }
  }
}

size_t small_smem_bytes(int Kp) {
  const int ks = Kp + 8;
  return (size_t)(kO * ks + kPx * ks + 8 * 16 * kSlabStride + 3 * (kPx + 2) * 8) * sizeof(bf16) + kO * sizeof(float) +
         (size_t)Kp * sizeof(short) + 16;
}

}  // namespace

bool small_in_conv_ok(int S, int O, int KH, int KW, int stride) {
  // OPT-IN (VCD_SMALL_CONV=1).  Measured on B200: 3 -> 128 fprop at 512^2 B=8 0.463 -> 0.326 ms; with the bias added after the
  // products (the default, as the im2col path's GEMM epilogue does) the 6-step AdamW trajectory test reads the same step-1
  // loss to six digits as the im2col path; with the bias seeding the accumulator ONE of 2 097 152 conv_in outputs differs by
  // a bf16 ulp, which re-draws the rounding noise of the network behind it and moved that test's post-nudge loss errors
  // from 2-5e-3 to 1.1e-2.  Opt-in only because the full GPU suite was last run without it.
  static int on = -1;
  if (on < 0) { const char* e = getenv("VCD_SMALL_CONV"); on = (e && e[0] == '1') ? 1 : 0; }
  if (!on) return false;
  const int taps = KH * KW;
  return stride == 1 && S >= 1 && S <= 8 && O == kO && taps <= 9 && taps * S <= kMaxKp;
}

// out[px][O] = bias + sum_{t,s} src[px + sign*(k - pad)][s] * pack[t][O][s]   (pack = w_fprop for fprop, w_dgrad for dgrad)
int small_in_conv_launch(const void* src, const void* pack, const float* bias, void* out, int N, int H, int W, int S, int O,
                         int KH, int KW, int pad_t, int pad_l, int sign, cudaStream_t st) {
  VCD_CHECK_ARG(small_in_conv_ok(S, O, KH, KW, 1), "small-channel conv: unsupported shape (S=%d, O=%d)", S, O);
  SmallTaps taps;
  taps.n = KH * KW;
  for (int kh = 0; kh < KH; ++kh)
    for (int kw = 0; kw < KW; ++kw) {
      taps.dh[kh * KW + kw] = sign * (kh - pad_t);
      taps.dw[kh * KW + kw] = sign * (kw - pad_l);
      VCD_CHECK_ARG(abs(taps.dh[kh * KW + kw]) <= 1 && abs(taps.dw[kh * KW + kw]) <= 1, "small-channel conv: taps beyond 3x3");
    }
  const int Kp = (taps.n * S + 15) / 16 * 16;
  const int segs = (W + kPx - 1) / kPx;
  const int64_t tiles = (int64_t)N * H * segs;
  VCD_CHECK_ARG(tiles < (1ll << 31), "small-channel conv: too many rows");
  const size_t smem = small_smem_bytes(Kp);
  bool& attr = *vcd_device_once(9);
  if (!attr) {
    VCD_CUDA(cudaFuncSetAttribute(small_in_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)small_smem_bytes(kMaxKp)));
    attr = true;
  }
  int64_t grid = (int64_t)vcd_num_sms() * 3;
  if (grid > tiles) grid = tiles;
  static int late = -1;
  if (late < 0) { const char* e = getenv("VCD_SMALL_CONV_LATE_BIAS"); late = (e && e[0] == '0') ? 0 : 1; }
  small_in_conv_kernel<<<(unsigned)grid, kThreads, smem, st>>>((const bf16*)src, (const bf16*)pack, bias, (bf16*)out, N, H, W,
                                                              S, O, Kp, taps, (int)tiles, late);
  VCD_LAUNCH_CHECK();
  return 0;
}
