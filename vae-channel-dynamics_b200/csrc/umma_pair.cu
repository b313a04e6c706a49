// umma_pair.cu — persistent CTA-pair (tcgen05.mma.cta_group::2) implicit-GEMM kernel for sm_100a with
// shared-memory halo reuse across convolution taps.  See umma_pair.cuh for the operand forms.
//
//   warp 0 (1 lane, both CTAs) : TMA producer — A halo box (own 128 pixels) + own half of the B rows, bytes counted
//                                on the LEADER CTA's mbarrier (cp.async.bulk.tensor...cta_group::2)
//   warp 1 (1 lane, leader)    : MMA issuer   — tcgen05.mma.cta_group::2.kind::f16, M = 256, N = BLOCK_N;
//                                tcgen05.commit...multicast::cluster frees stages / publishes accumulators in both CTAs
//   warp 2 (both CTAs)         : TMEM allocator (tcgen05.alloc.cta_group::2)
//   warps 4..7 (both CTAs)     : epilogue — tcgen05.ld of the CTA's own 128 TMEM lanes -> alpha/bias/residual ->
//                                bf16 -> 128-bit global stores; releases the accumulator stage on the leader's barrier
// Two accumulator stages (2 x BLOCK_N TMEM columns) overlap the epilogue of item i with the main loop of item i+1.
// Reference arithmetic replaced: the cuDNN/cuBLAS calls behind diffusers' Conv2d / Linear / attention,
// reached from sdxl_vae_wrapper.py:60,71 and train.py:299.
#include <stdlib.h>
#include <string.h>

#include "umma_pair.cuh"
#include "umma_ptx.cuh"

// Profiling knobs (skip MMAs / stores / B traffic) exist only in a -DVCD_PAIR_DEBUG build: in the product library the
// macro is the constant 0 and the branches vanish from the MMA / TMA loops.
#ifdef VCD_PAIR_DEBUG
#define VCD_PAIR_DBG(p, bit) (((p).dbg & (bit)) != 0)
#else
#define VCD_PAIR_DBG(p, bit) (false)
#endif

namespace {
using namespace umma;

constexpr int kThreads = 384;    // warps 0-2: TMA / MMA / TMEM alloc, warp 3 idle, warps 4-11: epilogue
constexpr int kEpiWarps = 8;
constexpr int kABytes = 23552;  // (8+2) x (16+2) rows x 128 B = 23040, rounded up to a multiple of 1024
constexpr int kBStages = 8;
constexpr int kBResidentBytes = 147456;  // 144 KB: 18 boxes of 64 rows x 128 B (3x3 taps x 2 K chunks, N = 128)
constexpr int kStoreSlabBytes = 2048;          // per epilogue warp: 32 rows x 64 B staging for full-sector global stores
constexpr int kChanAccBytes = 512 * 2 * 4;  // per-channel (sum g*x, sum g) of the fused GroupNorm backward, <= 512 channels

template <int BLOCK_N>
struct PCfg {
  static constexpr int kTapBytes = (BLOCK_N / 2) * 128;  // this CTA's half of one tap's B rows, 64 K-elements each
  static constexpr int kTapsPerStage = 256 / BLOCK_N;    // a B stage is always 16 KB: 512 MMA cycles per barrier wait
  static constexpr int kBBytes = kTapBytes * kTapsPerStage;
  static constexpr int kAStages = 3;
  static constexpr int kTmemCols = 2 * BLOCK_N;
  static constexpr int kSmemBytes =
      kAStages * kABytes + kBStages * kBBytes + 1024 /*align slack*/ + 512 /*barriers*/ + kChanAccBytes +
      kEpiWarps * kStoreSlabBytes;
};

struct TileCoord {
  int w0, h0, n;  // n = Nimg for a tile past the end (TMA then zero-fills the whole A box)
  int n_b;        // image whose B rows the pair uses (batched GEMM): never out of range
  bool valid;
};
__device__ __forceinline__ TileCoord decode_tile(const PairParams& p, int pair, int rank) {
  const int per_n = p.tiles_w * p.tiles_h;
  int n, lt;
  bool valid;
  if (p.pair_in_image) {
    const int ppn = (per_n + 1) >> 1;
    n = pair / ppn;
    lt = (pair - n * ppn) * 2 + rank;
    valid = lt < per_n;
  } else {
    const int t = pair * 2 + rank;
    valid = t < p.pix_tiles;
    n = t / per_n;
    lt = t - n * per_n;
  }
  TileCoord c;
  const int tw = lt % p.tiles_w, th = lt / p.tiles_w;
  c.w0 = tw * (p.mode ? 128 : 8);
  c.h0 = th * (p.mode ? 1 : 16);
  c.n = valid ? n : p.Nimg;  // an out-of-range image index makes TMA zero-fill the whole box
  c.n_b = n < p.Nimg ? n : p.Nimg - 1;
  c.valid = valid;
  return c;
}

// per-thread partial GroupNorm sums of one 32-column chunk: vals[2g] = sum, vals[2g+1] = sum of squares of the
// bf16-rounded outputs of group g (D = 1 << LOGD channels per group, 32 / D groups per chunk)
template <int LOGD>
__device__ __forceinline__ void gn_partials(const float* v, float* vals) {
  constexpr int D = 1 << LOGD;
#pragma unroll
  for (int g = 0; g < 32 / D; ++g) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const float r = __bfloat162float(__float2bfloat16_rn(v[g * D + j]));
      s += r;
      q = fmaf(r, r, q);
    }
    vals[2 * g] = s;
    vals[2 * g + 1] = q;
  }
}
// sum 16 per-lane values over the 32 lanes of the warp in 16 shuffles (recursive halving): afterwards lanes L and L^1
// both hold the warp total of value index (bit4, bit3, bit2, bit1 of L)
__device__ __forceinline__ float warp_reduce16(float* vals, int lane) {
#pragma unroll
  for (int step = 0; step < 4; ++step) {
    const int off = 16 >> step, half = 8 >> step;
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float keep = up ? vals[i + half] : vals[i];
      const float send = up ? vals[i] : vals[i + half];
      vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return vals[0] + __shfl_xor_sync(0xffffffffu, vals[0], 1);
}

// sum 32 per-lane values over the 32 lanes of the warp in 31 shuffles: lane L ends with the warp total of vals[L]
__device__ __forceinline__ float warp_reduce32(float* vals, int lane) {
#pragma unroll
  for (int step = 0; step < 5; ++step) {
    const int off = 16 >> step, half = 16 >> step;
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float keep = up ? vals[i + half] : vals[i];
      const float send = up ? vals[i] : vals[i + half];
      vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return vals[0];
}

// GNB: the epilogue is the fused GroupNorm backward prologue of a dgrad launch (no bias / residual / output sums)
template <int BLOCK_N, bool GNB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
umma_pair_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                 const __grid_constant__ PairParams p) {
  using C = PCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // streaming layout: [A ring: kAStages boxes][B ring: kBStages x 16 KB]; B-resident layout (p.b_resident): [B: ntaps*kc
  // boxes, <= 144 KB][A ring: 2 boxes] inside the same region
  const int n_a_stages = p.b_resident ? 2 : C::kAStages;
  const uint32_t b_base = p.b_resident ? smem_base : smem_base + C::kAStages * kABytes;
  const uint32_t a_base = p.b_resident ? smem_base + kBResidentBytes : smem_base;
  const uint32_t bar_base = smem_base + C::kAStages * kABytes + kBStages * C::kBBytes;
  auto afull = [&](int s) { return bar_base + 8u * s; };
  auto aempty = [&](int s) { return bar_base + 8u * (C::kAStages + s); };
  auto bfull = [&](int s) { return bar_base + 8u * (2 * C::kAStages + s); };
  auto bempty = [&](int s) { return bar_base + 8u * (2 * C::kAStages + kBStages + s); };
  auto tfull = [&](int s) { return bar_base + 8u * (2 * C::kAStages + 2 * kBStages + s); };
  auto tempty = [&](int s) { return bar_base + 8u * (2 * C::kAStages + 2 * kBStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * C::kAStages + 2 * kBStages + 4);
  float* chan_acc = reinterpret_cast<float*>(smem_raw + (bar_base + 512u - smem_u32(smem_raw)));  // [Nout][2]
  const uint32_t slab_base = bar_base + 512u + kChanAccBytes;  // kEpiWarps x kStoreSlabBytes
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  // a weight-gradient GEMM enqueued behind this kernel with VCD_WGRAD_OVERLAP_PREV may start filling SMs as soon as this
  // grid's clusters retire (no effect on plain launches)
  vcd_pdl_trigger();

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::kAStages; ++s) {
      mbar_init(afull(s), 1);
      mbar_init(aempty(s), 1);
    }
    for (int s = 0; s < kBStages; ++s) {
      mbar_init(bfull(s), 1);
      mbar_init(bempty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull(s), 1);
      mbar_init(tempty(s), 2 * kEpiWarps);  // epilogue warps of both CTAs (only the leader's copy is waited on)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (GNB)
    for (int i = threadIdx.x; i < 2 * p.Nout; i += kThreads) chan_acc[i] = 0.f;
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"((uint32_t)C::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ============================== TMA producer (both CTAs; warp-uniform, one elected lane issues) ==========
    if (elect_one_sync()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    }
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    if (p.b_resident && cluster_id < p.total_items) {
      if (elect_one_sync()) {
        if (rank == 0) mbar_arrive_expect_tx(bfull(0), 2u * (uint32_t)(p.ntaps * p.kc * C::kTapBytes));
        for (int kc = 0; kc < p.kc; ++kc)
          for (int tap = 0; tap < p.ntaps; ++tap)
            tma_load_5d_pair(b_base + (uint32_t)((kc * p.ntaps + tap) * C::kTapBytes), &mapB, bfull(0) & kPeerBitMask, kc * 64,
                             rank * (BLOCK_N / 2) + p.tap_brow[tap], 0, 0, 0);
      }
      __syncwarp();
    }
    for (int item = cluster_id; item < p.total_items; item += n_clusters) {
      const int nt = item % p.n_tiles;
      const TileCoord tc = decode_tile(p, item / p.n_tiles, rank);
      const int brow0 = nt * BLOCK_N + rank * (BLOCK_N / 2) + tc.n_b * p.b_batch_rows;
      for (int kc = 0; kc < p.kc; ++kc) {
        for (int g = 0; g < p.ngroups; ++g) {
          mbar_wait(aempty(sa), pa ^ 1u);
          if (elect_one_sync()) {
            if (rank == 0) mbar_arrive_expect_tx(afull(sa), 2u * p.a_box_bytes);
            tma_load_5d_pair(a_base + sa * kABytes, &mapA, afull(sa) & kPeerBitMask, kc * 64,
                             act_cw(tc.w0 + p.g_dw[g], p.g_plane[g], p.a_es), act_ch(tc.h0 + p.g_dh[g], p.g_plane[g], p.a_es),
                             act_cp(p.g_plane[g], p.a_es), tc.n);
          }
          __syncwarp();
          if (++sa == n_a_stages) { sa = 0; pa ^= 1u; }
          if (p.b_resident) continue;
          for (int tap = p.g_tap0[g]; tap < p.g_tap0[g + 1]; tap += C::kTapsPerStage) {
            const int nt_here = min(C::kTapsPerStage, p.g_tap0[g + 1] - tap);
            mbar_wait(bempty(sb), pb ^ 1u);
            if (elect_one_sync()) {
              if (VCD_PAIR_DBG(p, 8)) {  // profiling aid: no B traffic (stale shared memory is multiplied)
                if (rank == 0) mbar_arrive(bfull(sb));
              } else {
                if (rank == 0) mbar_arrive_expect_tx(bfull(sb), 2u * (uint32_t)(nt_here * C::kTapBytes));
                for (int u = 0; u < nt_here; ++u)
                  tma_load_5d_pair(b_base + sb * C::kBBytes + u * C::kTapBytes, &mapB, bfull(sb) & kPeerBitMask,
                                   kc * 64, brow0 + p.tap_brow[tap + u], 0, 0, 0);
              }
            }
            __syncwarp();
            if (++sb == kBStages) { sb = 0; pb ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ============================== MMA issuer (leader CTA; warp-uniform, one elected lane issues) ===========
    int sa = 0, sb = 0, acc = 0;
    uint32_t pa = 0, pb = 0, acc_phase = 0;
    const uint32_t a_hi = (uint32_t)((p.mode ? 1024 : p.box_w * 128) >> 4) | (1u << 14) | (2u << 29);
    const uint32_t b_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    if (p.b_resident && cluster_id < p.total_items) {
      mbar_wait(bfull(0), 0u);   // the resident B operand of both CTAs has landed
      tc_fence_after();
    }
    for (int item = cluster_id; item < p.total_items; item += n_clusters) {
      mbar_wait(tempty(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
      uint32_t accum = 0;
      for (int kc = 0; kc < p.kc; ++kc) {
        for (int g = 0; g < p.ngroups; ++g) {
          mbar_wait(afull(sa), pa);
          tc_fence_after();
          const uint32_t sa_addr = a_base + sa * kABytes;
          if (p.b_resident) {
            if (elect_one_sync()) {
              for (int tap = p.g_tap0[g]; tap < p.g_tap0[g + 1]; ++tap) {
                const uint32_t a0 = sa_addr + (uint32_t)p.tap_aoff[tap];
                const uint32_t b0 = b_base + (uint32_t)((kc * p.ntaps + tap) * C::kTapBytes);
                const uint32_t a_lo = ((a0 >> 4) & 0x3FFFu) | (1u << 16);
                const uint32_t b_lo = ((b0 >> 4) & 0x3FFFu) | (1u << 16);
                umma_bf16_pair(d_tmem, ((uint64_t)a_hi << 32) | a_lo, ((uint64_t)b_hi << 32) | b_lo, p.idesc, accum);
                umma_bf16_pair(d_tmem, ((uint64_t)a_hi << 32) | (a_lo + 2), ((uint64_t)b_hi << 32) | (b_lo + 2), p.idesc, 1u);
                umma_bf16_pair(d_tmem, ((uint64_t)a_hi << 32) | (a_lo + 4), ((uint64_t)b_hi << 32) | (b_lo + 4), p.idesc, 1u);
                umma_bf16_pair(d_tmem, ((uint64_t)a_hi << 32) | (a_lo + 6), ((uint64_t)b_hi << 32) | (b_lo + 6), p.idesc, 1u);
                accum = 1u;
              }
              umma_commit_pair(aempty(sa));
            }
            __syncwarp();
            accum = 1u;
            if (++sa == n_a_stages) { sa = 0; pa ^= 1u; }
            continue;
          }
          for (int tap = p.g_tap0[g]; tap < p.g_tap0[g + 1]; tap += C::kTapsPerStage) {
            const int nt_here = min(C::kTapsPerStage, p.g_tap0[g + 1] - tap);
            mbar_wait(bfull(sb), pb);
            tc_fence_after();
            if (elect_one_sync()) {
              for (int u = 0; u < nt_here; ++u) {
                const uint32_t a0 = sa_addr + (uint32_t)p.tap_aoff[tap + u];
                const uint32_t b0 = b_base + sb * C::kBBytes + u * C::kTapBytes;
                const uint32_t a_lo = ((a0 >> 4) & 0x3FFFu) | (1u << 16);
                const uint32_t b_lo = ((b0 >> 4) & 0x3FFFu) | (1u << 16);
                if (!VCD_PAIR_DBG(p, 2)) {
                  // K advances 16 bf16 = 32 B (descriptor units of 16 B: +2) inside the 128-byte swizzle row
                  umma_bf16_pair(d_tmem, ((uint64_t)a_hi << 32) | a_lo, ((uint64_t)b_hi << 32) | b_lo, p.idesc, accum);
                  umma_bf16_pair(d_tmem, ((uint64_t)a_hi << 32) | (a_lo + 2), ((uint64_t)b_hi << 32) | (b_lo + 2), p.idesc, 1u);
                  umma_bf16_pair(d_tmem, ((uint64_t)a_hi << 32) | (a_lo + 4), ((uint64_t)b_hi << 32) | (b_lo + 4), p.idesc, 1u);
                  umma_bf16_pair(d_tmem, ((uint64_t)a_hi << 32) | (a_lo + 6), ((uint64_t)b_hi << 32) | (b_lo + 6), p.idesc, 1u);
                }
                accum = 1u;
              }
              umma_commit_pair(bempty(sb));
            }
            __syncwarp();
            accum = 1u;
            if (++sb == kBStages) { sb = 0; pb ^= 1u; }
          }
          if (elect_one_sync()) umma_commit_pair(aempty(sa));
          __syncwarp();
          if (++sa == n_a_stages) { sa = 0; pa ^= 1u; }
        }
      }
      if (elect_one_sync()) umma_commit_pair(tfull(acc));
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else if (warp >= 4) {
    // ============================== epilogue (both CTAs) ==============================
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int chalf = (warp - 4) >> 2;      // which half of the BLOCK_N columns
    const int row = q * 32 + lane;
    const int tile_w = p.mode ? 128 : 8;
    const int wi = row % tile_w, hi = row / tile_w;
    int acc = 0;
    uint32_t acc_phase = 0;
    int cur_n = -1;  // image whose channel sums sit in chan_acc (fused GroupNorm backward)
    // all 8 epilogue warps walk the same items: named barrier 1 (256 threads) brackets the flush of chan_acc
    auto flush_chan_acc = [&]() {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int i = threadIdx.x - 128; i < 2 * p.Nout; i += 256) {
        const float t = chan_acc[i];
        if (t != 0.f) atomicAdd(p.gnb_dsdb + (long long)cur_n * 2 * p.Nout + i, t);
        chan_acc[i] = 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    };
    for (int item = cluster_id; item < p.total_items; item += n_clusters) {
      const int nt = item % p.n_tiles;
      const TileCoord tc = decode_tile(p, item / p.n_tiles, rank);
      if (GNB && tc.valid && tc.n != cur_n) {
        if (cur_n >= 0) flush_chan_acc();
        cur_n = tc.n;
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
      const int w = tc.w0 + wi, h = tc.h0 + hi;
      const bool valid = tc.valid && (w < p.W) && (h < p.H);
      const long long off =
          (long long)tc.n * p.out_sn + (long long)h * p.out_sh + (long long)w * p.out_sw + nt * BLOCK_N;
      // side input of the epilogue (residual of a fprop, GroupNorm input of a fused dgrad): register double buffer,
      // the first chunk is requested BEFORE waiting for the accumulator so its latency hides behind this item's MMAs
      const bf16* side = GNB ? p.gnb_x : p.residual;
      bf16x8 sb[2][4];
      auto side_load = [&](int ch, bf16x8* dst) {
        if (side != nullptr && valid && nt * BLOCK_N + ch * 32 + 32 <= p.Nout) {
#pragma unroll
          for (int g = 0; g < 4; ++g) dst[g] = ld8(side + off + ch * 32 + g * 8);
        }
      };
      side_load(chalf * (BLOCK_N / 64), sb[0]);
      // ... and the side input of the NEXT item of this cluster is pulled into L2 now (one prefetch per 128-byte line of this
      // warp's half of the thread's pixel row): its loads, issued a whole item later, then see L2 latency instead of HBM
      // latency (ncu source page: the unpack of the side input was the top long-scoreboard stall of the fused kernels)
      if (side != nullptr && p.side_prefetch && item + n_clusters < p.total_items) {
        const int item2 = item + n_clusters;
        const int nt2 = item2 % p.n_tiles;
        const TileCoord tc2 = decode_tile(p, item2 / p.n_tiles, rank);
        const int w2 = tc2.w0 + wi, h2 = tc2.h0 + hi;
        if (tc2.valid && w2 < p.W && h2 < p.H) {
          const bf16* pf = side + (long long)tc2.n * p.out_sn + (long long)h2 * p.out_sh + (long long)w2 * p.out_sw +
                           nt2 * BLOCK_N + chalf * (BLOCK_N / 2);
#pragma unroll
          for (int l = 0; l < BLOCK_N / 128; ++l)
            if (nt2 * BLOCK_N + chalf * (BLOCK_N / 2) + l * 64 + 64 <= p.Nout)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + l * 64));
        }
      }
      // transposed store (see process_chunk): in store instruction i lane l writes 16 B of row 8i + l/4
      const unsigned vmask = __ballot_sync(0xffffffffu, valid);
      long long off_t[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) off_t[i] = __shfl_sync(0xffffffffu, off, 8 * i + (lane >> 2));
      const uint32_t slab = slab_base + (uint32_t)(warp - 4) * kStoreSlabBytes;
      mbar_wait(tfull(acc), acc_phase);
      tc_fence_after();
      // one 32-column chunk; `cur` holds the chunk's side input (residual or GroupNorm input), already loaded
      auto process_chunk = [&](int ch, const bf16x8* cur) {
        uint32_t r[32];
        tmem_ld32(taddr + ch * 32, r);
        tmem_wait_ld();
        const int col0 = nt * BLOCK_N + ch * 32;
        // full chunk of an 8-element-aligned tensor: the 32 x 64 B block goes through shared memory so that every
        // global store instruction writes 8 rows x 64 contiguous bytes (whole sectors) instead of 32 rows x 16 B
        const bool staged = (p.Nout & 7) == 0 && col0 + 32 <= p.Nout && !VCD_PAIR_DBG(p, 16);
        float gv[GNB ? 1 : 16];
        if (!GNB) {
#pragma unroll
          for (int j = 0; j < 16; ++j) gv[j] = 0.f;
        }
        float v[32], xs[GNB ? 32 : 1];  // xs: GroupNorm input of the fused backward prologue, then g * x
        if (GNB) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = xs[j] = 0.f;
        }
        if (valid && col0 < p.Nout && !VCD_PAIR_DBG(p, 4)) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
          if constexpr (GNB) {  // v = dL/d act(GN(x)) -> g = v * SiLU'(a x + b)
#pragma unroll
            for (int g = 0; g < 4; ++g) unpack8(cur[g], xs + g * 8);
            if (p.gnb_act) {
              const float4* abp = reinterpret_cast<const float4*>(p.gnb_ab + ((long long)tc.n * p.Nout + col0) * 2);
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float4 ab = __ldg(abp + j);
                v[2 * j] *= silu_grad_f(fmaf(ab.x, xs[2 * j], ab.y));
                v[2 * j + 1] *= silu_grad_f(fmaf(ab.z, xs[2 * j + 1], ab.w));
              }
            }
          }
          if (!GNB && p.bias) {
            const float* bp = p.bias + col0;
            if (col0 + 32 <= p.Nout) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp) + j);
                v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.Nout) v[j] += __ldg(bp + j);
            }
          }
          if (!GNB && p.residual && col0 + 32 <= p.Nout) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float f[8];
              unpack8(cur[g], f);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[g * 8 + j] += f[j];
            }
          }
          if (staged) {
            // this lane's row (64 B) -> the warp's staging slab, 16-byte chunk c at slot c ^ ((lane >> 1) & 3): both
            // this write and the transposed read below are bank-conflict free
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const bf16x8 pk = pack8(v + g * 8);
              const uint32_t a = slab + (uint32_t)lane * 64u + (uint32_t)((g ^ ((lane >> 1) & 3)) * 16);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pk.u.x), "r"(pk.u.y), "r"(pk.u.z), "r"(pk.u.w)
                           : "memory");
            }
          } else {
            bf16* op = p.out + off + ch * 32;
            if ((p.Nout & 7) == 0) {
#pragma unroll
              for (int g = 0; g < 4; ++g)
                if (col0 + g * 8 < p.Nout) st8(op + g * 8, pack8(v + g * 8));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.Nout) op[j] = __float2bfloat16_rn(v[j]);
            }
          }
          if constexpr (!GNB) {
            if (p.gn_sums) {
              if (p.gn_logD == 2) gn_partials<2>(v, gv);
              else if (p.gn_logD == 3) gn_partials<3>(v, gv);
              else gn_partials<4>(v, gv);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) xs[j] *= v[j];
          }
        }
        if (staged && tc.valid && !VCD_PAIR_DBG(p, 4)) {  // warp-uniform
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = 8 * i + (lane >> 2), c = lane & 3;
            uint32_t x0, x1, x2, x3;
            const uint32_t a = slab + (uint32_t)r * 64u + (uint32_t)((c ^ ((r >> 1) & 3)) * 16);
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(a) : "memory");
            if ((vmask >> r) & 1u) {
              bf16x8 o;
              o.u = make_uint4(x0, x1, x2, x3);
              st8(p.out + off_t[i] + ch * 32 + c * 8, o);
            }
          }
          __syncwarp();
        }
        if constexpr (GNB) {
          if (tc.valid && col0 < p.Nout) {  // warp-uniform
          const float sgx = warp_reduce32(xs, lane);  // sum over the warp's 32 pixels of g * x, channel col0 + lane
          const float sg = warp_reduce32(v, lane);    //                                  of g
            atomicAdd(&chan_acc[(col0 + lane) * 2], sgx);
            atomicAdd(&chan_acc[(col0 + lane) * 2 + 1], sg);
          }
        } else if (p.gn_sums && tc.valid && col0 < p.Nout) {  // warp-uniform: all 32 lanes take part in the shuffles
          const float tot = warp_reduce16(gv, lane);
          const int id = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
          if ((lane & 1) == 0 && id < (64 >> p.gn_logD)) {
            const int n_img = p.mode ? (int)(((long long)tc.n * p.W + tc.w0) / p.gn_rows_per_img) : tc.n;
            const int group = (col0 >> p.gn_logD) + (id >> 1);
            atomicAdd(p.gn_sums + ((long long)n_img * p.gn_G + group) * 2 + (id & 1), (double)tot);
          }
        }
      };
      if (!VCD_PAIR_DBG(p, 1)) {
        constexpr int NCH = BLOCK_N / 64;  // chunks per warp
        const int ch0 = chalf * NCH;
#pragma unroll 1
        for (int i = 0; i < NCH; i += 2) {
          side_load(ch0 + i + 1, sb[1]);   // the next chunk's side input is in flight while this one is processed
          process_chunk(ch0 + i, sb[0]);
          if (i + 2 < NCH) side_load(ch0 + i + 2, sb[0]);
          process_chunk(ch0 + i + 1, sb[1]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (p.release_arrive) mbar_arrive_cluster(tempty(acc) & kPeerBitMask);
        else mbar_arrive_cluster_relaxed(tempty(acc) & kPeerBitMask);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (GNB && cur_n >= 0) flush_chan_acc();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the leader's MMAs read the peer's shared memory; nobody leaves before both are done
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::kTmemCols)
                 : "memory");
  }
}

// ====================================================================================================================
// CTA-pair weight-gradient kernel (see umma_pair.cuh PairWgradParams)
// ====================================================================================================================
template <int BLOCK_N>
struct WCfg {
  static constexpr int kABytes = 16384;                 // 128 output channels x 64 pixels (2 boxes)
  static constexpr int kBBytes = (BLOCK_N / 2) * 128;   // this CTA's half of the input channels x 64 pixels
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BLOCK_N == 256) ? 5 : 7;
  static constexpr int kTmemCols = 2 * BLOCK_N;
  static constexpr int kRedSlabBytes = 4096;   // per epilogue warp: 32 x 32 fp32 staging for the transposed reduction
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 512 + kEpiWarps * kRedSlabBytes;
};
struct WItem {
  int tap, nt, mp, k0, nk;
};
__device__ __forceinline__ WItem decode_witem(const PairWgradParams& p, int t) {
  WItem r;  // taps fastest: clusters running together share the same pixel range in L2
  r.tap = t % p.ntaps;
  int q = t / p.ntaps;
  r.nt = q % p.n_tiles;
  q /= p.n_tiles;
  r.mp = q % p.m_pairs;
  const int split = q / p.m_pairs;
  r.k0 = split * p.k_per_split;
  r.nk = min(r.k0 + p.k_per_split, p.k_tiles) - r.k0;
  return r;
}

template <int BLOCK_N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
umma_pair_wgrad_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                       const __grid_constant__ PairWgradParams p) {
  using C = WCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + C::kStages * C::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
  auto tfull = [&](int s) { return bar_base + 8u * (2 * C::kStages + s); };
  auto tempty = [&](int s) { return bar_base + 8u * (2 * C::kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * C::kStages + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull(s), 1);
      mbar_init(tempty(s), 2 * kEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"((uint32_t)C::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ---------------- TMA producer (both CTAs)
    if (elect_one_sync()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int item = cluster_id; item < p.total_items; item += n_clusters) {
      const WItem it = decode_witem(p, item);
      const int a_c0 = it.mp * 256 + rank * 128;
      const int b_c0 = it.nt * BLOCK_N + rank * (BLOCK_N / 2);
      for (int k = 0; k < it.nk; ++k) {
        const int t = it.k0 + k;
        const int tw = t % p.tiles_w, tq = t / p.tiles_w;
        const int w0 = tw * p.tile_w, h0 = (tq % p.tiles_h) * p.tile_h, n0 = (tq / p.tiles_h) * p.tile_n;
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (elect_one_sync()) {
          const uint32_t sa = smem_base + stage * C::kStageBytes;
          const uint32_t lbar = full_bar(stage) & kPeerBitMask;
          if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2u * C::kStageBytes);
#pragma unroll
          for (int b = 0; b < 2; ++b)
            tma_load_5d_pair(sa + b * 8192, &mapA, lbar, a_c0 + b * 64, act_cw(w0, p.tap_plane_a[it.tap], p.a_es),
                             act_ch(h0, p.tap_plane_a[it.tap], p.a_es), act_cp(p.tap_plane_a[it.tap], p.a_es), n0);
#pragma unroll
          for (int b = 0; b < BLOCK_N / 128; ++b)
            tma_load_5d_pair(sa + C::kABytes + b * 8192, &mapB, lbar, b_c0 + b * 64,
                             act_cw(w0 + p.tap_dw[it.tap], p.tap_plane[it.tap], p.b_es),
                             act_ch(h0 + p.tap_dh[it.tap], p.tap_plane[it.tap], p.b_es), act_cp(p.tap_plane[it.tap], p.b_es),
                             n0);
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ---------------- MMA issuer (leader)
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    for (int item = cluster_id; item < p.total_items; item += n_clusters) {
      const WItem it = decode_witem(p, item);
      mbar_wait(tempty(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
      for (int k = 0; k < it.nk; ++k) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t sa = smem_base + stage * C::kStageBytes;
          const uint32_t a_lo = ((sa >> 4) & 0x3FFFu) | ((8192u >> 4) << 16);               // LBO: next 64-channel box
          const uint32_t b_lo = (((sa + C::kABytes) >> 4) & 0x3FFFu) | ((8192u >> 4) << 16);
#pragma unroll
          for (int j = 0; j < 4; ++j)   // K = 16 pixels = 16 rows of 128 B per MMA
            umma_bf16_pair(d_tmem, ((uint64_t)hi << 32) | (a_lo + j * (2048u >> 4)),
                           ((uint64_t)hi << 32) | (b_lo + j * (2048u >> 4)), p.idesc, (k | j) != 0 ? 1u : 0u);
          umma_commit_pair(empty_bar(stage));
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
      }
      if (elect_one_sync()) umma_commit_pair(tfull(acc));
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue (both CTAs): fp32 accumulators -> red.global.add
    const int q = warp & 3, chalf = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = cluster_id; item < p.total_items; item += n_clusters) {
      const WItem it = decode_witem(p, item);
      mbar_wait(tfull(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
      const int co = it.mp * 256 + rank * 128 + row;
      const int co0 = it.mp * 256 + rank * 128 + q * 32;   // first row of this warp
      float* op = p.acc + ((long long)it.tap * p.Mout + co) * (long long)p.Nout + it.nt * BLOCK_N;
      float* op0 = p.acc + ((long long)it.tap * p.Mout + co0) * (long long)p.Nout + it.nt * BLOCK_N;
      const uint32_t slab = bar_base + 512u + (uint32_t)(warp - 4) * C::kRedSlabBytes;
#pragma unroll 1
      for (int ch = chalf * (BLOCK_N / 64); ch < (chalf + 1) * (BLOCK_N / 64); ++ch) {
        uint32_t r[32];
        tmem_ld32(taddr + ch * 32, r);
        tmem_wait_ld();
        if ((p.Nout & 3) == 0 && it.nt * BLOCK_N + ch * 32 + 32 <= p.Nout) {
          red_chunk_32x32(r, slab, op0 + ch * 32, p.Nout, p.Mout - co0, lane);
        } else if (co < p.Mout) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (it.nt * BLOCK_N + ch * 32 + j < p.Nout) atomicAdd(op + ch * 32 + j, __uint_as_float(r[j]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(tempty(acc) & kPeerBitMask);
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::kTmemCols)
                 : "memory");
  }
}

uint32_t pair_idesc(int n) {
  // kind::f16: D = f32, A = B = bf16, both K-major, N >> 3, M = 256 >> 4
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((256u >> 4) << 24);
}

}  // namespace

static int g_pair_on = -1;
bool pair_enabled() {
  if (g_pair_on < 0) {
    const char* e = getenv("VCD_PAIR");
    g_pair_on = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pair_on == 1;
}
extern "C" int vcd_set_pair_kernels(int enabled) {
  const int prev = pair_enabled() ? 1 : 0;
  g_pair_on = enabled ? 1 : 0;
  return prev;
}

bool pair_setup_halo(PairParams& p, int W, int H, int N, const PairTap* taps, int ntaps, int block_n, int* box_h_out) {
  if (W < 8 || H < 16 || ntaps < 1 || ntaps > 16) return false;
  p.mode = 0;
  if (p.a_es != 2) p.a_es = 1;
  p.W = W; p.H = H; p.Nimg = N;
  p.tiles_w = (W + 7) / 8;
  p.tiles_h = (H + 15) / 16;
  p.pix_tiles = p.tiles_w * p.tiles_h * N;
  p.pair_in_image = 0;
  p.pairs = (p.pix_tiles + 1) / 2;
  // groups = distinct planes, in order of first appearance
  int ng = 0, gmin_h[4], gmin_w[4], gmax_h[4], gmax_w[4], tap_group[16];
  for (int t = 0; t < ntaps; ++t) {
    int g = -1;
    for (int k = 0; k < ng; ++k)
      if (p.g_plane[k] == taps[t].plane) g = k;
    if (g < 0) {
      if (ng == 4) return false;
      g = ng++;
      p.g_plane[g] = taps[t].plane;
      gmin_h[g] = gmax_h[g] = taps[t].dh;
      gmin_w[g] = gmax_w[g] = taps[t].dw;
    }
    tap_group[t] = g;
    if (taps[t].dh < gmin_h[g]) gmin_h[g] = taps[t].dh;
    if (taps[t].dh > gmax_h[g]) gmax_h[g] = taps[t].dh;
    if (taps[t].dw < gmin_w[g]) gmin_w[g] = taps[t].dw;
    if (taps[t].dw > gmax_w[g]) gmax_w[g] = taps[t].dw;
  }
  int halo_h = 0, halo_w = 0;
  for (int g = 0; g < ng; ++g) {
    if (gmax_h[g] - gmin_h[g] > halo_h) halo_h = gmax_h[g] - gmin_h[g];
    if (gmax_w[g] - gmin_w[g] > halo_w) halo_w = gmax_w[g] - gmin_w[g];
  }
  if (halo_h > 2 || halo_w > 2) return false;
  p.ngroups = ng;
  p.box_w = 8 + halo_w;
  *box_h_out = 16 + halo_h;
  p.a_box_bytes = (uint32_t)(p.box_w * (16 + halo_h) * 128);
  int k = 0;
  for (int g = 0; g < ng; ++g) {
    p.g_dh[g] = gmin_h[g];
    p.g_dw[g] = gmin_w[g];
    p.g_tap0[g] = k;
    for (int t = 0; t < ntaps; ++t)
      if (tap_group[t] == g) {
        p.tap_aoff[k] = ((taps[t].dh - gmin_h[g]) * p.box_w + (taps[t].dw - gmin_w[g])) * 128;
        p.tap_brow[k] = taps[t].brow;
        ++k;
      }
  }
  p.g_tap0[ng] = k;
  p.ntaps = ntaps;
  p.idesc = pair_idesc(block_n);
  return true;
}

void pair_setup_rows(PairParams& p, int rows, int batch, int pair_in_image, int brow, int block_n) {
  p.mode = 1;
  p.a_es = 1;
  p.W = rows; p.H = 1; p.Nimg = batch;
  p.tiles_w = (rows + 127) / 128;
  p.tiles_h = 1;
  p.pix_tiles = p.tiles_w * batch;
  p.pair_in_image = pair_in_image;
  p.pairs = pair_in_image ? ((p.tiles_w + 1) / 2) * batch : (p.pix_tiles + 1) / 2;
  p.box_w = 128;
  p.a_box_bytes = 128 * 128;
  p.ngroups = 1;
  p.g_plane[0] = 0; p.g_dh[0] = 0; p.g_dw[0] = 0;
  p.g_tap0[0] = 0; p.g_tap0[1] = 1;
  p.ntaps = 1;
  p.tap_aoff[0] = 0;
  p.tap_brow[0] = brow;
  p.idesc = pair_idesc(block_n);
}

static long long g_pair_launches = 0;
extern "C" int64_t vcd_pair_kernel_launches(void) { return g_pair_launches; }

int pair_launch(const CUtensorMap& mapA, const CUtensorMap& mapB, PairParams& p, int block_n, cudaStream_t st) {
  p.total_items = p.pairs * p.n_tiles;
  if (p.total_items <= 0) return 0;
  ++g_pair_launches;
  {
    static int rel = -1;
    if (rel < 0) { const char* e = getenv("VCD_PAIR_RELEASE"); rel = (e && e[0] == '1') ? 1 : 0; }
    p.release_arrive = rel;
    static int pf = -1;
    if (pf < 0) { const char* e = getenv("VCD_PAIR_PREFETCH"); pf = (e && e[0] == '0') ? 0 : 1; }
    p.side_prefetch = pf;
#ifdef VCD_PAIR_DEBUG
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("VCD_PAIR_DBG"); dbg = e ? atoi(e) : 0; }
    p.dbg = dbg;
#else
    p.dbg = 0;
#endif
  }
  const int max_clusters = vcd_num_sms() / 2;
  const int clusters = p.total_items < max_clusters ? p.total_items : max_clusters;
  const int grid = clusters * 2;
  const bool gnb = p.gnb_x != nullptr;
  auto launch = [&](auto kernel, int smem) -> int {
    bool& done = *vcd_device_once(2 + (block_n == 256 ? 2 : 0) + (gnb ? 1 : 0));
    if (!done) {
      VCD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      done = true;
    }
    kernel<<<grid, kThreads, smem, st>>>(mapA, mapB, p);
    return 0;
  };
  int rc;
  if (block_n == 256) rc = gnb ? launch(umma_pair_kernel<256, true>, PCfg<256>::kSmemBytes)
                               : launch(umma_pair_kernel<256, false>, PCfg<256>::kSmemBytes);
  else if (block_n == 128) rc = gnb ? launch(umma_pair_kernel<128, true>, PCfg<128>::kSmemBytes)
                                    : launch(umma_pair_kernel<128, false>, PCfg<128>::kSmemBytes);
  else {
    vcd_set_error("pair_launch: BLOCK_N %d unsupported", block_n);
    return -1;
  }
  if (rc) return rc;
  VCD_LAUNCH_CHECK();
  return 0;
}

int pair_wgrad_launch(const CUtensorMap& mapA, const CUtensorMap& mapB, PairWgradParams& p, int block_n, cudaStream_t st,
                      bool overlap_prev) {
  p.total_items = p.ntaps * p.n_tiles * p.m_pairs * p.splits;
  if (p.total_items <= 0) return 0;
  if (p.a_es != 2) p.a_es = 1;
  if (p.b_es != 2) p.b_es = 1;
  ++g_pair_launches;
  // kind::f16: D = f32, A = B = bf16, both MN-major (bits 15, 16), N >> 3, M = 256 >> 4
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(block_n >> 3) << 17) |
            ((256u >> 4) << 24);
  const int max_clusters = vcd_num_sms() / 2;
  const int grid = 2 * (p.total_items < max_clusters ? p.total_items : max_clusters);
  bool* attr_set[2] = {vcd_device_once(6), vcd_device_once(7)};
  if (block_n == 256) {
    if (!*attr_set[0]) {
      VCD_CUDA(cudaFuncSetAttribute(umma_pair_wgrad_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    WCfg<256>::kSmemBytes));
      *attr_set[0] = true;
    }
    VCD_CUDA(vcd_launch(umma_pair_wgrad_kernel<256>, grid, kThreads, WCfg<256>::kSmemBytes, st, overlap_prev, mapA, mapB, p));
  } else if (block_n == 128) {
    if (!*attr_set[1]) {
      VCD_CUDA(cudaFuncSetAttribute(umma_pair_wgrad_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    WCfg<128>::kSmemBytes));
      *attr_set[1] = true;
    }
    VCD_CUDA(vcd_launch(umma_pair_wgrad_kernel<128>, grid, kThreads, WCfg<128>::kSmemBytes, st, overlap_prev, mapA, mapB, p));
  } else {
    vcd_set_error("pair_wgrad_launch: BLOCK_N %d unsupported", block_n);
    return -1;
  }
  VCD_LAUNCH_CHECK();
  return 0;
}
