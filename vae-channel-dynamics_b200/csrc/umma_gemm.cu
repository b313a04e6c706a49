// umma_gemm.cu — persistent, warp-specialised tcgen05 implicit-GEMM kernel for sm_100a.
//
//   warp 0 (1 lane)  : TMA producer  — cp.async.bulk.tensor.5d into a ring of 128B-swizzled stages
//   warp 1 (1 lane)  : MMA issuer    — tcgen05.mma.cta_group::1.kind::f16, fp32 accumulators in TMEM,
//                                      tcgen05.commit releases stages / publishes accumulators
//   warp 2           : TMEM allocator (tcgen05.alloc / dealloc)
//   warps 4..7       : epilogue      — tcgen05.ld (each warp its own 32 TMEM lanes) -> bias/residual/
//                                      bf16 pack -> global (form 0) or red.global.add.f32 (form 1)
// Two TMEM accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.
// See umma_gemm.cuh for the two operand forms.  Reference arithmetic replaced: the cuDNN/cuBLAS calls
// behind diffusers' Conv2d/Linear/attention, reached from sdxl_vae_wrapper.py:60,71 and
// train.py:299.
#include "umma_gemm.cuh"
#include "umma_ptx.cuh"

namespace {
using namespace umma;

constexpr int kStageA = 16384;  // 128 rows x 128 B (form 0) or 2 boxes of 64 x 128 B (form 1)
constexpr int kThreads = 256;
struct PixTile {
  int w0, h0, n0;
};
__device__ __forceinline__ PixTile decode_pix(const UmmaParams& p, int t) {
  PixTile r;
  int tw = t % p.tiles_w;
  int q = t / p.tiles_w;
  r.w0 = tw * p.tile_w;
  r.h0 = (q % p.tiles_h) * p.tile_h;
  r.n0 = (q / p.tiles_h) * p.tile_n;
  return r;
}

// form-1 tile -> (batch, tap, m tile, n tile, first K step, K steps)
struct RedTile {
  int batch, tap, mt, nt, k0, nk;
};
__device__ __forceinline__ RedTile decode_red(const UmmaParams& p, int t) {
  // taps fastest: the CTAs running at the same time work on the SAME pixel range for all taps / channel tiles, so
  // the activation and gradient tiles are read from DRAM once and served from L2 to the other taps
  RedTile r;
  r.tap = t % p.tap_items;
  int q = t / p.tap_items;
  r.nt = q % p.n_tiles;
  q /= p.n_tiles;
  r.mt = q % p.m_tiles;
  q /= p.m_tiles;
  int split = q % p.splits;
  r.batch = q / p.splits;
  r.k0 = split * p.k_per_split;
  int k1 = min(r.k0 + p.k_per_split, p.k_tiles);
  r.nk = k1 - r.k0;
  return r;
}

template <int BLOCK_N>
struct Cfg {
  static constexpr int kStageB = BLOCK_N * 128;
  static constexpr int kStageBytes = kStageA + kStageB;
  static constexpr int kStages = (BLOCK_N == 256) ? 4 : 6;
  static constexpr int kTmemCols = 2 * BLOCK_N;
  static constexpr int kRedSlabBytes = 4096;  // per epilogue warp: staging of the transposed fp32 reduction (form 1)
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + 4 * kRedSlabBytes;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(kThreads, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                 const __grid_constant__ UmmaParams p) {
  using C = Cfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + C::kStages * C::kStageBytes;
  // barrier layout (8 B each): full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], then tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * C::kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * C::kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * C::kStages + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  vcd_pdl_trigger();   // see umma_pair_kernel: lets a VCD_WGRAD_OVERLAP_PREV launch behind this kernel start early

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"((uint32_t)C::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ============================== TMA producer (warp-uniform loop, one elected lane issues) ==============
    if (elect_one_sync()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      if (p.form == 0) {
        const int nt = tile % p.n_tiles;
        const PixTile pt = decode_pix(p, tile / p.n_tiles);
        const int brow0 = nt * BLOCK_N + pt.n0 * p.b_batch_rows;
        for (int tap = 0; tap < p.ntaps; ++tap) {
          for (int kc = 0; kc < p.kc_per_tap; ++kc) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t sa = smem_base + stage * C::kStageBytes;
            if (elect_one_sync()) {
              mbar_arrive_expect_tx(full_bar(stage), C::kStageBytes);
              tma_load_5d(sa, &mapA, full_bar(stage), kc * 64, act_cw(pt.w0 + p.tap_dw[tap], p.tap_plane[tap], p.a_es),
                          act_ch(pt.h0 + p.tap_dh[tap], p.tap_plane[tap], p.a_es), act_cp(p.tap_plane[tap], p.a_es), pt.n0);
              tma_load_5d(sa + kStageA, &mapB, full_bar(stage), kc * 64, brow0 + p.tap_brow[tap], 0, 0, 0);
            }
            __syncwarp();
            if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
          }
        }
      } else {
        const RedTile rt = decode_red(p, tile);
        const int tiles_per_img = p.tiles_w * p.tiles_h;
        for (int k = 0; k < rt.nk; ++k) {
          // batched (attention): pixel tiles of image `batch` only; conv wgrad: all pixel tiles
          const int pix = (p.batches > 1 ? rt.batch * tiles_per_img : 0) + rt.k0 + k;
          const PixTile pt = decode_pix(p, pix);
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * C::kStageBytes;
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(full_bar(stage), C::kStageBytes);
            const int tap0 = p.tap_pairs ? rt.tap * 2 : rt.tap;
#pragma unroll
            for (int b = 0; b < 2; ++b)
              tma_load_5d(sa + b * 8192, &mapA, full_bar(stage), rt.mt * 128 + b * 64, act_cw(pt.w0, p.tap_plane_a[tap0], p.a_es),
                          act_ch(pt.h0, p.tap_plane_a[tap0], p.a_es), act_cp(p.tap_plane_a[tap0], p.a_es), pt.n0);
#pragma unroll
            for (int b = 0; b < BLOCK_N / 64; ++b) {
              // tap_pairs: boxes 0,1 = tap 2g (channels 0-63, 64-127), boxes 2,3 = tap 2g+1 (the last odd tap is loaded
              // twice; its duplicate columns are dropped by the epilogue)
              const int tap = p.tap_pairs ? min(tap0 + (b >> 1), p.ntaps - 1) : tap0;
              const int c0 = p.tap_pairs ? (b & 1) * 64 : rt.nt * BLOCK_N + b * 64;
              tma_load_5d(sa + kStageA + b * 8192, &mapB, full_bar(stage), c0,
                          act_cw(pt.w0 + p.tap_dw[tap], p.tap_plane[tap], p.b_es),
                          act_ch(pt.h0 + p.tap_dh[tap], p.tap_plane[tap], p.b_es), act_cp(p.tap_plane[tap], p.b_es), pt.n0);
            }
          }
          __syncwarp();
          if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer (warp-uniform loop, one elected lane issues) ==================
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int nk;
      if (p.form == 0) nk = p.ntaps * p.kc_per_tap;
      else nk = decode_red(p, tile).nk;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
      for (int k = 0; k < nk; ++k) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * C::kStageBytes;
        const uint32_t sb = sa + kStageA;
        if (elect_one_sync()) {
          const uint32_t a_lo = ((sa >> 4) & 0x3FFFu) | (p.a_lbo << 16);
          const uint32_t b_lo = ((sb >> 4) & 0x3FFFu) | (p.b_lbo << 16);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            umma_bf16(d_tmem, ((uint64_t)p.a_desc_hi << 32) | (a_lo + j * p.a_kstep),
                      ((uint64_t)p.b_desc_hi << 32) | (b_lo + j * p.b_kstep), p.idesc, (k | j) != 0 ? 1u : 0u);
          umma_commit(empty_bar(stage));  // stage reusable once these MMAs have read it
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
      }
      if (elect_one_sync()) umma_commit(tfull_bar(acc));  // accumulator complete
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else if (warp >= 4) {
    // ============================== epilogue ==============================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
      if (p.form == 0) {
        const int nt = tile % p.n_tiles;
        const PixTile pt = decode_pix(p, tile / p.n_tiles);
        const int wi = row % p.tile_w;
        const int rq = row / p.tile_w;
        const int hi = rq % p.tile_h;
        const int ni = rq / p.tile_h;
        const int w = pt.w0 + wi, h = pt.h0 + hi, n = pt.n0 + ni;
        const bool valid = (w < p.W) && (h < p.H) && (n < p.Nimg);
        const long long off = (long long)n * p.out_sn + (long long)h * p.out_sh + (long long)w * p.out_sw + nt * BLOCK_N;
#pragma unroll 1
        for (int ch = 0; ch < BLOCK_N / 32; ++ch) {
          uint32_t r[32];
          tmem_ld32(taddr + ch * 32, r);
          tmem_wait_ld();
          if (valid) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
            if (p.bias) {
              const float* bp = p.bias + nt * BLOCK_N + ch * 32;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (nt * BLOCK_N + ch * 32 + j < p.Nout) v[j] += __ldg(bp + j);
            }
            if (p.residual) {
              const bf16* rp = p.residual + off + ch * 32;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                float f[8];
                if (nt * BLOCK_N + ch * 32 + g * 8 + 8 > p.Nout) continue;
                unpack8(ld8(rp + g * 8), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[g * 8 + j] += f[j];
              }
            }
            bf16* op = p.out + off + ch * 32;
            const int col0 = nt * BLOCK_N + ch * 32;
            if ((p.Nout & 7) == 0) {
#pragma unroll
              for (int g = 0; g < 4; ++g)
                if (col0 + g * 8 < p.Nout) st8(op + g * 8, pack8(v + g * 8));
            } else {  // narrow outputs (e.g. decoder.conv_out, 3 channels): scalar stores
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.Nout) op[j] = __float2bfloat16_rn(v[j]);
            }
          }
        }
      } else {
        const RedTile rt = decode_red(p, tile);
#pragma unroll 1
        for (int ch = 0; ch < BLOCK_N / 32; ++ch) {
          uint32_t r[32];
          tmem_ld32(taddr + ch * 32, r);
          tmem_wait_ld();
          // tap_pairs: columns [0,128) belong to tap 2g, [128,256) to tap 2g+1 (Nout = 128)
          const int tap = p.tap_pairs ? rt.tap * 2 + (ch >> 2) : rt.tap;
          const int col0 = p.tap_pairs ? (ch & 3) * 32 : rt.nt * BLOCK_N + ch * 32;
          float* op = p.acc + (((long long)rt.batch * p.ntaps + tap) * p.Mout + rt.mt * 128 + row) * (long long)p.Nout + col0;
          if ((p.Nout & 3) == 0 && col0 + 32 <= p.Nout) {
            if (tap < p.ntaps) {  // warp-uniform
              const int row0 = rt.mt * 128 + q * 32;
              float* op0 = p.acc + (((long long)rt.batch * p.ntaps + tap) * p.Mout + row0) * (long long)p.Nout + col0;
              red_chunk_32x32(r, bar_base + 256u + (uint32_t)q * C::kRedSlabBytes, op0, p.Nout, p.Mout - row0, lane);
            }
          } else if (rt.mt * 128 + row < p.Mout && tap < p.ntaps) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.Nout) atomicAdd(op + j, __uint_as_float(r[j]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (p.release_arrive) mbar_arrive(tempty_bar(acc)); else mbar_arrive_relaxed(tempty_bar(acc)); }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::kTmemCols)
                 : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

}  // namespace

int make_act_map(CUtensorMap* m, const void* base, int C, int W, int H, int P, int N, int box_c, int box_w, int box_h,
                 int box_n, int es) {
  // the encode is a driver-API call: autograd's backward threads may not have bound the primary context yet
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(nullptr);
    ctx_bound = true;
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    vcd_set_error("cuTensorMapEncodeTiled entry point unavailable (driver too old?)");
    return -4;
  }
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)P, (cuuint64_t)N};
  cuuint64_t s1 = (cuuint64_t)C * 2, s2 = s1 * W, s3 = s2 * H, s4 = s3 * P;
  cuuint64_t strides[4] = {s1, s2, s3, s4};
  cuuint32_t box[5] = {(cuuint32_t)box_c, (cuuint32_t)(box_w * es), (cuuint32_t)(box_h * es), 1u, (cuuint32_t)box_n};
  cuuint32_t estr[5] = {1, (cuuint32_t)es, (cuuint32_t)es, 1, 1};
  if (box_w * es > 256 || box_h * es > 256) {
    vcd_set_error("TMA map: box %d x %d with element stride %d exceeds 256", box_w, box_h, es);
    return -4;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (s1 & 15)) {
    vcd_set_error("TMA map: base/stride not 16-byte aligned (C=%d)", C);
    return -4;
  }
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    vcd_set_error("cuTensorMapEncodeTiled failed (%d) dims=(%d,%d,%d,%d,%d) box=(%d,%d,%d,1,%d)", (int)r, C, W, H, P, N,
                  box_c, box_w, box_h, box_n);
    return -4;
  }
  return 0;
}

int umma_launch(const CUtensorMap& mapA, const CUtensorMap& mapB, const UmmaParams& p_in, int block_n, cudaStream_t st,
                bool overlap_prev) {
  UmmaParams p = p_in;
  static int rel = -1;
  if (rel < 0) { const char* e = getenv("VCD_GEMM_RELEASE"); rel = (e && e[0] == '1') ? 1 : 0; }
  p.release_arrive = rel;
  int grid = p.total_tiles < vcd_num_sms() ? p.total_tiles : vcd_num_sms();
  if (grid <= 0) return 0;
  if (block_n == 256) {
    bool& attr = *vcd_device_once(0);
    if (!attr) {
      VCD_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg<256>::kSmemBytes));
      attr = true;
    }
    VCD_CUDA(vcd_launch(umma_gemm_kernel<256>, grid, kThreads, Cfg<256>::kSmemBytes, st, overlap_prev, mapA, mapB, p));
  } else if (block_n == 128) {
    bool& attr = *vcd_device_once(1);
    if (!attr) {
      VCD_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg<128>::kSmemBytes));
      attr = true;
    }
    VCD_CUDA(vcd_launch(umma_gemm_kernel<128>, grid, kThreads, Cfg<128>::kSmemBytes, st, overlap_prev, mapA, mapB, p));
  } else {
    vcd_set_error("umma_launch: BLOCK_N %d unsupported", block_n);
    return -1;
  }
  VCD_LAUNCH_CHECK();
  return 0;
}
