// umma_ptx.cuh — inline-PTX wrappers shared by the tcgen05 kernels (umma_gemm.cu, umma_pair.cu):
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05.mma / commit / ld, fences.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace umma {

constexpr long long kSpinLimit = 4000000000LL;  // ~2 s at 2 GHz: turn a deadlock into a trap

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive without release semantics (TMEM accumulator hand-off, see mbar_arrive_cluster_relaxed)
__device__ __forceinline__ void mbar_arrive_relaxed(uint32_t bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait with a suspend-time hint: the warp may sleep in hardware up to `ns` nanoseconds.  Without the hint a waiting
// role warp re-issues the try_wait every ~70 cycles (ncu source page: 10 M loop iterations in a 0.4 ms wgrad kernel).
// Measured on B200 (same box, bench.py): a 20 us hint changes nothing per kernel and costs 0.3-0.6 ms per step, so the
// default build does NOT use it (-DVCD_WAIT_HINT_NS=20000 enables it).
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
#ifndef VCD_WAIT_HINT_NS
#define VCD_WAIT_HINT_NS 0
#endif
  while (!(VCD_WAIT_HINT_NS ? mbar_try_wait_hint(bar, parity, VCD_WAIT_HINT_NS) : mbar_try_wait(bar, parity))) {
    if (clock64() - t0 > kSpinLimit) {
      printf("vcd umma_gemm: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x,
             bar, parity);
      __trap();
    }
  }
}

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// one lane of the (fully active, converged) warp; the role loops stay warp-uniform so that descriptors and barrier
// addresses live in uniform registers and UTCHMMA / UTMALDG need no per-instruction R2UR waterfall
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xFFFFFFFF;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


// ---------------------------------------------------------------- transposed fp32 reduction of one 32 x 32 accumulator chunk
// `r` holds this lane's row (32 fp32 columns).  The rows go through a 4 KB per-warp shared-memory slab (16-byte slots
// XOR-swizzled by the row: conflict free both ways) so that each red.global.add.v4.f32 instruction adds 4 rows x 128
// contiguous bytes — 4 L2 lines per instruction instead of 32 lines per scalar RED.
//   dst_row0: address of (row 0, column 0) of the chunk; row_stride in floats; rows_valid: rows [0, rows_valid) exist
__device__ __forceinline__ void red_chunk_32x32(const uint32_t* r, uint32_t slab, float* dst_row0, long long row_stride,
                                                int rows_valid, int lane) {
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const uint32_t a = slab + (uint32_t)lane * 128u + (uint32_t)((g ^ (lane & 7)) * 16);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(r[4 * g]), "r"(r[4 * g + 1]), "r"(r[4 * g + 2]),
                 "r"(r[4 * g + 3])
                 : "memory");
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = 4 * i + (lane >> 3), c = lane & 7;
    float x0, x1, x2, x3;
    const uint32_t a = slab + (uint32_t)row * 128u + (uint32_t)((c ^ (row & 7)) * 16);
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x0), "=f"(x1), "=f"(x2), "=f"(x3) : "r"(a) : "memory");
    if (row < rows_valid) {
      float* d = dst_row0 + (long long)row * row_stride + c * 4;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(x0), "f"(x1), "f"(x2), "f"(x3) : "memory");
    }
  }
  __syncwarp();
}

// activation-map coordinates of pixel (w, h) of parity plane `plane` = ph*2 + pw: with element stride es = 2 the map
// covers the plain NHWC tensor and the plane is the coordinate parity; with es = 1 the plane is the 4th coordinate
__device__ __forceinline__ int act_cw(int w, int plane, int es) { return es == 2 ? 2 * w + (plane & 1) : w; }
__device__ __forceinline__ int act_ch(int h, int plane, int es) { return es == 2 ? 2 * h + (plane >> 1) : h; }
__device__ __forceinline__ int act_cp(int plane, int es) { return es == 2 ? 0 : plane; }

// ---------------------------------------------------------------- CTA-pair (cta_group::2) variants
// A shared::cta address is also a valid shared::cluster address of the executing CTA; clearing bit 24 names the
// same offset in the even (leader) CTA of the pair.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on a barrier that may live in the peer CTA (address in the shared::cluster window)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// Same arrive without release semantics.  Used where the barrier hands over a TMEM accumulator stage only (epilogue ->
// MMA issuer): the TMEM reads are ordered by tcgen05.wait::ld + tcgen05.fence::before_thread_sync, no generic-memory data
// travels through the barrier, and a cluster-scope RELEASE compiles to MEMBAR.ALL.GPU + ERRBAR — the epilogue warp then
// sits until all of its global stores of the tile have been acknowledged (ncu: 12-26 % of the kernel's stall samples).
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// TMA load issued by either CTA of a pair; the bytes are counted on the LEADER's barrier
__device__ __forceinline__ void tma_load_5d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1,
                                                 int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all prior MMAs of this thread -> one arrival on the barrier at the same offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"((uint16_t)3)
      : "memory");
}

}  // namespace umma
