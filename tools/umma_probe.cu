// umma_probe.cu — hardware probe (NOT part of libvcd_b200.so; tools/probe_umma.py builds it into tools/_probe.so): does a K-major SWIZZLE_128B UMMA A-descriptor work when
// its start address is shifted by whole 128-byte rows inside a TMA-written tile (start not 1024-byte aligned)?
// This is what lets ONE shared-memory input patch serve all 3x3 taps of an implicit-GEMM convolution.
//   D[128][N=64] = A[r0 : r0+128][0:64] * B[64][64]^T, A tile = 144 rows x 64 bf16 loaded by one TMA box.
#include "../vae-channel-dynamics_b200/csrc/umma_gemm.cuh"

namespace {

__device__ __forceinline__ uint32_t smem_u32p(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, float* out, int r0,
                  int base_offset_mode) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32p(raw) + 1023u) & ~1023u;
  const uint32_t sA = base;                 // 144 rows x 128 B = 18432
  const uint32_t sB = base + 18432;         // 18 KB: still 1024-byte aligned
  const uint32_t bar = sB + 8192;           // B: 64 rows x 128 B
  const uint32_t mma_bar = bar + 8;
  const uint32_t slot = bar + 16;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(raw + (slot - smem_u32p(raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mma_bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot_ptr;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(18432u + 8192u) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(sA), "l"(&mapA), "r"(bar), "r"(0), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(sB), "l"(&mapB), "r"(bar), "r"(0), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bar), "r"(0u) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a_start = sA + (uint32_t)r0 * 128u;
    uint32_t a_hi = hi;
    if (base_offset_mode) a_hi |= ((a_start >> 7) & 7u) << 17;   // descriptor bits [49,52)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    for (int j = 0; j < 4; ++j) {
      uint64_t ad = ((uint64_t)a_hi << 32) | (uint64_t)((((a_start + j * 32) >> 4) & 0x3FFFu) | (1u << 16));
      uint64_t bd = ((uint64_t)hi << 32) | (uint64_t)((((sB + j * 32) >> 4) & 0x3FFFu) | (1u << 16));
      uint32_t accum = j ? 1u : 0u;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(accum) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mma_bar) : "memory");
  }
  {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(mma_bar), "r"(0u) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[32];
  for (int ch = 0; ch < 2; ++ch) {
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + ch * 32;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + ch * 32 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
  }
}

}  // namespace

// A: bf16 [144][64], B: bf16 [64][64], out: fp32 [128][64]
extern "C" int vcd_debug_umma_shifted(const void* A, const void* B, float* out, int r0, int base_offset_mode,
                                      vcd_stream_t stream) {
  CUtensorMap mA, mB;
  int rc;
  if ((rc = make_act_map(&mA, A, 64, 144, 1, 1, 1, 64, 144, 1, 1))) return rc;
  if ((rc = make_act_map(&mB, B, 64, 64, 1, 1, 1, 64, 64, 1, 1))) return rc;
  VCD_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960));
  umma_probe_kernel<<<1, 128, 40960, as_stream(stream)>>>(mA, mB, out, r0, base_offset_mode);
  VCD_LAUNCH_CHECK();
  return 0;
}
