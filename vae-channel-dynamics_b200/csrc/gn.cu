// gn.cu — GroupNorm(32,C,eps)[+SiLU] forward/backward on bf16 NHWC with fused per-channel
// activation statistics (HBM-bound kernels; 16-byte vector access, fp32 math).
//
// Replaces [upstream] torch.nn.GroupNorm + F.silu reached from sdxl_vae_wrapper.py:60,71 and the
// hook arithmetic of src/tracking/monitor.py:64-75 (|t|.mean(dim=[0,2,3]) etc.), which in the
// reference is a second full read of the tensor plus a blocking D2H per hook.
//
// Thread mapping (all kernels): a pixel row of C channels is V = C/8 16-byte vectors; a 256-thread
// block covers PL = 256/V pixels per iteration; thread (pl, v) walks pixels pl, pl+PL, ... of its
// block's pixel range inside ONE image (blockIdx.y = n), so every warp reads contiguous 512 B.
#include <stdlib.h>

#include "common.cuh"
#include "conv_dispatch.h"

namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;
// cp.async staging depth (stages of 2 pixels): bytes in flight per thread = stages x 2 x streams x 16
constexpr int kApplyStages = 4, kReduceStages = 3, kApplyBwdStages = 3, kApplyBwdStagesRes = 2;

struct Map {
  int V, PL, v, pl, c0;
  int64_t p_begin, p_end;
};
__device__ __forceinline__ Map make_map(int C, int HW, int ppb) {  // ppb = pixels per block (host: gn_launch_shape)
  Map m;
  m.V = C >> 3;
  m.PL = kThreads / m.V;
  m.v = threadIdx.x % m.V;
  m.pl = threadIdx.x / m.V;
  m.c0 = m.v * 8;
  m.p_begin = (int64_t)blockIdx.x * ppb;
  m.p_end = min(m.p_begin + ppb, (int64_t)HW);
  return m;
}


// Block-wide per-channel reduction without atomics: thread (pl, v) deposits its 8 channels x K partials at
// red[k][j][pl][v] (v fastest: conflict-free), then one thread per channel sums the PL pixel lanes.
template <int K>
__device__ __forceinline__ void deposit(float* red, const Map& m, const float (*vals)[8]) {
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) red[((k * 8 + j) * m.PL + m.pl) * m.V + m.v] = vals[k][j];
}
__device__ __forceinline__ float lane_sum(const float* red, const Map& m, int k, int c) {
  const int v = c >> 3, j = c & 7;
  const float* p = red + ((k * 8 + j) * m.PL) * m.V + v;
  float t = 0.f;
  for (int q = 0; q < m.PL; ++q) t += p[q * m.V];
  return t;
}
__device__ __forceinline__ float lane_max(const float* red, const Map& m, int k, int c) {
  const int v = c >> 3, j = c & 7;
  const float* p = red + ((k * 8 + j) * m.PL) * m.V + v;
  float t = 0.f;
  for (int q = 0; q < m.PL; ++q) t = fmaxf(t, p[q * m.V]);
  return t;
}

// mean / rstd of every group of image n, computed once per block (fp64 sums -> fp32) into shared memory
__device__ __forceinline__ void load_group_stats(const double* sums, int n, int G, double cnt, float eps, float* s_mean,
                                                 float* s_rstd) {
  for (int g = threadIdx.x; g < G; g += kThreads) {
    double s = sums[((int64_t)n * G + g) * 2], q = sums[((int64_t)n * G + g) * 2 + 1];
    double mu = s / cnt;
    double var = q / cnt - mu * mu;
    if (var < 0.0) var = 0.0;
    s_mean[g] = (float)mu;
    s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
}
constexpr int kMaxGroups = 64;

// ---------------------------------------------------------------- cp.async staging
// ncu's source page on the register-staged kernels: 55-70 % of all warp-stall samples sit on the first use of a loaded
// vector (long scoreboard) with the issue slots 34-54 % busy and DRAM at 67-76 % — latency-bound on bytes in flight, and
// the registers (80 per thread for 3 resident blocks) allow only 64-96 bytes per thread.  cp.async (LDGSTS) keeps the loads
// out of the register file: every thread streams its 16-byte vectors of NS tensors through its OWN shared-memory slots,
// ST stages of U pixels deep (2-3x the bytes in flight), and reads a slot back only after cp.async.wait_group has retired
// the group that filled it — no block-level synchronisation, each thread only ever reads what it requested itself.
__device__ __forceinline__ uint32_t gn_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int NS, int U, int ST>
struct Pipe {
  static constexpr int kBytes = ST * U * NS * kThreads * 16;
  uint32_t base;
  __device__ __forceinline__ explicit Pipe(void* smem) : base(gn_smem_u32(smem) + threadIdx.x * 16u) {}
  __device__ __forceinline__ uint32_t slot(int st, int u, int s) const {
    return base + (uint32_t)(((st * U + u) * NS + s) * kThreads) * 16u;
  }
  __device__ __forceinline__ void issue(int st, int u, int s, const void* g) const {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot(st, u, s)), "l"(g) : "memory");
  }
  __device__ __forceinline__ void commit() const { asm volatile("cp.async.commit_group;" ::: "memory"); }
  __device__ __forceinline__ void wait_oldest() const { asm volatile("cp.async.wait_group %0;" ::"n"(ST - 1) : "memory"); }
  __device__ __forceinline__ void drain() const { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
  __device__ __forceinline__ bf16x8 read(int st, int u, int s) const {
    bf16x8 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.u.x), "=r"(v.u.y), "=r"(v.u.z), "=r"(v.u.w)
                 : "r"(slot(st, u, s))
                 : "memory");
    return v;
  }
};
// Walks the thread's pixels (m.p_begin + m.pl, step m.PL) through the pipe.  `src(s)` = base pointer of stream s (already
// offset to the thread's channels), `body(p, v)` consumes the NS vectors of pixel p.
template <int NS, int U, int ST, typename Src, typename Body>
__device__ __forceinline__ void stream_pixels(void* smem, const Map& m, int C, Src src, Body body) {
  Pipe<NS, U, ST> pipe(smem);
  const int64_t stride = (int64_t)m.PL;
  int64_t pq = m.p_begin + m.pl;          // next pixel to request
#pragma unroll
  for (int st = 0; st < ST; ++st) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (pq < m.p_end) {
#pragma unroll
        for (int s2 = 0; s2 < NS; ++s2) pipe.issue(st, u, s2, src(s2) + pq * C);
      }
      pq += stride;
    }
    pipe.commit();
  }
  int st = 0;
  for (int64_t p0 = m.p_begin + m.pl; p0 < m.p_end; p0 += (int64_t)U * stride) {
    pipe.wait_oldest();
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t p = p0 + (int64_t)u * stride;
      if (p < m.p_end) {
        bf16x8 v[NS];
#pragma unroll
        for (int s2 = 0; s2 < NS; ++s2) v[s2] = pipe.read(st, u, s2);
        body(p, v);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {          // refill the stage just consumed
      if (pq < m.p_end) {
#pragma unroll
        for (int s2 = 0; s2 < NS; ++s2) pipe.issue(st, u, s2, src(s2) + pq * C);
      }
      pq += stride;
    }
    pipe.commit();
    st = (st + 1 == ST) ? 0 : st + 1;
  }
  pipe.drain();
}

// All hot loops below use packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2, common.cuh): a thread's 8 channels are 4 pairs.
// SiLU and its derivative go through ONE tanh per element on the half-argument u = y/2 (constants pre-halved):
//     silu(y)  = y * sigmoid(y)            = u * (1 + t)                 t = tanh(u)
//     silu'(y) = s * (1 + y * (1 - s))     = 0.5 * (1 + t + u * (1 - t*t))
// Latency is hidden by occupancy (3 blocks of 256 threads per SM, 2 pixels = up to 96 bytes in flight per thread), not by
// register double buffering: the packed math leaves these kernels short of issue pressure, so more resident warps pay.

// five per-channel statistics of one value pair stream: sum, sum of squares, sum |.|, max |.|, #(|.| < tau)
struct Stat5 {
  f32x2 s[4], q[4], sa[4], mx[4], nz[4];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int j = 0; j < 4; ++j) s[j] = q[j] = sa[j] = mx[j] = nz[j] = 0ull;
  }
  // LINEAR: sum and sum of squares are NOT accumulated — the caller derives them from another Stat5 of an affinely
  // related stream (y = A x + B per thread-channel: sum y = A sum x + B n, sum y^2 = A^2 sum x^2 + 2AB sum x + B^2 n)
  template <bool NZ, bool LINEAR = false>
  __device__ __forceinline__ void add(int j, f32x2 y, float tau) {
    const f32x2 a = abs2(y);
    if (!LINEAR) {
      s[j] = add2(s[j], y);
      q[j] = fma2(y, y, q[j]);
    }
    sa[j] = add2(sa[j], a);
    mx[j] = max2(mx[j], a);
    if (NZ) {   // the near-zero count is compiled in only when a threshold is configured (tau > 0)
      float lo, hi;
      upk2(a, lo, hi);
      nz[j] = add2(nz[j], pk2(lo < tau ? 1.f : 0.f, hi < tau ? 1.f : 0.f));
    }
  }
  __device__ __forceinline__ void to_acc(float (*acc)[8]) const {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      upk2(s[j], acc[0][2 * j], acc[0][2 * j + 1]);
      upk2(q[j], acc[1][2 * j], acc[1][2 * j + 1]);
      upk2(sa[j], acc[2][2 * j], acc[2][2 * j + 1]);
      upk2(mx[j], acc[3][2 * j], acc[3][2 * j + 1]);
      upk2(nz[j], acc[4][2 * j], acc[4][2 * j + 1]);
    }
  }
};
// block-level reduction of a Stat5 into the [5][C] slot (fp32 atomics: one per channel and statistic per block)
__device__ __forceinline__ void flush_stat5(const Stat5& st, float* red, const Map& m, float* __restrict__ cstats, int C) {
  float acc[5][8];
  st.to_acc(acc);
  deposit<5>(red, m, acc);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kThreads) {
    atomicAdd(&cstats[0 * C + c], lane_sum(red, m, 0, c));
    atomicAdd(&cstats[1 * C + c], lane_sum(red, m, 1, c));
    atomicAdd(&cstats[2 * C + c], lane_sum(red, m, 2, c));
    atomic_max_nonneg(&cstats[3 * C + c], lane_max(red, m, 3, c));
    const float z = lane_sum(red, m, 4, c);
    if (z != 0.f) atomicAdd(&cstats[4 * C + c], z);
  }
  __syncthreads();
}

// ---------------------------------------------------------------- pass 1: group sums (+ input stats)
template <bool STATS, bool NZ>
__global__ void __launch_bounds__(kThreads, STATS ? 2 : 3) gn_stats_kernel(const bf16* __restrict__ x, double* __restrict__ sums,
                                                               float* __restrict__ cstats, float near_zero, int HW,
                                                               int C, int G, int ppb) {
  extern __shared__ float red[];  // [K][8][PL][V] with K = 2 (5 when STATS), then [C][2] channel sums
  constexpr int K = STATS ? 5 : 2;
  const int n = blockIdx.y;
  Map m = make_map(C, HW, ppb);
  Stat5 st;
  st.init();
  const bf16* xb = x + (int64_t)n * HW * C + m.c0;
  stream_pixels<1, 2, kApplyStages>(red, m, C, [&](int) { return xb; }, [&](int64_t, const bf16x8* v) {
    f32x2 f[4];
    unpack8x(v[0], f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (STATS) {
        st.add<NZ>(j, f[j], near_zero);
      } else {
        st.s[j] = add2(st.s[j], f[j]);
        st.q[j] = fma2(f[j], f[j], st.q[j]);
      }
    }
  });
  __syncthreads();   // the reduction below reuses the staging memory
  float acc[5][8];
  st.to_acc(acc);
  deposit<K>(red, m, acc);
  __syncthreads();
  float* chan = red + K * 8 * kThreads;  // [C][2]
  for (int c = threadIdx.x; c < C; c += kThreads) {
    float s = lane_sum(red, m, 0, c), q = lane_sum(red, m, 1, c);
    chan[c * 2] = s;
    chan[c * 2 + 1] = q;
    if (STATS) {
      atomicAdd(&cstats[0 * C + c], s);
      atomicAdd(&cstats[1 * C + c], q);
      atomicAdd(&cstats[2 * C + c], lane_sum(red, m, 2, c));
      atomic_max_nonneg(&cstats[3 * C + c], lane_max(red, m, 3, c));
      const float z = lane_sum(red, m, 4, c);
      if (z != 0.f) atomicAdd(&cstats[4 * C + c], z);
    }
  }
  __syncthreads();
  const int D = C / G;
  for (int g = threadIdx.x; g < G; g += kThreads) {
    double gs = 0.0, gq = 0.0;
    for (int j = 0; j < D; ++j) {
      gs += (double)chan[(g * D + j) * 2];
      gq += (double)chan[(g * D + j) * 2 + 1];
    }
    atomicAdd(&sums[((int64_t)n * G + g) * 2], gs);
    atomicAdd(&sums[((int64_t)n * G + g) * 2 + 1], gq);
  }
}

// per-thread affine constants of the thread's 8 channels: y = a x + b (ACT: pre-halved, u = y/2 = ka x + kb)
template <bool ACT>
__device__ __forceinline__ void load_affine(const void* gamma, const void* beta, int pdt, const float* s_mean,
                                            const float* s_rstd, int c0, int D, f32x2* ka, f32x2* kb) {
  const float h = ACT ? 0.5f : 1.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float a[2], b[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = c0 + 2 * j + e;
      a[e] = s_rstd[c / D] * load_param(gamma, pdt, c);
      b[e] = load_param(beta, pdt, c) - s_mean[c / D] * a[e];
    }
    ka[j] = pk2(h * a[0], h * a[1]);
    kb[j] = pk2(h * b[0], h * b[1]);
  }
}

// ---------------------------------------------------------------- pass 2: normalise (+ input / output stats, SiLU)
// SIN : per-channel statistics of the INPUT x  (capture_point "input" of this GroupNorm == "output" of the layer that
//       produced x, e.g. encoder.conv_in) — the pass reads x anyway;
// SOUT: per-channel statistics of y = gamma*xhat + beta BEFORE SiLU (capture_point "output").
template <bool ACT, bool SIN, bool SOUT, bool NZ>
__global__ void __launch_bounds__(kThreads, (SIN || SOUT) ? 2 : 3) gn_apply_kernel(const bf16* __restrict__ x, const double* __restrict__ sums,
                                                               const void* __restrict__ gamma, const void* __restrict__ beta,
                                                               int pdt, bf16* __restrict__ out, float* __restrict__ cstats_in,
                                                               float* __restrict__ cstats_out, float near_zero, float eps,
                                                               int HW, int C, int G, int ppb) {
  extern __shared__ float sm[];  // cp.async staging, then [5][8][PL][V] for the statistics reduction
  const int n = blockIdx.y;
  Map m = make_map(C, HW, ppb);
  const int D = C / G;
  const double cnt = (double)D * (double)HW;
  __shared__ float s_mean[kMaxGroups], s_rstd[kMaxGroups];
  load_group_stats(sums, n, G, cnt, eps, s_mean, s_rstd);
  f32x2 ka[4], kb[4];
  load_affine<ACT>(gamma, beta, pdt, s_mean, s_rstd, m.c0, D, ka, kb);
  Stat5 sin, sout;
  if (SIN) sin.init();
  if (SOUT) sout.init();
  const int64_t base = (int64_t)n * HW * C + m.c0;
  float npix = 0.f;   // pixels this thread processed (SIN && SOUT: output sums are derived from the input sums)
  const bf16* xb = x + base;
  stream_pixels<1, 2, kApplyStages>(sm, m, C, [&](int) { return xb; }, [&](int64_t p, const bf16x8* v) {
    f32x2 f[4];
    unpack8x(v[0], f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (SIN) sin.add<NZ>(j, f[j], near_zero);
      const f32x2 w = fma2(ka[j], f[j], kb[j]);       // ACT: u = y/2, else y
      if (SOUT) sout.add<NZ, SIN && SOUT>(j, ACT ? add2(w, w) : w, near_zero);
      f[j] = ACT ? fma2(w, tanh2(w), w) : w;           // silu(y) = u + u*tanh(u)
    }
    st8(out + base + p * C, pack8x(f));
    if (SIN && SOUT) npix += 1.f;
  });
  if (SIN || SOUT) __syncthreads();   // the statistics reduction below reuses the staging memory
  if (SIN && SOUT) {
    const f32x2 two = dup2(ACT ? 2.f : 1.f), n2 = dup2(npix);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const f32x2 A = mul2(ka[j], two), B = mul2(kb[j], two);          // y = A x + B (ka, kb are halved under ACT)
      sout.s[j] = fma2(A, sin.s[j], mul2(B, n2));
      const f32x2 AB2 = mul2(mul2(A, B), dup2(2.f));
      sout.q[j] = fma2(mul2(A, A), sin.q[j], fma2(AB2, sin.s[j], mul2(mul2(B, B), n2)));
    }
  }
  if (SIN) flush_stat5(sin, sm, m, cstats_in, C);
  if (SOUT) flush_stat5(sout, sm, m, cstats_out, C);
}

// g * silu'(y) for one pair: gg = (g/2) * (1 + t + u (1 - t^2)),  u = y/2, t = tanh(u)
__device__ __forceinline__ f32x2 silu_grad_times(f32x2 g, f32x2 u) {
  const f32x2 t = tanh2(u);
  const f32x2 v = fma2(neg2(t), t, dup2(1.f));
  const f32x2 r = fma2(u, v, t);
  const f32x2 gh = mul2(g, dup2(0.5f));
  return fma2(gh, r, gh);
}

// ---------------------------------------------------------------- backward pass 1: ds/db per (n, c)
template <bool ACT>
__global__ void __launch_bounds__(kThreads, 3) gn_bwd_reduce_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dout,
                                                                    const double* __restrict__ sums,
                                                                    const void* __restrict__ gamma,
                                                                    const void* __restrict__ beta, int pdt,
                                                                    float* __restrict__ dsdb, float eps, int HW, int C, int G,
                                                                    int ppb) {
  extern __shared__ float sm[];  // [2][8][PL][V]
  const int n = blockIdx.y;
  Map m = make_map(C, HW, ppb);
  const int D = C / G;
  const double cnt = (double)D * (double)HW;
  __shared__ float s_mean[kMaxGroups], s_rstd[kMaxGroups];
  load_group_stats(sums, n, G, cnt, eps, s_mean, s_rstd);
  f32x2 ka[4], kb[4], ds[4], db[4];
  load_affine<ACT>(gamma, beta, pdt, s_mean, s_rstd, m.c0, D, ka, kb);
#pragma unroll
  for (int j = 0; j < 4; ++j) ds[j] = db[j] = 0ull;
  const int64_t base = (int64_t)n * HW * C + m.c0;
  const bf16 *xb = x + base, *gb = dout + base;
  stream_pixels<2, 2, kReduceStages>(sm, m, C, [&](int s2) { return s2 == 0 ? xb : gb; }, [&](int64_t, const bf16x8* v) {
    f32x2 f[4], g[4];
    unpack8x(v[0], f);
    unpack8x(v[1], g);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const f32x2 gg = ACT ? silu_grad_times(g[j], fma2(ka[j], f[j], kb[j])) : g[j];
      ds[j] = fma2(gg, f[j], ds[j]);
      db[j] = add2(db[j], gg);
    }
  });
  __syncthreads();   // the reduction below reuses the staging memory
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    upk2(ds[j], acc[0][2 * j], acc[0][2 * j + 1]);
    upk2(db[j], acc[1][2 * j], acc[1][2 * j + 1]);
  }
  deposit<2>(sm, m, acc);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kThreads) {
    atomicAdd(&dsdb[((int64_t)n * C + c) * 2], lane_sum(sm, m, 0, c));
    atomicAdd(&dsdb[((int64_t)n * C + c) * 2 + 1], lane_sum(sm, m, 1, c));
  }
}

// ---------------------------------------------------------------- backward pass 2: dx
//   dx = a * gg + c2 * x + c3 (+ dres)      gg = dout * silu'(y) (ACT) | dout
//   ACT: a * gg = (a/2) g (1 + r) = k + k r  with k = ka * g  — the SiLU factor is folded into the final FFMA2
template <bool ACT, bool HAS_RES, bool WIDE>
__global__ void __launch_bounds__(kThreads, 3) gn_bwd_apply_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dout,
                                                                   const double* __restrict__ sums,
                                                                   const void* __restrict__ gamma,
                                                                   const void* __restrict__ beta, int pdt,
                                                                   const float* __restrict__ dsdb, bf16* __restrict__ dx,
                                                                   const bf16* __restrict__ dres,
                                                                   float* __restrict__ colsum, void* __restrict__ dgamma,
                                                                   void* __restrict__ dbeta, int N, float eps,
                                                                   int HW, int C, int G, int ppb) {
  extern __shared__ float sm[];  // [1][8][PL][V] when colsum
  const int n = blockIdx.y;
  Map m = make_map(C, HW, ppb);
  const int D = C / G;
  const double cnt = (double)D * (double)HW;
  __shared__ float s_mean[kMaxGroups], s_rstd[kMaxGroups];
  load_group_stats(sums, n, G, cnt, eps, s_mean, s_rstd);
  // c2 / c3 are per-GROUP constants.  WIDE (D = C/G >= 4 and a multiple of 4, every layer of the VAE): pairs 0-1 and pairs
  // 2-3 of the thread's 8 channels each share one value (8 registers instead of 16); otherwise one value per pair (D even)
  f32x2 ka[4], kb[4], c2[4], c3[4], cs[4];
  load_affine<ACT>(gamma, beta, pdt, s_mean, s_rstd, m.c0, D, ka, kb);
  {
    int prev_g = -1;
    float pc2 = 0.f, pc3 = 0.f, t2[8], t3[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = m.c0 + j, g = c / D;
      if (g != prev_g) {
        const float mean = s_mean[g], rstd = s_rstd[g];
        float S1 = 0.f, S2 = 0.f;
        for (int k = 0; k < D; ++k) {
          const int cc = g * D + k;
          const float gm = load_param(gamma, pdt, cc);
          S1 += gm * dsdb[((int64_t)n * C + cc) * 2];
          S2 += gm * dsdb[((int64_t)n * C + cc) * 2 + 1];
        }
        const float inv = 1.f / (float)cnt;
        pc2 = (S2 * mean - S1) * rstd * rstd * rstd * inv;
        pc3 = -pc2 * mean - S2 * rstd * inv;
        prev_g = g;
      }
      t2[j] = pc2;
      t3[j] = pc3;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int src = WIDE ? (j >> 1) * 4 : 2 * j;
      c2[j] = dup2(t2[src]);
      c3[j] = dup2(t3[src]);
      cs[j] = 0ull;
    }
  }
  const int64_t base = (int64_t)n * HW * C + m.c0;
  const bf16 *xb = x + base, *gb = dout + base, *rb = HAS_RES ? dres + base : nullptr;
  stream_pixels<HAS_RES ? 3 : 2, 2, HAS_RES ? kApplyBwdStagesRes : kApplyBwdStages>(
      sm, m, C, [&](int s2) { return s2 == 0 ? xb : (s2 == 1 ? gb : rb); }, [&](int64_t p, const bf16x8* v) {
        f32x2 f[4], g[4], r[4];
        unpack8x(v[0], f);
        unpack8x(v[1], g);
        if (HAS_RES) unpack8x(v[HAS_RES ? 2 : 1], r);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          f32x2 e = fma2(c2[WIDE ? (j >> 1) * 2 : j], f[j], c3[WIDE ? (j >> 1) * 2 : j]);
          if (HAS_RES) e = add2(e, r[j]);
          f32x2 d;
          if (ACT) {
            const f32x2 uu = fma2(ka[j], f[j], kb[j]);
            const f32x2 t = tanh2(uu);
            const f32x2 vv = fma2(neg2(t), t, dup2(1.f));
            const f32x2 rr = fma2(uu, vv, t);
            const f32x2 k = mul2(ka[j], g[j]);
            d = fma2(k, rr, add2(e, k));
          } else {
            d = fma2(ka[j], g[j], e);
          }
          cs[j] = add2(cs[j], d);
          g[j] = d;
        }
        st8(dx + base + p * C, pack8x(g));
      });
  if (colsum) __syncthreads();   // the column-sum reduction below reuses the staging memory
  if (colsum) {
    float acc[1][8];
#pragma unroll
    for (int j = 0; j < 4; ++j) upk2(cs[j], acc[0][2 * j], acc[0][2 * j + 1]);
    deposit<1>(sm, m, acc);
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kThreads) atomicAdd(&colsum[c], lane_sum(sm, m, 0, c));
  }
  // parameter gradients (vcd_gn_param_grad's arithmetic) by the first block: they need only sums and dsdb, complete
  // before this kernel started — one launch less per layer
  if (dgamma && blockIdx.x == 0 && blockIdx.y == 0) {
    for (int c = threadIdx.x; c < C; c += kThreads) {
      const int g = c / D;
      float dg = 0.f, dbv = 0.f;
      for (int i = 0; i < N; ++i) {
        const double sg = sums[((int64_t)i * G + g) * 2], qg = sums[((int64_t)i * G + g) * 2 + 1];
        const double mu = sg / cnt;
        double var = qg / cnt - mu * mu;
        if (var < 0.0) var = 0.0;
        const float mean = (float)mu, rstd = (float)(1.0 / sqrt(var + (double)eps));
        const float ds = dsdb[((int64_t)i * C + c) * 2], db = dsdb[((int64_t)i * C + c) * 2 + 1];
        dg += (ds - mean * db) * rstd;
        dbv += db;
      }
      store_param(dgamma, pdt, c, dg);
      store_param(dbeta, pdt, c, dbv);
    }
  }
}

__global__ void gn_param_grad_kernel(const double* __restrict__ sums, const float* __restrict__ dsdb, void* dgamma,
                                     void* dbeta, int pdt, float eps, int N, int HW, int C, int G) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int D = C / G;
  const double cnt = (double)D * (double)HW;
  float dg = 0.f, dbv = 0.f;
  for (int n = 0; n < N; ++n) {
    const int g = c / D;
    double sg = sums[((int64_t)n * G + g) * 2], qg = sums[((int64_t)n * G + g) * 2 + 1];
    double mu = sg / cnt, var = qg / cnt - mu * mu;
    if (var < 0.0) var = 0.0;
    const float mean = (float)mu, rstd = (float)(1.0 / sqrt(var + (double)eps));
    float ds = dsdb[((int64_t)n * C + c) * 2], db = dsdb[((int64_t)n * C + c) * 2 + 1];
    dg += (ds - mean * db) * rstd;
    dbv += db;
  }
  store_param(dgamma, pdt, c, dg);
  store_param(dbeta, pdt, c, dbv);
}

// ab[n][c] = (a, b) with y = a * x + b the affine GroupNorm of image n, channel c (used by the conv dgrad epilogue that
// applies SiLU' and reduces the GroupNorm backward sums, umma_pair.cu)
__global__ void gn_ab_kernel(const double* __restrict__ sums, const void* __restrict__ gamma, const void* __restrict__ beta,
                             int pdt, float eps, int N, int HW, int C, int G, float* __restrict__ ab,
                             float* __restrict__ zero_dsdb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  if (zero_dsdb) zero_dsdb[2 * i] = zero_dsdb[2 * i + 1] = 0.f;
  const int n = i / C, c = i - n * C, D = C / G, g = c / D;
  const double cnt = (double)D * (double)HW;
  const double mu = sums[((int64_t)n * G + g) * 2] / cnt;
  double var = sums[((int64_t)n * G + g) * 2 + 1] / cnt - mu * mu;
  if (var < 0.0) var = 0.0;
  const float a = (float)(1.0 / sqrt(var + (double)eps)) * load_param(gamma, pdt, c);
  ab[2 * i] = a;
  ab[2 * i + 1] = load_param(beta, pdt, c) - (float)mu * a;
}

int check_shape(int C, int G) {
  int V = C / 8;
  if (C % 8 != 0 || V < 1 || V > kThreads || (kThreads % V) != 0) {
    vcd_set_error("GroupNorm kernels need C %% 8 == 0 and C/8 a divisor of 256 (got C=%d)", C);
    return -1;
  }
  if (G <= 0 || C % G != 0 || G > kMaxGroups || ((C / G) & 1)) {
    vcd_set_error("GroupNorm: need C divisible by G with an even number of channels per group (C=%d, G=%d)", C, G);
    return -1;
  }
  return 0;
}
// Launch shape: blocks of `ppb` pixels inside one image.  The grid is sized to `bps` resident blocks per SM times a
// whole number of waves (no ragged last wave), with at most ~2048 pixels-iterations per thread block so that small
// tensors still spread over all SMs.
struct GnShape {
  dim3 grid;
  int ppb;
};
GnShape gn_launch_shape(int N, int HW, int C, int bps) {
  const int PL = kThreads / (C / 8);
  const int64_t resident = (int64_t)vcd_num_sms() * bps;
  int64_t waves = 1;
  // at most 64 loop iterations (pixels per thread) per block
  while (ceil_div64((int64_t)N * HW, resident * waves) > (int64_t)PL * 64) ++waves;
  int64_t per_img = (resident * waves) / N;
  if (per_img < 1) per_img = 1;
  int64_t ppb = ceil_div64(HW, per_img);
  ppb = ceil_div64(ppb, PL) * PL;
  GnShape r;
  r.ppb = (int)ppb;
  r.grid = dim3((unsigned)ceil_div64(HW, ppb), (unsigned)N);
  return r;
}

}  // namespace

// (measured in round 1: asking for the GEMM kernels' maximum shared-memory carveout here made the step 1.4 % slower —
// the streaming kernels do profit from L1 — so no carveout preference is set)
int gn_make_ab(const double* sums, const void* gamma, const void* beta, int pdt, float eps, int N, int HW, int C, int G,
               float* ab, float* zero_dsdb, cudaStream_t st) {
  if (check_shape(C, G)) return -1;
  gn_ab_kernel<<<(N * C + 255) / 256, 256, 0, st>>>(sums, gamma, beta, pdt, eps, N, HW, C, G, ab, zero_dsdb);
  VCD_LAUNCH_CHECK();
  return 0;
}

static size_t stats_smem(int K, int C) {   // cp.async staging, reused for the [K][8][kThreads] + [C][2] floats of the reduction
  const size_t red = (size_t)(K * 8 * kThreads + 2 * C) * sizeof(float), pipe = Pipe<1, 2, kApplyStages>::kBytes;
  return red > pipe ? red : pipe;
}

extern "C" int vcd_gn_stats(const void* x, double* sums, float* chan_stats_in, float near_zero, int N, int HW, int C,
                            int G, vcd_stream_t stream) {
  return gn_stats_launch(x, sums, chan_stats_in, near_zero, N, HW, C, G, as_stream(stream), false);
}
int gn_stats_launch(const void* x, double* sums, float* chan_stats_in, float near_zero, int N, int HW, int C, int G,
                    cudaStream_t st, bool prezeroed) {
  if (check_shape(C, G)) return -1;
  if (!prezeroed) VCD_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * N * G, st));
  const GnShape sh = gn_launch_shape(N, HW, C, chan_stats_in ? 2 : 3);
  if (chan_stats_in && near_zero > 0.f)
    gn_stats_kernel<true, true><<<sh.grid, kThreads, stats_smem(5, C), st>>>(
        (const bf16*)x, sums, chan_stats_in, near_zero, HW, C, G, sh.ppb);
  else if (chan_stats_in)
    gn_stats_kernel<true, false><<<sh.grid, kThreads, stats_smem(5, C), st>>>(
        (const bf16*)x, sums, chan_stats_in, near_zero, HW, C, G, sh.ppb);
  else
    gn_stats_kernel<false, false><<<sh.grid, kThreads, stats_smem(2, C), st>>>(
        (const bf16*)x, sums, nullptr, near_zero, HW, C, G, sh.ppb);
  VCD_LAUNCH_CHECK();
  return 0;
}

extern "C" int vcd_gn_apply_fwd(const void* x, const double* sums, const void* gamma, const void* beta, int param_dtype,
                                void* out, float* chan_stats_in, float* chan_stats_out, float near_zero, float eps,
                                int act_silu, int N, int HW, int C, int G, vcd_stream_t stream) {
  if (check_shape(C, G)) return -1;
  cudaStream_t st = as_stream(stream);
  const bool sin = chan_stats_in != nullptr, sout = chan_stats_out != nullptr;
  const GnShape sh = gn_launch_shape(N, HW, C, (sin || sout) ? 2 : 3);
  size_t smem = Pipe<1, 2, kApplyStages>::kBytes;
  if ((sin || sout) && smem < 5 * 8 * kThreads * sizeof(float)) smem = 5 * 8 * kThreads * sizeof(float);
#define VCD_GN_APPLY(ACT, SIN, SOUT)                                                                                    \
  do {                                                                                                                  \
    if ((SIN || SOUT) && near_zero > 0.f)                                                                               \
      gn_apply_kernel<ACT, SIN, SOUT, (SIN || SOUT)><<<sh.grid, kThreads, smem, st>>>(                                  \
          (const bf16*)x, sums, gamma, beta, param_dtype, (bf16*)out, chan_stats_in, chan_stats_out, near_zero, eps, HW, C, G, \
          sh.ppb);                                                                                                      \
    else                                                                                                                \
      gn_apply_kernel<ACT, SIN, SOUT, false><<<sh.grid, kThreads, smem, st>>>(                                          \
          (const bf16*)x, sums, gamma, beta, param_dtype, (bf16*)out, chan_stats_in, chan_stats_out, near_zero, eps, HW, C, G, \
          sh.ppb);                                                                                                      \
  } while (0)
  const int sel = (act_silu ? 4 : 0) | (sin ? 2 : 0) | (sout ? 1 : 0);
  switch (sel) {
    case 0: VCD_GN_APPLY(false, false, false); break;
    case 1: VCD_GN_APPLY(false, false, true); break;
    case 2: VCD_GN_APPLY(false, true, false); break;
    case 3: VCD_GN_APPLY(false, true, true); break;
    case 4: VCD_GN_APPLY(true, false, false); break;
    case 5: VCD_GN_APPLY(true, false, true); break;
    case 6: VCD_GN_APPLY(true, true, false); break;
    default: VCD_GN_APPLY(true, true, true); break;
  }
#undef VCD_GN_APPLY
  VCD_LAUNCH_CHECK();
  return 0;
}

extern "C" int vcd_gn_bwd_reduce(const void* x, const void* dout, const double* sums, const void* gamma, const void* beta,
                                 int param_dtype, float* dsdb, float eps, int act_silu, int N, int HW, int C, int G,
                                 vcd_stream_t stream) {
  if (check_shape(C, G)) return -1;
  cudaStream_t st = as_stream(stream);
  const bool prezeroed = (act_silu & VCD_ACC_PREZEROED) != 0;
  act_silu &= 1;
  if (!prezeroed) VCD_CUDA(cudaMemsetAsync(dsdb, 0, sizeof(float) * 2 * N * C, st));
  const GnShape sh = gn_launch_shape(N, HW, C, 3);
  const size_t smem = Pipe<2, 2, kReduceStages>::kBytes;   // >= the [2][8][kThreads] floats of the final reduction
  VCD_CUDA(cudaFuncSetAttribute(gn_bwd_reduce_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VCD_CUDA(cudaFuncSetAttribute(gn_bwd_reduce_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (act_silu)
    gn_bwd_reduce_kernel<true><<<sh.grid, kThreads, smem, st>>>((const bf16*)x, (const bf16*)dout, sums, gamma, beta,
                                                                param_dtype, dsdb, eps, HW, C, G, sh.ppb);
  else
    gn_bwd_reduce_kernel<false><<<sh.grid, kThreads, smem, st>>>((const bf16*)x, (const bf16*)dout, sums, gamma, beta,
                                                                 param_dtype, dsdb, eps, HW, C, G, sh.ppb);
  VCD_LAUNCH_CHECK();
  return 0;
}

extern "C" int vcd_gn_bwd_apply(const void* x, const void* dout, const double* sums, const void* gamma, const void* beta,
                                int param_dtype, const float* dsdb, void* dx, const void* dres, float* dx_colsum,
                                void* dgamma, void* dbeta, float eps, int act_silu, int N, int HW, int C, int G,
                                vcd_stream_t stream) {
  if (check_shape(C, G)) return -1;
  cudaStream_t st = as_stream(stream);
  const bool prezeroed = (act_silu & VCD_ACC_PREZEROED) != 0;
  act_silu &= 1;
  if (dx_colsum && !prezeroed) VCD_CUDA(cudaMemsetAsync(dx_colsum, 0, sizeof(float) * C, st));
  const size_t smem = dres ? Pipe<3, 2, kApplyBwdStagesRes>::kBytes : Pipe<2, 2, kApplyBwdStages>::kBytes;   // >= 8 * kThreads floats
  const GnShape sh = gn_launch_shape(N, HW, C, 3);
  const bool wide = ((C / G) & 3) == 0;
#define VCD_GN_BWD(ACT, RES)                                                                                              \
  do {                                                                                                                    \
    VCD_CUDA(cudaFuncSetAttribute(gn_bwd_apply_kernel<ACT, RES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
    VCD_CUDA(cudaFuncSetAttribute(gn_bwd_apply_kernel<ACT, RES, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    if (wide)                                                                                                             \
      gn_bwd_apply_kernel<ACT, RES, true><<<sh.grid, kThreads, smem, st>>>(                                               \
          (const bf16*)x, (const bf16*)dout, sums, gamma, beta, param_dtype, dsdb, (bf16*)dx, (const bf16*)dres, dx_colsum, \
          dgamma, dbeta, N, eps, HW, C, G, sh.ppb);                                                                       \
    else                                                                                                                  \
      gn_bwd_apply_kernel<ACT, RES, false><<<sh.grid, kThreads, smem, st>>>(                                              \
          (const bf16*)x, (const bf16*)dout, sums, gamma, beta, param_dtype, dsdb, (bf16*)dx, (const bf16*)dres, dx_colsum, \
          dgamma, dbeta, N, eps, HW, C, G, sh.ppb);                                                                       \
  } while (0)
  if (act_silu) {
    if (dres) VCD_GN_BWD(true, true); else VCD_GN_BWD(true, false);
  } else {
    if (dres) VCD_GN_BWD(false, true); else VCD_GN_BWD(false, false);
  }
#undef VCD_GN_BWD
  VCD_LAUNCH_CHECK();
  return 0;
}

extern "C" int vcd_gn_param_grad(const double* sums, const float* dsdb, void* dgamma, void* dbeta, int param_dtype,
                                 float eps, int N, int HW, int C, int G, vcd_stream_t stream) {
  if (check_shape(C, G)) return -1;
  gn_param_grad_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(sums, dsdb, dgamma, dbeta, param_dtype, eps, N, HW,
                                                                       C, G);
  VCD_LAUNCH_CHECK();
  return 0;
}
