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

namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

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

// ---------------------------------------------------------------- pass 1: group sums (+ input stats)
template <bool STATS>
__global__ void __launch_bounds__(kThreads) gn_stats_kernel(const bf16* __restrict__ x, double* __restrict__ sums,
                                                            float* __restrict__ cstats, float near_zero, int HW,
                                                            int C, int G, int ppb) {
  extern __shared__ float red[];  // [K][8][PL][V] with K = 2 (5 when STATS), then [C][2] channel sums
  constexpr int K = STATS ? 5 : 2;
  const int n = blockIdx.y;
  Map m = make_map(C, HW, ppb);
  float acc[K][8];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[k][j] = 0.f;
  const bf16* xb = x + (int64_t)n * HW * C + m.c0;
  for (int64_t p0 = m.p_begin + m.pl; p0 < m.p_end; p0 += (int64_t)kUnroll * m.PL) {
    bf16x8 v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t p = p0 + (int64_t)u * m.PL;
      if (p < m.p_end) v[u] = ld8(xb + p * C);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if (p0 + (int64_t)u * m.PL >= m.p_end) break;
      float f[8];
      unpack8(v[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[0][j] += f[j];
        acc[1][j] += f[j] * f[j];
        if (STATS) {
          float a = fabsf(f[j]);
          acc[2][j] += a;
          acc[3][j] = fmaxf(acc[3][j], a);
          acc[4][j] += (a < near_zero) ? 1.f : 0.f;
        }
      }
    }
  }
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
      atomicAdd(&cstats[4 * C + c], lane_sum(red, m, 4, c));
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

// ---------------------------------------------------------------- pass 2: normalise (+ output stats, SiLU)
template <bool STATS>
__global__ void __launch_bounds__(kThreads) gn_apply_kernel(const bf16* __restrict__ x, const double* __restrict__ sums,
                                                            const void* __restrict__ gamma, const void* __restrict__ beta,
                                                            int pdt, bf16* __restrict__ out, float* __restrict__ cstats,
                                                            float near_zero, float eps, int act, int HW, int C, int G, int ppb) {
  extern __shared__ float sm[];  // [5][8][PL][V] when STATS
  const int n = blockIdx.y;
  Map m = make_map(C, HW, ppb);
  const int D = C / G;
  const double cnt = (double)D * (double)HW;
  __shared__ float s_mean[kMaxGroups], s_rstd[kMaxGroups];
  load_group_stats(sums, n, G, cnt, eps, s_mean, s_rstd);
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int c = m.c0 + j;
    a[j] = s_rstd[c / D] * load_param(gamma, pdt, c);
    b[j] = load_param(beta, pdt, c) - s_mean[c / D] * a[j];
  }
  float s[8], q[8], sa[8], mx[8], nz[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = sa[j] = mx[j] = nz[j] = 0.f;
  const int64_t base = (int64_t)n * HW * C + m.c0;
  for (int64_t p0 = m.p_begin + m.pl; p0 < m.p_end; p0 += (int64_t)kUnroll * m.PL) {
    bf16x8 v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t p = p0 + (int64_t)u * m.PL;
      if (p < m.p_end) v[u] = ld8(x + base + p * C);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t p = p0 + (int64_t)u * m.PL;
      if (p >= m.p_end) break;
      float f[8];
      unpack8(v[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float y = fmaf(a[j], f[j], b[j]);
        if (STATS) {
          float ab = fabsf(y);
          s[j] += y;
          q[j] += y * y;
          sa[j] += ab;
          mx[j] = fmaxf(mx[j], ab);
          nz[j] += (ab < near_zero) ? 1.f : 0.f;
        }
        f[j] = act ? silu_f(y) : y;
      }
      st8(out + base + p * C, pack8(f));
    }
  }
  if (STATS) {
    float acc[5][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[0][j] = s[j]; acc[1][j] = q[j]; acc[2][j] = sa[j]; acc[3][j] = mx[j]; acc[4][j] = nz[j]; }
    deposit<5>(sm, m, acc);
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kThreads) {
      atomicAdd(&cstats[0 * C + c], lane_sum(sm, m, 0, c));
      atomicAdd(&cstats[1 * C + c], lane_sum(sm, m, 1, c));
      atomicAdd(&cstats[2 * C + c], lane_sum(sm, m, 2, c));
      atomic_max_nonneg(&cstats[3 * C + c], lane_max(sm, m, 3, c));
      atomicAdd(&cstats[4 * C + c], lane_sum(sm, m, 4, c));
    }
  }
}

// ---------------------------------------------------------------- backward pass 1: ds/db per (n, c)
__global__ void __launch_bounds__(kThreads, 2) gn_bwd_reduce_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dout,
                                                                 const double* __restrict__ sums,
                                                                 const void* __restrict__ gamma,
                                                                 const void* __restrict__ beta, int pdt,
                                                                 float* __restrict__ dsdb, float eps, int act, int HW,
                                                                 int C, int G, int ppb) {
  extern __shared__ float sm[];  // [2][8][PL][V]
  const int n = blockIdx.y;
  Map m = make_map(C, HW, ppb);
  const int D = C / G;
  const double cnt = (double)D * (double)HW;
  __shared__ float s_mean[kMaxGroups], s_rstd[kMaxGroups];
  load_group_stats(sums, n, G, cnt, eps, s_mean, s_rstd);
  float a[8], b[8], ds[8], db[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int c = m.c0 + j;
    a[j] = s_rstd[c / D] * load_param(gamma, pdt, c);
    b[j] = load_param(beta, pdt, c) - s_mean[c / D] * a[j];
    ds[j] = db[j] = 0.f;
  }
  const int64_t base = (int64_t)n * HW * C + m.c0;
  // register double buffer: the loads of the next stage are in flight while this stage is reduced
  constexpr int U = 2;
  const int64_t S = (int64_t)U * m.PL;
  bf16x8 bx[2][U], bg[2][U];
  auto load = [&](int64_t q, bf16x8* vx, bf16x8* vg) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t pp = q + (int64_t)u * m.PL;
      if (pp < m.p_end) {
        vx[u] = ld8(x + base + pp * C);
        vg[u] = ld8(dout + base + pp * C);
      }
    }
  };
  auto compute = [&](int64_t q, const bf16x8* vx, const bf16x8* vg) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (q + (int64_t)u * m.PL >= m.p_end) break;
      float f[8], g[8];
      unpack8(vx[u], f);
      unpack8(vg[u], g);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float gg = g[j];
        if (act) gg *= silu_grad_f(fmaf(a[j], f[j], b[j]));
        ds[j] = fmaf(gg, f[j], ds[j]);
        db[j] += gg;
      }
    }
  };
  int64_t pq = m.p_begin + m.pl;
  if (pq < m.p_end) {
    load(pq, bx[0], bg[0]);
    while (true) {
      if (pq + S < m.p_end) load(pq + S, bx[1], bg[1]);
      compute(pq, bx[0], bg[0]);
      pq += S;
      if (pq >= m.p_end) break;
      if (pq + S < m.p_end) load(pq + S, bx[0], bg[0]);
      compute(pq, bx[1], bg[1]);
      pq += S;
      if (pq >= m.p_end) break;
    }
  }
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[0][j] = ds[j]; acc[1][j] = db[j]; }
  deposit<2>(sm, m, acc);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kThreads) {
    atomicAdd(&dsdb[((int64_t)n * C + c) * 2], lane_sum(sm, m, 0, c));
    atomicAdd(&dsdb[((int64_t)n * C + c) * 2 + 1], lane_sum(sm, m, 1, c));
  }
}

// ---------------------------------------------------------------- backward pass 2: dx
template <bool HAS_RES>
__global__ void __launch_bounds__(kThreads, 2) gn_bwd_apply_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dout,
                                                                const double* __restrict__ sums,
                                                                const void* __restrict__ gamma,
                                                                const void* __restrict__ beta, int pdt,
                                                                const float* __restrict__ dsdb, bf16* __restrict__ dx,
                                                                const bf16* __restrict__ dres,
                                                                float* __restrict__ colsum, void* __restrict__ dgamma,
                                                                void* __restrict__ dbeta, int N, float eps, int act,
                                                                int HW, int C, int G, int ppb) {
  extern __shared__ float sm[];  // [1][8][PL][V] when colsum
  constexpr int kU = 2;          // fewer pixels in flight than the other passes: three streams per pixel
  const int n = blockIdx.y;
  Map m = make_map(C, HW, ppb);
  float cs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) cs[j] = 0.f;
  const int D = C / G;
  const double cnt = (double)D * (double)HW;
  __shared__ float s_mean[kMaxGroups], s_rstd[kMaxGroups];
  load_group_stats(sums, n, G, cnt, eps, s_mean, s_rstd);
  float a[8], b[8], c2[8], c3[8];
  int prev_g = -1;
  float pc2 = 0.f, pc3 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int c = m.c0 + j;
    int g = c / D;
    const float mean = s_mean[g], rstd = s_rstd[g];
    a[j] = rstd * load_param(gamma, pdt, c);
    b[j] = load_param(beta, pdt, c) - mean * a[j];
    if (g != prev_g) {
      float S1 = 0.f, S2 = 0.f;
      for (int k = 0; k < D; ++k) {
        int cc = g * D + k;
        float gm = load_param(gamma, pdt, cc);
        S1 += gm * dsdb[((int64_t)n * C + cc) * 2];
        S2 += gm * dsdb[((int64_t)n * C + cc) * 2 + 1];
      }
      float inv = 1.f / (float)cnt;
      pc2 = (S2 * mean - S1) * rstd * rstd * rstd * inv;
      pc3 = -pc2 * mean - S2 * rstd * inv;
      prev_g = g;
    }
    c2[j] = pc2;
    c3[j] = pc3;
  }
  const int64_t base = (int64_t)n * HW * C + m.c0;
  // register double buffer: the loads of the next stage are in flight while this stage is computed and stored
  const int64_t S = (int64_t)kU * m.PL;
  bf16x8 bx[2][kU], bg[2][kU], br[2][kU];
  auto load = [&](int64_t q, bf16x8* vx, bf16x8* vg, bf16x8* vr) {
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t pp = q + (int64_t)u * m.PL;
      if (pp < m.p_end) {
        vx[u] = ld8(x + base + pp * C);
        vg[u] = ld8(dout + base + pp * C);
        if (HAS_RES) vr[u] = ld8(dres + base + pp * C);
      }
    }
  };
  auto compute = [&](int64_t q, const bf16x8* vx, const bf16x8* vg, const bf16x8* vr) {
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t pp = q + (int64_t)u * m.PL;
      if (pp >= m.p_end) break;
      float f[8], g[8], r[8];
      unpack8(vx[u], f);
      unpack8(vg[u], g);
      if (HAS_RES) unpack8(vr[u], r);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float gg = g[j];
        if (act) gg *= silu_grad_f(fmaf(a[j], f[j], b[j]));
        float d = fmaf(a[j], gg, fmaf(c2[j], f[j], c3[j]));
        if (HAS_RES) d += r[j];
        g[j] = d;
        cs[j] += d;
      }
      st8(dx + base + pp * C, pack8(g));
    }
  };
  int64_t pq = m.p_begin + m.pl;
  if (pq < m.p_end) {
    load(pq, bx[0], bg[0], br[0]);
    while (true) {
      if (pq + S < m.p_end) load(pq + S, bx[1], bg[1], br[1]);
      compute(pq, bx[0], bg[0], br[0]);
      pq += S;
      if (pq >= m.p_end) break;
      if (pq + S < m.p_end) load(pq + S, bx[0], bg[0], br[0]);
      compute(pq, bx[1], bg[1], br[1]);
      pq += S;
      if (pq >= m.p_end) break;
    }
  }
  if (colsum) {
    float acc[1][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[0][j] = cs[j];
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
                             int pdt, float eps, int N, int HW, int C, int G, float* __restrict__ ab) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
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
  if (G <= 0 || C % G != 0 || G > kMaxGroups) {
    vcd_set_error("GroupNorm: C=%d not divisible by G=%d", C, G);
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

// The GEMM kernels around these kernels run with the maximum shared-memory carveout (~206 KB dynamic shared memory);
// asking for the same carveout here avoids an SM-wide L1/shared-memory reconfiguration at every kernel boundary
// (these kernels stream and do not rely on L1).  VCD_CARVEOUT=0 disables it (A/B measurement).
static void prefer_max_shared_once() {
  static bool done = false;
  if (done) return;
  done = true;
  // measured on B200 (bench.py, 20 steps, A/B twice): 92.4 ms/step with the preference, 91.1 ms without — the streaming
  // kernels do profit from L1, so the preference is opt-in only
  const char* e = getenv("VCD_CARVEOUT");
  if (!(e && e[0] == '1')) return;
  const int c = cudaSharedmemCarveoutMaxShared;
  cudaFuncSetAttribute(gn_ab_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, c);
  cudaFuncSetAttribute(gn_stats_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, c);
  cudaFuncSetAttribute(gn_stats_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, c);
  cudaFuncSetAttribute(gn_apply_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, c);
  cudaFuncSetAttribute(gn_apply_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, c);
  cudaFuncSetAttribute(gn_bwd_reduce_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, c);
  cudaFuncSetAttribute(gn_bwd_apply_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, c);
  cudaFuncSetAttribute(gn_bwd_apply_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, c);
  cudaFuncSetAttribute(gn_param_grad_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, c);
}

int gn_make_ab(const double* sums, const void* gamma, const void* beta, int pdt, float eps, int N, int HW, int C, int G,
               float* ab, cudaStream_t st) {
  if (check_shape(C, G)) return -1;
  prefer_max_shared_once();
  gn_ab_kernel<<<(N * C + 255) / 256, 256, 0, st>>>(sums, gamma, beta, pdt, eps, N, HW, C, G, ab);
  VCD_LAUNCH_CHECK();
  return 0;
}

extern "C" int vcd_gn_stats(const void* x, double* sums, float* chan_stats_in, float near_zero, int N, int HW, int C,
                            int G, vcd_stream_t stream) {
  if (check_shape(C, G)) return -1;
  prefer_max_shared_once();
  cudaStream_t st = as_stream(stream);
  VCD_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * N * G, st));
  const GnShape sh = gn_launch_shape(N, HW, C, 4);
  if (chan_stats_in)
    gn_stats_kernel<true><<<sh.grid, kThreads, (5 * 8 * kThreads + 2 * C) * sizeof(float), st>>>(
        (const bf16*)x, sums, chan_stats_in, near_zero, HW, C, G, sh.ppb);
  else
    gn_stats_kernel<false><<<sh.grid, kThreads, (2 * 8 * kThreads + 2 * C) * sizeof(float), st>>>(
        (const bf16*)x, sums, nullptr, near_zero, HW, C, G, sh.ppb);
  VCD_LAUNCH_CHECK();
  return 0;
}

extern "C" int vcd_gn_apply_fwd(const void* x, const double* sums, const void* gamma, const void* beta, int param_dtype,
                                void* out, float* chan_stats_out, float near_zero, float eps, int act_silu, int N,
                                int HW, int C, int G, vcd_stream_t stream) {
  if (check_shape(C, G)) return -1;
  prefer_max_shared_once();
  cudaStream_t st = as_stream(stream);
  const GnShape sh = gn_launch_shape(N, HW, C, 3);
  if (chan_stats_out)
    gn_apply_kernel<true><<<sh.grid, kThreads, 5 * 8 * kThreads * sizeof(float), st>>>(
        (const bf16*)x, sums, gamma, beta, param_dtype, (bf16*)out, chan_stats_out, near_zero, eps, act_silu, HW, C, G, sh.ppb);
  else
    gn_apply_kernel<false><<<sh.grid, kThreads, 0, st>>>((const bf16*)x, sums, gamma, beta, param_dtype, (bf16*)out, nullptr,
                                                         near_zero, eps, act_silu, HW, C, G, sh.ppb);
  VCD_LAUNCH_CHECK();
  return 0;
}

extern "C" int vcd_gn_bwd_reduce(const void* x, const void* dout, const double* sums, const void* gamma, const void* beta,
                                 int param_dtype, float* dsdb, float eps, int act_silu, int N, int HW, int C, int G,
                                 vcd_stream_t stream) {
  if (check_shape(C, G)) return -1;
  prefer_max_shared_once();
  cudaStream_t st = as_stream(stream);
  VCD_CUDA(cudaMemsetAsync(dsdb, 0, sizeof(float) * 2 * N * C, st));
  const GnShape sh = gn_launch_shape(N, HW, C, 2);
  gn_bwd_reduce_kernel<<<sh.grid, kThreads, 2 * 8 * kThreads * sizeof(float), st>>>(
      (const bf16*)x, (const bf16*)dout, sums, gamma, beta, param_dtype, dsdb, eps, act_silu, HW, C, G, sh.ppb);
  VCD_LAUNCH_CHECK();
  return 0;
}

extern "C" int vcd_gn_bwd_apply(const void* x, const void* dout, const double* sums, const void* gamma, const void* beta,
                                int param_dtype, const float* dsdb, void* dx, const void* dres, float* dx_colsum,
                                void* dgamma, void* dbeta, float eps, int act_silu, int N, int HW, int C, int G,
                                vcd_stream_t stream) {
  if (check_shape(C, G)) return -1;
  prefer_max_shared_once();
  if (dx_colsum) VCD_CUDA(cudaMemsetAsync(dx_colsum, 0, sizeof(float) * C, as_stream(stream)));
  const size_t smem = dx_colsum ? 8 * kThreads * sizeof(float) : 0;
  const GnShape sh = gn_launch_shape(N, HW, C, 2);
  if (dres)
    gn_bwd_apply_kernel<true><<<sh.grid, kThreads, smem, as_stream(stream)>>>(
        (const bf16*)x, (const bf16*)dout, sums, gamma, beta, param_dtype, dsdb, (bf16*)dx, (const bf16*)dres, dx_colsum,
        dgamma, dbeta, N, eps, act_silu, HW, C, G, sh.ppb);
  else
    gn_bwd_apply_kernel<false><<<sh.grid, kThreads, smem, as_stream(stream)>>>(
        (const bf16*)x, (const bf16*)dout, sums, gamma, beta, param_dtype, dsdb, (bf16*)dx, nullptr, dx_colsum, dgamma, dbeta,
        N, eps, act_silu, HW, C, G, sh.ppb);
  VCD_LAUNCH_CHECK();
  return 0;
}

extern "C" int vcd_gn_param_grad(const double* sums, const float* dsdb, void* dgamma, void* dbeta, int param_dtype,
                                 float eps, int N, int HW, int C, int G, vcd_stream_t stream) {
  if (check_shape(C, G)) return -1;
  prefer_max_shared_once();
  gn_param_grad_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(sums, dsdb, dgamma, dbeta, param_dtype, eps, N, HW,
                                                                       C, G);
  VCD_LAUNCH_CHECK();
  return 0;
}
