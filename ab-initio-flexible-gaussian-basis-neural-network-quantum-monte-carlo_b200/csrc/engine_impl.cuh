// engine_impl.cuh -- sm_100a kernels of the walker engine for one (n_elec, n_atoms) pair.
// Included by the generated per-system translation units (build/inst_N_A.cu).
//
// Thread mapping (v1): one thread = one electron configuration.  A VMC sweep evaluates
// B*(N+1) configurations, the ccECP energy B*50*N*A more, so even the smallest system fills
// the 148 SMs many times over.  Parameters (<= ~13k doubles) are staged once per CTA in
// shared memory and read as warp-wide broadcasts.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <string.h>

#include <mutex>
#include <utility>

#include "fastmath.cuh"
#include "ops_table.h"
#include "psi_core.cuh"
#include "deriv_split.cuh"

namespace aiqmc {

constexpr int kThreads = 128;       // threads per CTA for the per-configuration kernels
constexpr int kRedThreads = 256;
#ifndef AIQMC_QUAD_MINB
#define AIQMC_QUAD_MINB 1            // min resident CTAs/SM requested for the value-only kernels
#endif


__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block sum; result valid in thread 0.
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* red /* THREADS/32 doubles */) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < THREADS / 32; ++i) r += red[i];
  return r;
}

template <int NE, int NA>
__device__ __forceinline__ const double* stage_params(const double* __restrict__ params, double* sP) {
  constexpr int total = make_layout(NE, NA).total;
  for (int i = threadIdx.x; i < total; i += blockDim.x) sP[i] = params[i];
  if (threadIdx.x < kExpTab) g_exp_tab[threadIdx.x] = exp2((double)threadIdx.x * (1.0 / kExpTab));
  __syncthreads();
  return sP;
}

// ---------------------------------------------------------------------------------------
// signed_network forward / gradient / forward-Laplacian on arbitrary configurations
// ---------------------------------------------------------------------------------------
template <int NE, int NA>
__global__ void __launch_bounds__(kThreads) k_psi(AiqmcSystem sys, const double* __restrict__ params,
                                                  const double* __restrict__ pos, int64_t n_cfg,
                                                  double* __restrict__ phase, double* __restrict__ logabs) {
  extern __shared__ double sP[];
  const double* P = stage_params<NE, NA>(params, sP);
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_cfg) return;
  double x[3 * NE];
  for (int q = 0; q < 3 * NE; ++q) x[q] = pos[t * 3 * NE + q];
  double ph, la;
  Psi<NE, NA>::eval_value(sys, P, x, ph, la);
  phase[t] = ph;
  logabs[t] = la;
}

// ---------------------------------------------------------------------------------------
// VMC sweep (VMC/VMCmcstep.py:28-111), split at the two batch-global limdrift sums (quirk Q6)
// ---------------------------------------------------------------------------------------
struct SweepWs {           // carved out of the caller's workspace
  double* grad;            // (B,3N)   grad log|psi| at x1
  double* logabs1;         // (B)
  double* logabs2;         // (B,N)
  double* gnew;            // (B,N,3)  components 3i..3i+2 of grad log|psi| at x2[b,i]
  double* xprop;           // (B,N,3)  proposed position of electron i
  double* partials;        // (nblocks_max, 4)
  double* scal;            // [0]=v2(grad x1) [1]=v2(grad x2) [2]=sum x_new [3]=sum x_prop
  double* dcache;          // (DerivCache::SIZE_GRAD, B*N) structure-of-arrays derivative cache
};

__host__ __device__ inline int64_t align256(int64_t v) { return (v + 255) & ~int64_t(255); }

inline int64_t deriv_cache_doubles(int n, int a, bool lap) {
  const int qm = (3 * a + 2) > 5 ? (3 * a + 2) : 5;
  return 3 * n + 12 * n * n + 24 * n + 4 * a * n + 8 * a + 12 * n + 16 + 3 * n * qm + 2 * n * n + 8 * n + 7 * n +
         (lap ? 8 * n * n : 0);
}
// The derivative cache is filled and consumed in chunks of configurations so that it stays bounded whatever the
// batch.  A chunk is also the grid of the one-thread-per-configuration primal pass, so it must be large enough to
// fill 148 SMs: the budget is 1 GiB for small systems and grows with the per-configuration record (C6H6: 152 kB)
// up to 16 GiB of the 180 GB -- with the former flat 1 GiB a C6H6 chunk was 7 040 threads (55 CTAs) and the sweep
// took 394 instead of 262 ms at 2 368 walkers.
#ifndef AIQMC_DERIV_CACHE_MAX_GIB
#define AIQMC_DERIV_CACHE_MAX_GIB 16
#endif
inline int64_t deriv_chunk(int n, int a, bool lap) {
  const int64_t per_cfg = deriv_cache_doubles(n, a, lap) * 8;
  int64_t budget = per_cfg * 131072;
  if (budget < (int64_t(1) << 30)) budget = int64_t(1) << 30;
  if (budget > (int64_t(AIQMC_DERIV_CACHE_MAX_GIB) << 30)) budget = int64_t(AIQMC_DERIV_CACHE_MAX_GIB) << 30;
  int64_t c = budget / per_cfg;
  c &= ~int64_t(127);
  return c < 128 ? 128 : c;
}
inline int64_t deriv_cache_bytes(int n, int a, bool lap, int64_t n_cfg) {
  const int64_t c = deriv_chunk(n, a, lap);
  return align256(deriv_cache_doubles(n, a, lap) * (n_cfg < c ? n_cfg : c) * 8);
}
// rows of block partials one tangent sweep over n_cfg configurations produces (chunked launches)
inline int64_t tangent_rows(int n, int a, bool lap, int64_t n_cfg) {
  const int64_t c = deriv_chunk(n, a, lap);
  const int64_t full = n_cfg / c, rest = n_cfg - full * c;
  const int64_t groups = 3 * n <= 24 ? 1 : (3 * n + 14) / 15;          // >= tan_groups<n>()
  return (full * ((c + 31) / 32) + (rest + 31) / 32) * groups;
}
constexpr int kCoopMaxGrid = 148 * 16;     // persistent grids of the lane-per-electron kernels never exceed this
inline int64_t sweep_partial_rows(int n, int a, int64_t B) {
  const int64_t r = tangent_rows(n, a, false, B * n), r3 = (B * n + kRedThreads - 1) / kRedThreads;
  const int64_t m = r > r3 ? r : r3;
  return (m > kCoopMaxGrid ? m : kCoopMaxGrid) + 1;
}
inline int64_t psi_ws_bytes(int n, int a, int64_t n_cfg, int with_lap) {
  return deriv_cache_bytes(n, a, with_lap != 0, n_cfg) +
         (with_lap ? align256((n_cfg < deriv_chunk(n, a, true) ? n_cfg : deriv_chunk(n, a, true)) * 3 * n * 8) : 0);
}

inline int64_t sweep_ws_bytes(int n, int a, int64_t B) {
  return align256(B * 3 * n * 8) + align256(B * 8) + align256(B * n * 8) + 2 * align256(B * n * 3 * 8) +
         align256(sweep_partial_rows(n, a, B) * 4 * 8) + 256 + deriv_cache_bytes(n, a, false, B * n);
}

inline SweepWs carve_sweep_ws(void* ws, int n, int a, int64_t B) {
  char* p = (char*)ws;
  SweepWs w;
  w.grad = (double*)p; p += align256(B * 3 * n * 8);
  w.logabs1 = (double*)p; p += align256(B * 8);
  w.logabs2 = (double*)p; p += align256(B * n * 8);
  w.gnew = (double*)p; p += align256(B * n * 3 * 8);
  w.xprop = (double*)p; p += align256(B * n * 3 * 8);
  w.partials = (double*)p; p += align256(sweep_partial_rows(n, a, B) * 4 * 8);
  w.scal = (double*)p; p += 256;
  w.dcache = (double*)p;
  return w;
}

__device__ __forceinline__ double taueff_of(double v2, double tau, double acyrus) {
  return (sqrt(1.0 + 2.0 * tau * acyrus * v2) - 1.0) / (acyrus * v2);   // VMCmcstep.py:11-14
}

// ---------------------------------------------------------------------------------------
// two-pass derivatives (deriv_split.cuh)
// ---------------------------------------------------------------------------------------
struct MovedSrc {          // SRC == 1: configuration (b,i) = walker b with electron i at its proposed position
  const double* grad;      // (B,3N) grad log|psi| at x1
  const double* gauss1;    // (B,3N)
  const double* scal;      // [0] = v2 of grad(x1)
  double* xprop;           // (B,N,3) out
  double tau, acyrus;
};

// primal pass: one thread per configuration -> structure-of-arrays derivative cache dc[slot * n_cfg + cfg]
#ifndef AIQMC_PRIMAL_MINB
#define AIQMC_PRIMAL_MINB 1
#endif
template <int NE, int NA, bool LAP, int SRC>
__global__ void __launch_bounds__(kThreads, AIQMC_PRIMAL_MINB) k_primal(AiqmcSystem sys, const double* __restrict__ params,
                                                     const double* __restrict__ pos, int64_t cfg0, int64_t n_cfg,
                                                     MovedSrc ms, double* __restrict__ dc, double* __restrict__ mc,
                                                     double* __restrict__ phase, double* __restrict__ logabs) {
  extern __shared__ double sP[];
  const double* P = stage_params<NE, NA>(params, sP);
  const int64_t tl = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // index inside this chunk
  if (tl >= n_cfg) return;
  const int64_t t = cfg0 + tl;                                           // global configuration index
  double x[3 * NE];
  if (SRC == 0) {
    for (int q = 0; q < 3 * NE; ++q) x[q] = pos[t * 3 * NE + q];
  } else {
    const int64_t b = t / NE;
    const int i = (int)(t - b * NE);
    const double te = taueff_of(ms.scal[0], ms.tau, ms.acyrus);
    double xn[3];
    for (int c = 0; c < 3; ++c) {
      // g = grad_eff * tstep + gauss ; x2 = x1 + g on electron i only   (VMCmcstep.py:60-78)
      const double step = (ms.grad[b * 3 * NE + 3 * i + c] * te) * ms.tau + ms.gauss1[b * 3 * NE + 3 * i + c];
      xn[c] = step + pos[b * 3 * NE + 3 * i + c];
      ms.xprop[t * 3 + c] = xn[c];
    }
    // no dynamically indexed stores into x (nvcc 12.9 miscompiled that pattern in k_ecp_quad)
    for (int e = 0; e < NE; ++e)
      for (int c = 0; c < 3; ++c) x[3 * e + c] = (e == i) ? xn[c] : pos[b * 3 * NE + 3 * e + c];
  }
  double ph, la;
  DerivSplit<NE, NA>::template primal<LAP>(sys, P, x, dc + tl, n_cfg, mc ? mc + t * MoveCache<NE, NA>::SIZE : nullptr, ph, la);
  if (phase) phase[t] = ph;
  logabs[t] = la;
}

// gradient by the fused forward + reverse (adjoint) sweep: one thread per configuration, nothing cached in HBM.
// OUT == 0: grad (n_cfg,3N); OUT == 1: configurations are (walker, moved electron i), gnew (n_cfg,3) keeps electron
// i's components.  Always: block partial of sum g^2 over all 3N components (limdrift's batch-global v2, quirk Q6).
#ifndef AIQMC_GR_MINB
#define AIQMC_GR_MINB 1
#endif
template <int NE, int NA, int SRC, int OUT>
__global__ void __launch_bounds__(kThreads, AIQMC_GR_MINB) k_grad_reverse(AiqmcSystem sys, const double* __restrict__ params,
                                                           const double* __restrict__ pos, int64_t n_cfg, MovedSrc ms,
                                                           double* __restrict__ phase, double* __restrict__ logabs,
                                                           double* __restrict__ gout, double* __restrict__ partials,
                                                           int pcol) {
  extern __shared__ double sP[];
  __shared__ double red[kThreads / 32];
  const double* P = stage_params<NE, NA>(params, sP);
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double g2 = 0.0;
  if (t < n_cfg) {
    double x[3 * NE], g[3 * NE];
    int i = 0;
    if (SRC == 0) {
      for (int q = 0; q < 3 * NE; ++q) x[q] = pos[t * 3 * NE + q];
    } else {
      const int64_t b = t / NE;
      i = (int)(t - b * NE);
      const double te = taueff_of(ms.scal[0], ms.tau, ms.acyrus);
      double xn[3];
      for (int c = 0; c < 3; ++c) {
        // g = grad_eff * tstep + gauss ; x2 = x1 + g on electron i only   (VMCmcstep.py:60-78)
        const double step = (ms.grad[b * 3 * NE + 3 * i + c] * te) * ms.tau + ms.gauss1[b * 3 * NE + 3 * i + c];
        xn[c] = step + pos[b * 3 * NE + 3 * i + c];
        ms.xprop[t * 3 + c] = xn[c];
      }
      for (int e = 0; e < NE; ++e)      // no dynamically indexed stores into x (nvcc 12.9 miscompiled that pattern)
        for (int c = 0; c < 3; ++c) x[3 * e + c] = (e == i) ? xn[c] : pos[b * 3 * NE + 3 * e + c];
    }
    double ph, la;
    DerivSplit<NE, NA>::grad_reverse(sys, P, x, ph, la, g);
    if (phase) phase[t] = ph;
    logabs[t] = la;
    for (int q = 0; q < 3 * NE; ++q) g2 += g[q] * g[q];
    if (OUT == 0) {
      for (int q = 0; q < 3 * NE; ++q) gout[t * 3 * NE + q] = g[q];
    } else {
      for (int e = 0; e < NE; ++e)
        if (e == i)
          for (int c = 0; c < 3; ++c) gout[t * 3 + c] = g[3 * e + c];
    }
  }
  if (partials) {
    const double s = block_sum<kThreads>(g2, red);
    if (threadIdx.x == 0) partials[blockIdx.x * 4 + pcol] = s;
  }
}

// reverse sweep on the derivative cache (deriv_split.cuh: grad_reverse_cached): one thread per configuration of the
// chunk; outputs and block partials as k_grad_reverse
template <int NE, int NA, int OUT>
__global__ void __launch_bounds__(kThreads) k_reverse_cached(AiqmcSystem sys, const double* __restrict__ params,
                                                             const double* __restrict__ dc, int64_t cfg0, int64_t n_cfg,
                                                             double* __restrict__ gout, double* __restrict__ partials,
                                                             int pcol) {
  extern __shared__ double sP[];
  __shared__ double red[kThreads / 32];
  const double* P = stage_params<NE, NA>(params, sP);
  const int64_t tl = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double g2 = 0.0;
  if (tl < n_cfg) {
    double g[3 * NE];
    DerivSplit<NE, NA>::grad_reverse_cached(sys, P, dc + tl, n_cfg, g);
    for (int q = 0; q < 3 * NE; ++q) g2 += g[q] * g[q];
    const int64_t t = cfg0 + tl;
    if (OUT == 0) {
      for (int q = 0; q < 3 * NE; ++q) gout[t * 3 * NE + q] = g[q];
    } else {
      const int i = (int)(t % NE);
      for (int e = 0; e < NE; ++e)
        if (e == i)
          for (int c = 0; c < 3; ++c) gout[t * 3 + c] = g[3 * e + c];
    }
  }
  if (partials) {
    const double s = block_sum<kThreads>(g2, red);
    if (threadIdx.x == 0) partials[blockIdx.x * 4 + pcol] = s;
  }
}

// tangent pass: one thread per (configuration, electron, direction).  A CTA owns a tile of 32 consecutive
// configurations and kTanWarps<N> of the 3N (electron, direction) pairs: warp = one pair, lane = one
// configuration -> coalesced cache reads, no divergence, and the tile's cache lines are fetched from DRAM once
// and re-used by the other warps through L1/L2 (with the pairs spread over the grid the cache was re-read
// 3N times: 5.9 GB per sweep at N=4).
// OUT == 0: grad (n_cfg,3N) [+ lap_parts (3N,lap_stride)];  OUT == 1: configurations are (walker, moved
// electron i): only electron i's components are kept, gnew (n_cfg,3).
// Always: block partial of sum g^2 -> partials[block*4 + pcol]   (limdrift's batch-global v2, quirk Q6).
template <int NE> constexpr int tan_warps() { return 3 * NE <= 24 ? 3 * NE : (3 * NE % 15 == 0 ? 15 : 18); }
template <int NE> constexpr int tan_groups() { return (3 * NE + tan_warps<NE>() - 1) / tan_warps<NE>(); }

template <int NE, int NA, bool LAP, int OUT>
__global__ void __launch_bounds__(32 * tan_warps<NE>()) k_tangent(AiqmcSystem sys, const double* __restrict__ params,
                                                      const double* __restrict__ dc, int64_t cfg0, int64_t n_cfg,
                                                      double* __restrict__ gout, double* __restrict__ lap_parts,
                                                      int64_t lap_stride, double* __restrict__ partials, int pcol) {
  extern __shared__ double sP[];
  __shared__ double red[tan_warps<NE>()];
  const double* P = stage_params<NE, NA>(params, sP);
  const int64_t tile = (int64_t)blockIdx.x / tan_groups<NE>();
  const int grp = (int)(blockIdx.x - tile * tan_groups<NE>());
  const int64_t cfg = tile * 32 + (threadIdx.x & 31);
  const int ed = grp * tan_warps<NE>() + (threadIdx.x >> 5);
#ifndef AIQMC_NO_TAN_PREFETCH
  {   // pull the tile's cache rows (32 configurations x 8 B = two 128-byte lines per slot) into L1 up front:
      // the pass is otherwise bound by the latency of ~250 dependent-issue loads per thread
    constexpr int kSlots = LAP ? DerivCache<NE, NA>::SIZE_LAP : DerivCache<NE, NA>::SIZE_GRAD;
    const int64_t c0 = tile * 32;
    for (int q = threadIdx.x; q < 2 * kSlots; q += blockDim.x) {
      const int64_t c = c0 + 16 * (q & 1);
      if (c < n_cfg) asm volatile("prefetch.global.L1 [%0];" ::"l"(dc + (int64_t)(q >> 1) * n_cfg + c));
    }
  }
#endif
  double g2 = 0.0;
  if (cfg < n_cfg && ed < 3 * NE) {
    const int e = ed / 3, dir = ed - 3 * e;
    double g, l2 = 0.0;
    DerivSplit<NE, NA>::template tangent<LAP>(sys, P, dc + cfg, n_cfg, e, dir, g, l2);
    g2 = g * g;
    const int64_t gc = cfg0 + cfg;                                       // global configuration index
    if (OUT == 0) {
      gout[gc * 3 * NE + ed] = g;
      if (LAP) lap_parts[(int64_t)ed * lap_stride + gc] = l2;
    } else {
      if (e == (int)(gc % NE)) gout[gc * 3 + dir] = g;
    }
  }
  if (partials) {
    const double s = block_sum<32 * tan_warps<NE>()>(g2, red);
    if (threadIdx.x == 0) partials[blockIdx.x * 4 + pcol] = s;
  }
}

// lap[cfg0 + c] = sum_q parts[q * stride + c]  (fixed order)
static __global__ void k_sum_lap_parts(int nq, const double* __restrict__ parts, int64_t stride, int64_t n,
                                       double* __restrict__ lap) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  double s = 0.0;
  for (int q = 0; q < nq; ++q) s += parts[(int64_t)q * stride + c];
  lap[c] = s;
}

// sums `nparts` rows of partials[:, col] in a fixed order -> out
static __global__ void __launch_bounds__(kRedThreads) k_reduce_partials(const double* __restrict__ partials, int nparts,
                                                                 int col0, int ncols, double* __restrict__ out,
                                                                 int out0) {
  __shared__ double red[kRedThreads / 32];
  for (int c = 0; c < ncols; ++c) {
    double v = 0.0;
    for (int i = threadIdx.x; i < nparts; i += kRedThreads) v += partials[i * 4 + col0 + c];
    const double s = block_sum<kRedThreads>(v, red);
    if (threadIdx.x == 0) out[out0 + c] = s;
  }
}

// accept/reject (VMCmcstep.py:80-109, walkers_accept :18-25); HBM-bound, 1 thread per (b,i), 44N B/walker.
// A CTA owns kRedThreads consecutive (walker, electron) pairs: its slices of grad, gnew, xprop, pos, gauss2 (compact
// form), log|psi(x2)| and rnd are seven CONTIGUOUS blocks of HBM.  Full tiles are staged by seven TMA bulk copies
// (cp.async.bulk, SASS UBLKCP) signalled on one mbarrier -- no thread issues a global load for them -- and the
// accepted coordinates leave by a bulk store of the tile; the ragged last tile and unaligned callers take plain loads.
struct AcceptTile {
  double grad[kRedThreads * 3], gnew[kRedThreads * 3], xprop[kRedThreads * 3], pos[kRedThreads * 3], g2[kRedThreads * 3];
  double la2[kRedThreads], rnd[kRedThreads];
};
static __global__ void __launch_bounds__(kRedThreads) k_sweep_accept(int n, double* __restrict__ pos,
                                                              const double* __restrict__ gauss2,
                                                              const double* __restrict__ rnd, int64_t B, double tau,
                                                              double acyrus, int signed_ratio, int g2_compact, int use_tma,
                                                              uint8_t* __restrict__ accept,
                                                              double* __restrict__ grad_eff_old, SweepWs w) {
  __shared__ double red[kRedThreads / 32];
  __shared__ __align__(16) AcceptTile tile;
  __shared__ __align__(8) unsigned long long bar;
  const int64_t t0 = (int64_t)blockIdx.x * kRedThreads;
  const int64_t t = t0 + threadIdx.x;
  const bool staged = use_tma && g2_compact && (t0 + kRedThreads <= B * n);
  if (staged) {
    const unsigned b32 = (unsigned)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b32));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      constexpr unsigned b3 = kRedThreads * 3 * sizeof(double), b1 = kRedThreads * sizeof(double);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b32), "r"(5 * b3 + 2 * b1) : "memory");
      auto bulk = [&](void* dst, const void* src, unsigned bytes) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"(b32) : "memory");
      };
      bulk(tile.grad, w.grad + t0 * 3, b3);
      bulk(tile.gnew, w.gnew + t0 * 3, b3);
      bulk(tile.xprop, w.xprop + t0 * 3, b3);
      bulk(tile.pos, pos + t0 * 3, b3);
      bulk(tile.g2, gauss2 + t0 * 3, b3);
      bulk(tile.la2, w.logabs2 + t0, b1);
      bulk(tile.rnd, rnd + t0, b1);
    }
    __syncthreads();
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "AQ_ACC_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
        "@p bra AQ_ACC_READY;\n"
        "bra AQ_ACC_WAIT;\n"
        "AQ_ACC_READY:\n"
        "}\n" ::"r"(b32) : "memory");
  }
  double s_new = 0.0, s_prop = 0.0;
  if (t < B * n) {
    const int64_t b = t / n;
    const int i = (int)(t - b * n);
    const int lt = threadIdx.x;
    const double te_old = taueff_of(w.scal[0], tau, acyrus), te_new = taueff_of(w.scal[1], tau, acyrus);
    double tp = 0.0;
    for (int c = 0; c < 3; ++c) {
      const double ge = (staged ? tile.grad[lt * 3 + c] : w.grad[b * 3 * n + 3 * i + c]) * te_old;
      const double gn = (staged ? tile.gnew[lt * 3 + c] : w.gnew[t * 3 + c]) * te_new;
      // only the diagonal 3-blocks of the reference's (B,N,3N) array are ever read (VMCmcstep.py:86-94);
      // g2_compact: the caller passes just those, (B,N,3)
      const double g2 = staged ? tile.g2[lt * 3 + c] : (g2_compact ? gauss2[t * 3 + c] : gauss2[(b * n + i) * 3 * n + 3 * i + c]);
      const double fwd = g2 * g2;
      const double bw = g2 + (ge + gn) * tau;
      tp += exp((fwd - bw * bw) / (2.0 * tau));
      if (grad_eff_old) grad_eff_old[b * 3 * n + 3 * i + c] = ge;
    }
    const double ratio = exp((staged ? tile.la2[lt] : w.logabs2[t]) - w.logabs1[b]);
    double acc = fabs(ratio) * fabs(ratio) * tp;
    if (signed_ratio) acc *= (ratio > 0.0) ? 1.0 : (ratio < 0.0 ? -1.0 : 0.0);
    const bool ok = acc > (staged ? tile.rnd[lt] : rnd[t]);
    for (int c = 0; c < 3; ++c) {
      const double xp = staged ? tile.xprop[lt * 3 + c] : w.xprop[t * 3 + c];
      const double xo = staged ? tile.pos[lt * 3 + c] : pos[b * 3 * n + 3 * i + c];
      const double xn = ok ? xp : xo;
      if (staged) tile.pos[lt * 3 + c] = xn;
      else if (ok) pos[b * 3 * n + 3 * i + c] = xp;
      s_new += xn;
      s_prop += xp;
    }
    if (accept) accept[t] = ok ? 1 : 0;
  }
  if (staged) {                         // the tile of new positions goes back by ONE bulk store
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                   ::"l"(pos + t0 * 3), "r"((unsigned)__cvta_generic_to_shared(tile.pos)), "r"((unsigned)(kRedThreads * 3 * sizeof(double))) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
  const double a = block_sum<kRedThreads>(s_new, red);
  const double c2 = block_sum<kRedThreads>(s_prop, red);
  if (threadIdx.x == 0) { w.partials[blockIdx.x * 4 + 2] = a; w.partials[blockIdx.x * 4 + 3] = c2; }
}

// ---------------------------------------------------------------------------------------
// local energy (Energy/hamiltonian.py:236-260, Energy/pphamiltonian.py:130-190)
// ---------------------------------------------------------------------------------------
struct EnergyWs {
  double* base;      // (B)    KE + Coulomb (+ local ECP channel)
  double* logabs;    // (B)    denominator of the ECP "ratio" (quirk Q12)
  double* phase;     // (B)
  double* gnorm;     // (B,4)  Frobenius norm of each rotated point group (quirk Q14)
  double* vl;        // (B,N,A,4)  v_l(r_ia)
  double* epp;       // (B,2)  non-local energy accumulator (re, im)
  double* cache;     // (B, MoveCache::SIZE) per-walker single-electron-move cache (quadrature kernels)
  double* grad;      // (B,3N) grad log|psi|
  double* lap_parts; // (3N,B) d2 log|psi| / dx_q^2
  double* dcache;    // (DerivCache::SIZE_LAP, B) structure-of-arrays derivative cache
};

inline int64_t move_cache_doubles(int n, int a) {     // == MoveCache<n, a>::SIZE (psi_core.cuh)
  return (12 * n * n + 24 * n + 4 * a * n + 8 * a + 9 * n + 4 + 4 + 3 * n + 15 + 4 * n * a + 1) & ~(int64_t)1;
}

inline int64_t energy_ws_bytes(int n, int a, int64_t B, int with_ecp) {
  int64_t s = align256(B * 8) + 2 * align256(B * 8) + 2 * align256(B * 3 * n * 8) + deriv_cache_bytes(n, a, true, B);
  if (with_ecp) s += align256(B * move_cache_doubles(n, a) * 8);
  if (with_ecp) s += align256(B * 4 * 8) + align256(B * n * a * 4 * 8 + 64 * 8 * 8) + align256(B * 2 * 8);
  return s;
}

inline int64_t tmove_ws_bytes(int n, int a, int64_t B) {
  return energy_ws_bytes(n, a, B, 1) + align256(B * n * a * AIQMC_NQUAD * 4 * 8);
}

inline EnergyWs carve_energy_ws(void* ws, int n, int a, int64_t B, int with_ecp) {
  char* p = (char*)ws;
  EnergyWs w;
  w.base = (double*)p; p += align256(B * 8);
  w.logabs = (double*)p; p += align256(B * 8);
  w.phase = (double*)p; p += align256(B * 8);
  w.grad = (double*)p; p += align256(B * 3 * n * 8);
  w.lap_parts = (double*)p; p += align256(B * 3 * n * 8);
  w.dcache = (double*)p; p += deriv_cache_bytes(n, a, true, B);
  w.gnorm = w.vl = w.epp = w.cache = nullptr;
  if (with_ecp) {
    w.gnorm = (double*)p; p += align256(B * 4 * 8);
    w.vl = (double*)p; p += align256(B * n * a * 4 * 8 + 64 * 8 * 8);
    w.epp = (double*)p; p += align256(B * 2 * 8);
    w.cache = (double*)p;
  }
  return w;
}

static __constant__ AiqmcEcp c_ecp;   // one ECP table per translation unit (per system instantiation)

__device__ __forceinline__ int quad_group(int p) { return p < 6 ? 0 : p < 18 ? 1 : p < 26 ? 2 : 3; }

// the quadrature record at the tail of a walker's MoveCache (psi_core.cuh: MoveCache::QR)
template <int NE, int NA>
__device__ __forceinline__ void write_quad_record(double* __restrict__ mc, const double* __restrict__ x,
                                                  const double* __restrict__ rot9, const double* __restrict__ gnorm4,
                                                  double logabs, double phase, const double* __restrict__ vl) {
  using MC = MoveCache<NE, NA>;
  for (int q = 0; q < 3 * NE; ++q) mc[MC::QR + q] = x[q];
  for (int q = 0; q < 9; ++q) mc[MC::QR + 3 * NE + q] = rot9[q];
  for (int q = 0; q < 4; ++q) mc[MC::QR + 3 * NE + 9 + q] = gnorm4[q];
  mc[MC::QR + 3 * NE + 13] = logabs;
  mc[MC::QR + 3 * NE + 14] = phase;
  for (int q = 0; q < 4 * NE * NA; ++q) mc[MC::QR_VL + q] = vl[q];
}

// everything of E_L except the non-local quadrature, from the two derivative passes' outputs:
// one thread per walker
template <int NE, int NA, bool ECP>
__global__ void __launch_bounds__(kThreads) k_energy_rest(AiqmcSystem sys, const double* __restrict__ params,
                                                          const double* __restrict__ pos,
                                                          const double* __restrict__ rot, int64_t B,
                                                          double* __restrict__ e_out, EnergyWs w) {
  extern __shared__ double sP[];
  const double* P = stage_params<NE, NA>(params, sP);
  constexpr LayoutC<NE, NA> L{};
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double x[3 * NE];
  for (int q = 0; q < 3 * NE; ++q) x[q] = pos[b * 3 * NE + q];
  double g2 = 0.0, lp = 0.0;
  for (int q = 0; q < 3 * NE; ++q) {                                // fixed order: deterministic
    const double g = w.grad[b * 3 * NE + q];
    g2 += g * g;
    lp += w.lap_parts[(int64_t)q * B + b];
  }
  double e = -0.5 * lp - 0.5 * g2;                                // pphamiltonian.py:100-102
  for (int i = 0; i < NE; ++i)                                    // potential_electron_electron
    for (int j = i + 1; j < NE; ++j) {
      const double dx = x[3 * i] - x[3 * j], dy = x[3 * i + 1] - x[3 * j + 1], dz = x[3 * i + 2] - x[3 * j + 2];
      e += 1.0 / sqrt(dx * dx + dy * dy + dz * dz);
    }
  for (int a = 0; a < NA; ++a)                                    // potential_nuclear_nuclear
    for (int c = a + 1; c < NA; ++c) {
      const double dx = P[L.atoms + 3 * a] - P[L.atoms + 3 * c], dy = P[L.atoms + 3 * a + 1] - P[L.atoms + 3 * c + 1],
                   dz = P[L.atoms + 3 * a + 2] - P[L.atoms + 3 * c + 2];
      e += P[L.charges + a] * P[L.charges + c] / sqrt(dx * dx + dy * dy + dz * dz);
    }
  for (int i = 0; i < NE; ++i)
    for (int a = 0; a < NA; ++a) {
      const double dx = x[3 * i] - P[L.atoms + 3 * a], dy = x[3 * i + 1] - P[L.atoms + 3 * a + 1],
                   dz = x[3 * i + 2] - P[L.atoms + 3 * a + 2];
      const double r = sqrt(dx * dx + dy * dy + dz * dz);
      e += -P[L.charges + a] / r;            // potential_electron_nuclear / ECP local part1 (-Z_eff/r)
      if (ECP) {
        for (int k = 0; k < c_ecp.k_loc; ++k)                    // pseudopotential.py:95-101, r^(n-2)
          e += c_ecp.local_coes[a][k] * pow(r, c_ecp.rn_local[a][k] - 2.0) * exp(-c_ecp.local_exps[a][k] * r * r);
        for (int l = 0; l < AIQMC_ECP_MAX_L; ++l) {              // pseudopotential.py:150, r^n
          double v = 0.0;
          if (l < c_ecp.n_l)
            for (int k = 0; k < c_ecp.k_nl; ++k)
              v += c_ecp.non_local_coes[a][l][k] * pow(r, c_ecp.rn_non_local[a][l][k]) *
                   exp(-c_ecp.non_local_exps[a][l][k] * r * r);
          w.vl[((b * NE + i) * NA + a) * 4 + l] = v;
        }
      }
    }
  if (ECP) {
    w.base[b] = e;
    w.epp[2 * b] = 0.0;
    w.epp[2 * b + 1] = 0.0;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int p = 0; p < AIQMC_NQUAD; ++p) {
      double n2 = 0.0;
      for (int l = 0; l < 3; ++l) {
        double v = 0.0;
        for (int k = 0; k < 3; ++k) v += c_ecp.quad_pts[p][k] * rot[b * 9 + 3 * k + l];
        n2 += v * v;
      }
      acc[quad_group(p)] += n2;
    }
    for (int q = 0; q < 4; ++q) { acc[q] = sqrt(acc[q]); w.gnorm[4 * b + q] = acc[q]; }
    write_quad_record<NE, NA>(w.cache + b * MoveCache<NE, NA>::SIZE, x, rot + b * 9, acc, w.logabs[b], w.phase[b],
                              w.vl + b * NE * NA * 4);
  } else {
    e_out[b] = e;
  }
}

__device__ __forceinline__ void tmove_point_out(double* __restrict__ out, double v0, double v1, double v2, double v3,
                                                double cs, double rr, double ri, double tau);

// one thread = one quadrature point of one (electron, atom) pair of one walker
template <int NE, int NA>
__global__ void __launch_bounds__(kThreads, AIQMC_QUAD_MINB) k_ecp_quad(AiqmcSystem sys, const double* __restrict__ params,
                                                       const double* __restrict__ pos,
                                                       const double* __restrict__ rot, int64_t B, EnergyWs w,
                                                       double* __restrict__ tm_out, double tm_tau) {
  extern __shared__ double sP[];
  const double* P = stage_params<NE, NA>(params, sP);
  constexpr LayoutC<NE, NA> L{};
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per_walker = (int64_t)NE * NA * AIQMC_NQUAD;
  if (t >= B * per_walker) return;
  const int64_t b = t / per_walker;
  int rem = (int)(t - b * per_walker);
  const int i = rem / (NA * AIQMC_NQUAD);
  rem -= i * NA * AIQMC_NQUAD;
  const int a = rem / AIQMC_NQUAD, p = rem - a * AIQMC_NQUAD;
  const double* vl = w.vl + ((b * NE + i) * NA + a) * 4;
  bool any = false;
  for (int l = 0; l < AIQMC_ECP_MAX_L; ++l) any = any || (vl[l] != 0.0);
  if (!any && !tm_out) return;            // exact zero coefficient: contributes exactly 0
  double ae[3], nh[3];
  for (int c = 0; c < 3; ++c) ae[c] = pos[b * 3 * NE + 3 * i + c] - P[L.atoms + 3 * a + c];
  const double r = sqrt(ae[0] * ae[0] + ae[1] * ae[1] + ae[2] * ae[2]);
  for (int l = 0; l < 3; ++l) {           // Points = O @ rot   (pseudopotential.py:236-240)
    double v = 0.0;
    for (int k = 0; k < 3; ++k) v += c_ecp.quad_pts[p][k] * rot[b * 9 + 3 * k + l];
    nh[l] = v;
  }
  double dot = 0.0;
  for (int c = 0; c < 3; ++c) dot += ae[c] * (r * nh[c]);
  // moved configuration, built without dynamically indexed stores (keeps x in registers);
  // quirk Q13: electron i is placed at the absolute position r_ia * n_hat
  double x[3 * NE];
  for (int e = 0; e < NE; ++e)
    for (int c = 0; c < 3; ++c) x[3 * e + c] = (e == i) ? r * nh[c] : pos[b * 3 * NE + 3 * e + c];
  const double cs = dot / (r * (r * w.gnorm[4 * b + quad_group(p)]));   // quirk Q14
  double ph, la;
  Psi<NE, NA>::eval_value(sys, P, x, ph, la);
  // ratio = log psi(x') / log psi(x) * weight, complex logs (quirk Q12)
  const double dr = w.logabs[b], di = w.phase[b];
  const double inv = c_ecp.quad_wts[p] / (dr * dr + di * di);
  const double rr = (la * dr + ph * di) * inv, ri = (ph * dr - la * di) * inv;
  const double k4 = 0.07957747154594767;  // 1/(4 pi)
  const double pl[4] = {k4, 3.0 * k4 * cs, 5.0 * k4 * 0.5 * (3.0 * cs * cs - 1.0),
                        7.0 * k4 * 0.5 * (5.0 * cs * cs * cs - 3.0 * cs)};
  double f = 0.0;
  for (int l = 0; l < AIQMC_ECP_MAX_L; ++l) f += vl[l] * pl[l];
#ifdef AIQMC_DEBUG_QUAD
  if (t < 64) { double* d = w.vl + B * NE * NA * 4 + t * 8; d[0] = la; d[1] = ph; d[2] = cs; d[3] = f; d[4] = rr; d[5] = ri; d[6] = r; d[7] = dr; }
#endif
  if (tm_out) {
    tmove_point_out(tm_out + t * 4, vl[0], vl[1], vl[2], vl[3], cs, rr, ri, tm_tau);
    return;
  }
  atomicAdd(&w.epp[2 * b], f * rr);
  atomicAdd(&w.epp[2 * b + 1], f * ri);
}

// T-move preparation (DMC/Tmoves.py:32-66): the single-electron-move cache, log psi of the walker, v_l tables and
// point-group norms -- the energy stage without any derivative.  One thread per walker.
template <int NE, int NA>
__global__ void __launch_bounds__(kThreads) k_tmove_prep(AiqmcSystem sys, const double* __restrict__ params,
                                                         const double* __restrict__ pos,
                                                         const double* __restrict__ rot, int64_t B, EnergyWs w) {
  extern __shared__ double sP[];
  const double* P = stage_params<NE, NA>(params, sP);
  constexpr LayoutC<NE, NA> L{};
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double x[3 * NE];
  for (int q = 0; q < 3 * NE; ++q) x[q] = pos[b * 3 * NE + q];
  {
    typename Psi<NE, NA>::Primal pr;
    cplx M[NE * NE];
    double* mc = w.cache + b * MoveCache<NE, NA>::SIZE;
    Psi<NE, NA>::forward(sys, P, x, pr, M, mc + MoveCache<NE, NA>::HP, 1);
    double ld, ph;
    Psi<NE, NA>::lu_logdet(M, ld, ph);
    Psi<NE, NA>::write_cache(pr, ld + pr.jastrow, ph, mc);
    w.logabs[b] = ld + pr.jastrow;
    w.phase[b] = ph;
    DerivSplit<NE, NA>::keep_alive(&pr); DerivSplit<NE, NA>::keep_alive(M);
  }
  for (int i = 0; i < NE; ++i)
    for (int a = 0; a < NA; ++a) {
      const double dx = x[3 * i] - P[L.atoms + 3 * a], dy = x[3 * i + 1] - P[L.atoms + 3 * a + 1],
                   dz = x[3 * i + 2] - P[L.atoms + 3 * a + 2];
      const double r = sqrt(dx * dx + dy * dy + dz * dz);
      for (int l = 0; l < AIQMC_ECP_MAX_L; ++l) {              // pseudopotential.py:150, r^n
        double v = 0.0;
        if (l < c_ecp.n_l)
          for (int k = 0; k < c_ecp.k_nl; ++k)
            v += c_ecp.non_local_coes[a][l][k] * pow(r, c_ecp.rn_non_local[a][l][k]) *
                 exp(-c_ecp.non_local_exps[a][l][k] * r * r);
        w.vl[((b * NE + i) * NA + a) * 4 + l] = v;
      }
    }
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int p = 0; p < AIQMC_NQUAD; ++p) {
    double n2 = 0.0;
    for (int l = 0; l < 3; ++l) {
      double v = 0.0;
      for (int k = 0; k < 3; ++k) v += c_ecp.quad_pts[p][k] * rot[b * 9 + 3 * k + l];
      n2 += v * v;
    }
    acc[quad_group(p)] += n2;
  }
  for (int q = 0; q < 4; ++q) { acc[q] = sqrt(acc[q]); w.gnorm[4 * b + q] = acc[q]; }
  write_quad_record<NE, NA>(w.cache + b * MoveCache<NE, NA>::SIZE, x, rot + b * 9, acc, w.logabs[b], w.phase[b],
                            w.vl + b * NE * NA * 4);
}

// T-move selection (DMC/Tmoves.py:114-222), one thread per (walker, electron).  tm (B,N,A,50,4) holds
// [fwd.re, fwd.im, ratio.re, ratio.im] per quadrature point.  Complex comparisons are lexicographic (quirk Q25);
// the selection is jnp.searchsorted's fixed-trip-count bisection on the (unsorted) complex cdf; the back
// amplitude indexes the ELECTRON axis with the move index and the slices 1:19:55:79:151 are the reference's
// hard-coded A = 3 layout (quirk Q18) -- both replicated.
template <int NE, int NA>
__global__ void __launch_bounds__(kThreads) k_tmove_select(const double* __restrict__ pos, const double* __restrict__ rot,
                                                           const double* __restrict__ params,
                                                           const double* __restrict__ tm, const double* __restrict__ u,
                                                           const double* __restrict__ rnd, int64_t B,
                                                           double* __restrict__ pos_out, double* __restrict__ acceptance,
                                                           int32_t* __restrict__ selected) {
  constexpr int N = NE, A = NA, PW = A * AIQMC_NQUAD, NMOV = 1 + PW;
  constexpr LayoutC<NE, NA> L{};
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * N) return;
  const int64_t b = t / N;
  const int i = (int)(t - b * N);
  const double* tw = tm + b * N * PW * 4;
  // norm = 1 + sum_g w_g sum fwd_g
  cplx norm = {1.0, 0.0};
  for (int e = 0; e < N; ++e)
    for (int a = 0; a < A; ++a)
      for (int p = 0; p < AIQMC_NQUAD; ++p) {
        const double* q = tw + ((e * A + a) * AIQMC_NQUAD + p) * 4;
        norm.re += c_ecp.quad_wts[p] * q[0];
        norm.im += c_ecp.quad_wts[p] * q[1];
      }
  const cplx ninv = cinv(norm);
  auto total = [&](int e, int m) -> cplx {       // forward_probability_output_total_final[e][m]
    if (m == 0) return cplx{1.0, 0.0};
    const double* q = tw + (e * PW + (m - 1)) * 4;
    return cplx{q[0], q[1]};
  };
  auto cdf_at = [&](int m) -> cplx {             // cumsum(total[i] / norm)[m]
    cplx s = {0.0, 0.0};
    for (int k = 0; k <= m; ++k) s = cadd(s, cmul(total(i, k), ninv));
    return s;
  };
  // cumulative sums once, in order (matches jnp.cumsum's sequential association)
  // bisection with ceil(log2(NMOV + 1)) trips
  const double r = u[b] + 1.0;
  int low = 0, high = NMOV;
  int trips = 0;
  for (int v = NMOV; v > 0; v >>= 1) ++trips;     // bit length of NMOV == ceil(log2(NMOV + 1))
  // running prefix for the bisection: recomputing the prefix is O(NMOV) per probe; NMOV <= 1 + 50 A
  if constexpr (NMOV <= 128) {
    // the whole running sum once (same sequential association as cdf_at), then the probes read it: the bisection made
    // ~8 probes of ~NMOV/2 complex multiply-adds each
    cplx cdf[NMOV];
    cplx run = {0.0, 0.0};
    for (int k = 0; k < NMOV; ++k) { run = cadd(run, cmul(total(i, k), ninv)); cdf[k] = run; }
    for (int it = 0; it < trips; ++it) {
      const int mid = (low + high) / 2;
      const cplx c = cdf[mid < NMOV ? mid : NMOV - 1];
      const bool le = (r < c.re) || (r == c.re && 0.0 <= c.im);   // r + 0j <= cdf[mid], lexicographic
      if (le) high = mid; else low = mid;
    }
  } else {
    for (int it = 0; it < trips; ++it) {
      const int mid = (low + high) / 2;
      const cplx c = cdf_at(mid < NMOV ? mid : NMOV - 1);
      const bool le = (r < c.re) || (r == c.re && 0.0 <= c.im);   // r + 0j <= cdf[mid], lexicographic
      if (le) high = mid; else low = mid;
    }
  }
  int mv = high < NMOV ? high : 0;
  if (selected) selected[t] = mv;
  // selected coordinates and ratio
  double xo[3], xs[3];
  for (int c = 0; c < 3; ++c) { xo[c] = pos[b * 3 * N + 3 * i + c]; xs[c] = xo[c]; }
  cplx rsel = {1.0, 0.0};
  if (mv > 0) {
    const int a = (mv - 1) / AIQMC_NQUAD, p = (mv - 1) - a * AIQMC_NQUAD;
    double ae[3];
    for (int c = 0; c < 3; ++c) ae[c] = xo[c] - params[L.atoms + 3 * a + c];
    const double rr = sqrt(ae[0] * ae[0] + ae[1] * ae[1] + ae[2] * ae[2]);
    for (int l = 0; l < 3; ++l)          // quirk Q13: the electron is placed at r_ia * n_hat (atom position not added)
      xs[l] = rr * (c_ecp.quad_pts[p][0] * rot[b * 9 + l] + c_ecp.quad_pts[p][1] * rot[b * 9 + 3 + l] +
                    c_ecp.quad_pts[p][2] * rot[b * 9 + 6 + l]);
    const double* q = tw + (i * PW + (mv - 1)) * 4;
    rsel = cplx{q[2], q[3]};
  }
  const cplx rinv = cinv(rsel);
  // back amplitudes: t_amp[move] indexes the electron axis (clamped), quirk Q18
  const int erow = mv < N - 1 ? mv : N - 1;
  cplx bn = {1.0, 0.0};
  const int lo[4] = {1, 19, 55, 79}, hi[4] = {19, 55, 79, 151};
  const int gfirst[4] = {0, 6, 18, 26};
  for (int g = 0; g < 4; ++g) {
    cplx sg = {0.0, 0.0};
    for (int m = lo[g]; m < hi[g] && m < NMOV; ++m) sg = cadd(sg, cmul(total(erow, m), rinv));
    bn.re += c_ecp.quad_wts[gfirst[g]] * sg.re;
    bn.im += c_ecp.quad_wts[gfirst[g]] * sg.im;
  }
  const cplx ratio = cmul(norm, cinv(bn));
  const double acc = ratio.re;
  acceptance[t] = acc;
  const bool ok = acc > rnd[t];
  for (int c = 0; c < 3; ++c) pos_out[b * 3 * N + 3 * i + c] = ok ? xs[c] : xo[c];
}

static __global__ void k_energy_final(int64_t B, double* __restrict__ e_l, EnergyWs w) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  e_l[2 * b] = w.base[b] + w.epp[2 * b];
  e_l[2 * b + 1] = w.epp[2 * b + 1];
}

}  // namespace aiqmc
#include "ecp_coop.cuh"
#include "ecp_pt.cuh"
#include "ecp_grp.cuh"
#include "coop_grad.cuh"
#include "coop_lap.cuh"
#include "param_grad.cuh"
namespace aiqmc {

// ---------------------------------------------------------------------------------------
// parameter gradient of sum_w (alpha_w log|psi_w| + beta_w phase_w)  (Loss/pploss.py:186-223, param_grad.cuh)
// One warp per CTA, one walker per lane, grid-stride over the batch; every contribution is reduced across the
// warp in a fixed order and added to the CTA's shared-memory accumulator by lane 0 (deterministic, no atomics);
// the per-CTA partials are summed by k_reduce_param_partials.
// ---------------------------------------------------------------------------------------
struct WarpSink {
  double* acc;
  __device__ __forceinline__ void add(int off, double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) acc[off] += v;
  }
};
constexpr int kPgradMaxBlocks = 148 * 16;
inline int64_t pgrad_blocks(int64_t B) {
  const int64_t nb = (B + 31) / 32;
  return nb < 1 ? 1 : (nb > kPgradMaxBlocks ? kPgradMaxBlocks : nb);
}
constexpr int kPgradCachedBlocks = 148;      // CTAs per chunk of the N > 16 path (parameters + accumulator fill an SM's shared memory)
inline int64_t pgrad_cached_rows(int n, int a, int64_t B) {
  const int64_t c = deriv_chunk(n, a, false);
  return ((B + c - 1) / c) * kPgradCachedBlocks;
}
inline int64_t pgrad_ws_bytes(int n, int a, int64_t B) {
  if (n <= 16) return align256(pgrad_blocks(B) * make_layout(n, a).total * 8);
  return align256(pgrad_cached_rows(n, a, B) * make_layout(n, a).total * 8) + deriv_cache_bytes(n, a, false, B) + align256(B * 8);
}

template <int NE, int NA>
__global__ void __launch_bounds__(32) k_param_grad(AiqmcSystem sys, const double* __restrict__ params,
                                                   const double* __restrict__ pos, int64_t B,
                                                   const double* __restrict__ alpha, const double* __restrict__ beta,
                                                   double* __restrict__ phase, double* __restrict__ logabs,
                                                   double* __restrict__ partials) {
  constexpr int total = make_layout(NE, NA).total;
  extern __shared__ double sP[];
  const double* P = stage_params<NE, NA>(params, sP);
  double* acc = sP + total;
  for (int q = threadIdx.x; q < total; q += 32) acc[q] = 0.0;
  __syncwarp();
  WarpSink sink{acc};
  for (int64_t b0 = (int64_t)blockIdx.x * 32; b0 < B; b0 += (int64_t)gridDim.x * 32) {
    const int64_t b = b0 + threadIdx.x;
    const bool valid = b < B;
    const int64_t bb = valid ? b : B - 1;                // tail lanes redo the last walker with a zero seed
    double x[3 * NE];
    for (int q = 0; q < 3 * NE; ++q) x[q] = pos[bb * 3 * NE + q];
    double ph, la;
    ParamGrad<NE, NA>::run(sys, P, x, valid ? alpha[b] : 0.0, valid ? beta[b] : 0.0, ph, la, sink);
    if (valid && phase) phase[b] = ph;
    if (valid && logabs) logabs[b] = la;
  }
  __syncwarp();
  for (int q = threadIdx.x; q < total; q += 32) partials[(int64_t)blockIdx.x * total + q] = acc[q];
}

// N > 16: the same sweep run on the derivative cache of a chunk of walkers (k_primal<false> wrote it)
template <int NE, int NA>
__global__ void __launch_bounds__(32) k_param_grad_cached(AiqmcSystem sys, const double* __restrict__ params,
                                                          const double* __restrict__ dc, int64_t cfg0, int64_t n_cfg,
                                                          const double* __restrict__ alpha, const double* __restrict__ beta,
                                                          double* __restrict__ partials) {
  constexpr int total = make_layout(NE, NA).total;
  extern __shared__ double sP[];
  const double* P = stage_params<NE, NA>(params, sP);
  double* acc = sP + total;
  for (int q = threadIdx.x; q < total; q += 32) acc[q] = 0.0;
  __syncwarp();
  WarpSink sink{acc};
  for (int64_t b0 = (int64_t)blockIdx.x * 32; b0 < n_cfg; b0 += (int64_t)gridDim.x * 32) {
    const int64_t tl = b0 + threadIdx.x;
    const bool valid = tl < n_cfg;
    const int64_t tt = valid ? tl : n_cfg - 1;            // tail lanes redo the last walker with a zero seed
    ParamGrad<NE, NA>::run_cached(sys, P, dc + tt, n_cfg, valid ? alpha[cfg0 + tl] : 0.0, valid ? beta[cfg0 + tl] : 0.0, sink);
  }
  __syncwarp();
  for (int q = threadIdx.x; q < total; q += 32) partials[(int64_t)blockIdx.x * total + q] = acc[q];
}

static __global__ void k_reduce_param_partials(const double* __restrict__ partials, int nblk, int total,
                                               double* __restrict__ out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= total) return;
  double s = 0.0;
  for (int k = 0; k < nblk; ++k) s += partials[(int64_t)k * total + q];      // fixed order
  out[q] = s;
}

// ---------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------
#define AQ_CUDA_OK(call)                                   \
  do {                                                     \
    cudaError_t e_ = (call);                               \
    if (e_ != cudaSuccess) { g_last_cuda_error = (int)e_; return AIQMC_E_CUDA; } \
  } while (0)

extern std::atomic<int> g_last_cuda_error;
extern std::atomic<int64_t> g_launch_count;      // kernels launched by this library (aiqmc_launch_count)

template <int NE, int NA>
struct Launch {
  static constexpr int kSmem = make_layout(NE, NA).total * 8;
  static constexpr bool kCoop = (NE <= 16 && NA <= 4);   // lane-per-electron quadrature kernel available
  static constexpr bool kGrp = (NE >= 5 && NE <= 32);    // packed lane-per-electron groups (ecp_grp.cuh)
  static constexpr bool kPt = (NE <= 4 && make_layout(NE, NA).total <= kConstParMax);   // thread-per-point kernel

  template <class K>
  static cudaError_t prep(K kernel) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  }

  // This instantiation's __constant__ copies (c_ecp, c_par, c_uni) are ONE set per device, shared by every engine
  // and every stream of the process.  ConstScope makes their use safe: it serialises the host threads of the process
  // on a mutex, keeps the "what was uploaded last" record PER DEVICE (a second engine on another GPU re-uploads),
  // and orders streams -- a caller on stream S first waits for the event recorded behind the previous user's last
  // kernel (uploaded table still being read, or upload still in flight, on another stream), and records its own
  // event when it leaves.  Two engines with different parameters on two streams therefore interleave correctly
  // (tests/test_gpu_streams.py); they serialise only at the kernels that read constant memory.
  struct ConstSlot {
    bool has_ev = false, ecp_valid = false;
    cudaEvent_t ev{};
    cudaStream_t last{};
    AiqmcEcp ecp;
  };
  struct ConstScope {
    static constexpr int kMaxDev = 64;
    std::unique_lock<std::mutex> lock;
    ConstSlot* slot = nullptr;
    cudaStream_t st;
    static std::mutex& mu() { static std::mutex m; return m; }
    explicit ConstScope(cudaStream_t s) : lock(mu()), st(s) {
      static ConstSlot slots[kMaxDev];
      int dev = 0;
      if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return;
      slot = &slots[dev];
      if (slot->has_ev && slot->last != st) cudaStreamWaitEvent(st, slot->ev, 0);
    }
    ~ConstScope() {
      if (!slot) return;
      if (!slot->has_ev && cudaEventCreateWithFlags(&slot->ev, cudaEventDisableTiming) == cudaSuccess) slot->has_ev = true;
      if (slot->has_ev) cudaEventRecord(slot->ev, st);
      slot->last = st;
    }
    cudaError_t upload_ecp(const AiqmcEcp* ecp) {
      if (slot && slot->ecp_valid && memcmp(&slot->ecp, ecp, sizeof(AiqmcEcp)) == 0) return cudaSuccess;
      const cudaError_t e = cudaMemcpyToSymbolAsync(c_ecp, ecp, sizeof(AiqmcEcp), 0, cudaMemcpyHostToDevice, st);
      if (e != cudaSuccess) return e;
      if (slot) { slot->ecp = *ecp; slot->ecp_valid = true; }
      return cudaSuccess;
    }
  };

  // One chunked sweep of the two derivative passes over n_cfg configurations (deriv_split.cuh).
  //   SRC/OUT as in k_primal / k_tangent.  partials (may be null): block partials of sum g^2 go to rows
  //   [*rows, ...) column pcol; *rows is advanced.
  template <bool LAP, int SRC, int OUT>
  static int deriv(const AiqmcSystem* sys, const double* params, const double* pos, int64_t n_cfg, MovedSrc ms,
                   double* dcache, double* mc, double* phase, double* logabs, double* gout, double* lap_parts,
                   int64_t lap_stride, double* partials, int pcol, int64_t* rows, cudaStream_t st) {
    AQ_CUDA_OK(prep(k_primal<NE, NA, LAP, SRC>));
    AQ_CUDA_OK(prep(k_tangent<NE, NA, LAP, OUT>));
    const int64_t chunk = deriv_chunk(NE, NA, LAP);
    for (int64_t c0 = 0; c0 < n_cfg; c0 += chunk) {
      const int64_t nc = (n_cfg - c0 < chunk) ? n_cfg - c0 : chunk;
      const unsigned gp = (unsigned)((nc + kThreads - 1) / kThreads);
      const unsigned gt = (unsigned)(((nc + 31) / 32) * tan_groups<NE>());
      ++g_launch_count;
      k_primal<NE, NA, LAP, SRC><<<gp, kThreads, kSmem, st>>>(*sys, params, pos, c0, nc, ms, dcache, mc, phase, logabs);
      ++g_launch_count;
#ifndef AIQMC_NO_REVERSE_CACHED
      if constexpr (!LAP) {     // gradient only: ONE reverse sweep per configuration instead of 3N forward tangents
        AQ_CUDA_OK(prep(k_reverse_cached<NE, NA, OUT>));
        k_reverse_cached<NE, NA, OUT><<<gp, kThreads, kSmem, st>>>(*sys, params, dcache, c0, nc, gout,
                                                                  partials ? partials + *rows * 4 : nullptr, pcol);
        if (rows) *rows += gp;
        continue;
      }
#endif
      k_tangent<NE, NA, LAP, OUT><<<gt, 32 * tan_warps<NE>(), kSmem, st>>>(*sys, params, dcache, c0, nc, gout, lap_parts, lap_stride,
                                                              partials ? partials + *rows * 4 : nullptr, pcol);
      if (rows) *rows += gt;
    }
    AQ_CUDA_OK(cudaGetLastError());
    return AIQMC_OK;
  }

  // gradient + per-coordinate second derivatives (+ MoveCache) of n_cfg configurations: the two-phase shared-memory
  // kernel of coop_lap.cuh for N <= 16, the two-pass HBM-cache path beyond
  static constexpr bool kLapCoop = (NE <= 16);
  static int lap_pass(const AiqmcSystem* sys, const double* params, const double* pos, int64_t n_cfg, double* dcache,
                 double* mc, double* phase, double* logabs, double* gout, double* lap_parts, int64_t lap_stride,
                 cudaStream_t st) {
#ifndef AIQMC_LAP_TWO_PASS
    if constexpr (kLapCoop) {
      using CL = CoopLapCfg<NE, NA>;
      AQ_CUDA_OK(cudaFuncSetAttribute(k_lap_coop<NE, NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, CL::kBytes));
      AQ_CUDA_OK(cudaFuncSetAttribute(k_lap_coop<NE, NA>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      cudaSharedmemCarveoutMaxShared));
      unsigned g = 1;
      const int rc = coop_grid(k_lap_coop<NE, NA>, CL::T, CL::kBytes, (n_cfg + CL::NG - 1) / CL::NG, &g);
      if (rc != AIQMC_OK) return rc;
      ++g_launch_count;
      k_lap_coop<NE, NA><<<g, CL::T, CL::kBytes, st>>>(*sys, params, pos, n_cfg, mc, phase, logabs, gout, lap_parts, lap_stride);
      AQ_CUDA_OK(cudaGetLastError());
      return AIQMC_OK;
    }
#endif
    MovedSrc ms{};
    return deriv<true, 0, 0>(sys, params, pos, n_cfg, ms, dcache, mc, phase, logabs, gout, lap_parts, lap_stride, nullptr, 0,
                             nullptr, st);
  }

  // spin layouts the quadrature kernel is specialised for (ecp_pt.cuh): 0 = up-first, 1 = alternating; -1 = other
  static int spin_layout(const AiqmcSystem* sys) {
    constexpr int NUP = (NE + 1) / 2;
    if (sys->n_up != NUP || sys->n_dn != NE - NUP || sys->n_up_rows != NUP) return -1;
    bool l0 = true, l1 = true;
    for (int k = 0; k < NE; ++k) {
      l0 = l0 && sys->sigma[k] == pt_sigma<NE, 0>(k);
      l1 = l1 && sys->sigma[k] == pt_sigma<NE, 1>(k);
    }
    return l0 ? 0 : (l1 ? 1 : -1);
  }
  template <int LAYOUT, int... I>
  static void launch_pt_layout(const AiqmcSystem* sys, const double* pos, const double* rot, int64_t B,
                               const EnergyWs& w, double* tm_out, double tm_tau, cudaStream_t st) {
    constexpr int WPC = AIQMC_PT_WPCI;
    (k_ecp_pt<NE, NA, WPC, I, LAYOUT><<<(unsigned)((B + WPC - 1) / WPC), pt_threads<NE, NA, WPC, I>(), 0, st>>>(
         *sys, pos, rot, B, w.cache, w, tm_out, tm_tau), ...);
  }
  template <int... I>
  static void launch_pt(const AiqmcSystem* sys, const double* pos, const double* rot, int64_t B, const EnergyWs& w,
                        double* tm_out, double tm_tau, cudaStream_t st, std::integer_sequence<int, I...>) {
    g_launch_count += sizeof...(I);
    const int layout = spin_layout(sys);
    if (layout == 0) launch_pt_layout<0, I...>(sys, pos, rot, B, w, tm_out, tm_tau, st);
    else if (layout == 1) launch_pt_layout<1, I...>(sys, pos, rot, B, w, tm_out, tm_tau, st);
    else launch_pt_layout<-1, I...>(sys, pos, rot, B, w, tm_out, tm_tau, st);
  }

  // packed-group quadrature / T-move amplitudes: stage the electron-independent weights in constant memory, launch
  static int launch_grp(const AiqmcSystem* sys, const double* params, const double* pos, const double* rot, int64_t B,
                        const EnergyWs& w, double* tm_out, double tm_tau, cudaStream_t st) {
    if constexpr (kGrp) {
      using U = UniLayout<NE, NA>;
      using CF = GrpCfg<NE, NA>;
      for (int l = 0; l < 3; ++l)
        AQ_CUDA_OK(cudaMemcpyToSymbolAsync(c_uni, params + U::seg_src(l), U::seg_len(l) * sizeof(double),
                                           U::seg_dst(l) * sizeof(double), cudaMemcpyDeviceToDevice, st));
      AQ_CUDA_OK(cudaMemcpyToSymbolAsync(c_uni, params + U::K.y_w, 6 * NE * sizeof(double), U::y_w * sizeof(double),
                                         cudaMemcpyDeviceToDevice, st));
      auto launch = [&](auto kern) -> int {
        AQ_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::kBytes));
        AQ_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        ++g_launch_count;
        kern<<<(unsigned)B, CF::T, CF::kBytes, st>>>(*sys, params, pos, rot, B, w.cache, w, tm_out, tm_tau);
        return AIQMC_OK;
      };
      int rc;
      if constexpr (CF::kFlat) rc = launch(k_ecp_grp<NE, NA>);             // flat point loop for N <= 16
      else rc = launch(k_ecp_grp_rows<NE, NA>);                            // per-electron loop beyond (C6H6)
      if (rc != AIQMC_OK) return rc;
      AQ_CUDA_OK(cudaGetLastError());
    }
    return AIQMC_OK;
  }

  static constexpr bool kReverse = (NE <= 16);   // fused forward + reverse gradient, lane per electron (coop_grad.cuh)

  // persistent grid of a lane-per-electron kernel: every CTA resident at once, tiles dealt round-robin
  template <class K>
  static int coop_grid(K kernel, int threads, int smem_bytes, int64_t tiles, unsigned* grid) {
    int dev = 0, sms = 148, per_sm = 1;
    AQ_CUDA_OK(cudaGetDevice(&dev));
    AQ_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    AQ_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem_bytes));
    int64_t g = (int64_t)sms * (per_sm < 1 ? 1 : per_sm);
    if (g > kCoopMaxGrid) g = kCoopMaxGrid;
    if (g > tiles) g = tiles;
    *grid = (unsigned)(g < 1 ? 1 : g);
    return AIQMC_OK;
  }

  // gradient of n_cfg configurations: fused reverse sweep for N <= 16, the two-pass path beyond
  template <int SRC, int OUT>
  static int grad(const AiqmcSystem* sys, const double* params, const double* pos, int64_t n_cfg, MovedSrc ms,
                  double* dcache, double* phase, double* logabs, double* gout, double* partials, int pcol,
                  int64_t* rows, cudaStream_t st) {
    if constexpr (kReverse) {
#ifdef AIQMC_GRAD_THREAD_PER_CFG      // round-1 kernel (one thread per configuration), kept for A/B timing only
      AQ_CUDA_OK(prep(k_grad_reverse<NE, NA, SRC, OUT>));
      const unsigned g = (unsigned)((n_cfg + kThreads - 1) / kThreads);
      ++g_launch_count;
      k_grad_reverse<NE, NA, SRC, OUT><<<g, kThreads, kSmem, st>>>(*sys, params, pos, n_cfg, ms, phase, logabs, gout,
                                                                    partials ? partials + *rows * 4 : nullptr, pcol);
#else
      using CG = CoopGradCfg<NE, NA>;
      AQ_CUDA_OK(cudaFuncSetAttribute(k_grad_coop<NE, NA, SRC, OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, CG::kBytes));
      unsigned g = 1;
      const int rc = coop_grid(k_grad_coop<NE, NA, SRC, OUT>, CG::T, CG::kBytes, (n_cfg + CG::NG - 1) / CG::NG, &g);
      if (rc != AIQMC_OK) return rc;
      ++g_launch_count;
      k_grad_coop<NE, NA, SRC, OUT><<<g, CG::T, CG::kBytes, st>>>(*sys, params, pos, n_cfg, ms, phase, logabs, gout,
                                                                   partials ? partials + *rows * 4 : nullptr, pcol);
#endif
      if (rows) *rows += g;
      AQ_CUDA_OK(cudaGetLastError());
      return AIQMC_OK;
    } else {
      return deriv<false, SRC, OUT>(sys, params, pos, n_cfg, ms, dcache, nullptr, phase, logabs, gout, nullptr, 0,
                                    partials, pcol, rows, st);
    }
  }

  static int psi(const AiqmcSystem* sys, const double* params, const double* pos, int64_t n_cfg, int mode,
                 double* phase, double* logabs, double* grad, double* lap, void* ws, int64_t ws_bytes,
                 cudaStream_t st) {
    if (n_cfg <= 0) return AIQMC_OK;
    if (mode == 0) {
      const unsigned grid = (unsigned)((n_cfg + kThreads - 1) / kThreads);
      AQ_CUDA_OK(prep(k_psi<NE, NA>));
      ++g_launch_count;
      k_psi<NE, NA><<<grid, kThreads, kSmem, st>>>(*sys, params, pos, n_cfg, phase, logabs);
      AQ_CUDA_OK(cudaGetLastError());
      return AIQMC_OK;
    }
    if (!ws || ws_bytes < psi_ws_bytes(NE, NA, n_cfg, mode == 2)) return AIQMC_E_WORKSPACE;
    MovedSrc ms{};
    double* dcache = (double*)ws;
    if (mode == 1)
      return Launch::grad<0, 0>(sys, params, pos, n_cfg, ms, dcache, phase, logabs, grad, nullptr, 0, nullptr, st);
    // Laplacian: the per-coordinate second derivatives of one chunk are summed right after its tangent pass
    double* parts = (double*)((char*)ws + deriv_cache_bytes(NE, NA, true, n_cfg));
    const int64_t chunk = deriv_chunk(NE, NA, true);
    for (int64_t c0 = 0; c0 < n_cfg; c0 += chunk) {
      const int64_t nc = (n_cfg - c0 < chunk) ? n_cfg - c0 : chunk;
      const int rc = lap_pass(sys, params, pos + c0 * 3 * NE, nc, dcache, nullptr, phase + c0, logabs + c0, grad + c0 * 3 * NE,
                         parts, nc, st);
      if (rc != AIQMC_OK) return rc;
      ++g_launch_count;
      k_sum_lap_parts<<<(unsigned)((nc + 255) / 256), 256, 0, st>>>(3 * NE, parts, nc, nc, lap + c0);
    }
    AQ_CUDA_OK(cudaGetLastError());
    return AIQMC_OK;
  }

  static int sweep(const AiqmcSystem* sys, const double* params, double* pos, const double* gauss1,
                   const double* gauss2, const double* rnd, int64_t B, double tau, double acyrus, int signed_ratio,
                   int g2_compact, uint8_t* accept, double* grad_eff_old, double* aux_out, void* ws, int64_t ws_bytes,
                   cudaStream_t st) {
    if (B <= 0) return AIQMC_OK;
    if (ws_bytes < sweep_ws_bytes(NE, NA, B)) return AIQMC_E_WORKSPACE;
    SweepWs w = carve_sweep_ws(ws, NE, NA, B);
    const int64_t n2 = B * NE;                                              // single-electron-moved configurations
    const unsigned g3 = (unsigned)((B * NE + kRedThreads - 1) / kRedThreads);
    MovedSrc ms{w.grad, gauss1, w.scal, w.xprop, tau, acyrus};
    // grad log|psi| at x1 and its batch-global square sum (limdrift, quirk Q6)
    int64_t rows = 0;
    int rc = grad<0, 0>(sys, params, pos, B, ms, w.dcache, nullptr, w.logabs1, w.grad, w.partials, 0, &rows, st);
    if (rc != AIQMC_OK) return rc;
    ++g_launch_count;
    k_reduce_partials<<<1, kRedThreads, 0, st>>>(w.partials, (int)rows, 0, 1, w.scal, 0);
    // the N single-electron-moved configurations of every walker
    rows = 0;
    rc = grad<1, 1>(sys, params, pos, n2, ms, w.dcache, nullptr, w.logabs2, w.gnew, w.partials, 1, &rows, st);
    if (rc != AIQMC_OK) return rc;
    ++g_launch_count;
    k_reduce_partials<<<1, kRedThreads, 0, st>>>(w.partials, (int)rows, 1, 1, w.scal, 1);
    ++g_launch_count;
    const int use_tma = ((((uintptr_t)pos) | ((uintptr_t)gauss2) | ((uintptr_t)rnd)) & 15) == 0;   // bulk copies need 16-byte alignment
    k_sweep_accept<<<g3, kRedThreads, 0, st>>>(NE, pos, gauss2, rnd, B, tau, acyrus, signed_ratio, g2_compact, use_tma, accept,
                                               grad_eff_old, w);
    ++g_launch_count;
    k_reduce_partials<<<1, kRedThreads, 0, st>>>(w.partials, (int)g3, 2, 2, w.scal, 2);
    if (aux_out) {
      // aux_out = [sum x_new, sum x_prop, v2_old, v2_new]
      AQ_CUDA_OK(cudaMemcpyAsync(aux_out, w.scal + 2, 2 * sizeof(double), cudaMemcpyDeviceToDevice, st));
      AQ_CUDA_OK(cudaMemcpyAsync(aux_out + 2, w.scal, 2 * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    AQ_CUDA_OK(cudaGetLastError());
    return AIQMC_OK;
  }

  static int energy(const AiqmcSystem* sys, const AiqmcEcp* ecp, const double* params, const double* pos,
                    const double* rot, int64_t B, double* e_l, void* ws, int64_t ws_bytes, int stages, cudaStream_t st) {
    if (B <= 0) return AIQMC_OK;
    const int with_ecp = ecp != nullptr;
    if (ws_bytes < energy_ws_bytes(NE, NA, B, with_ecp)) return AIQMC_E_WORKSPACE;
    EnergyWs w = carve_energy_ws(ws, NE, NA, B, with_ecp);
    const unsigned g1 = (unsigned)((B + kThreads - 1) / kThreads);
    if (!with_ecp) {
      AQ_CUDA_OK(prep(k_energy_rest<NE, NA, false>));
      const int rc = lap_pass(sys, params, pos, B, w.dcache, nullptr, w.phase, w.logabs, w.grad, w.lap_parts, B, st);
      if (rc != AIQMC_OK) return rc;
      ++g_launch_count;
      k_energy_rest<NE, NA, false><<<g1, kThreads, kSmem, st>>>(*sys, params, pos, nullptr, B, e_l, w);
    } else {
      ConstScope cs(st);
      AQ_CUDA_OK(cs.upload_ecp(ecp));
      if (stages & 1) {
        AQ_CUDA_OK(prep(k_energy_rest<NE, NA, true>));
        const int rc = lap_pass(sys, params, pos, B, w.dcache, w.cache, w.phase, w.logabs, w.grad, w.lap_parts, B, st);
        if (rc != AIQMC_OK) return rc;
        ++g_launch_count;
        k_energy_rest<NE, NA, true><<<g1, kThreads, kSmem, st>>>(*sys, params, pos, rot, B, e_l, w);
      }
      if (stages & 2) {
        // stage bits 8 / 16 force the thread-per-point full-evaluation kernel / the lane-per-electron
        // kernel (cross-checks and systems the faster ones do not cover)
        bool done = false;
        if constexpr (kPt) {
          if (!(stages & (8 | 16))) {
            AQ_CUDA_OK(cudaMemcpyToSymbolAsync(c_par, params, make_layout(NE, NA).total * sizeof(double), 0,
                                               cudaMemcpyDeviceToDevice, st));
#ifdef AIQMC_PT_SINGLE_LAUNCH
            constexpr int WPC = AIQMC_PT_WPC;
            ++g_launch_count;
            k_ecp_pt<NE, NA, WPC, -1, -1><<<(unsigned)((B + WPC - 1) / WPC), pt_threads<NE, NA, WPC, -1>(), 0, st>>>(
                *sys, pos, rot, B, w.cache, w, nullptr, 0.0);
#else
            launch_pt(sys, pos, rot, B, w, nullptr, 0.0, st, std::make_integer_sequence<int, NE>{});   // one launch per moved electron
#endif
            done = true;
          }
        }
        if constexpr (kGrp) {
          if (!done && !(stages & (8 | 16))) {
            const int rc = launch_grp(sys, params, pos, rot, B, w, nullptr, 0.0, st);
            if (rc != AIQMC_OK) return rc;
            done = true;
          }
        }
        if constexpr (kCoop) {
          if (!done && !(stages & 8)) {
            const int T = coop_threads<NE, NA>();
            const int smem = CoopSmem<NE, NA>::doubles(T / GroupSize<NE>::G) * 8;
            AQ_CUDA_OK(cudaFuncSetAttribute(k_ecp_coop<NE, NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            AQ_CUDA_OK(cudaFuncSetAttribute(k_ecp_coop<NE, NA>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                            cudaSharedmemCarveoutMaxShared));
            ++g_launch_count;
            k_ecp_coop<NE, NA><<<(unsigned)B, T, smem, st>>>(*sys, params, pos, rot, B, w.cache, w);
            done = true;
          }
        }
        if (!done) {   // thread-per-point full evaluation (N > 16, or forced)
          AQ_CUDA_OK(prep(k_ecp_quad<NE, NA>));
          const int64_t nt = B * NE * NA * AIQMC_NQUAD;
          const unsigned g2 = (unsigned)((nt + kThreads - 1) / kThreads);
          ++g_launch_count;
          k_ecp_quad<NE, NA><<<g2, kThreads, kSmem, st>>>(*sys, params, pos, rot, B, w, nullptr, 0.0);
        }
      }
      if (stages & 4) ++g_launch_count;
      if (stages & 4) k_energy_final<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(B, e_l, w);
    }
    AQ_CUDA_OK(cudaGetLastError());
    return AIQMC_OK;
  }

  // T-moves (DMC/Tmoves.py:32-225): prep -> per-point amplitudes (the quadrature kernel in record mode) -> select
  static int tmove(const AiqmcSystem* sys, const AiqmcEcp* ecp, const double* params, const double* pos,
                   const double* rot, const double* u, const double* rnd, int64_t B, double tau, double* pos_out,
                   double* acceptance, int32_t* selected, void* ws, int64_t ws_bytes, cudaStream_t st) {
    if (B <= 0) return AIQMC_OK;
    if (ws_bytes < tmove_ws_bytes(NE, NA, B)) return AIQMC_E_WORKSPACE;
    EnergyWs w = carve_energy_ws(ws, NE, NA, B, 1);
    double* tm = (double*)((char*)ws + energy_ws_bytes(NE, NA, B, 1));
    ConstScope cs(st);
    AQ_CUDA_OK(cs.upload_ecp(ecp));
    const unsigned g1 = (unsigned)((B + kThreads - 1) / kThreads);
    AQ_CUDA_OK(prep(k_tmove_prep<NE, NA>));
    ++g_launch_count;
    k_tmove_prep<NE, NA><<<g1, kThreads, kSmem, st>>>(*sys, params, pos, rot, B, w);
    if constexpr (kPt) {
      AQ_CUDA_OK(cudaMemcpyToSymbolAsync(c_par, params, make_layout(NE, NA).total * sizeof(double), 0,
                                         cudaMemcpyDeviceToDevice, st));
      launch_pt(sys, pos, rot, B, w, tm, tau, st, std::make_integer_sequence<int, NE>{});
    } else if constexpr (kGrp) {
      const int rc = launch_grp(sys, params, pos, rot, B, w, tm, tau, st);
      if (rc != AIQMC_OK) return rc;
    } else {
      AQ_CUDA_OK(prep(k_ecp_quad<NE, NA>));
      const int64_t nt = B * NE * NA * AIQMC_NQUAD;
      ++g_launch_count;
      k_ecp_quad<NE, NA><<<(unsigned)((nt + kThreads - 1) / kThreads), kThreads, kSmem, st>>>(*sys, params, pos, rot, B, w,
                                                                                          tm, tau);
    }
    ++g_launch_count;
    k_tmove_select<NE, NA><<<(unsigned)((B * NE + kThreads - 1) / kThreads), kThreads, 0, st>>>(
        pos, rot, params, tm, u, rnd, B, pos_out, acceptance, selected);
    AQ_CUDA_OK(cudaGetLastError());
    return AIQMC_OK;
  }


  static constexpr bool kPgrad = (NE <= 16);   // fused forward + reverse per thread; beyond: primal pass + sweep on the cache
  static int param_grad(const AiqmcSystem* sys, const double* params, const double* pos, int64_t B, const double* alpha,
                        const double* beta, double* grad_out, double* phase, double* logabs, void* ws, int64_t ws_bytes,
                        cudaStream_t st) {
    if constexpr (kPgrad) {
      constexpr int total = make_layout(NE, NA).total;
      if (B <= 0) {
        AQ_CUDA_OK(cudaMemsetAsync(grad_out, 0, total * sizeof(double), st));
        return AIQMC_OK;
      }
      if (!ws || ws_bytes < pgrad_ws_bytes(NE, NA, B)) return AIQMC_E_WORKSPACE;
      const int nblk = (int)pgrad_blocks(B);
      AQ_CUDA_OK(cudaFuncSetAttribute(k_param_grad<NE, NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kSmem));
      g_launch_count += 2;
      k_param_grad<NE, NA><<<nblk, 32, 2 * kSmem, st>>>(*sys, params, pos, B, alpha, beta, phase, logabs, (double*)ws);
      k_reduce_param_partials<<<(total + 255) / 256, 256, 0, st>>>((const double*)ws, nblk, total, grad_out);
      AQ_CUDA_OK(cudaGetLastError());
      return AIQMC_OK;
    } else {
      // N > 16: primal pass into the derivative cache chunk by chunk, the sweep on the cache, one reduction at the end
      constexpr int total = make_layout(NE, NA).total;
      if (B <= 0) {
        AQ_CUDA_OK(cudaMemsetAsync(grad_out, 0, total * sizeof(double), st));
        return AIQMC_OK;
      }
      if (!ws || ws_bytes < pgrad_ws_bytes(NE, NA, B)) return AIQMC_E_WORKSPACE;
      double* partials = (double*)ws;
      double* dcache = (double*)((char*)ws + align256(pgrad_cached_rows(NE, NA, B) * total * 8));
      double* la_tmp = (double*)((char*)dcache + deriv_cache_bytes(NE, NA, false, B));
      AQ_CUDA_OK(prep(k_primal<NE, NA, false, 0>));
      AQ_CUDA_OK(cudaFuncSetAttribute(k_param_grad_cached<NE, NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kSmem));
      const int64_t chunk = deriv_chunk(NE, NA, false);
      MovedSrc ms{};
      int rows = 0;
      for (int64_t c0 = 0; c0 < B; c0 += chunk) {
        const int64_t nc = (B - c0 < chunk) ? B - c0 : chunk;
        const unsigned gp = (unsigned)((nc + kThreads - 1) / kThreads);
        int64_t nb = (nc + 31) / 32;
        if (nb > kPgradCachedBlocks) nb = kPgradCachedBlocks;
        g_launch_count += 2;
        k_primal<NE, NA, false, 0><<<gp, kThreads, kSmem, st>>>(*sys, params, pos, c0, nc, ms, dcache, nullptr, phase,
                                                                 logabs ? logabs : la_tmp);
        k_param_grad_cached<NE, NA><<<(unsigned)nb, 32, 2 * kSmem, st>>>(*sys, params, dcache, c0, nc, alpha, beta,
                                                                          partials + (int64_t)rows * total);
        rows += (int)nb;
      }
      ++g_launch_count;
      k_reduce_param_partials<<<(total + 255) / 256, 256, 0, st>>>(partials, rows, total, grad_out);
      AQ_CUDA_OK(cudaGetLastError());
      return AIQMC_OK;
    }
  }

  static const OpsTable* table() {
    static const OpsTable t = {NE, NA, &Launch::psi, &Launch::sweep, &Launch::energy, &Launch::tmove, &Launch::param_grad};
    return &t;
  }
};

}  // namespace aiqmc
