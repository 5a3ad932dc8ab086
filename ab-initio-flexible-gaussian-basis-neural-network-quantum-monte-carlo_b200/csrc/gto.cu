// gto.cu -- contracted Gaussian primitives times real solid harmonics with analytic gradient and Laplacian
// (AIQMC/Gaussian_orbitals.py:11-13 evaluated one point at a time in a Python loop; conventions of
// ferminet/utils/gto.py:117-135,338-389).
//
//   AO(r) = f(d2) S_lm(d),  d = r - R,  d2 = |d|^2,  f = sum_p c_p exp(-a_p d2),  S_lm = |d|^l Y_lm (harmonic polynomial)
//   grad AO = f grad S + 2 f' S d                       f'  = df/d(d2)  = -sum a_p c_p exp(-a_p d2)
//   lap  AO = S (4 l f' + 6 f' + 4 f'' d2)              f'' = sum a_p^2 c_p exp(-a_p d2);  lap S = 0, d.grad S = l S
//
// One thread per point; the shell table sits in constant memory, so exponents / coefficients are broadcast operands
// and the loop over shells is uniform across the warp.  HBM-bound: 24 B in, 40 B out per (point, AO); exponentials
// are the in-house fexp (10 FP64 ops) shared by value / gradient / Laplacian of a primitive.
#include <cuda_runtime.h>
#include <atomic>
#include <mutex>
#include <string.h>
#include <type_traits>

#include "../../include/aiqmc_b200.h"
#include "fastmath.cuh"

namespace aiqmc {
extern std::atomic<int> g_last_cuda_error;
extern std::atomic<int64_t> g_launch_count;

struct GtoTable {
  int32_t n_shells, n_centres;
  AiqmcGtoShell shells[AIQMC_GTO_MAX_SHELLS];
  double centres[AIQMC_GTO_MAX_CENTRES][3];
};
static __constant__ GtoTable c_gto;

// real solid harmonics S_lm(x,y,z) and their gradients, m = -l..l, l <= 3 (closed forms of gto.py:117-135)
template <int L>
__device__ __forceinline__ void solid_harmonics(double x, double y, double z, double* __restrict__ s, double (*g)[3]) {
  if (L == 0) {
    s[0] = 0.28209479177387814;
    g[0][0] = g[0][1] = g[0][2] = 0.0;
  } else if (L == 1) {
    const double c = 0.48860251190291992;
    s[0] = c * y; s[1] = c * z; s[2] = c * x;
    g[0][0] = 0; g[0][1] = c; g[0][2] = 0;
    g[1][0] = 0; g[1][1] = 0; g[1][2] = c;
    g[2][0] = c; g[2][1] = 0; g[2][2] = 0;
  } else if (L == 2) {
    const double a = 1.0925484305920792, b = 0.31539156525252005, c = 0.54627421529603959;
    s[0] = a * x * y;              g[0][0] = a * y; g[0][1] = a * x; g[0][2] = 0;
    s[1] = a * y * z;              g[1][0] = 0; g[1][1] = a * z; g[1][2] = a * y;
    s[2] = b * (2 * z * z - x * x - y * y); g[2][0] = -2 * b * x; g[2][1] = -2 * b * y; g[2][2] = 4 * b * z;
    s[3] = a * x * z;              g[3][0] = a * z; g[3][1] = 0; g[3][2] = a * x;
    s[4] = c * (x * x - y * y);    g[4][0] = 2 * c * x; g[4][1] = -2 * c * y; g[4][2] = 0;
  } else {
    const double a = 0.59004358992664352;   // 1/4 sqrt(35/(2 pi))
    const double b = 2.8906114426405538;    // 1/2 sqrt(105/pi)
    const double c = 0.45704579946446577;   // 1/4 sqrt(21/(2 pi))
    const double d = 0.37317633259011546;   // 1/4 sqrt(7/pi)
    const double e = 1.4453057213202769;    // 1/4 sqrt(105/pi)
    const double x2 = x * x, y2 = y * y, z2 = z * z;
    s[0] = a * y * (3 * x2 - y2);        g[0][0] = 6 * a * x * y; g[0][1] = a * (3 * x2 - 3 * y2); g[0][2] = 0;
    s[1] = b * x * y * z;                g[1][0] = b * y * z; g[1][1] = b * x * z; g[1][2] = b * x * y;
    s[2] = c * y * (4 * z2 - x2 - y2);   g[2][0] = -2 * c * x * y; g[2][1] = c * (4 * z2 - x2 - 3 * y2); g[2][2] = 8 * c * y * z;
    s[3] = d * z * (2 * z2 - 3 * x2 - 3 * y2); g[3][0] = -6 * d * x * z; g[3][1] = -6 * d * y * z; g[3][2] = d * (6 * z2 - 3 * x2 - 3 * y2);
    s[4] = c * x * (4 * z2 - x2 - y2);   g[4][0] = c * (4 * z2 - 3 * x2 - y2); g[4][1] = -2 * c * x * y; g[4][2] = 8 * c * x * z;
    s[5] = e * z * (x2 - y2);            g[5][0] = 2 * e * x * z; g[5][1] = -2 * e * y * z; g[5][2] = e * (x2 - y2);
    s[6] = a * x * (x2 - 3 * y2);        g[6][0] = a * (3 * x2 - 3 * y2); g[6][1] = -6 * a * x * y; g[6][2] = 0;
  }
}

// radial sums f, f', f'' of one shell at squared distance d2 (arguments of the exponentials are <= 0: one-sided clamp)
template <int UNR>
__device__ __forceinline__ void shell_radial(const AiqmcGtoShell& sh, double d2, const double* __restrict__ tab, double& f,
                                             double& f1, double& f2) {
  f = 0.0; f1 = 0.0; f2 = 0.0;
#pragma unroll UNR
  for (int p = 0; p < sh.n_prim; ++p) {
    const double a = sh.alpha[p];
    const double e = sh.coef[p] * fexp_nonpos(-a * d2, tab);
    f += e;
    f1 -= a * e;
    f2 += a * a * e;
  }
}

template <int L>
__device__ __forceinline__ void shell_out(const AiqmcGtoShell& sh, double dx, double dy, double dz, double f, double f1,
                                          double f2, double d2, int64_t t, int nao, double* __restrict__ val,
                                          double* __restrict__ grad, double* __restrict__ lap) {
  constexpr int M = 2 * L + 1;
  double s[M], g[M][3];
  solid_harmonics<L>(dx, dy, dz, s, g);
  const double lf = (4.0 * L + 6.0) * f1 + 4.0 * f2 * d2;
#pragma unroll
  for (int m = 0; m < M; ++m) {
    const int64_t col = t * nao + sh.ao_offset + m;
    val[col] = f * s[m];
    if (grad) {
      grad[col * 3 + 0] = f * g[m][0] + 2.0 * f1 * s[m] * dx;
      grad[col * 3 + 1] = f * g[m][1] + 2.0 * f1 * s[m] * dy;
      grad[col * 3 + 2] = f * g[m][2] + 2.0 * f1 * s[m] * dz;
    }
    if (lap) lap[col] = s[m] * lf;
  }
}

__global__ void __launch_bounds__(128) k_gto_eval(const double* __restrict__ points, int64_t n, int nao,
                                                  double* __restrict__ val, double* __restrict__ grad,
                                                  double* __restrict__ lap) {
  if (threadIdx.x < kExpTab) g_exp_tab[threadIdx.x] = exp2((double)threadIdx.x * (1.0 / kExpTab));
  __syncthreads();
  const double* tab = g_exp_tab;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const double x = points[3 * t], y = points[3 * t + 1], z = points[3 * t + 2];
  for (int sidx = 0; sidx < c_gto.n_shells; ++sidx) {
    const AiqmcGtoShell& sh = c_gto.shells[sidx];
    const double dx = x - c_gto.centres[sh.centre][0], dy = y - c_gto.centres[sh.centre][1],
                 dz = z - c_gto.centres[sh.centre][2];
    const double d2 = dx * dx + dy * dy + dz * dz;
    double f = 0.0, f1 = 0.0, f2 = 0.0;
    for (int p = 0; p < sh.n_prim; ++p) {
      const double a = sh.alpha[p];
      const double e = sh.coef[p] * fexp(-a * d2, tab);
      f += e;
      f1 -= a * e;
      f2 += a * a * e;
    }
    switch (sh.l) {            // uniform across the grid
      case 0: shell_out<0>(sh, dx, dy, dz, f, f1, f2, d2, t, nao, val, grad, lap); break;
      case 1: shell_out<1>(sh, dx, dy, dz, f, f1, f2, d2, t, nao, val, grad, lap); break;
      case 2: shell_out<2>(sh, dx, dy, dz, f, f1, f2, d2, t, nao, val, grad, lap); break;
      default: shell_out<3>(sh, dx, dy, dz, f, f1, f2, d2, t, nao, val, grad, lap); break;
    }
  }
}
// Streaming kernel (the default): PERSISTENT warps, each looping over 32-point tiles with its own shared-memory tile
// [point][AO] (rows of odd length: conflict-free).  The tile is a CONTIGUOUS block of val / grad / lap in HBM because the
// outputs are point-major, so it leaves by three TMA bulk stores (cp.async.bulk shared -> global) issued by one lane.
// The direct kernel above writes 65 doubles per thread at a stride of nao*8 bytes between lanes (one 32-byte sector per
// store): 24 B in + 520 B out per point at C/cc-pVDZ ran at a few hundred GB/s.  History: a CTA-per-128-point-tile
// kernel with per-thread 16-byte stores reached 2.9 TB/s, with bulk stores and 32-point tiles 3.7 TB/s; ncu on it: 34 % of the stall samples sat on the first instruction that needs the coordinates (a
// DRAM round trip queued behind 0.76 GB of stores) and 9 % on the bulk-store drain before the CTA could retire.  Here the
// NEXT tile's coordinates are fetched into registers before the current tile is evaluated, and a warp waits for its
// previous bulk store only when it is about to overwrite the tile -- after the next coordinates have been requested.
template <int UNR>
__global__ void __launch_bounds__(128) k_gto_eval_stream(const double* __restrict__ points, int64_t n, int nao, int warps,
                                                         double* __restrict__ val, double* __restrict__ grad,
                                                         double* __restrict__ lap) {
  extern __shared__ __align__(128) double s_gto[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < kExpTab) g_exp_tab[threadIdx.x] = exp2((double)threadIdx.x * (1.0 / kExpTab));
  __syncthreads();
  if (warp >= warps) return;
  const double* tab = g_exp_tab;
  const int per_warp = ((160 * nao + 96) + 15) & ~15;   // doubles; keeps every warp's tile 128-byte aligned
  double* sval = s_gto + (size_t)warp * per_warp;       // [32][nao]
  double* sgrad = sval + 32 * nao;                      // [32][nao][3]
  double* slap = sgrad + 96 * nao;                      // [32][nao]
  double* spts = slap + 32 * nao;                       // [32][3]
  const int64_t ntiles = (n + 31) / 32;
  const int64_t stride = (int64_t)gridDim.x * warps;
  int64_t tile = (int64_t)blockIdx.x * warps + warp;
  const bool aligned = ((((uintptr_t)val) | ((uintptr_t)grad) | ((uintptr_t)lap)) & 15) == 0 && (nao * 32 % 2 == 0);
  double pr[3];
  auto fetch = [&](int64_t t) {
    const int64_t base = 96 * t;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int64_t q = base + 32 * c + lane;
      pr[c] = (t < ntiles && q < 3 * n) ? points[q] : 0.0;
    }
  };
  fetch(tile);
  bool pending = false;
  for (; tile < ntiles; tile += stride) {
    const int64_t t0 = tile * 32;
    const int np = (int)((n - t0) < 32 ? (n - t0) : 32);
    if (pending) {                                         // the previous tile must have left before it is overwritten
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      pending = false;
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 3; ++c) spts[32 * c + lane] = pr[c];
    __syncwarp();
    const double x = spts[3 * lane], y = spts[3 * lane + 1], z = spts[3 * lane + 2];
    fetch(tile + stride);
    if (lane < np) {
      for (int sidx = 0; sidx < c_gto.n_shells; ++sidx) {
        const AiqmcGtoShell& sh = c_gto.shells[sidx];
        const double dx = x - c_gto.centres[sh.centre][0], dy = y - c_gto.centres[sh.centre][1],
                     dz = z - c_gto.centres[sh.centre][2];
        const double d2 = dx * dx + dy * dy + dz * dz;
        double f, f1, f2;
        shell_radial<UNR>(sh, d2, tab, f, f1, f2);
        switch (sh.l) {            // uniform across the grid
          case 0: shell_out<0>(sh, dx, dy, dz, f, f1, f2, d2, lane, nao, sval, grad ? sgrad : nullptr, lap ? slap : nullptr); break;
          case 1: shell_out<1>(sh, dx, dy, dz, f, f1, f2, d2, lane, nao, sval, grad ? sgrad : nullptr, lap ? slap : nullptr); break;
          case 2: shell_out<2>(sh, dx, dy, dz, f, f1, f2, d2, lane, nao, sval, grad ? sgrad : nullptr, lap ? slap : nullptr); break;
          default: shell_out<3>(sh, dx, dy, dz, f, f1, f2, d2, lane, nao, sval, grad ? sgrad : nullptr, lap ? slap : nullptr); break;
        }
      }
    }
    if (aligned && np == 32) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy tile writes -> visible to the TMA engine
      __syncwarp();
      if (lane == 0) {
        auto bulk = [&](const double* src, double* dst, int count) {
          const uint32_t s32 = (uint32_t)__cvta_generic_to_shared(src);
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(s32), "r"((uint32_t)(count * 8))
                       : "memory");
        };
        bulk(sval, val + t0 * nao, 32 * nao);
        if (grad) bulk(sgrad, grad + t0 * nao * 3, 96 * nao);
        if (lap) bulk(slap, lap + t0 * nao, 32 * nao);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      pending = true;
    } else {                                               // ragged last tile / unaligned outputs: plain stores
      __syncwarp();
      for (int q = lane; q < np * nao; q += 32) val[t0 * nao + q] = sval[q];
      if (grad) for (int q = lane; q < 3 * np * nao; q += 32) grad[t0 * nao * 3 + q] = sgrad[q];
      if (lap) for (int q = lane; q < np * nao; q += 32) lap[t0 * nao + q] = slap[q];
      __syncwarp();
    }
  }
  if (pending && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
}  // namespace aiqmc

extern "C" int aiqmc_gto_eval(const AiqmcGtoShell* shells, int32_t n_shells, const double* centres, int32_t n_centres,
                              const double* points, int64_t n_points, int32_t nao, double* val, double* grad,
                              double* lap, void* stream) {
  using namespace aiqmc;
  if (!shells || !centres || n_shells < 1 || n_shells > AIQMC_GTO_MAX_SHELLS || n_centres < 1 ||
      n_centres > AIQMC_GTO_MAX_CENTRES || n_points < 0 || nao < 1)
    return AIQMC_E_BADARG;
  if (n_points > 0 && (!points || !val)) return AIQMC_E_BADARG;
  static GtoTable h[64];                 // what each DEVICE's constant copy holds (a second GPU re-uploads)
  static bool valid[64];
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return AIQMC_E_CUDA;
  GtoTable t;
  memset(&t, 0, sizeof(t));
  t.n_shells = n_shells;
  t.n_centres = n_centres;
  for (int s = 0; s < n_shells; ++s) {
    const AiqmcGtoShell& sh = shells[s];
    if (sh.l < 0 || sh.l > 3 || sh.n_prim < 1 || sh.n_prim > AIQMC_GTO_MAX_PRIM || sh.centre < 0 ||
        sh.centre >= n_centres || sh.ao_offset < 0 || sh.ao_offset + 2 * sh.l + 1 > nao)
      return AIQMC_E_BADARG;
    t.shells[s] = sh;
  }
  memcpy(t.centres, centres, sizeof(double) * 3 * n_centres);
  if (n_points == 0) return AIQMC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (!valid[dev] || memcmp(&h[dev], &t, sizeof(t)) != 0) {
    const cudaError_t e = cudaMemcpyToSymbolAsync(c_gto, &t, sizeof(t), 0, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; return AIQMC_E_CUDA; }
    cudaStreamSynchronize(st);       // `t` is a stack object: the copy must have read it before we return
    h[dev] = t;
    valid[dev] = true;
  }
  ++g_launch_count;
  auto go_stream = [&](auto unr_c) -> int {
    constexpr int kUnr = decltype(unr_c)::value;
    const size_t per_warp = (size_t)(((160 * nao + 96) + 15) & ~15) * sizeof(double);
    int warps = (int)((200 * 1024) / per_warp);
    if (warps < 1) return 1;                              // the tile of one warp does not fit: direct kernel
    warps = warps > 4 ? 4 : warps;
    const size_t smem = per_warp * warps;
    int sms = 148, per_sm = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const cudaError_t ea = cudaFuncSetAttribute(k_gto_eval_stream<kUnr>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ea != cudaSuccess) { g_last_cuda_error = (int)ea; return AIQMC_E_CUDA; }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gto_eval_stream<kUnr>, 128, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    const int64_t ntiles = (n_points + 31) / 32;
    int64_t grid = (int64_t)sms * per_sm;                 // persistent: every resident warp loops over tiles
    if (grid * warps > ntiles) grid = (ntiles + warps - 1) / warps;
    k_gto_eval_stream<kUnr><<<(unsigned)grid, 128, smem, st>>>(points, n_points, nao, warps, val, grad, lap);
    return 0;
  };
  // short runs are latency-bound (4 primitives in flight per thread: 46 vs 52 us at 393,216 points), long runs are
  // issue-bound (rolled loop: 544 vs 617 us at 6.3 M points)
  int rc = n_points <= (1 << 20) ? go_stream(std::integral_constant<int, 4>{}) : go_stream(std::integral_constant<int, 1>{});
  if (rc < 0) return rc;
  if (rc == 1) k_gto_eval<<<(unsigned)((n_points + 127) / 128), 128, 0, st>>>(points, n_points, nao, val, grad, lap);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_last_cuda_error = (int)e; return AIQMC_E_CUDA; }
  return AIQMC_OK;
}
