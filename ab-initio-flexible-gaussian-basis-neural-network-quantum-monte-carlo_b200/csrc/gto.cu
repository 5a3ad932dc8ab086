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
// Tiled variant: a CTA owns 128 consecutive points.  The coordinates arrive by coalesced loads through shared memory,
// every thread evaluates its point into a shared-memory tile [point][AO] (rows of odd length: conflict-free), and the
// tile -- a CONTIGUOUS block of val / grad / lap in HBM because the outputs are point-major -- leaves by coalesced
// 16-byte stores.  The direct kernel above writes 65 doubles per thread at a stride of nao*8 bytes between lanes (one
// 32-byte sector per store): 24 B in + 520 B out per point at C/cc-pVDZ ran at a few hundred GB/s.
constexpr int kGtoTile = 128;
__global__ void __launch_bounds__(kGtoTile) k_gto_eval_tiled(const double* __restrict__ points, int64_t n, int nao,
                                                             double* __restrict__ val, double* __restrict__ grad,
                                                             double* __restrict__ lap) {
  extern __shared__ __align__(16) double s_gto[];
  double* spts = s_gto;                                 // [128][3]
  double* sval = spts + 3 * kGtoTile;                   // [128][nao]
  double* sgrad = sval + kGtoTile * nao;                // [128][nao][3]
  double* slap = sgrad + 3 * kGtoTile * nao;            // [128][nao]
  if (threadIdx.x < kExpTab) g_exp_tab[threadIdx.x] = exp2((double)threadIdx.x * (1.0 / kExpTab));
  const int64_t t0 = (int64_t)blockIdx.x * kGtoTile;
  const int np = (int)((n - t0) < kGtoTile ? (n - t0) : kGtoTile);
  for (int q = threadIdx.x; q < 3 * np; q += kGtoTile) spts[q] = points[3 * t0 + q];
  __syncthreads();
  const double* tab = g_exp_tab;
  if ((int)threadIdx.x < np) {
    const double x = spts[3 * threadIdx.x], y = spts[3 * threadIdx.x + 1], z = spts[3 * threadIdx.x + 2];
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
        case 0: shell_out<0>(sh, dx, dy, dz, f, f1, f2, d2, threadIdx.x, nao, sval, grad ? sgrad : nullptr, lap ? slap : nullptr); break;
        case 1: shell_out<1>(sh, dx, dy, dz, f, f1, f2, d2, threadIdx.x, nao, sval, grad ? sgrad : nullptr, lap ? slap : nullptr); break;
        case 2: shell_out<2>(sh, dx, dy, dz, f, f1, f2, d2, threadIdx.x, nao, sval, grad ? sgrad : nullptr, lap ? slap : nullptr); break;
        default: shell_out<3>(sh, dx, dy, dz, f, f1, f2, d2, threadIdx.x, nao, sval, grad ? sgrad : nullptr, lap ? slap : nullptr); break;
      }
    }
  }
  __syncthreads();
  auto stream_out = [&](const double* src, double* dst, int64_t count) {        // dst + t0*... is 16-byte aligned when count is even
    if ((count & 1) == 0 && (((uintptr_t)dst) & 15) == 0) {
      const double2* s2 = reinterpret_cast<const double2*>(src);
      double2* d2p = reinterpret_cast<double2*>(dst);
      for (int64_t q = threadIdx.x; q < count / 2; q += kGtoTile) d2p[q] = s2[q];
    } else {
      for (int64_t q = threadIdx.x; q < count; q += kGtoTile) dst[q] = src[q];
    }
  };
  stream_out(sval, val + t0 * nao, (int64_t)np * nao);
  if (grad) stream_out(sgrad, grad + t0 * nao * 3, (int64_t)np * nao * 3);
  if (lap) stream_out(slap, lap + t0 * nao, (int64_t)np * nao);
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
  const size_t smem = (size_t)(3 * kGtoTile + 5 * kGtoTile * nao) * sizeof(double);
  if (smem <= 200 * 1024) {              // the output tile fits shared memory: coalesced stores
    const cudaError_t ea = cudaFuncSetAttribute(k_gto_eval_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ea != cudaSuccess) { g_last_cuda_error = (int)ea; return AIQMC_E_CUDA; }
    k_gto_eval_tiled<<<(unsigned)((n_points + kGtoTile - 1) / kGtoTile), kGtoTile, smem, st>>>(points, n_points, nao, val, grad, lap);
  } else {
    k_gto_eval<<<(unsigned)((n_points + 127) / 128), 128, 0, st>>>(points, n_points, nao, val, grad, lap);
  }
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_last_cuda_error = (int)e; return AIQMC_E_CUDA; }
  return AIQMC_OK;
}
