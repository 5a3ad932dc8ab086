// rng.cu -- counter-based random inputs of the walker path, generated on the device ("throughput mode",
// SURVEY.md section 8b "Ownership").  The reference draws inside its jitted graph:
//   gauss1 (B,3N), gauss2 (B,N,3N)   jax.random.normal   VMC/VMCmcstep.py:58,83
//   rnd (B,N)                        jax.random.uniform  VMC/VMCmcstep.py:19-20
//   rot (3,3) per walker             jax.random.orthogonal  pseudopotential/pseudopotential.py:233-235
//   T-move u (B), rnd (B,N)          DMC/Tmoves.py:146,216-217
// jax's threefry streams cannot be reproduced without jax ("parity unpinned" for RNG streams; parity mode passes
// explicit arrays instead).  Here every draw is Philox4x32-10 (Salmon et al., SC'11) keyed by the 64-bit seed with
// counter (walker id lo, walker id hi, step, slot): a walker's numbers depend only on (seed, GLOBAL walker id, step),
// never on the batch size or on how walkers are sharded over GPUs.  oracle/philox.py restates the same mapping in
// numpy (uniforms and integer streams bit-exact, normals to libm rounding).
//
// Of gauss2 only the diagonal 3-blocks gauss2[b,i,3i:3i+3] are ever read (VMCmcstep.py:86-94 indexes [b,i,i,d]); the
// generator therefore writes the compact (B,N,3) form that aiqmc_vmc_sweep_compact consumes: 3/4 of the reference's
// normal draws for that array are never made.
#include <cuda_runtime.h>
#include <atomic>
#include <math.h>
#include <stdint.h>

#include "../../include/aiqmc_b200.h"

namespace aiqmc {
extern std::atomic<int> g_last_cuda_error;
extern std::atomic<int64_t> g_launch_count;

struct U4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c.x, p1 = (uint64_t)M1 * c.z;
    const U4 n = {(uint32_t)(p1 >> 32) ^ c.y ^ k0, (uint32_t)p1, (uint32_t)(p0 >> 32) ^ c.w ^ k1, (uint32_t)p0};
    c = n;
    k0 += W0;
    k1 += W1;
  }
  return c;
}

// 53-bit uniform in [0,1) from two words
__host__ __device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
  return (double)(((uint64_t)(hi >> 5) << 26) | (uint64_t)(lo >> 6)) * (1.0 / 9007199254740992.0);
}
// Box-Muller pair from four words; the radius uniform is shifted to (0,1]
__device__ __forceinline__ void normal_pair(const U4& r, double& a, double& b) {
  const double u1 = 1.0 - u53(r.x, r.y), u2 = u53(r.z, r.w);
  const double rad = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  a = rad * c;
  b = rad * s;
}

// slot layout of the counter's 4th word
constexpr uint32_t kSlotSweep = 0u;        // 4*i + c, c = 0..3
constexpr uint32_t kSlotRot = 0x10000u;    // + c, c = 0..4
constexpr uint32_t kSlotUniform = 0x20000u;  // + tag * 0x1000 + col

// 1 thread = (walker, electron): 6 normals (3 of gauss1, 3 of the gauss2 diagonal block) and one uniform
__global__ void k_rng_sweep(uint64_t seed, uint32_t step, int64_t walker0, int64_t B, int n, double scale,
                            double* __restrict__ gauss1, double* __restrict__ gauss2c, double* __restrict__ rnd) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * n) return;
  const int64_t b = t / n;
  const int i = (int)(t - b * n);
  const uint64_t w = (uint64_t)(walker0 + b);
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  const U4 base = {(uint32_t)w, (uint32_t)(w >> 32), step, kSlotSweep + 4u * (uint32_t)i};
  double v[6];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    U4 ctr = base;
    ctr.w += (uint32_t)c;
    normal_pair(philox4x32_10(ctr, k0, k1), v[2 * c], v[2 * c + 1]);
  }
  U4 ctr = base;
  ctr.w += 3u;
  const U4 r = philox4x32_10(ctr, k0, k1);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    gauss1[b * 3 * n + 3 * i + c] = scale * v[c];
    gauss2c[t * 3 + c] = scale * v[3 + c];
  }
  rnd[t] = u53(r.x, r.y);
}

// 1 thread = walker: 9 normals -> modified Gram-Schmidt QR -> Q * sign(diag R) (Haar measure; the convention of
// numpy.linalg.qr-based stand-ins for jax.random.orthogonal)
__global__ void k_rng_rot(uint64_t seed, uint32_t step, int64_t walker0, int64_t B, double* __restrict__ rot) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const uint64_t w = (uint64_t)(walker0 + b);
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  double g[10];
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    const U4 ctr = {(uint32_t)w, (uint32_t)(w >> 32), step, kSlotRot + (uint32_t)c};
    normal_pair(philox4x32_10(ctr, k0, k1), g[2 * c], g[2 * c + 1]);
  }
  // columns a_j = g[row*3 + j]; Gram-Schmidt on columns; with R's diagonal positive by construction the result
  // equals Q * sign(diag R) of a Householder QR
  double q[3][3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    double v0 = g[j], v1 = g[3 + j], v2 = g[6 + j];
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      if (p < j) {
        const double d = q[0][p] * v0 + q[1][p] * v1 + q[2][p] * v2;
        v0 -= d * q[0][p]; v1 -= d * q[1][p]; v2 -= d * q[2][p];
      }
    }
    const double inv = 1.0 / sqrt(v0 * v0 + v1 * v1 + v2 * v2);
    q[0][j] = v0 * inv; q[1][j] = v1 * inv; q[2][j] = v2 * inv;
  }
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) rot[b * 9 + 3 * r + c] = q[r][c];
}

// 1 thread = (walker, column): one uniform in [0,1)
__global__ void k_rng_uniform(uint64_t seed, uint32_t step, int64_t walker0, int64_t B, int cols, uint32_t tag,
                              double* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * cols) return;
  const int64_t b = t / cols;
  const int c = (int)(t - b * cols);
  const uint64_t w = (uint64_t)(walker0 + b);
  const U4 ctr = {(uint32_t)w, (uint32_t)(w >> 32), step, kSlotUniform + tag * 0x1000u + (uint32_t)c};
  const U4 r = philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
  out[t] = u53(r.x, r.y);
}
}  // namespace aiqmc

using aiqmc::g_last_cuda_error;
using aiqmc::g_launch_count;
#define AQ_RNG_OK()                                                                      \
  do {                                                                                   \
    cudaError_t e_ = cudaGetLastError();                                                 \
    if (e_ != cudaSuccess) { g_last_cuda_error = (int)e_; return AIQMC_E_CUDA; }         \
  } while (0)

extern "C" {

int aiqmc_rng_sweep(uint64_t seed, uint32_t step, int64_t walker0, int64_t n_walkers, int32_t n_elec, double tstep,
                    double* gauss1, double* gauss2c, double* rnd, void* stream) {
  if (n_walkers < 0 || walker0 < 0 || n_elec < 1 || n_elec > AIQMC_MAX_ELEC || !(tstep > 0.0)) return AIQMC_E_BADARG;
  if (n_walkers == 0) return AIQMC_OK;
  if (!gauss1 || !gauss2c || !rnd) return AIQMC_E_BADARG;
  const int64_t nt = n_walkers * n_elec;
  ++g_launch_count;
  aiqmc::k_rng_sweep<<<(unsigned)((nt + 255) / 256), 256, 0, (cudaStream_t)stream>>>(seed, step, walker0, n_walkers, n_elec,
                                                                                     sqrt(tstep), gauss1, gauss2c, rnd);
  AQ_RNG_OK();
  return AIQMC_OK;
}

int aiqmc_rng_rotations(uint64_t seed, uint32_t step, int64_t walker0, int64_t n_walkers, double* rot, void* stream) {
  if (n_walkers < 0 || walker0 < 0) return AIQMC_E_BADARG;
  if (n_walkers == 0) return AIQMC_OK;
  if (!rot) return AIQMC_E_BADARG;
  ++g_launch_count;
  aiqmc::k_rng_rot<<<(unsigned)((n_walkers + 255) / 256), 256, 0, (cudaStream_t)stream>>>(seed, step, walker0, n_walkers, rot);
  AQ_RNG_OK();
  return AIQMC_OK;
}

int aiqmc_rng_uniform(uint64_t seed, uint32_t step, int64_t walker0, int64_t n_walkers, int32_t cols, uint32_t tag,
                      double* out, void* stream) {
  if (n_walkers < 0 || walker0 < 0 || cols < 1 || cols > 0x1000 || tag > 0xf) return AIQMC_E_BADARG;
  if (n_walkers == 0) return AIQMC_OK;
  if (!out) return AIQMC_E_BADARG;
  const int64_t nt = n_walkers * cols;
  ++g_launch_count;
  aiqmc::k_rng_uniform<<<(unsigned)((nt + 255) / 256), 256, 0, (cudaStream_t)stream>>>(seed, step, walker0, n_walkers, cols, tag, out);
  AQ_RNG_OK();
  return AIQMC_OK;
}

}  // extern "C"
