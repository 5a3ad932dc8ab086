// fastmath.cuh -- double-precision exp / tanh for the FP64-pipe-bound kernels.
//
// tanh and exp are ~75 % of the FP64 instructions of a psi evaluation (308 tanh at N=4), so
// they are hand-written instead of calling libm:
//   fexp(x):  x = (64k + j) ln2/64 + r, |r| <= ln2/128;  exp(x) = 2^k * T[j] * (1 + r + ... + r^5/120)
//             T[j] = 2^(j/64) from a 64-entry table in shared memory (LDS, off the FP64 pipe).
//             ~9 FP64 ops.  Max relative error measured against libm: < 3e-16 (tests/test_fastmath.py).
//   ftanh(x): (1 - e) / (1 + e), e = fexp(-2|x|); reciprocal from an FP32 MUFU.RCP seed + two
//             Newton steps.  ~17 FP64 ops.  Max ABSOLUTE error < 3e-16; relative accuracy is lost for
//             |x| < 1e-8 by design (only absolute accuracy enters log|psi| and E_L).
// Both are __host__ __device__ so the host test build exercises the same arithmetic.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define AQF_HD __host__ __device__ __forceinline__
#else
#define AQF_HD inline
#endif

namespace aiqmc {

constexpr int kExpTab = 64;

AQF_HD double fexp(double x, const double* __restrict__ tab) {
  x = fmin(fmax(x, -700.0), 700.0);
  const double kMagic = 6755399441055744.0;              // 1.5 * 2^52: rint via add
  const double kInv = 92.332482616893656877;             // 64 / ln2
  const double kLn2Hi = 1.0830424696223417675e-02;       // ln2/64 head (trailing bits zero)
  const double kLn2Lo = 2.5728046223276688017e-14;       // ln2/64 tail
  double t = fma(x, kInv, kMagic);
  int64_t bits;
#ifdef __CUDA_ARCH__
  bits = __double_as_longlong(t);
#else
  memcpy(&bits, &t, 8);
#endif
  const int n = (int)(uint32_t)bits;                     // low 32 bits hold rint(x*64/ln2)
  const double nf = t - kMagic;
  double r = fma(nf, -kLn2Hi, x);
  r = fma(nf, -kLn2Lo, r);
  double q = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  q = fma(r, q, 1.0 / 6.0);
  q = fma(r, q, 0.5);
  q = fma(r, q, 1.0);
  q = q * r;                                             // exp(r) - 1
  const double tj = tab[n & (kExpTab - 1)];
  double res = fma(tj, q, tj);
  const int k = n >> 6;
#ifdef __CUDA_ARCH__
  res = __hiloint2double(__double2hiint(res) + (k << 20), __double2loint(res));
#else
  res = ldexp(res, k);
#endif
  return res;
}

AQF_HD double frcp12(double d) {                         // 1/d for d in [1,2]
#ifdef __CUDA_ARCH__
  double y = (double)__frcp_rn((float)d);
#else
  double y = (double)(1.0f / (float)d);
#endif
  double e = fma(-d, y, 1.0);
  y = fma(y, e, y);
  e = fma(-d, y, 1.0);
  y = fma(y, e, y);
  return y;
}

AQF_HD double ftanh(double x, const double* __restrict__ tab) {
  const double a = fmin(fabs(x), 20.0);
  const double e = fexp(-2.0 * a, tab);
  const double t = (1.0 - e) * frcp12(1.0 + e);
  return copysign(t, x);
}

// host-side table (also used to fill the per-CTA shared copy)
inline const double* host_exp_table() {
  static double tab[kExpTab];
  static bool init = false;
  if (!init) {
    for (int j = 0; j < kExpTab; ++j) tab[j] = exp2((double)j / kExpTab);
    init = true;
  }
  return tab;
}

}  // namespace aiqmc
