// fastmath.cuh -- double-precision exp / tanh / reciprocal / rsqrt for the FP64-pipe-bound kernels.
//
// tanh is ~60 % of the FP64 instructions of a psi evaluation (174 tanh per single-electron move at
// N=4), so the transcendental bodies are hand-written to a fixed FP64-instruction budget instead of
// calling libm (which also converts through FP32 and branches):
//   fexp(x):   x = (16k + j) ln2/16 + r, |r| <= ln2/32;  exp(x) = 2^k * T[j] * P6(r),  T[j] = 2^(j/16).
//              The table has 16 doubles = exactly one 128-byte row of shared-memory banks, so a warp's
//              divergent lookups are conflict-free by construction (a 64-entry table measured 4-6
//              wavefronts per LDS).  10 FP64 ops.  Relative error < 1e-15.
//   ftanh(x):  2/(1+e) - 1 with e = exp(-2|x|) (degree-5 polynomial, single-constant reduction),
//              reciprocal from the MUFU.RCP64H seed + one cubic step.  13 FP64 ops + one DMUL for -2|x| (the |.| is
//              an operand modifier; the integer-op form cost three more issue slots), sign by integer ops on the high word.  ABSOLUTE error < 2e-13 (relative accuracy is lost for
//              |x| < 1e-8 by design: only absolute accuracy enters log|psi| and E_L).
//   ftanh_n<NV, ACC>: NV of them with interleaved steps (ILP); ACC = 1 is an 11-op variant (< 5e-11).
//   frcp(d), frsqrt(x): MUFU.RCP64H / MUFU.RSQ64H seed + one cubic step (3 / 5 FP64 ops), ~1 ulp.
// Everything is __host__ __device__ so the host test build exercises the same arithmetic (the host
// replaces the MUFU seeds by a float-precision division / sqrt).
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

constexpr int kExpTab = 16;
#ifdef __CUDACC__
// 2^(j/16), filled by every kernel prologue.  Device code indexes this array directly (not through the `tab`
// pointer argument) so the lookup is one LDS with an immediate base instead of generic-pointer arithmetic.
__shared__ double g_exp_tab[kExpTab];
#endif
#ifdef __CUDA_ARCH__
#define AQF_TAB(tab, j) g_exp_tab[j]
#else
#define AQF_TAB(tab, j) (tab)[j]
#endif

// Polynomial / reduction constants.  On the device they live in constant memory so that each one is a
// c[3][imm] operand of the DFMA that uses it (as 64-bit literals every use costs two IMAD.MOVs).
struct FmK { double inv, ln2, ln2hi, ln2lo, c2, c3, c4, c5, c6; };
#define AQF_FMK_INIT {23.083120654223414, 0.04332169878499658, 0.043321698784978935, 1.7647056601894736e-14, \
                      0.5, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0}
#ifdef __CUDACC__
static __constant__ FmK c_fmk = AQF_FMK_INIT;
#endif
AQF_HD const FmK& fmk() {
#ifdef __CUDA_ARCH__
  return c_fmk;
#else
  static const FmK k = AQF_FMK_INIT;
  return k;
#endif
}

AQF_HD int32_t hi_word(double x) {
#ifdef __CUDA_ARCH__
  return __double2hiint(x);
#else
  int64_t b; memcpy(&b, &x, 8); return (int32_t)(b >> 32);
#endif
}
AQF_HD int32_t lo_word(double x) {
#ifdef __CUDA_ARCH__
  return __double2loint(x);
#else
  int64_t b; memcpy(&b, &x, 8); return (int32_t)(uint32_t)b;
#endif
}
AQF_HD double make_double(int32_t hi, int32_t lo) {
#ifdef __CUDA_ARCH__
  return __hiloint2double(hi, lo);
#else
  int64_t b = ((int64_t)hi << 32) | (uint32_t)lo; double x; memcpy(&x, &b, 8); return x;
#endif
}

// 1/d, d finite, normal and nonzero.  Seed relative error <= 2^-20 -> cubic step -> ~2^-58.
AQF_HD double frcp(double d) {
  double y;
#ifdef __CUDA_ARCH__
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
#else
  y = (double)(1.0f / (float)d);
#endif
  const double e = fma(-d, y, 1.0);
  const double t = fma(e, e, e);
  return fma(y, t, y);
}

// 1/sqrt(x), x > 0 finite normal.
AQF_HD double frsqrt(double x) {
  double y;
#ifdef __CUDA_ARCH__
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#else
  y = (double)(1.0f / sqrtf((float)x));
#endif
  const double t = x * y;
  const double e = fma(-t, y, 1.0);
  double c = fma(e, 0.375, 0.5);
  c = c * e;
  return fma(y, c, y);
}

template <bool NONPOS>
AQF_HD double fexp_t(double x, const double* __restrict__ tab) {
  if (NONPOS) x = x < -700.0 ? -700.0 : x;               // caller guarantees x <= 0: one compare + select
  else x = fmin(fmax(x, -700.0), 700.0);
  const double kMagic = 6755399441055744.0;              // 1.5 * 2^52: rint via add
  const FmK& K = fmk();                                  // inv = 16/ln2, ln2hi + ln2lo = ln2/16 (head has 12 zero bits)
  const double t = fma(x, K.inv, kMagic);
  const int n = lo_word(t);                              // low 32 bits hold rint(x*16/ln2)
  const double nf = t - kMagic;
  double r = fma(-nf, K.ln2hi, x);
  r = fma(-nf, K.ln2lo, r);
  double q = fma(r, K.c6, K.c5);
  q = fma(r, q, K.c4);
  q = fma(r, q, K.c3);
  q = fma(r, q, 0.5);
  q = fma(r, q, 1.0);
  q = q * r;                                             // exp(r) - 1
  const double tj = AQF_TAB(tab, n & (kExpTab - 1));
  const double res = fma(tj, q, tj);
  const int k = n >> 4;
#ifdef __CUDA_ARCH__
  return make_double(hi_word(res) + (k << 20), lo_word(res));
#else
  return ldexp(res, k);
#endif
}

AQF_HD double fexp(double x, const double* __restrict__ tab) { return fexp_t<false>(x, tab); }
AQF_HD double fexp_nonpos(double x, const double* __restrict__ tab) { return fexp_t<true>(x, tab); }

// n tanh evaluations with the steps interleaved in source order, so the FP64 pipe always has n independent
// dependency chains in flight (a single ftanh is a 13-deep chain of dependent DFMAs).
//   ACC = 0: degree-5 polynomial, cubic reciprocal step:     13 FP64 ops, absolute error < 2e-13
//   ACC = 1: degree-4 polynomial, quadratic reciprocal step: 11 FP64 ops, absolute error < 5e-11
//            (value-only quadrature kernels: 1e-10 on log psi is 5 orders below the 1e-5 Ha tolerance)
// Domain: |x| < 4e7 (beyond that the int32 reduction index overflows; tanh saturates at |x| ~ 19).
template <int NV, int ACC>
AQF_HD void ftanh_n(const double* __restrict__ x, double* __restrict__ out, const double* __restrict__ tab) {
  const double kMagic = 6755399441055744.0;
  const FmK& K = fmk();
  double u[NV], t[NV], r[NV], p[NV], d[NV], y[NV];
  int n[NV];
  uint32_t sg[NV];               // sign of x, taken first so that x itself is dead once u exists (one register, not two)
#ifdef __CUDACC__
#pragma unroll
#endif
  for (int i = 0; i < NV; ++i) {
    const uint32_t hx = (uint32_t)hi_word(x[i]);
    sg[i] = hx & 0x80000000u;
#if !defined(AIQMC_TANH_INTABS) && defined(__CUDA_ARCH__)
    u[i] = -2.0 * fabs(x[i]);    // one DMUL with the |.| operand modifier instead of three integer ops and a move
#else
    // u = -2|x| by integer ops (exponent + 1, sign set); x = 0 gives a harmless |u| <= 2^-1021
    u[i] = make_double((int32_t)(((hx & 0x7fffffffu) + 0x00100000u) | 0x80000000u), lo_word(x[i]));
#endif
  }
#ifdef __CUDACC__
#pragma unroll
#endif
  for (int i = 0; i < NV; ++i) t[i] = fma(u[i], K.inv, kMagic);
#ifdef __CUDACC__
#pragma unroll
#endif
  for (int i = 0; i < NV; ++i) { n[i] = lo_word(t[i]); t[i] = t[i] - kMagic; }
#ifdef __CUDACC__
#pragma unroll
#endif
  for (int i = 0; i < NV; ++i) r[i] = fma(-t[i], K.ln2, u[i]);
  if (ACC == 0) {
#ifdef __CUDACC__
#pragma unroll
#endif
    for (int i = 0; i < NV; ++i) p[i] = fma(r[i], K.c5, K.c4);
#ifdef __CUDACC__
#pragma unroll
#endif
    for (int i = 0; i < NV; ++i) p[i] = fma(r[i], p[i], K.c3);
  } else {
#ifdef __CUDACC__
#pragma unroll
#endif
    for (int i = 0; i < NV; ++i) p[i] = fma(r[i], K.c4, K.c3);
  }
#ifdef __CUDACC__
#pragma unroll
#endif
  for (int i = 0; i < NV; ++i) p[i] = fma(r[i], p[i], 0.5);
#ifdef __CUDACC__
#pragma unroll
#endif
  for (int i = 0; i < NV; ++i) p[i] = fma(r[i], p[i], 1.0);
#ifdef __CUDACC__
#pragma unroll
#endif
  for (int i = 0; i < NV; ++i) p[i] = fma(r[i], p[i], 1.0);              // exp(r), |r| <= ln2/32
#ifdef __CUDACC__
#pragma unroll
#endif
  for (int i = 0; i < NV; ++i) {
    const double tj = AQF_TAB(tab, n[i] & (kExpTab - 1));
    int k = n[i] >> 4;
    k = k < -64 ? -64 : k;                                               // e < 2^-64 no longer changes 1 + e
    d[i] = fma(make_double(hi_word(tj) + (k << 20), lo_word(tj)), p[i], 1.0);   // 1 + exp(-2|x|) in (1, 2]
  }
#ifdef __CUDACC__
#pragma unroll
#endif
  for (int i = 0; i < NV; ++i) {
#ifdef __CUDA_ARCH__
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y[i]) : "d"(d[i]));
#else
    y[i] = (double)(1.0f / (float)d[i]);
#endif
  }
#ifdef __CUDACC__
#pragma unroll
#endif
  for (int i = 0; i < NV; ++i) d[i] = fma(-d[i], y[i], 1.0);
  if (ACC == 0) {
#ifdef __CUDACC__
#pragma unroll
#endif
    for (int i = 0; i < NV; ++i) d[i] = fma(d[i], d[i], d[i]);
  }
#ifdef __CUDACC__
#pragma unroll
#endif
  for (int i = 0; i < NV; ++i) y[i] = fma(y[i], d[i], y[i]);
#ifdef __CUDACC__
#pragma unroll
#endif
  for (int i = 0; i < NV; ++i) {
    const double th = fma(2.0, y[i], -1.0);
    out[i] = make_double((int32_t)((uint32_t)hi_word(th) | sg[i]), lo_word(th));
  }
}

AQF_HD double ftanh(double x, const double* __restrict__ tab) {
  double o;
  ftanh_n<1, 0>(&x, &o, tab);
  return o;
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
