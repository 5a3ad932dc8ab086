// fastmath.cuh -- double-precision exp / tanh / reciprocal / rsqrt for the FP64-pipe-bound kernels.
//
// tanh is ~60 % of the FP64 instructions of a psi evaluation (174 tanh per single-electron move at
// N=4), so the transcendental bodies are hand-written to a fixed FP64-instruction budget instead of
// calling libm (which also converts through FP32 and branches):
//   fexp(x):   x = (16k + j) ln2/16 + r, |r| <= ln2/32;  exp(x) = 2^k * T[j] * P6(r),  T[j] = 2^(j/16).
//              The table has 16 doubles = exactly one 128-byte row of shared-memory banks, so a warp's
//              divergent lookups are conflict-free by construction (a 64-entry table measured 4-6
//              wavefronts per LDS).  10 FP64 ops.  Relative error < 1e-15.
//   ftanh(x):  1 - 2/(1 + exp(2x)) for either sign of x: no |x|, no sign fix-up (only ABSOLUTE accuracy enters log|psi|
//              and E_L; the sign of a result below 1e-13 is not defined).  The reduction works on x itself
//              (2x = n ln2/T + 2 r2, single constant) and the polynomial is exp(2 r2) in r2; reciprocal from the
//              MUFU.RCP64H seed + one cubic step.  13 FP64 ops, absolute error < 2e-13.  (Round 1: exp(-2|x|), 14 ops
//              plus two integer ops for the sign.)
//   ftanh_n<NV, ACC>: NV of them with interleaved steps (ILP); ACC = 1 is the 9-op variant of the quadrature kernels
//              (512-entry table, quadratic, one Newton step: < 5e-11).
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

// ACC = 1 tanh of the quadrature kernels: a 512-entry table 2^(j/512) (4 kB) shrinks the reduced argument to
// |2 r2| <= ln2/1024, where a QUADRATIC is enough for the same accuracy ((ln2/1024)^3/6 < 5.2e-11 on the exponential,
// half of that on tanh): two polynomial terms fewer than with the 16-entry table.  The lookup is no longer
// conflict-free (32 lanes over 16 bank pairs: ~4-5 wavefronts).  k_ecp_pt leaves the shared-memory pipe mostly idle
// (LSU wavefronts 47 % with the table) and is bound by the FP64 pipe and the issue slots: 16 entries / quartic 4.92 ms,
// 64 entries / cubic 4.81 ms, sign-free form 4.63, 512 entries / quadratic 4.56, one-instruction clamp 4.48 (carbon,
// 65,536 walkers).  k_ecp_grp<10,2> is closer to its shared-memory limit (LSU wavefronts 84 % with the table): there
// the big table is worth 1.2 % (122.8 vs 124.3 ms, A/B on one box).
#ifndef AIQMC_TANH_TAB64
#define AIQMC_TANH_TAB64 1
#endif
#ifndef AIQMC_TANH_I2F
#define AIQMC_TANH_I2F 0      // 1: rint(u*inv) back to double by I2F.F64 instead of `t - magic` (measured: 4.85 vs 4.81 ms, no gain)
#endif
constexpr int kExpTab64 = 512;       // (the name dates from the 64-entry version)
constexpr int kExpTab64Log2 = 9;
#ifdef __CUDACC__
__shared__ double g_exp_tab64[kExpTab64];            // filled by the prologue of every kernel that evaluates ACC = 1 tanh
#endif
inline const double* host_exp_table64() {
  static double tab[kExpTab64];
  static bool init = false;
  if (!init) {
    for (int j = 0; j < kExpTab64; ++j) tab[j] = exp2((double)j / kExpTab64 - 64.0);
    init = true;
  }
  return tab;
}
#ifdef __CUDACC__
// Prologue fill of g_exp_tab64 = 2^(j/512 - 64): 2^(a/16 - 64) from 16 literals (index uniform per warp: a broadcast
// constant load) times exp((j & 31) ln2/512) by a degree-7 polynomial -- a dozen instructions per entry instead of a
// libm exp2 (the exp2 fill was 1.9 % of k_ecp_pt's instructions).
static __constant__ double c_exp16_m64[16] = {
    0x1.0000000000000p-64, 0x1.0b5586cf9890fp-64, 0x1.172b83c7d517bp-64, 0x1.2387a6e756238p-64, 0x1.306fe0a31b715p-64,
    0x1.3dea64c123422p-64, 0x1.4bfdad5362a27p-64, 0x1.5ab07dd485429p-64, 0x1.6a09e667f3bcdp-64, 0x1.7a11473eb0187p-64,
    0x1.8ace5422aa0dbp-64, 0x1.9c49182a3f090p-64, 0x1.ae89f995ad3adp-64, 0x1.c199bdd85529cp-64, 0x1.d5818dcfba487p-64,
    0x1.ea4afa2a490dap-64};
__device__ __forceinline__ void fill_tanh_table(int tid, int nthreads) {
  static_assert(kExpTab64 == 512, "16 x 32 factorisation");
  for (int j = tid; j < kExpTab64; j += nthreads) {
    const double x = (double)(j & 31) * (0.69314718055994530942 / 512.0);      // <= 0.042: x^8/8! < 3e-16
    double p = 1.0 / 5040.0;
    p = fma(x, p, 1.0 / 720.0);
    p = fma(x, p, 1.0 / 120.0);
    p = fma(x, p, 1.0 / 24.0);
    p = fma(x, p, 1.0 / 6.0);
    p = fma(x, p, 0.5);
    p = fma(x, p, 1.0);
    p = fma(x, p, 1.0);
    g_exp_tab64[j] = c_exp16_m64[j >> 5] * p;
  }
}
#endif
#ifdef __CUDA_ARCH__
#define AQF_TAB64(j) g_exp_tab64[j]
#else
#define AQF_TAB64(j) host_exp_table64()[j]
#endif

// Polynomial / reduction constants.  On the device they live in constant memory so that each one is a
// c[3][imm] operand of the DFMA that uses it (as 64-bit literals every use costs two IMAD.MOVs).
// tanh: inv16x2 = 2*16/ln2, hl16 = ln2/32, inv64x2 = 2*512/ln2, hl64 = ln2/1024, e3..e5 = 2^k/k! (exp(2 r) coefficients)
struct FmK { double inv, ln2, ln2hi, ln2lo, c2, c3, c4, c5, c6, inv16x2, hl16, inv64x2, hl64, e3, e4, e5; };
#define AQF_FMK_INIT {23.083120654223414, 0.04332169878499658, 0.043321698784978935, 1.7647056601894736e-14, \
                      0.5, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, \
                      46.166241308446828, 0.021660849392498290, 1477.3197218702985, 0.0006769015435155716, \
                      4.0 / 3.0, 2.0 / 3.0, 4.0 / 15.0}
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
//   ACC = 0: 16-entry table, degree-5 polynomial, cubic reciprocal step:  13 FP64 ops, absolute error < 2e-13
//   ACC = 1: 512-entry table, quadratic, quadratic reciprocal step:         9 FP64 ops, absolute error < 5e-11
//            (value-only quadrature kernels: 1e-10 on log psi is 5 orders below the 1e-5 Ha tolerance)
// Domain: |x| < 1e6 (beyond that the int32 reduction index overflows; tanh saturates at |x| ~ 19).
// RAW = true returns y = 1/(1 + exp(2x)) instead of tanh(x) = 1 - 2y, for callers that fold the affine map into their
// own next FMA (tanh_res in psi_core.cuh: (h + tanh z)/sqrt2 = fma(y, -sqrt2, fma(h, 1/sqrt2, 1/sqrt2)), one op fewer).
template <int NV, int ACC, bool RAW = false>
AQF_HD void ftanh_n(const double* __restrict__ x, double* __restrict__ out, const double* __restrict__ tab) {
  // tanh(x) = 1 - 2/(1 + exp(2x)) for either sign of x (no |x|, no sign fix-up: for x << 0 the exponential vanishes, for
  // x >> 0 the reciprocal does; ABSOLUTE accuracy is what log|psi| and E_L see).  The reduction works on x itself:
  // 2x = n ln2/T + 2 r2 with n = rint(2x T/ln2), |r2| <= ln2/(4T), and the polynomial is exp(2 r2) in r2, so the doubling
  // costs nothing.  T = 16 (conflict-free table) or 64 (ACC = 1, see above).
  constexpr double kMagic = 6755399441055744.0;
  const FmK& K = fmk();
  constexpr bool kT64 = (ACC == 1) && (AIQMC_TANH_TAB64 != 0);
  double t[NV], r[NV], p[NV], d[NV], y[NV];
  int n[NV];
  // kT64: the magic constant carries a bias of 64 * 512, so the exponent k arrives as k + 64 >= 0 for every argument
  // that matters and ONE relu-min clamps it to [0, 128]; the table holds 2^(j/512 - 64) to take the bias out again
  constexpr double kMagicB = kT64 ? kMagic + 64.0 * kExpTab64 : kMagic;
#ifdef __CUDACC__
#pragma unroll
#endif
  for (int i = 0; i < NV; ++i) t[i] = fma(x[i], kT64 ? K.inv64x2 : K.inv16x2, kMagicB);
#ifdef __CUDACC__
#pragma unroll
#endif
  for (int i = 0; i < NV; ++i) { n[i] = lo_word(t[i]); t[i] = t[i] - kMagicB; }
#ifdef __CUDACC__
#pragma unroll
#endif
  for (int i = 0; i < NV; ++i) r[i] = fma(-t[i], kT64 ? K.hl64 : K.hl16, x[i]);       // r2 = x - n ln2/(2T)
  // exp(2 r2) = 1 + 2 r2 + 2 r2^2 + 4/3 r2^3 + 2/3 r2^4 + 4/15 r2^5
  if (kT64) {
#ifdef __CUDACC__
#pragma unroll
#endif
    for (int i = 0; i < NV; ++i) p[i] = 2.0;
  } else {
    if (ACC == 0) {
#ifdef __CUDACC__
#pragma unroll
#endif
      for (int i = 0; i < NV; ++i) p[i] = fma(r[i], K.e5, K.e4);
#ifdef __CUDACC__
#pragma unroll
#endif
      for (int i = 0; i < NV; ++i) p[i] = fma(r[i], p[i], K.e3);
    } else {
#ifdef __CUDACC__
#pragma unroll
#endif
      for (int i = 0; i < NV; ++i) p[i] = fma(r[i], K.e4, K.e3);
    }
#ifdef __CUDACC__
#pragma unroll
#endif
    for (int i = 0; i < NV; ++i) p[i] = fma(r[i], p[i], 2.0);
  }
#ifdef __CUDACC__
#pragma unroll
#endif
  for (int i = 0; i < NV; ++i) p[i] = fma(r[i], p[i], 2.0);
#ifdef __CUDACC__
#pragma unroll
#endif
  for (int i = 0; i < NV; ++i) p[i] = fma(r[i], p[i], 1.0);              // exp(2 r2)
#ifdef __CUDACC__
#pragma unroll
#endif
  for (int i = 0; i < NV; ++i) {
    const double tj = kT64 ? AQF_TAB64(n[i] & (kExpTab64 - 1)) : AQF_TAB(tab, n[i] & (kExpTab - 1));
    int k = kT64 ? (n[i] >> kExpTab64Log2) : (n[i] >> 4);
    if (kT64) {
#ifdef __CUDA_ARCH__
      k = __vimin_s32_relu(k, 128);                                      // max(min(k + 64, 128), 0) in one VIMNMX.RELU
#else
      k = k < 0 ? 0 : (k > 128 ? 128 : k);
#endif
    } else {
      k = k < -64 ? -64 : (k > 64 ? 64 : k);                             // beyond 2^+-64 the result is +-1 to the last bit
    }
    d[i] = fma(make_double(hi_word(tj) + (k << 20), lo_word(tj)), p[i], 1.0);   // 1 + exp(2x)
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
  for (int i = 0; i < NV; ++i) out[i] = RAW ? y[i] : fma(-2.0, y[i], 1.0);
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
