// psi_core.cuh -- per-configuration evaluation of the AIQMCrelease3 wavefunction
// (value / gradient / forward-Laplacian), written once as __host__ __device__ code so the
// same arithmetic can be unit-tested on a CPU build (tests/hostcore) and run in the sm_100a
// kernels (kernels.cu).  One thread evaluates one electron configuration; all scratch is
// per-thread (registers / local memory), parameters are read through `P` (shared or global).
//
// What is evaluated (citations under /root/reference/AIQMCrelease3/):
//   features            wavefunction_Ynlm/nn.py:106-137
//   Ynlm stream         nn.py:156-193,313-341         (quirk Q3: x[3] clamps to x[2])
//   one/two-e streams   nn.py:142-153,280-311 + network_blocks.py:106-133
//   orbitals, envelope  nn.py:409-506, envelope.py:26-30 (quirk Q4 row/electron mix kept)
//   Jastrows            Jastrow.py:23-52,74-93  (exp(J/N) on every element == +J on log|det|)
//   slogdet             network_blocks.py:138-206
//
// Derivatives are NOT done the reference's way (reverse-mode grad + 3N jvps, O(N^4)).
// For every electron e a 3-direction second-order jet is pushed through the parts of the
// network that depend on r_e (2N-1 pair chains, electron e's Ynlm/envelope, all one-electron
// rows), and the determinant is differentiated analytically through M^-1 using the low-rank
// structure  dM[k,j] = sum_c dh[sigma_k,c] W[c,j] E[k,j] + delta_ke P[e,j] dE[e,j]:
//   d log det   = tr(X),                  X = dM M^-1 = dh . T + delta_ke S1
//   d2 log det  = tr(M^-1 d2M) - tr(X X)
// with T[k,c,l] = sum_j W[c,j] E[k,j] Minv[j,l] built once per configuration (4N^3).
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/aiqmc_b200.h"
#include "fastmath.cuh"

#ifdef __CUDACC__
#define AQ_HD __host__ __device__ __forceinline__
#define AQ_HD_NOINLINE __host__ __device__ __noinline__
#else
#define AQ_HD inline
#define AQ_HD_NOINLINE
#endif

namespace aiqmc {

constexpr double kInvSqrt2 = 0.70710678118654752440;
constexpr double kPi = 3.14159265358979323846;

// ---------------------------------------------------------------------------------------
// parameter layout (single source of truth; exported through aiqmc_param_layout)
// ---------------------------------------------------------------------------------------
AQ_HD constexpr AiqmcLayout make_layout(int N, int A) {
  AiqmcLayout L{};
  int o = 0;
  const int d[3] = {12 * A + 8, 20, 20};
  const int ky[3] = {4 * A + 2, 6, 6};
  for (int l = 0; l < 3; ++l) {
    L.conv_w[l] = o; o += N * d[l];
    L.conv_b[l] = o; o += N * (d[l] / 4);
    L.sing_w[l] = o; o += (d[l] / 4) * 4;
    L.sing_b[l] = o; o += 4;
    if (l < 2) { L.dbl_w[l] = o; o += 16; L.dbl_b[l] = o; o += 4; }
    L.yn_w[l] = o; o += ky[l] * 6;
    L.yn_b[l] = o; o += 6;
  }
  for (int s = 0; s < 2; ++s) { L.orb_w[s] = o; o += 4 * 2 * N; L.orb_b[s] = o; o += 2 * N; }
  L.y_w = o; o += 6 * N;
  L.jas_alpha = o; o += N * N;
  L.jas_cusp = o; o += N * N;
  L.jas_beta = o; o += N * A;
  L.jas_c34 = o; o += A;
  L.jas_c14 = o; o += A;
  L.env_pi = o; o += N * A * 3;
  L.env_sx = o; o += N * A * 3;
  L.env_alpha = o; o += N;
  L.env_beta = o; o += N * A;
  L.atoms = o; o += A * 3;
  L.charges = o; o += A;
  L.total = o;
  return L;
}


// Compile-time view of the layout.  NOTE: a function-local `constexpr AiqmcLayout` that is indexed
// with a runtime layer number gets materialised in local memory, and nvcc 12.9's stack colouring
// then overlapped it with the caller's gradient buffer (observed on sm_100a: the first 19 doubles
// of grad came back as layout offsets).  The selectors below keep every offset an immediate.
template <int V0, int V1, int V2> struct Sel3 {
  AQ_HD constexpr int operator[](int l) const { return l == 0 ? V0 : (l == 1 ? V1 : V2); }
};
template <int V0, int V1> struct Sel2 {
  AQ_HD constexpr int operator[](int l) const { return l == 0 ? V0 : V1; }
};
template <int NE, int NA>
struct LayoutC {
  static constexpr AiqmcLayout K = make_layout(NE, NA);
  Sel3<K.conv_w[0], K.conv_w[1], K.conv_w[2]> conv_w;
  Sel3<K.conv_b[0], K.conv_b[1], K.conv_b[2]> conv_b;
  Sel3<K.sing_w[0], K.sing_w[1], K.sing_w[2]> sing_w;
  Sel3<K.sing_b[0], K.sing_b[1], K.sing_b[2]> sing_b;
  Sel2<K.dbl_w[0], K.dbl_w[1]> dbl_w;
  Sel2<K.dbl_b[0], K.dbl_b[1]> dbl_b;
  Sel3<K.yn_w[0], K.yn_w[1], K.yn_w[2]> yn_w;
  Sel3<K.yn_b[0], K.yn_b[1], K.yn_b[2]> yn_b;
  Sel2<K.orb_w[0], K.orb_w[1]> orb_w;
  Sel2<K.orb_b[0], K.orb_b[1]> orb_b;
  static constexpr int y_w = K.y_w, jas_alpha = K.jas_alpha, jas_cusp = K.jas_cusp, jas_beta = K.jas_beta,
                       jas_c34 = K.jas_c34, jas_c14 = K.jas_c14, env_pi = K.env_pi, env_sx = K.env_sx,
                       env_alpha = K.env_alpha, env_beta = K.env_beta, atoms = K.atoms, charges = K.charges,
                       total = K.total;
};

// ---------------------------------------------------------------------------------------
// scalars: double, or a jet carrying d/dx_c and d2/dx_c^2 for the 3 coordinates of ONE electron
// ---------------------------------------------------------------------------------------
template <bool LAP, int ND = 3>
struct Jet {
  double v;
  double d[ND];
  double s[ND];
};

template <class S> struct ScalarOps;

template <> struct ScalarOps<double> {
  static AQ_HD double cst(double x) { return x; }
  static AQ_HD double val(double x) { return x; }
};
template <bool LAP, int ND> struct ScalarOps<Jet<LAP, ND>> {
  static AQ_HD Jet<LAP, ND> cst(double x) { Jet<LAP, ND> r; r.v = x; for (int c = 0; c < ND; ++c) { r.d[c] = 0; r.s[c] = 0; } return r; }
  static AQ_HD double val(const Jet<LAP, ND>& x) { return x.v; }
};

template <bool L, int ND> AQ_HD Jet<L, ND> operator+(const Jet<L, ND>& a, const Jet<L, ND>& b) {
  Jet<L, ND> r; r.v = a.v + b.v;
  for (int c = 0; c < ND; ++c) { r.d[c] = a.d[c] + b.d[c]; if (L) r.s[c] = a.s[c] + b.s[c]; else r.s[c] = 0; }
  return r;
}
template <bool L, int ND> AQ_HD Jet<L, ND> operator-(const Jet<L, ND>& a, const Jet<L, ND>& b) {
  Jet<L, ND> r; r.v = a.v - b.v;
  for (int c = 0; c < ND; ++c) { r.d[c] = a.d[c] - b.d[c]; if (L) r.s[c] = a.s[c] - b.s[c]; else r.s[c] = 0; }
  return r;
}
template <bool L, int ND> AQ_HD Jet<L, ND> operator+(const Jet<L, ND>& a, double b) { Jet<L, ND> r = a; r.v += b; return r; }
template <bool L, int ND> AQ_HD Jet<L, ND> operator-(const Jet<L, ND>& a, double b) { Jet<L, ND> r = a; r.v -= b; return r; }
template <bool L, int ND> AQ_HD Jet<L, ND> operator*(const Jet<L, ND>& a, double b) {
  Jet<L, ND> r; r.v = a.v * b;
  for (int c = 0; c < ND; ++c) { r.d[c] = a.d[c] * b; r.s[c] = L ? a.s[c] * b : 0.0; }
  return r;
}
template <bool L, int ND> AQ_HD Jet<L, ND> operator*(double b, const Jet<L, ND>& a) { return a * b; }
template <bool L, int ND> AQ_HD Jet<L, ND> operator*(const Jet<L, ND>& a, const Jet<L, ND>& b) {
  Jet<L, ND> r; r.v = a.v * b.v;
  for (int c = 0; c < ND; ++c) {
    r.d[c] = a.d[c] * b.v + a.v * b.d[c];
    r.s[c] = L ? (a.s[c] * b.v + 2.0 * a.d[c] * b.d[c] + a.v * b.s[c]) : 0.0;
  }
  return r;
}
// r = f(u) given f, f', f'' at u.v
template <bool L, int ND> AQ_HD Jet<L, ND> chain(const Jet<L, ND>& u, double f, double f1, double f2) {
  Jet<L, ND> r; r.v = f;
  for (int c = 0; c < ND; ++c) { r.d[c] = f1 * u.d[c]; r.s[c] = L ? (f2 * u.d[c] * u.d[c] + f1 * u.s[c]) : 0.0; }
  return r;
}

// exp / tanh: in-house fexp/ftanh (fastmath.cuh) unless AIQMC_LIBM is defined.  On the device the
// 64-entry 2^(j/64) table lives in shared memory (filled by stage_params in every kernel).
AQ_HD const double* exp_tab() {
#ifdef __CUDA_ARCH__
  return g_exp_tab;
#else
  return host_exp_table();
#endif
}
#ifdef AIQMC_LIBM
#define AQ_TANH(x) tanh(x)
#define AQ_EXP(x) exp(x)
#elif defined(AIQMC_TANH_NOINLINE) && defined(__CUDACC__)
// out-of-line transcendental bodies: the fully unrolled kernels are 170-400 kB of SASS and stall on
// instruction fetch ("no_instructions"); one shared body per function shrinks them ~3x.
static __device__ __noinline__ double ftanh_ool(double x) { return ftanh(x, g_exp_tab); }
static __device__ __noinline__ double fexp_ool(double x) { return fexp(x, g_exp_tab); }
AQ_HD double tanh_dispatch(double x) {
#ifdef __CUDA_ARCH__
  return ftanh_ool(x);
#else
  return ftanh(x, host_exp_table());
#endif
}
AQ_HD double exp_dispatch(double x) {
#ifdef __CUDA_ARCH__
  return fexp_ool(x);
#else
  return fexp(x, host_exp_table());
#endif
}
#define AQ_TANH(x) tanh_dispatch(x)
#define AQ_EXP(x) exp_dispatch(x)
#else
#define AQ_TANH(x) ftanh(x, exp_tab())
#define AQ_EXP(x) fexp(x, exp_tab())
#endif

AQ_HD double s_tanh(double x) { return AQ_TANH(x); }
AQ_HD double s_exp(double x) { return AQ_EXP(x); }
#ifdef AIQMC_LIBM
AQ_HD double s_rsqrt(double x) { return 1.0 / sqrt(x); }
AQ_HD double s_inv(double x) { return 1.0 / x; }
#else
AQ_HD double s_rsqrt(double x) { return frsqrt(x); }
AQ_HD double s_inv(double x) { return frcp(x); }
#endif
AQ_HD double s_sqrt(double x) { return x * s_rsqrt(x); }                      // x > 0
template <bool L, int ND> AQ_HD Jet<L, ND> s_tanh(const Jet<L, ND>& u) { double t = AQ_TANH(u.v); double g = 1.0 - t * t; return chain(u, t, g, -2.0 * t * g); }
template <bool L, int ND> AQ_HD Jet<L, ND> s_exp(const Jet<L, ND>& u) { double e = AQ_EXP(u.v); return chain(u, e, e, e); }
template <bool L, int ND> AQ_HD Jet<L, ND> s_sqrt(const Jet<L, ND>& u) { double i = s_rsqrt(u.v); return chain(u, u.v * i, 0.5 * i, -0.25 * i * i * i); }
template <bool L, int ND> AQ_HD Jet<L, ND> s_inv(const Jet<L, ND>& u) { double f = s_inv(u.v); return chain(u, f, -f * f, 2.0 * f * f * f); }
// r = sqrt(r2) and 1/r from one reciprocal square root
AQ_HD void s_sqrt_inv(double r2, double& r, double& ri) { ri = s_rsqrt(r2); r = r2 * ri; }
template <bool L, int ND> AQ_HD void s_sqrt_inv(const Jet<L, ND>& r2, Jet<L, ND>& r, Jet<L, ND>& ri) {
  const double i = s_rsqrt(r2.v), i2 = i * i;
  r = chain(r2, r2.v * i, 0.5 * i, -0.25 * i * i2);
  ri = chain(r2, i, -0.5 * i * i2, 0.75 * i * i2 * i2);
}

#if defined(AIQMC_TANH_OOL) && defined(__CUDACC__)
// experiment: one out-of-line body per (NV, ACC) instead of an inlined copy at each of the ~40 call sites
template <int NV, int ACC>
static __device__ __noinline__ void tanh_ool(const double* __restrict__ z, double* __restrict__ out) {
  ftanh_n<NV, ACC>(z, out, g_exp_tab);
}
#endif
#ifndef AIQMC_TANH_CH
#define AIQMC_TANH_CH 8            // tanh evaluations interleaved per batch (ILP against register pressure)
#endif
// NV tanh at once: doubles go through the interleaved ftanh_n (ILP), jets one by one.
template <int NV, int ACC>
AQ_HD void tanhv(const double* __restrict__ z, double* __restrict__ out) {
#if defined(AIQMC_TANH_OOL) && defined(__CUDA_ARCH__)
  if constexpr (NV <= 8) { tanh_ool<NV, ACC>(z, out); return; }
#endif
#ifdef AIQMC_LIBM
  for (int i = 0; i < NV; ++i) out[i] = tanh(z[i]);
#else
  constexpr int CH = AIQMC_TANH_CH;
  if constexpr (NV <= CH) {
    ftanh_n<NV, ACC>(z, out, exp_tab());
  } else {
    ftanh_n<CH, ACC>(z, out, exp_tab());
    tanhv<NV - CH, ACC>(z + CH, out + CH);
  }
#endif
}
template <int NV, int ACC, bool L, int ND>
AQ_HD void tanhv(const Jet<L, ND>* __restrict__ z, Jet<L, ND>* __restrict__ out) {
  for (int i = 0; i < NV; ++i) out[i] = s_tanh(z[i]);
}
#ifndef AIQMC_TANH_RESFUSE
#define AIQMC_TANH_RESFUSE 0   // measured A/B on one box: k_ecp_pt 4.54 ms fused vs 4.50 ms unfused (fewer FP64 ops, but more
#endif                         // spill traffic at the 128-register cap); k_ecp_grp<10,2> 131.5 vs 132.4 ms.  Off by default.
// out[i] = (h[i] + tanh(z[i])) / sqrt(2): the residual update of every stream (nn.py:305-309, 327-341).  On plain
// doubles the tanh's last FMA (1 - 2y) is folded into the update: 2 FP64 ops after the reciprocal instead of 3.
// `out` may alias `h`.
template <int NV, int ACC>
AQ_HD void tanh_res(const double* __restrict__ z, const double* h, double* out) {
#if defined(AIQMC_LIBM) || (defined(AIQMC_TANH_OOL) && defined(__CUDA_ARCH__)) || !AIQMC_TANH_RESFUSE
  double t[NV];
  tanhv<NV, ACC>(z, t);
  for (int i = 0; i < NV; ++i) out[i] = (h[i] + t[i]) * kInvSqrt2;
#else
  constexpr double kSqrt2 = 1.41421356237309504880;
  constexpr int CH = 8;
  double y[NV];
  if constexpr (NV <= CH) {
    ftanh_n<NV, ACC, true>(z, y, exp_tab());
  } else {
    ftanh_n<CH, ACC, true>(z, y, exp_tab());
    ftanh_n<NV - CH, ACC, true>(z + CH, y + CH, exp_tab());
    static_assert(NV <= 2 * CH, "tanh_res: at most 16 values");
  }
#ifdef __CUDACC__
#pragma unroll
#endif
  for (int i = 0; i < NV; ++i) out[i] = fma(y[i], -kSqrt2, fma(h[i], kInvSqrt2, kInvSqrt2));
#endif
}
template <int NV, int ACC, bool L, int ND>
AQ_HD void tanh_res(const Jet<L, ND>* __restrict__ z, const Jet<L, ND>* h, Jet<L, ND>* out) {
  for (int i = 0; i < NV; ++i) out[i] = (h[i] + s_tanh(z[i])) * kInvSqrt2;
}

struct cplx { double re, im; };
AQ_HD cplx cmul(cplx a, cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
AQ_HD cplx cadd(cplx a, cplx b) { return {a.re + b.re, a.im + b.im}; }
AQ_HD cplx csub(cplx a, cplx b) { return {a.re - b.re, a.im - b.im}; }
AQ_HD cplx cscale(cplx a, double s) { return {a.re * s, a.im * s}; }
AQ_HD cplx cinv(cplx a) { double n = 1.0 / (a.re * a.re + a.im * a.im); return {a.re * n, -a.im * n}; }
AQ_HD void cfma(cplx& acc, cplx a, cplx b) { acc.re += a.re * b.re - a.im * b.im; acc.im += a.re * b.im + a.im * b.re; }
AQ_HD void cfms(cplx& acc, cplx a, cplx b) { acc.re -= a.re * b.re - a.im * b.im; acc.im -= a.re * b.im + a.im * b.re; }

// Per-walker cache consumed by the single-electron-move kernels (ecp_coop.cuh): everything of the
// base configuration that survives when ONE electron is displaced.
template <int NE, int NA>
struct MoveCache {
  static constexpr int HP = 0;                            // [3][N][N][4]  pair chains h_two^l[i,j,:]
  static constexpr int GS = HP + 3 * NE * NE * 4;         // [3][2][N][4]  block sums over i of h_two^l[i,j,:]
  static constexpr int H0 = GS + 3 * 2 * NE * 4;          // [N][4A]       layer-0 one-electron features
  static constexpr int G0M = H0 + NE * 4 * NA;            // [2][4A]       block MEANS of H0
  static constexpr int Y = G0M + 2 * 4 * NA;              // [N][6]        Ynlm stream output
  static constexpr int ENV = Y + NE * 6;                  // [N]
  static constexpr int JAE = ENV + NE;                    // [N]
  static constexpr int JEE = JAE + NE;                    // [N]           sum_{k != i} u_ee(r_ik)
  static constexpr int MISC = JEE + NE;                   // [4] total Jastrow, log|psi|, phase, 0
  // "quadrature record": the walker-local inputs of the quadrature kernels, appended so that ONE bulk copy of the
  // record brings everything a CTA needs (written by k_energy_rest / k_tmove_prep)
  static constexpr int QR = MISC + 4;                     // [3N] positions [9] rotation [4] group norms [2] log|psi|, phase
  static constexpr int QR_VL = QR + 3 * NE + 15;          // [N][A][4] v_l(r_ia)
  static constexpr int SIZE = (QR_VL + 4 * NE * NA + 1) & ~1;   // even: records stay 16-byte aligned
};

// ---------------------------------------------------------------------------------------
template <int NE, int NA>
struct Psi {
  static constexpr int N = NE, A = NA;
  static constexpr int D0 = 12 * NA + 8, Q0 = 3 * NA + 2, K0 = 4 * NA + 2;
  static constexpr bool kSmall = (NE <= 8);

  // ---- electron-local quantities: h0 = [r_ea, ae] (4A), Ynlm stream output y[6], envelope,
  //      electron-nucleus Jastrow term.  S = double or Jet.
  template <class S, int ACC = 0>
  static AQ_HD void electron_local(const double* __restrict__ P, int e, const S xe[3], S* __restrict__ h0,
                                   S y[6], S& env, S& jae) {
    using Op = ScalarOps<S>;
    constexpr LayoutC<NE, NA> L{};
    const double c0 = 0.28209479177387814;        // 1/2 sqrt(1/pi)
    const double c1 = 0.48860251190291992;        // sqrt(3/(4 pi))
    const double k15h = 1.0925484305920792;       // 1/2 sqrt(15/pi)
    const double k5q = 0.31539156525252005;       // 1/4 sqrt(5/pi)
    const double k15q = 0.54627421529603959;      // 1/4 sqrt(15/pi)
    const double k35 = 0.59004358992664352;       // 1/4 sqrt(35/(2pi))
    const double k105h = 2.8906114426405538;      // 1/2 sqrt(105/pi)
    const double k21 = 0.45704579946446577;       // 1/4 sqrt(21/(2pi))
    const double k7q = 0.37317633259011546;       // 1/4 sqrt(7/pi)
    const double k105q = 1.4453057213202769;      // 1/4 sqrt(105/pi)
    S z[6];
    for (int m = 0; m < 6; ++m) z[m] = Op::cst(P[L.yn_b[0] + m]);
    S sum_df = Op::cst(0.0), sum_sp = Op::cst(0.0);
    S y0keep[6];
    env = Op::cst(0.0);
    jae = Op::cst(0.0);
    const double* W0 = P + L.yn_w[0];
    for (int a = 0; a < A; ++a) {
      S ae[3];
      for (int c = 0; c < 3; ++c) ae[c] = xe[c] - P[L.atoms + 3 * a + c];
      S r2 = ae[0] * ae[0] + ae[1] * ae[1] + ae[2] * ae[2];
      S r, ri;
      s_sqrt_inv(r2, r, ri);
      h0[4 * a] = r;
      for (int c = 0; c < 3; ++c) h0[4 * a + 1 + c] = ae[c];
      S t0 = ae[0] * ri, t1 = ae[1] * ri, t2 = ae[2] * ri;
      S sp[4] = {Op::cst(c0), t0 * c1, t1 * c1, t2 * c1};
      for (int q = 0; q < 4; ++q) {
        sum_sp = sum_sp + sp[q];
        if (A == 1) y0keep[q] = sp[q];
        for (int m = 0; m < 6; ++m) z[m] = z[m] + sp[q] * W0[(4 * a + q) * 6 + m];
      }
      // d/f block (nn.py:182-193): unit-vector components over r^2 / r^3 (quirk Q3)
      S ri2 = ri * ri, ri3 = ri2 * ri;
      S t00 = t0 * t0, t11 = t1 * t1, t22 = t2 * t2;
      S d2 = (t0 * t1) * k15h + (t1 * t2) * k15h + (t22 * 3.0 - r2) * k5q + (t0 * t2) * k15h + (t00 - t11) * k15q;
      S f3 = (t1 * (t00 * 3.0 - t11)) * k35 + (t0 * t1 * t2) * k105h + (t1 * (t22 * 5.0 - r2)) * k21 +
             (t22 * t2 * 5.0 - t2 * r2 * 3.0) * k7q + (t0 * (t22 * 5.0 - r2)) * k21 +
             ((t00 - t11) * t2) * k105q + (t0 * (t00 - t11 * 3.0)) * k35;
      sum_df = sum_df + d2 * ri2 + f3 * ri3;
      // envelope (envelope.py:26-30), row = this electron
      env = env + s_exp(r2 * (-P[L.env_beta + e * A + a])) * P[L.env_alpha + e];
      for (int c = 0; c < 3; ++c)
        env = env + s_exp(ae[c] * (-P[L.env_pi + (e * A + a) * 3 + c])) * P[L.env_sx + (e * A + a) * 3 + c];
      // electron-nucleus Pade term (Jastrow.py:84)
      double beta = P[L.jas_beta + e * A + a];
      S ex = s_exp(r * (-P[L.jas_c14 + a] * beta));
      jae = jae + (ex - 1.0) * (P[L.jas_c34 + a] / (2.0 * beta));
    }
    S mdf = sum_df * (1.0 / (12.0 * A)), msp = sum_sp * (1.0 / (4.0 * A));
    for (int m = 0; m < 6; ++m) z[m] = z[m] + mdf * W0[(4 * A) * 6 + m] + msp * W0[(4 * A + 1) * 6 + m];
    if (A == 1) { y0keep[4] = mdf; y0keep[5] = msp; }
    if (A == 1) tanh_res<6, ACC>(z, y0keep, y);                      // residual only if 4A+2 == 6 (quirk Q5)
    else tanhv<6, ACC>(z, y);
    for (int l = 1; l < 3; ++l) {
      const double* W = P + L.yn_w[l];
      S zz[6];
      for (int m = 0; m < 6; ++m) zz[m] = Op::cst(P[L.yn_b[l] + m]);
      for (int q = 0; q < 6; ++q)
        for (int m = 0; m < 6; ++m) zz[m] = zz[m] + y[q] * W[q * 6 + m];
      tanh_res<6, ACC>(zz, y, y);
    }
  }

  // ---- two-electron chain for one ordered pair: h0=[r,d], h1, h2 (nn.py:305-309)
  template <class S, int ACC = 0>
  static AQ_HD void pair_chain(const double* __restrict__ P, const S d[3], bool diag, S h0[4], S h1[4], S h2[4]) {
    using Op = ScalarOps<S>;
    constexpr LayoutC<NE, NA> L{};
    if (diag) {
      h0[0] = Op::cst(0.0);
    } else {
      h0[0] = s_sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    }
    for (int c = 0; c < 3; ++c) h0[1 + c] = d[c];
    const S* in = h0;
    S* out = h1;
    for (int l = 0; l < 2; ++l) {
      const double* W = P + L.dbl_w[l];
      S z[4];
      for (int m = 0; m < 4; ++m) z[m] = Op::cst(P[L.dbl_b[l] + m]);
      for (int q = 0; q < 4; ++q)
        for (int m = 0; m < 4; ++m) z[m] = z[m] + in[q] * W[q * 4 + m];
      tanh_res<4, ACC>(z, in, out);
      in = h1;
      out = h2;
    }
  }

  // ---- one-electron layer for electron k: conv (grouped mean of 4) -> tanh -> linear -> tanh
  //      xin = [h_k (DIN), g_up (DIN), g_dn (DIN), G_up[k]/n_up (4), G_dn[k]/n_dn (4)]
  template <int DIN, class S, int ACC = 0>
  static AQ_HD void one_layer(const double* __restrict__ P, int l, int k, const S* __restrict__ hk,
                              const S* __restrict__ gup, const S* __restrict__ gdn, const S Gu[4], const S Gd[4],
                              S hout[4], double* __restrict__ rec = nullptr, int64_t rec_stride = 0) {
    using Op = ScalarOps<S>;
    constexpr LayoutC<NE, NA> L{};
    constexpr int DTOT = 3 * DIN + 8, Q = DTOT / 4;
    const double* cw = P + L.conv_w[l] + k * DTOT;
    const double* cb = P + L.conv_b[l] + k * Q;
    const double* sw = P + L.sing_w[l];
    S z[4], pre[Q], t[Q];
    for (int m = 0; m < 4; ++m) z[m] = Op::cst(P[L.sing_b[l] + m]);
    for (int q = 0; q < Q; ++q) {
      S acc = Op::cst(0.0);
      for (int c = 0; c < 4; ++c) {
        const int idx = 4 * q + c;
        const S& x = idx < DIN ? hk[idx] : idx < 2 * DIN ? gup[idx - DIN] : idx < 3 * DIN ? gdn[idx - 2 * DIN]
                     : idx < 3 * DIN + 4 ? Gu[idx - 3 * DIN] : Gd[idx - 3 * DIN - 4];
        acc = acc + x * cw[idx];
      }
      pre[q] = acc * 0.25 + cb[q];
    }
    tanhv<Q, ACC>(pre, t);
    if (rec)
      for (int q = 0; q < Q; ++q) rec[q * rec_stride] = Op::val(t[q]);   // first-stage tanh outputs (deriv_split.cuh)
    for (int q = 0; q < Q; ++q)
      for (int m = 0; m < 4; ++m) z[m] = z[m] + t[q] * sw[q * 4 + m];
    if (DIN == 4) tanh_res<4, ACC>(z, hk, hout);                     // residual only if shapes match (Q5)
    else tanhv<4, ACC>(z, hout);
  }

  // ---- complex LU (value only): log|det| and phase; destroys m
  static AQ_HD void lu_logdet(cplx* __restrict__ m, double& logabs, double& phase) {
    double mant = 1.0; int ex = 0; cplx ph = {1.0, 0.0};
    for (int k = 0; k < N; ++k) {
      int p = k; double best = m[k * N + k].re * m[k * N + k].re + m[k * N + k].im * m[k * N + k].im;
      for (int r = k + 1; r < N; ++r) {
        double v = m[r * N + k].re * m[r * N + k].re + m[r * N + k].im * m[r * N + k].im;
        if (v > best) { best = v; p = r; }
      }
      if (p != k) {
        for (int j = k; j < N; ++j) { cplx t = m[k * N + j]; m[k * N + j] = m[p * N + j]; m[p * N + j] = t; }
        ph.re = -ph.re; ph.im = -ph.im;
      }
      cplx piv = m[k * N + k];
      double a = sqrt(best);
      int e2; mant = frexp(mant * a, &e2); ex += e2;
      ph = cscale(cmul(ph, piv), 1.0 / a);
      cplx pinv = cinv(piv);
      for (int r = k + 1; r < N; ++r) {
        cplx f = cmul(m[r * N + k], pinv);
        for (int j = k + 1; j < N; ++j) cfms(m[r * N + j], f, m[k * N + j]);
      }
    }
    logabs = log(mant) + ex * 0.69314718055994530942;
    phase = atan2(ph.im, ph.re);
  }

  // ---- in-place Gauss-Jordan inverse with partial pivoting; also log|det| and phase
  static AQ_HD void gj_inverse(cplx* __restrict__ m, double& logabs, double& phase) {
    double mant = 1.0; int ex = 0; cplx ph = {1.0, 0.0};
    int piv[N];
    for (int k = 0; k < N; ++k) {
      int p = k; double best = m[k * N + k].re * m[k * N + k].re + m[k * N + k].im * m[k * N + k].im;
      for (int r = k + 1; r < N; ++r) {
        double v = m[r * N + k].re * m[r * N + k].re + m[r * N + k].im * m[r * N + k].im;
        if (v > best) { best = v; p = r; }
      }
      piv[k] = p;
      if (p != k) {
        for (int j = 0; j < N; ++j) { cplx t = m[k * N + j]; m[k * N + j] = m[p * N + j]; m[p * N + j] = t; }
        ph.re = -ph.re; ph.im = -ph.im;
      }
      cplx pv = m[k * N + k];
      double a = sqrt(best);
      int e2; mant = frexp(mant * a, &e2); ex += e2;
      ph = cscale(cmul(ph, pv), 1.0 / a);
      cplx pinv = cinv(pv);
      m[k * N + k] = {1.0, 0.0};
      for (int j = 0; j < N; ++j) m[k * N + j] = cmul(m[k * N + j], pinv);
      for (int r = 0; r < N; ++r) {
        if (r == k) continue;
        cplx f = m[r * N + k];
        m[r * N + k] = {0.0, 0.0};
        for (int j = 0; j < N; ++j) cfms(m[r * N + j], f, m[k * N + j]);
      }
    }
    for (int k = N - 1; k >= 0; --k) {
      int p = piv[k];
      if (p != k)
        for (int r = 0; r < N; ++r) { cplx t = m[r * N + k]; m[r * N + k] = m[r * N + p]; m[r * N + p] = t; }
    }
    logabs = log(mant) + ex * 0.69314718055994530942;
    phase = atan2(ph.im, ph.re);
  }

  static AQ_HD int spin_block(const AiqmcSystem& sys, int i) { return i < sys.n_up ? 0 : 1; }

  // Primal state kept for the derivative pass.
  struct Primal {
    double G[3][2][N][4];     // sums over i in spin block of h_two^l[i,j,:]
    double h0[N][4 * A];      // layer-0 one-electron features
    double g0[2][4 * A];      // block means of h0
    double h[4][N][4];        // h[l] = one-electron stream after layer l-1 (h[1..3]); h[0] unused
    double g[3][2][4];        // block means of h[1], h[2] (index l = 1,2)
    double y[N][6];
    double env[N];
    double jae[N];            // per-electron electron-nucleus Jastrow term
    double jee[N];            // per-electron sum of the electron-electron Jastrow terms it takes part in
    double jastrow;
  };

  // ---- forward pass up to the orbital matrix; fills `pr`, returns M (row-major N x N)
  static constexpr int QM = (3 * NA + 2) > 5 ? (3 * NA + 2) : 5;     // first-stage width of a one-electron layer
  // hp: optional pair-chain cache [3][N][N][4] with element stride hp_stride;
  // t1: optional record of the first-stage tanh outputs [3][N][QM] with element stride t1_stride.
  static AQ_HD void forward(const AiqmcSystem& sys, const double* __restrict__ P, const double* __restrict__ x,
                            Primal& pr, cplx* __restrict__ M, double* __restrict__ hp = nullptr, int64_t hp_stride = 1,
                            double* __restrict__ t1 = nullptr, int64_t t1_stride = 1) {
    constexpr LayoutC<NE, NA> L{};
    const double inv_nup = 1.0 / sys.n_up, inv_ndn = 1.0 / sys.n_dn;
    double jas = 0.0;
    for (int e = 0; e < N; ++e) {
      double xe[3] = {x[3 * e], x[3 * e + 1], x[3 * e + 2]};
      double jae;
      electron_local<double>(P, e, xe, pr.h0[e], pr.y[e], pr.env[e], jae);
      pr.jae[e] = jae;
      pr.jee[e] = 0.0;
      jas += jae;
    }
    for (int s = 0; s < 2; ++s)
      for (int q = 0; q < 4 * A; ++q) pr.g0[s][q] = 0.0;
    for (int e = 0; e < N; ++e) {
      int s = spin_block(sys, e);
      for (int q = 0; q < 4 * A; ++q) pr.g0[s][q] += pr.h0[e][q];
    }
    for (int q = 0; q < 4 * A; ++q) { pr.g0[0][q] *= inv_nup; pr.g0[1][q] *= inv_ndn; }
    for (int l = 0; l < 3; ++l)
      for (int s = 0; s < 2; ++s)
        for (int j = 0; j < N; ++j)
          for (int c = 0; c < 4; ++c) pr.G[l][s][j][c] = 0.0;
    for (int i = 0; i < N; ++i) {
      const int s = spin_block(sys, i);
      for (int j = 0; j < N; ++j) {
        double d[3] = {x[3 * j] - x[3 * i], x[3 * j + 1] - x[3 * i + 1], x[3 * j + 2] - x[3 * i + 2]};
        double a0[4], a1[4], a2[4];
        pair_chain<double>(P, d, i == j, a0, a1, a2);
        for (int c = 0; c < 4; ++c) { pr.G[0][s][j][c] += a0[c]; pr.G[1][s][j][c] += a1[c]; pr.G[2][s][j][c] += a2[c]; }
        if (hp) {   // pair-chain cache for the single-electron-move kernels: hp[l][i][j][c]
          for (int c = 0; c < 4; ++c) {
            hp[(((0 * N + i) * N + j) * 4 + c) * hp_stride] = a0[c];
            hp[(((1 * N + i) * N + j) * 4 + c) * hp_stride] = a1[c];
            hp[(((2 * N + i) * N + j) * 4 + c) * hp_stride] = a2[c];
          }
        }
        if (i < j) {   // electron-electron Pade term (Jastrow.py:23-41)
          const double r = a0[0];
          const double u = P[L.jas_cusp + i * N + j] * r * s_inv(1.0 + P[L.jas_alpha + i * N + j] * r);
          jas += u;
          pr.jee[i] += u;
          pr.jee[j] += u;
        }
      }
    }
    pr.jastrow = jas;
    // layer 0
    for (int k = 0; k < N; ++k) {
      double Gu[4], Gd[4];
      for (int c = 0; c < 4; ++c) { Gu[c] = pr.G[0][0][k][c] * inv_nup; Gd[c] = pr.G[0][1][k][c] * inv_ndn; }
      one_layer<4 * A, double>(P, 0, k, pr.h0[k], pr.g0[0], pr.g0[1], Gu, Gd, pr.h[1][k],
                               t1 ? t1 + (int64_t)(k * QM) * t1_stride : nullptr, t1_stride);
    }
    for (int l = 1; l < 3; ++l) {
      for (int s = 0; s < 2; ++s) for (int c = 0; c < 4; ++c) pr.g[l][s][c] = 0.0;
      for (int k = 0; k < N; ++k) { int s = spin_block(sys, k); for (int c = 0; c < 4; ++c) pr.g[l][s][c] += pr.h[l][k][c]; }
      for (int c = 0; c < 4; ++c) { pr.g[l][0][c] *= inv_nup; pr.g[l][1][c] *= inv_ndn; }
      for (int k = 0; k < N; ++k) {
        double Gu[4], Gd[4];
        for (int c = 0; c < 4; ++c) { Gu[c] = pr.G[l][0][k][c] * inv_nup; Gd[c] = pr.G[l][1][k][c] * inv_ndn; }
        one_layer<4, double>(P, l, k, pr.h[l][k], pr.g[l][0], pr.g[l][1], Gu, Gd, pr.h[l + 1][k],
                             t1 ? t1 + (int64_t)((l * N + k) * QM) * t1_stride : nullptr, t1_stride);
      }
    }
    // orbital matrix M[k,j] = P[k,j] * env[k] * Yo[k,j]
    for (int k = 0; k < N; ++k) {
      const int e = sys.sigma[k], s = k < sys.n_up_rows ? 0 : 1;
      const double* W = P + L.orb_w[s];
      const double* B = P + L.orb_b[s];
      for (int j = 0; j < N; ++j) {
        double pre = B[2 * j], pim = B[2 * j + 1];
        for (int c = 0; c < 4; ++c) { pre += pr.h[3][e][c] * W[c * 2 * N + 2 * j]; pim += pr.h[3][e][c] * W[c * 2 * N + 2 * j + 1]; }
        double yo = 0.0;
        for (int m = 0; m < 6; ++m) yo += pr.y[k][m] * P[L.y_w + m * N + j];
        const double ev = pr.env[k] * yo;
        M[k * N + j] = {pre * ev, pim * ev};
      }
    }
  }

  static AQ_HD void write_cache(const Primal& pr, double logabs, double phase, double* __restrict__ cache) {
    using MC = MoveCache<NE, NA>;
    for (int l = 0; l < 3; ++l)
      for (int s = 0; s < 2; ++s)
        for (int j = 0; j < N; ++j)
          for (int c = 0; c < 4; ++c) cache[MC::GS + ((l * 2 + s) * N + j) * 4 + c] = pr.G[l][s][j][c];
    for (int s = 0; s < 2; ++s)
      for (int q = 0; q < 4 * A; ++q) cache[MC::G0M + s * 4 * A + q] = pr.g0[s][q];
    for (int e = 0; e < N; ++e) {
      for (int q = 0; q < 4 * A; ++q) cache[MC::H0 + e * 4 * A + q] = pr.h0[e][q];
      for (int m = 0; m < 6; ++m) cache[MC::Y + e * 6 + m] = pr.y[e][m];
      cache[MC::ENV + e] = pr.env[e];
      cache[MC::JAE + e] = pr.jae[e];
      cache[MC::JEE + e] = pr.jee[e];
    }
    cache[MC::MISC + 0] = pr.jastrow;
    cache[MC::MISC + 1] = logabs;
    cache[MC::MISC + 2] = phase;
    cache[MC::MISC + 3] = 0.0;
  }

  // ---- value only
  static AQ_HD void eval_value(const AiqmcSystem& sys, const double* __restrict__ P, const double* __restrict__ x,
                               double& phase, double& logabs) {
    Primal pr;
    cplx M[N * N];
    forward(sys, P, x, pr, M);
    double ld;
    lu_logdet(M, ld, phase);
    logabs = ld + pr.jastrow;
#ifdef __CUDA_ARCH__
    asm volatile("" ::"l"(&pr), "l"(M) : "memory");   // forbid stack-slot sharing of live arrays (DESIGN.md toolchain notes)
#endif
  }

  // ---- value + gradient (+ Laplacian of log|psi| if LAP)
  template <bool LAP>
  static AQ_HD void eval_deriv(const AiqmcSystem& sys, const double* __restrict__ P, const double* __restrict__ x,
                               double& phase, double& logabs, double* __restrict__ grad, double& lap,
                               double* __restrict__ cache = nullptr) {
    using J = Jet<LAP>;
    constexpr LayoutC<NE, NA> L{};
    using Op = ScalarOps<J>;
    Primal pr;
    cplx Mi[N * N];
    forward(sys, P, x, pr, Mi, cache ? cache + MoveCache<NE, NA>::HP : nullptr, 1);
    double ld;
    gj_inverse(Mi, ld, phase);
    logabs = ld + pr.jastrow;
    if (cache) write_cache(pr, logabs, phase, cache);
    const double inv_n[2] = {1.0 / sys.n_up, 1.0 / sys.n_dn};

    // G[k,c] = sum_j Wc[c,j] E[k,j] Minv[j,k];  T[k,c,l] likewise for every column l (LAP only)
    cplx Gm[N][4];
    cplx T[LAP ? N * 4 * N : 1];
    for (int k = 0; k < N; ++k) {
      const int s = k < sys.n_up_rows ? 0 : 1;
      const double* W = P + L.orb_w[s];
      for (int c = 0; c < 4; ++c) {
        Gm[k][c] = {0.0, 0.0};
        if (LAP) for (int l = 0; l < N; ++l) T[(k * 4 + c) * N + l] = {0.0, 0.0};
      }
      for (int j = 0; j < N; ++j) {
        double yo = 0.0;
        for (int m = 0; m < 6; ++m) yo += pr.y[k][m] * P[L.y_w + m * N + j];
        const double ev = pr.env[k] * yo;
        for (int c = 0; c < 4; ++c) {
          cplx w = {W[c * 2 * N + 2 * j] * ev, W[c * 2 * N + 2 * j + 1] * ev};
          cfma(Gm[k][c], w, Mi[j * N + k]);
          if (LAP) for (int l = 0; l < N; ++l) cfma(T[(k * 4 + c) * N + l], w, Mi[j * N + l]);
        }
      }
    }

    double lap_acc = 0.0;
    for (int e = 0; e < N; ++e) {
      const int se = spin_block(sys, e);
      // --- electron-local jets
      J xe[3];
      for (int c = 0; c < 3; ++c) { xe[c] = Op::cst(x[3 * e + c]); xe[c].d[c] = 1.0; }
      J h0e[4 * A], ye[6], enve, jaee;
      electron_local<J>(P, e, xe, h0e, ye, enve, jaee);
      double dJ[3], sJ[3];
      for (int c = 0; c < 3; ++c) { dJ[c] = jaee.d[c]; sJ[c] = jaee.s[c]; }
      // --- pair chains touching e: derivative parts of G
      //     rowJ[l][j]: change of G[l][se][j] from pair (e,j);  colJ[l][s]: change of G[l][s][e]
      J rowJ[3][N][4];
      J colJ[3][2][4];
      for (int l = 0; l < 3; ++l) {
        for (int c = 0; c < 4; ++c) { colJ[l][0][c] = Op::cst(0.0); colJ[l][1][c] = Op::cst(0.0); rowJ[l][e][c] = Op::cst(0.0); }
      }
      for (int j = 0; j < N; ++j) {
        if (j == e) continue;
        J d[3], a0[4], a1[4], a2[4];
        for (int c = 0; c < 3; ++c) { d[c] = Op::cst(x[3 * j + c] - x[3 * e + c]); d[c].d[c] = -1.0; }
        pair_chain<J>(P, d, false, a0, a1, a2);
        for (int c = 0; c < 4; ++c) { rowJ[0][j][c] = a0[c]; rowJ[1][j][c] = a1[c]; rowJ[2][j][c] = a2[c]; }
        {   // e-e Jastrow derivative wrt r_e
          const int lo = e < j ? e : j, hi = e < j ? j : e;
          const double cu = P[L.jas_cusp + lo * N + hi], al = P[L.jas_alpha + lo * N + hi];
          J r = a0[0];
          J u = (r * cu) * s_inv(r * al + 1.0);
          for (int c = 0; c < 3; ++c) { dJ[c] += u.d[c]; sJ[c] += u.s[c]; }
        }
        for (int c = 0; c < 3; ++c) { d[c].v = -d[c].v; d[c].d[c] = 1.0; }
        pair_chain<J>(P, d, false, a0, a1, a2);
        const int sj = spin_block(sys, j);
        for (int c = 0; c < 4; ++c) { colJ[0][sj][c] = colJ[0][sj][c] + a0[c]; colJ[1][sj][c] = colJ[1][sj][c] + a1[c]; colJ[2][sj][c] = colJ[2][sj][c] + a2[c]; }
      }
      // the jets above carry primal values too; only their derivative parts are used below.

      // --- one-electron stream jets for every electron k
      J hj[2][N][4];
      {   // layer 0
        J gup[4 * A], gdn[4 * A];
        for (int q = 0; q < 4 * A; ++q) {
          gup[q] = Op::cst(pr.g0[0][q]); gdn[q] = Op::cst(pr.g0[1][q]);
          J& gs = se == 0 ? gup[q] : gdn[q];
          for (int c = 0; c < 3; ++c) { gs.d[c] = h0e[q].d[c] * inv_n[se]; gs.s[c] = h0e[q].s[c] * inv_n[se]; }
        }
        for (int k = 0; k < N; ++k) {
          J hk[4 * A], Gu[4], Gd[4];
          for (int q = 0; q < 4 * A; ++q) hk[q] = (k == e) ? h0e[q] : Op::cst(pr.h0[k][q]);
          g_jets(pr, 0, k, e, se, rowJ[0][k], colJ[0], inv_n, Gu, Gd);
          one_layer<4 * A, J>(P, 0, k, hk, gup, gdn, Gu, Gd, hj[0][k]);
        }
      }
      int cur = 0;
      for (int l = 1; l < 3; ++l) {
        J gm[2][4];
        for (int s = 0; s < 2; ++s) for (int c = 0; c < 4; ++c) gm[s][c] = Op::cst(0.0);
        for (int k = 0; k < N; ++k) { int s = spin_block(sys, k); for (int c = 0; c < 4; ++c) gm[s][c] = gm[s][c] + hj[cur][k][c]; }
        for (int c = 0; c < 4; ++c) { gm[0][c] = gm[0][c] * inv_n[0]; gm[1][c] = gm[1][c] * inv_n[1]; }
        for (int k = 0; k < N; ++k) {
          J Gu[4], Gd[4];
          g_jets(pr, l, k, e, se, rowJ[l][k], colJ[l], inv_n, Gu, Gd);
          one_layer<4, J>(P, l, k, hj[cur][k], gm[0], gm[1], Gu, Gd, hj[cur ^ 1][k]);
        }
        cur ^= 1;
      }
      // hj[cur][k][c] = jets of the final one-electron features

      // --- row e of E = env * Yo as jets, and P[e,:] (row e reads h of electron sigma[e])
      const int srow = e < sys.n_up_rows ? 0 : 1;
      const double* W = P + L.orb_w[srow];
      const double* B = P + L.orb_b[srow];
      const int eh = sys.sigma[e];
      for (int dir = 0; dir < 3; ++dir) {
        // S1[l] = sum_j P[e,j] dE[e,j] Minv[j,l];  S2 = sum_j (2 dP dE + P d2E) Minv[j,e]
        cplx S1[LAP ? N : 1];
        cplx S1e = {0.0, 0.0}, S2 = {0.0, 0.0};
        if (LAP) for (int l = 0; l < N; ++l) S1[l] = {0.0, 0.0};
        for (int j = 0; j < N; ++j) {
          J yo = Op::cst(0.0);
          for (int m = 0; m < 6; ++m) yo = yo + ye[m] * P[L.y_w + m * N + j];
          J E = enve * yo;
          cplx p = {B[2 * j], B[2 * j + 1]}, dp = {0.0, 0.0};
          for (int c = 0; c < 4; ++c) {
            const double wr = W[c * 2 * N + 2 * j], wi = W[c * 2 * N + 2 * j + 1];
            p.re += pr.h[3][eh][c] * wr; p.im += pr.h[3][eh][c] * wi;
            dp.re += hj[cur][eh][c].d[dir] * wr; dp.im += hj[cur][eh][c].d[dir] * wi;
          }
          cplx pdE = cscale(p, E.d[dir]);
          cfma(S1e, pdE, Mi[j * N + e]);
          if (LAP) {
            for (int l = 0; l < N; ++l) cfma(S1[l], pdE, Mi[j * N + l]);
            cplx t2 = cadd(cscale(dp, 2.0 * E.d[dir]), cscale(p, E.s[dir]));
            cfma(S2, t2, Mi[j * N + e]);
          }
        }
        // gradient: Re tr(X) = Re [ sum_k sum_c dh[sigma_k,c] G[k,c] + S1[e] ]
        double gsum = S1e.re;
        for (int k = 0; k < N; ++k) {
          const int ek = sys.sigma[k];
          for (int c = 0; c < 4; ++c) gsum += hj[cur][ek][c].d[dir] * Gm[k][c].re;
        }
        grad[3 * e + dir] = gsum + dJ[dir];
        if (LAP) {
          double l2 = S2.re + sJ[dir];
          for (int k = 0; k < N; ++k) {
            const int ek = sys.sigma[k];
            for (int c = 0; c < 4; ++c) l2 += hj[cur][ek][c].s[dir] * Gm[k][c].re;
          }
          // X[k,l] = sum_c dh[sigma_k,c] T[k,c,l] + delta_ke S1[l];  subtract Re tr(X X)
          cplx X[N * N];
          for (int k = 0; k < N; ++k) {
            const int ek = sys.sigma[k];
            for (int l = 0; l < N; ++l) {
              cplx acc = (k == e) ? S1[l] : cplx{0.0, 0.0};
              for (int c = 0; c < 4; ++c) {
                const double dh = hj[cur][ek][c].d[dir];
                acc.re += dh * T[(k * 4 + c) * N + l].re; acc.im += dh * T[(k * 4 + c) * N + l].im;
              }
              X[k * N + l] = acc;
            }
          }
          double trxx = 0.0;
          for (int k = 0; k < N; ++k)
            for (int l = 0; l < N; ++l) trxx += X[k * N + l].re * X[l * N + k].re - X[k * N + l].im * X[l * N + k].im;
          lap_acc += l2 - trxx;
        }
      }
    }
    lap = lap_acc;
  }

  // jets of G_l[.][k]/n for electron k when electron e moves
  template <class J>
  static AQ_HD void g_jets(const Primal& pr, int l, int k, int e, int se, const J* rowk, const J (*col)[4],
                           const double inv_n[2], J Gu[4], J Gd[4]) {
    using Op = ScalarOps<J>;
    for (int c = 0; c < 4; ++c) {
      Gu[c] = Op::cst(pr.G[l][0][k][c] * inv_n[0]);
      Gd[c] = Op::cst(pr.G[l][1][k][c] * inv_n[1]);
      if (k == e) {
        for (int q = 0; q < 3; ++q) {
          Gu[c].d[q] = col[0][c].d[q] * inv_n[0]; Gu[c].s[q] = col[0][c].s[q] * inv_n[0];
          Gd[c].d[q] = col[1][c].d[q] * inv_n[1]; Gd[c].s[q] = col[1][c].s[q] * inv_n[1];
        }
      } else {
        J& g = se == 0 ? Gu[c] : Gd[c];
        for (int q = 0; q < 3; ++q) { g.d[q] = rowk[c].d[q] * inv_n[se]; g.s[q] = rowk[c].s[q] * inv_n[se]; }
      }
    }
  }
};

}  // namespace aiqmc
