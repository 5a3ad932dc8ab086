// ecp_pt.cuh -- the ccECP non-local quadrature (pseudopotential.py:272-318, pp_energy_test.py:45-105),
// B200 version 3 for small systems (N <= 4): ONE THREAD PER QUADRATURE POINT on top of the
// single-electron-move cache.
//
// Why a third version: ncu on k_ecp_coop (lane-per-electron) showed 69 % of the issued instructions were
// not FP64 (shuffles, selects, LDS of weights) and every lane recomputed electron i's local part, so the
// FP64 pipe sat at 35 %.  Here a thread owns a whole point, nothing is computed twice and nothing is
// exchanged between lanes:
//   * parameters live in __constant__ memory; all offsets are immediates, so a weight is a c[3][imm]
//     operand of the DFMA itself -- no load instruction at all;
//   * the walker's MoveCache sits in shared memory; lanes of a warp share (walker, electron) except at
//     50-point boundaries, so cache reads are broadcasts;
//   * of the network only what depends on the displaced electron i is recomputed: i's local part, the
//     2(N-1) pair chains through i, the N one-electron rows (via cached block sums) and the determinant;
//   * determinants of order <= 4 by cofactor expansion of complex 2x2 minors (no pivot search, no
//     division, 120 FP64 ops at N=4);
//   * one CTA = WPC walkers; per-point contributions are staged in shared memory and summed in a fixed
//     order (deterministic, no atomics).
#pragma once
#include "psi_core.cuh"

namespace aiqmc {

#ifndef AIQMC_PT_WPC
#define AIQMC_PT_WPC 1             // walkers per CTA
#endif
#ifndef AIQMC_PT_MINB
#define AIQMC_PT_MINB 2            // resident CTAs/SM the register allocator must allow
#endif
#ifndef AIQMC_PT_ACC
#define AIQMC_PT_ACC 1             // tanh variant of the value-only quadrature (fastmath.cuh: 1 = 9-op, < 5e-11)
#endif
constexpr int kAcc = AIQMC_PT_ACC;
constexpr int kConstParMax = 3072;           // doubles of packed parameters kept in constant memory (24 kB)
static __constant__ double c_par[kConstParMax];

#ifndef AIQMC_PT_WPCI
#define AIQMC_PT_WPCI 5            // walkers per CTA of the per-electron specialisations: 5 x 50 points = 8 warps (97.6 % of lanes busy, no warp-allocation waste)
#endif
// IFIX >= 0: the kernel handles the points of electron IFIX only (one launch per electron): every `k == i`
// test, cache offset and chain skip becomes a compile-time decision -- no selects, no branches.
// IFIX < 0: one launch, i decoded per thread (kept for cross-checks).
template <int NE, int NA, int WPC, int IFIX>
constexpr int pt_points() { return (IFIX >= 0 ? 1 : NE) * NA * AIQMC_NQUAD; }
template <int NE, int NA, int WPC, int IFIX>
constexpr int pt_threads() { return ((WPC * pt_points<NE, NA, WPC, IFIX>() + 31) / 32) * 32; }

AQ_HD cplx cminor(cplx a0, cplx a1, cplx b0, cplx b1) { return csub(cmul(a0, b1), cmul(a1, b0)); }

// determinant of an N x N complex matrix held in registers (rows m[k][.]), N <= 4
template <int N>
AQ_HD cplx det_small(const cplx (*m)[N]) {
  if constexpr (N == 1) {
    return m[0][0];
  } else if constexpr (N == 2) {
    return cminor(m[0][0], m[0][1], m[1][0], m[1][1]);
  } else if constexpr (N == 3) {
    cplx d = cmul(m[0][0], cminor(m[1][1], m[1][2], m[2][1], m[2][2]));
    cfms(d, m[0][1], cminor(m[1][0], m[1][2], m[2][0], m[2][2]));
    cfma(d, m[0][2], cminor(m[1][0], m[1][1], m[2][0], m[2][1]));
    return d;
  } else {
    static_assert(N == 4, "det_small: N <= 4");
    const cplx s01 = cminor(m[0][0], m[0][1], m[1][0], m[1][1]), s02 = cminor(m[0][0], m[0][2], m[1][0], m[1][2]),
               s03 = cminor(m[0][0], m[0][3], m[1][0], m[1][3]), s12 = cminor(m[0][1], m[0][2], m[1][1], m[1][2]),
               s13 = cminor(m[0][1], m[0][3], m[1][1], m[1][3]), s23 = cminor(m[0][2], m[0][3], m[1][2], m[1][3]);
    const cplx c01 = cminor(m[2][0], m[2][1], m[3][0], m[3][1]), c02 = cminor(m[2][0], m[2][2], m[3][0], m[3][2]),
               c03 = cminor(m[2][0], m[2][3], m[3][0], m[3][3]), c12 = cminor(m[2][1], m[2][2], m[3][1], m[3][2]),
               c13 = cminor(m[2][1], m[2][3], m[3][1], m[3][3]), c23 = cminor(m[2][2], m[2][3], m[3][2], m[3][3]);
    cplx d = cmul(s01, c23);
    cfms(d, s02, c13);
    cfma(d, s03, c12);
    cfma(d, s12, c03);
    cfms(d, s13, c02);
    cfma(d, s23, c01);
    return d;
  }
}

// T-move record of one quadrature point (DMC/Tmoves.py:88-112): amplitude = ratio * sum_l (exp(-tau v_l) - 1) P_l(cos),
// kept only if it is "> 0" in jnp's lexicographic complex order (quirk Q25); out = [fwd.re, fwd.im, ratio.re, ratio.im]
__device__ __forceinline__ void tmove_point_out(double* __restrict__ out, double v0, double v1, double v2, double v3,
                                                double cs, double rr, double ri, double tau) {
  const double k4 = 0.07957747154594767;   // 1/(4 pi)
  const double ws = (exp(-tau * v0) - 1.0) * k4 + (exp(-tau * v1) - 1.0) * (3.0 * k4 * cs) +
                    (exp(-tau * v2) - 1.0) * (2.5 * k4 * (3.0 * cs * cs - 1.0)) +
                    (exp(-tau * v3) - 1.0) * (3.5 * k4 * (5.0 * cs * cs * cs - 3.0 * cs));
  const double tr = rr * ws, ti = ri * ws;
  const bool keep = (tr > 0.0) || (tr == 0.0 && ti > 0.0);
  out[0] = keep ? tr : 0.0;
  out[1] = keep ? ti : 0.0;
  out[2] = rr;
  out[3] = ri;
}

// Orbital-matrix row (nn.py:432-504): out[j] = (h . W[:, j] + b[j]) * env * (y . Yw[:, j]); SROW picks the
// spin block's weights at compile time so every weight stays an immediate constant-bank operand.
template <int NE, int NA, int SROW>
__device__ __forceinline__ void orbital_row(const double* __restrict__ P, const double hs[4], const double yr[6],
                                            double envr, cplx out[NE]) {
  constexpr LayoutC<NE, NA> L{};
  constexpr int N = NE;
  const double* W = P + L.orb_w[SROW];
  const double* Bv = P + L.orb_b[SROW];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double pre = Bv[2 * j], pim = Bv[2 * j + 1];
#pragma unroll
    for (int c = 0; c < 4; ++c) { pre += hs[c] * W[c * 2 * N + 2 * j]; pim += hs[c] * W[c * 2 * N + 2 * j + 1]; }
    double yo = 0.0;
#pragma unroll
    for (int m = 0; m < 6; ++m) yo += yr[m] * P[L.y_w + m * N + j];
    const double evv = envr * yo;
    out[j] = {pre * evv, pim * evv};
  }
}

#ifdef AIQMC_PT_MAXREG
#define AIQMC_PT_BOUNDS __maxnreg__(AIQMC_PT_MAXREG)
#else
#define AIQMC_PT_BOUNDS __launch_bounds__((pt_threads<NE, NA, WPC, IFIX>()), AIQMC_PT_MINB)
#endif
// LAYOUT >= 0: the spin layout is compile-time too: symmetric-feature blocks [0,NUP) / [NUP,N) with NUP = ceil(N/2)
// and orbital rows either in electron order (LAYOUT 0: spins up-first, sigma = identity) or up-spins-then-down-spins
// of an alternating spin list (LAYOUT 1: sigma = 0,2,4,..,1,3,5,.. -- the reference's examples, e.g.
// example/single_atom_C: spins = [1,-1,1,-1]).  The ~650 FSELs per point that pick spin blocks disappear.  The
// launcher checks the system against these layouts and falls back to LAYOUT = -1 (run-time layout) otherwise.
template <int NE, int LAYOUT>
__host__ __device__ constexpr int pt_sigma(int k) {
  constexpr int NUP = (NE + 1) / 2;
  return LAYOUT == 0 ? k : (k < NUP ? 2 * k : 2 * (k - NUP) + 1);
}
template <int NE, int NA, int WPC, int IFIX, int LAYOUT>
__global__ void AIQMC_PT_BOUNDS
k_ecp_pt(AiqmcSystem sys, const double* __restrict__ pos, const double* __restrict__ rot, int64_t B,
         const double* __restrict__ cache_all, EnergyWs w, double* __restrict__ tm_out, double tm_tau) {
  constexpr int N = NE, A = NA;
  constexpr int E = pt_points<NE, NA, WPC, IFIX>();
  using MC = MoveCache<NE, NA>;
  constexpr LayoutC<NE, NA> L{};
  static_assert(MC::SIZE % 2 == 0, "MoveCache records must keep 16-byte alignment");
  __shared__ __align__(16) double sC[WPC][MC::SIZE];
  __shared__ double sAcc[WPC][E][2];
  __shared__ __align__(8) unsigned long long sBar;
  const int tid = threadIdx.x;
  const int64_t b0 = (int64_t)blockIdx.x * WPC;
  // The CTA's WPC MoveCache records -- network intermediates AND the walker's positions / rotation / norms /
  // denominator / v_l tables (MoveCache::QR) -- are one contiguous, 16-byte aligned block of HBM (array-of-structs,
  // consecutive walkers): it is staged by ONE TMA bulk copy (cp.async.bulk, SASS UBLKCP) signalled on an mbarrier;
  // no thread issues a global load in this kernel.
  {
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&sBar);
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      const int64_t nw = (B - b0 < WPC) ? (B - b0) : WPC;
      const unsigned bytes = (unsigned)(nw * MC::SIZE * sizeof(double));
      const unsigned dst = (unsigned)__cvta_generic_to_shared(&sC[0][0]);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst), "l"(cache_all + b0 * MC::SIZE), "r"(bytes), "r"(bar) : "memory");
    }
    if (tid < kExpTab) g_exp_tab[tid] = exp2((double)tid * (1.0 / kExpTab));
    if (kAcc == 1 && AIQMC_TANH_TAB64) fill_tanh_table(tid, (int)blockDim.x);
    __syncthreads();                     // publishes the mbarrier initialisation to the waiting threads
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "AQ_WAIT_CACHE:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
        "@p bra AQ_CACHE_READY;\n"
        "bra AQ_WAIT_CACHE;\n"
        "AQ_CACHE_READY:\n"
        "}\n" ::"r"(bar) : "memory");
  }

  const double* P = c_par;
  const int wl = tid / E;
  const int ev = tid - wl * E;
  const int64_t b = b0 + wl;
  if (wl < WPC && b < B) {
    const double* C = sC[wl];
    const double* X = C + MC::QR;
    const int i = IFIX >= 0 ? IFIX : ev / (A * AIQMC_NQUAD);
    const int a = (IFIX >= 0 ? ev : ev - i * A * AIQMC_NQUAD) / AIQMC_NQUAD;
    const int p = (IFIX >= 0 ? ev : ev - i * A * AIQMC_NQUAD) - a * AIQMC_NQUAD;
    const double* vl = C + MC::QR_VL + (i * A + a) * 4;
    const double v0 = vl[0], v1 = vl[1], v2 = vl[2], v3 = vl[3];
    double out_re = 0.0, out_im = 0.0;
    if (tm_out || !(v0 == 0.0 && v1 == 0.0 && v2 == 0.0 && v3 == 0.0)) {   // exact zero channel contributes exactly 0
      constexpr int NUP = (N + 1) / 2;
      const int n_up = LAYOUT >= 0 ? NUP : sys.n_up, n_dn = LAYOUT >= 0 ? N - NUP : sys.n_dn;
      const int n_up_rows = LAYOUT >= 0 ? NUP : sys.n_up_rows;
      const double inv_n[2] = {1.0 / n_up, 1.0 / n_dn};
      const int si = i < n_up ? 0 : 1;
      // ---- rotated point and cos(theta)  (quirks Q13, Q14)
      double ae[3], xn[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) ae[c] = X[3 * i + c] - P[L.atoms + 3 * a + c];
      const double r = sqrt(ae[0] * ae[0] + ae[1] * ae[1] + ae[2] * ae[2]);
      double dot = 0.0;
#pragma unroll
      for (int l = 0; l < 3; ++l) {
        const double nh = c_ecp.quad_pts[p][0] * X[3 * N + l] + c_ecp.quad_pts[p][1] * X[3 * N + 3 + l] +
                          c_ecp.quad_pts[p][2] * X[3 * N + 6 + l];
        xn[l] = r * nh;
        dot += ae[l] * xn[l];
      }
      const double cs = dot / (r * (r * X[3 * N + 9 + quad_group(p)]));

      // ---- electron i at its new position: features, Ynlm stream, envelope, e-n Jastrow
      double h0n[4 * A], yn[6], envn, jaen;
      Psi<NE, NA>::template electron_local<double, kAcc>(P, i, xn, h0n, yn, envn, jaen);

      // ---- level-0 pair features through i: row (i,k): d = x_k - x_i', column (k,i): -d; e-e Jastrow
      double cr[N][4], cc[N][4];
      double jee = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const bool diag = (k == i);
        double d[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) d[c] = diag ? 0.0 : X[3 * k + c] - xn[c];
        const double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        const double rik = diag ? 0.0 : r2 * s_rsqrt(diag ? 1.0 : r2);
        cr[k][0] = rik; cc[k][0] = rik;
#pragma unroll
        for (int c = 0; c < 3; ++c) { cr[k][1 + c] = d[c]; cc[k][1 + c] = -d[c]; }
        const int lo = i < k ? i : k, hi = i < k ? k : i;
        jee += P[L.jas_cusp + lo * N + hi] * rik * s_inv(1.0 + P[L.jas_alpha + lo * N + hi] * rik);   // 0 on the diagonal
      }

      // ---- one-electron stream of all N electrons; the pair chains advance with the layers
      double h[N][4];
#ifndef AIQMC_PT_NO_ROLL   // layers share one copy of the row code (7.8 k instead of 10 k SASS instructions, 3 % faster)
#pragma unroll 1
#else
#pragma unroll
#endif
      for (int l = 0; l < 3; ++l) {
        // column block sums for electron i: G'_l[s][i] = sum_{k in s, k != i} h'_l[k,i] + [s == s_i] h_l[i,i]
        double su[4], sd[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const double dg = C[MC::HP + ((l * N + i) * N + i) * 4 + c];
          su[c] = si == 0 ? dg : 0.0;
          sd[c] = si == 0 ? 0.0 : dg;
#pragma unroll
          for (int k = 0; k < N; ++k) {
            const double v = (k == i) ? 0.0 : cc[k][c];
            if (k < n_up) su[c] += v; else sd[c] += v;
          }
        }
        double g0u[4 * A], g0d[4 * A], gm[2][4];
        if (l == 0) {
#pragma unroll
          for (int q = 0; q < 4 * A; ++q) {
            const double dh = h0n[q] - C[MC::H0 + i * 4 * A + q];
            g0u[q] = C[MC::G0M + q] + (si == 0 ? dh * inv_n[0] : 0.0);
            g0d[q] = C[MC::G0M + 4 * A + q] + (si == 1 ? dh * inv_n[1] : 0.0);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            double u = 0.0, d = 0.0;
#pragma unroll
            for (int k = 0; k < N; ++k) { if (k < n_up) u += h[k][c]; else d += h[k][c]; }
            gm[0][c] = u * inv_n[0];
            gm[1][c] = d * inv_n[1];
          }
        }
#pragma unroll
        for (int k = 0; k < N; ++k) {
          const bool diag = (k == i);
          double Gu[4], Gd[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            double gu = C[MC::GS + ((l * 2 + 0) * N + k) * 4 + c], gd = C[MC::GS + ((l * 2 + 1) * N + k) * 4 + c];
            const double delta = cr[k][c] - C[MC::HP + ((l * N + i) * N + k) * 4 + c];
            if (si == 0) gu += delta; else gd += delta;
            Gu[c] = (diag ? su[c] : gu) * inv_n[0];
            Gd[c] = (diag ? sd[c] : gd) * inv_n[1];
          }
          if (l == 0) {
            double hk[4 * A];
#pragma unroll
            for (int q = 0; q < 4 * A; ++q) hk[q] = diag ? h0n[q] : C[MC::H0 + k * 4 * A + q];
            Psi<NE, NA>::template one_layer<4 * A, double, kAcc>(P, 0, k, hk, g0u, g0d, Gu, Gd, h[k]);
          } else {
            double hn[4];
            Psi<NE, NA>::template one_layer<4, double, kAcc>(P, l, k, h[k], gm[0], gm[1], Gu, Gd, hn);
#pragma unroll
            for (int c = 0; c < 4; ++c) h[k][c] = hn[c];
          }
        }
        if (l < 2) {   // advance the 2(N-1) pair chains through double-layer l (nn.py:305-309)
          const double* W = P + L.dbl_w[l];
#pragma unroll
          for (int k = 0; k < N; ++k) {
            if (k != i) {          // i is warp-uniform except at 50-point boundaries: a real skip, not a predicate
              double z[8], t[8];               // row chain in z[0..3], column chain in z[4..7]
#pragma unroll
              for (int m = 0; m < 4; ++m) { z[m] = P[L.dbl_b[l] + m]; z[4 + m] = z[m]; }
#pragma unroll
              for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int m = 0; m < 4; ++m) { z[m] += cr[k][q] * W[q * 4 + m]; z[4 + m] += cc[k][q] * W[q * 4 + m]; }
#pragma unroll
              for (int m = 0; m < 4; ++m) { t[m] = cr[k][m]; t[4 + m] = cc[k][m]; }
              tanh_res<8, kAcc>(z, t, t);
#pragma unroll
              for (int m = 0; m < 4; ++m) { cr[k][m] = t[m]; cc[k][m] = t[4 + m]; }
            }
          }
        }
      }

      // ---- orbital matrix: row k reads h of electron sigma[k], envelope / Ynlm of electron k (quirk Q4)
      cplx M[N][N];
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const bool diag = (k == i);
        const int e = LAYOUT >= 0 ? pt_sigma<NE, LAYOUT>(k) : sys.sigma[k];
        double hs[4], yr[6];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          double v = h[0][c];
#pragma unroll
          for (int q = 1; q < N; ++q) v = (e == q) ? h[q][c] : v;
          hs[c] = v;
        }
#pragma unroll
        for (int m = 0; m < 6; ++m) yr[m] = diag ? yn[m] : C[MC::Y + k * 6 + m];
        const double envr = diag ? envn : C[MC::ENV + k];
        if (k < n_up_rows) orbital_row<NE, NA, 0>(P, hs, yr, envr, M[k]);
        else orbital_row<NE, NA, 1>(P, hs, yr, envr, M[k]);
      }
      const cplx det = det_small<N>(M);
      const double la = 0.5 * log(det.re * det.re + det.im * det.im) + C[MC::MISC + 0] + (jee - C[MC::JEE + i]) +
                        (jaen - C[MC::JAE + i]);
      const double pha = atan2(det.im, det.re);
      // ratio = log psi(x') / log psi(x) * weight with complex logs (quirk Q12)
      const double den_r = X[3 * N + 13], den_i = X[3 * N + 14];
      const double wq = c_ecp.quad_wts[p] / (den_r * den_r + den_i * den_i);
      const double rr = (la * den_r + pha * den_i) * wq, ri = (pha * den_r - la * den_i) * wq;
      const double k4 = 0.07957747154594767;   // 1/(4 pi)
      const double f = v0 * k4 + v1 * (3.0 * k4 * cs) + v2 * (2.5 * k4 * (3.0 * cs * cs - 1.0)) +
                       v3 * (3.5 * k4 * (5.0 * cs * cs * cs - 3.0 * cs));
      out_re = f * rr;
      out_im = f * ri;
      if (tm_out) tmove_point_out(tm_out + (((b * N + i) * A + a) * AIQMC_NQUAD + p) * 4, v0, v1, v2, v3, cs, rr, ri, tm_tau);
    }
    sAcc[wl][ev][0] = out_re;
    sAcc[wl][ev][1] = out_im;
  }
  __syncthreads();
  // fixed-order sum of the E contributions of each walker: warp wv handles walker wv
  const int wv = tid >> 5, lane = tid & 31;
  if (!tm_out && wv < WPC && b0 + wv < B) {
    double sr = 0.0, si2 = 0.0;
    for (int q = lane; q < E; q += 32) { sr += sAcc[wv][q][0]; si2 += sAcc[wv][q][1]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sr += __shfl_xor_sync(0xffffffffu, sr, o);
      si2 += __shfl_xor_sync(0xffffffffu, si2, o);
    }
    if (lane == 0) {
      if (IFIX <= 0) { w.epp[2 * (b0 + wv)] = sr; w.epp[2 * (b0 + wv) + 1] = si2; }
      else {   // one add per walker per launch, launches stream-ordered: the sum order is fixed; RED (no return value)
               // lets the CTA retire without waiting for a load of the accumulator
        atomicAdd(&w.epp[2 * (b0 + wv)], sr);
        atomicAdd(&w.epp[2 * (b0 + wv) + 1], si2);
      }
    }
  }
}

}  // namespace aiqmc
