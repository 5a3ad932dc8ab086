// deriv_split.cuh -- gradient / Laplacian of log|psi| as TWO passes (B200 version 2 of the derivative path;
// replaces the one-thread-does-everything eval_deriv in the kernels, which needed 255 registers plus a
// 12 kB stack frame per thread and ran at 2 warps per scheduler).
//
//   primal pass   (one thread per configuration): forward network, Gauss-Jordan inverse, the contractions of
//                 M^-1 with the orbital weights (Gm, and T for the Laplacian), and a structure-of-arrays
//                 cache of every intermediate a derivative needs -- including the OUTPUTS of the tanh layers;
//   tangent pass  (one thread per configuration x electron x Cartesian direction): pushes a first (and, for
//                 the Laplacian, second) derivative through the parts of the network that depend on that
//                 coordinate.  Because tanh' = 1 - t^2 and every t is either cached or recoverable from the
//                 cached layer outputs (t = sqrt2 * h_{l+1} - h_l), this pass evaluates NO transcendental
//                 except the 18 tanh / 5 exp of the moved electron's own Ynlm / envelope part; it is pure
//                 DFMA work on ~60 live doubles, so it runs at full occupancy.
// The cache is laid out [slot][configuration]: the primal pass writes and the tangent pass reads fully
// coalesced (a warp = 32 configurations, same electron, same direction -> no divergence either).
//
// The mathematics is that of Psi::eval_deriv (psi_core.cuh header): d log det = tr(X),
// d2 log det = tr(M^-1 d2M) - tr(X X),  X = dh . T + delta_ke S1.
#pragma once
#include "psi_core.cuh"

// The tangent pass keeps ~170 doubles of per-thread state in small arrays indexed by loop variables; nvcc leaves loops
// with large inlined bodies rolled, so those arrays live in local memory (1.3 kB of stack per thread at N = 4).
// Measured alternative (AIQMC_TANGENT_UNROLL): full unrolling makes every index a compile-time constant, but the state
// does not fit 168 registers (3.5 kB of spills) and at 255 registers the kernel lost a third of its occupancy and
// 3.5 minutes of compile time per instantiation: 1.52 ms against 0.97 ms for the kinetic stage of carbon.  Rolled it is.
#if defined(__CUDACC__) && defined(AIQMC_TANGENT_UNROLL)
#define AQ_PRAGMA_(x) _Pragma(#x)
#define AQ_UNROLL_N AQ_PRAGMA_(unroll (NE <= 6 ? 64 : 1))
#else
#define AQ_UNROLL_N
#endif

namespace aiqmc {

template <int NE, int NA>
struct DerivCache {
  static constexpr int QM = Psi<NE, NA>::QM;
  static constexpr int X = 0;                               // [3N]            positions
  static constexpr int HP = X + 3 * NE;                     // [3][N][N][4]    pair chains h_two^l[i,j,:]
  static constexpr int GS = HP + 3 * NE * NE * 4;           // [3][2][N][4]    block sums over i of h_two^l[i,j,:]
  static constexpr int H0 = GS + 3 * 2 * NE * 4;            // [N][4A]         layer-0 one-electron features
  static constexpr int G0M = H0 + NE * 4 * NA;              // [2][4A]         block means of H0
  static constexpr int H = G0M + 2 * 4 * NA;                // [3][N][4]       one-electron stream after layer 0,1,2
  static constexpr int GM = H + 3 * NE * 4;                 // [2][2][4]       block means of H[0], H[1]
  static constexpr int T1 = GM + 2 * 2 * 4;                 // [3][N][QM]      first-stage tanh outputs
  static constexpr int MI = T1 + 3 * NE * QM;               // [N][N][2]       M^-1
  static constexpr int GMAT = MI + NE * NE * 2;             // [N][4][2]       Gm[k][c] = sum_j W[c,j] E[k,j] Minv[j,k]
  static constexpr int YV = GMAT + NE * 4 * 2;              // [N][6]          Ynlm stream outputs
  static constexpr int ENVV = YV + NE * 6;                  // [N]             envelopes
  static constexpr int SIZE_GRAD = ENVV + NE;
  static constexpr int TT = SIZE_GRAD;                      // [N][4][N][2]    T[k][c][l] (Laplacian only)
  static constexpr int SIZE_LAP = TT + NE * 4 * NE * 2;
  static constexpr int size(bool lap) { return lap ? SIZE_LAP : SIZE_GRAD; }
};

template <int NE, int NA>
struct DerivSplit {
  using PS = Psi<NE, NA>;
  using DC = DerivCache<NE, NA>;
  static constexpr int N = NE, A = NA, QM = PS::QM;
  static constexpr double kSqrt2 = 1.41421356237309504880;

  // ---- primal pass: fills dc[slot * stride] (SoA); optionally the AoS MoveCache `mc` of the quadrature kernels
  template <bool LAP>
  static AQ_HD void primal(const AiqmcSystem& sys, const double* __restrict__ P, const double* __restrict__ x,
                           double* __restrict__ dc, int64_t stride, double* __restrict__ mc, double& phase,
                           double& logabs) {
    constexpr LayoutC<NE, NA> L{};
    typename PS::Primal pr;
    cplx Mi[N * N];
    PS::forward(sys, P, x, pr, Mi, dc + (int64_t)DC::HP * stride, stride, dc + (int64_t)DC::T1 * stride, stride);
    double ld;
    PS::gj_inverse(Mi, ld, phase);
    logabs = ld + pr.jastrow;
    if (mc) {
      using MC = MoveCache<NE, NA>;
      for (int q = 0; q < 3 * N * N * 4; ++q) mc[MC::HP + q] = dc[(int64_t)(DC::HP + q) * stride];
      PS::write_cache(pr, logabs, phase, mc);
    }
    for (int q = 0; q < 3 * N; ++q) dc[(int64_t)(DC::X + q) * stride] = x[q];
    for (int l = 0; l < 3; ++l)
      for (int s = 0; s < 2; ++s)
        for (int j = 0; j < N; ++j)
          for (int c = 0; c < 4; ++c) dc[(int64_t)(DC::GS + ((l * 2 + s) * N + j) * 4 + c) * stride] = pr.G[l][s][j][c];
    for (int e = 0; e < N; ++e)
      for (int q = 0; q < 4 * A; ++q) dc[(int64_t)(DC::H0 + e * 4 * A + q) * stride] = pr.h0[e][q];
    for (int s = 0; s < 2; ++s)
      for (int q = 0; q < 4 * A; ++q) dc[(int64_t)(DC::G0M + s * 4 * A + q) * stride] = pr.g0[s][q];
    for (int l = 0; l < 3; ++l)
      for (int k = 0; k < N; ++k)
        for (int c = 0; c < 4; ++c) dc[(int64_t)(DC::H + (l * N + k) * 4 + c) * stride] = pr.h[l + 1][k][c];
    for (int l = 1; l < 3; ++l)
      for (int s = 0; s < 2; ++s)
        for (int c = 0; c < 4; ++c) dc[(int64_t)(DC::GM + ((l - 1) * 2 + s) * 4 + c) * stride] = pr.g[l][s][c];
    for (int q = 0; q < N * N; ++q) {
      dc[(int64_t)(DC::MI + 2 * q) * stride] = Mi[q].re;
      dc[(int64_t)(DC::MI + 2 * q + 1) * stride] = Mi[q].im;
    }
    for (int k = 0; k < N; ++k) {
      for (int m = 0; m < 6; ++m) dc[(int64_t)(DC::YV + k * 6 + m) * stride] = pr.y[k][m];
      dc[(int64_t)(DC::ENVV + k) * stride] = pr.env[k];
    }
    // Gm[k,c] = sum_j W[c,j] E[k,j] Minv[j,k];  T[k,c,l] likewise for every column l (Laplacian only)
    for (int k = 0; k < N; ++k) {
      const int s = k < sys.n_up_rows ? 0 : 1;
      const double* W = P + L.orb_w[s];
      cplx Gm[4];
      cplx T[LAP ? 4 * N : 1];
      for (int c = 0; c < 4; ++c) {
        Gm[c] = {0.0, 0.0};
        if (LAP) for (int l = 0; l < N; ++l) T[c * N + l] = {0.0, 0.0};
      }
      for (int j = 0; j < N; ++j) {
        double yo = 0.0;
        for (int m = 0; m < 6; ++m) yo += pr.y[k][m] * P[L.y_w + m * N + j];
        const double ev = pr.env[k] * yo;
        for (int c = 0; c < 4; ++c) {
          const cplx w = {W[c * 2 * N + 2 * j] * ev, W[c * 2 * N + 2 * j + 1] * ev};
          cfma(Gm[c], w, Mi[j * N + k]);
          if (LAP) for (int l = 0; l < N; ++l) cfma(T[c * N + l], w, Mi[j * N + l]);
        }
      }
      for (int c = 0; c < 4; ++c) {
        dc[(int64_t)(DC::GMAT + (k * 4 + c) * 2) * stride] = Gm[c].re;
        dc[(int64_t)(DC::GMAT + (k * 4 + c) * 2 + 1) * stride] = Gm[c].im;
        if (LAP)
          for (int l = 0; l < N; ++l) {
            dc[(int64_t)(DC::TT + ((k * 4 + c) * N + l) * 2) * stride] = T[c * N + l].re;
            dc[(int64_t)(DC::TT + ((k * 4 + c) * N + l) * 2 + 1) * stride] = T[c * N + l].im;
          }
      }
    }
    keep_alive(&pr); keep_alive(Mi);      // see grad_reverse: forbid stack-slot sharing of live arrays
  }

  // tangent (first / second derivative along ONE coordinate) of a 4-vector
  struct Tan4 { double d[4]; double s[4]; };

  // out = tanh-layer(in): z = in . W (tangents only), t = cached tanh output, residual (in + t) / sqrt2
  template <bool LAP>
  static AQ_HD void chain_layer(const double* __restrict__ W, const Tan4& in, const double t[4], Tan4& out) {
    for (int m = 0; m < 4; ++m) {
      double zd = 0.0, zs = 0.0;
      for (int q = 0; q < 4; ++q) { zd += in.d[q] * W[q * 4 + m]; if (LAP) zs += in.s[q] * W[q * 4 + m]; }
      const double g = 1.0 - t[m] * t[m];
      out.d[m] = (in.d[m] + g * zd) * kInvSqrt2;
      out.s[m] = LAP ? (in.s[m] + g * (zs - 2.0 * t[m] * zd * zd)) * kInvSqrt2 : 0.0;
    }
  }

  // ---- tangent pass for electron e, direction dir: g = d log|psi| / dx_{e,dir};
  //      lap = d^2 log|psi| / dx_{e,dir}^2 (LAP only)
  template <bool LAP>
  static AQ_HD void tangent(const AiqmcSystem& sys, const double* __restrict__ P, const double* __restrict__ dc,
                            int64_t stride, int e, int dir, double& g_out, double& lap_out) {
    using J = Jet<LAP, 1>;
    using Op = ScalarOps<J>;
    constexpr LayoutC<NE, NA> L{};
    auto ld = [&](int slot) -> double { return dc[(int64_t)slot * stride]; };
    const int se = e < sys.n_up ? 0 : 1;
    const double inv_n[2] = {1.0 / sys.n_up, 1.0 / sys.n_dn};
    const double inv_se = inv_n[se];

    // ---- electron-local jets (the only transcendental work of this pass)
    double xe[3];
    J xj[3];
    AQ_UNROLL_N
    for (int c = 0; c < 3; ++c) {
      xe[c] = ld(DC::X + 3 * e + c);
      xj[c] = Op::cst(xe[c]);
      xj[c].d[0] = (c == dir) ? 1.0 : 0.0;
    }
    J h0e[4 * A], ye[6], enve, jaee;
    PS::template electron_local<J>(P, e, xj, h0e, ye, enve, jaee);
    double dJ = jaee.d[0], sJ = jaee.s[0];

    // ---- level-0 tangents of the pair chains through e: row (e,k): d = x_k - x_e, column (k,e): d = x_e - x_k
    Tan4 cr[N], cc[N];
    AQ_UNROLL_N
    for (int k = 0; k < N; ++k) {
      if (k == e) continue;
      double dd = 0.0;
      AQ_UNROLL_N
      for (int c = 0; c < 3; ++c) { const double d = ld(DC::X + 3 * k + c) - xe[c]; dd = (c == dir) ? d : dd; }
      const double r = ld(DC::HP + ((0 * N + e) * N + k) * 4);
      const double ri = s_inv(r);
      const double rt = -dd * ri;                                 // dr/dx_{e,dir}
      const double rs = (1.0 - rt * rt) * ri;                     // d2r/dx_{e,dir}^2
      cr[k].d[0] = rt; cc[k].d[0] = rt;
      cr[k].s[0] = LAP ? rs : 0.0; cc[k].s[0] = LAP ? rs : 0.0;
      AQ_UNROLL_N
      for (int c = 0; c < 3; ++c) {
        cr[k].d[1 + c] = (c == dir) ? -1.0 : 0.0;
        cc[k].d[1 + c] = (c == dir) ? 1.0 : 0.0;
        cr[k].s[1 + c] = 0.0; cc[k].s[1 + c] = 0.0;
      }
      const int lo = e < k ? e : k, hi = e < k ? k : e;           // e-e Pade term (Jastrow.py:23-41)
      const double cu = P[L.jas_cusp + lo * N + hi], al = P[L.jas_alpha + lo * N + hi];
      const double q = s_inv(1.0 + al * r);
      const double u1 = cu * q * q, u2 = -2.0 * al * u1 * q;
      dJ += u1 * rt;
      if (LAP) sJ += u2 * rt * rt + u1 * rs;
    }

    // ---- one-electron stream tangents for every electron k; the pair chains advance with the layers
    double hd[N][4], hs[N][4];
    AQ_UNROLL_N
    for (int l = 0; l < 3; ++l) {
      // block sums of the column chains: tangent of G_l[s][e]
      double sud[2][4], sus[2][4];
      AQ_UNROLL_N
      for (int s = 0; s < 2; ++s) for (int c = 0; c < 4; ++c) { sud[s][c] = 0.0; sus[s][c] = 0.0; }
      AQ_UNROLL_N
      for (int k = 0; k < N; ++k) {
        if (k == e) continue;
        AQ_UNROLL_N
        for (int c = 0; c < 4; ++c) {      // no runtime-indexed register arrays
          if (k < sys.n_up) { sud[0][c] += cc[k].d[c]; if (LAP) sus[0][c] += cc[k].s[c]; }
          else { sud[1][c] += cc[k].d[c]; if (LAP) sus[1][c] += cc[k].s[c]; }
        }
      }
      // tangents of the block means fed to every row
      double gmd[2][4 * A > 4 ? 4 * A : 4], gms[2][4 * A > 4 ? 4 * A : 4];
      constexpr int DIN0 = 4 * A;
      if (l == 0) {
        AQ_UNROLL_N
        for (int q = 0; q < DIN0; ++q) {
          gmd[0][q] = se == 0 ? h0e[q].d[0] * inv_se : 0.0; gmd[1][q] = se == 1 ? h0e[q].d[0] * inv_se : 0.0;
          gms[0][q] = (LAP && se == 0) ? h0e[q].s[0] * inv_se : 0.0; gms[1][q] = (LAP && se == 1) ? h0e[q].s[0] * inv_se : 0.0;
        }
      } else {
        AQ_UNROLL_N
        for (int c = 0; c < 4; ++c) {
          double u = 0.0, d = 0.0, us = 0.0, ds = 0.0;
          AQ_UNROLL_N
          for (int k = 0; k < N; ++k) {
            if (k < sys.n_up) { u += hd[k][c]; if (LAP) us += hs[k][c]; } else { d += hd[k][c]; if (LAP) ds += hs[k][c]; }
          }
          gmd[0][c] = u * inv_n[0]; gmd[1][c] = d * inv_n[1];
          gms[0][c] = us * inv_n[0]; gms[1][c] = ds * inv_n[1];
        }
      }
      AQ_UNROLL_N
      for (int k = 0; k < N; ++k) {
        if (l == 0) row_layer<LAP, DIN0>(P, 0, k, e, se, inv_n, dc, stride, h0e, hd[k], hs[k], gmd, gms, cr[k], sud, sus);
        else row_layer<LAP, 4>(P, l, k, e, se, inv_n, dc, stride, h0e, hd[k], hs[k], gmd, gms, cr[k], sud, sus);
      }
      if (l < 2) {   // advance the tangents of the 2(N-1) pair chains through double-layer l
        const double* W = P + L.dbl_w[l];
        AQ_UNROLL_N
        for (int k = 0; k < N; ++k) {
          if (k == e) continue;
          double tr[4], tc[4];
          AQ_UNROLL_N
          for (int m = 0; m < 4; ++m) {
            tr[m] = kSqrt2 * ld(DC::HP + (((l + 1) * N + e) * N + k) * 4 + m) - ld(DC::HP + ((l * N + e) * N + k) * 4 + m);
            tc[m] = kSqrt2 * ld(DC::HP + (((l + 1) * N + k) * N + e) * 4 + m) - ld(DC::HP + ((l * N + k) * N + e) * 4 + m);
          }
          Tan4 nr, nc;
          chain_layer<LAP>(W, cr[k], tr, nr);
          chain_layer<LAP>(W, cc[k], tc, nc);
          cr[k] = nr; cc[k] = nc;
        }
      }
    }

    // ---- row e of E = env * Yo as jets; P[e,:] reads h of electron sigma[e]  (quirk Q4)
    const int srow = e < sys.n_up_rows ? 0 : 1;
    const double* W = P + L.orb_w[srow];
    const double* Bv = P + L.orb_b[srow];
    const int eh = sys.sigma[e];
    double h3[4], h3d[4];
    AQ_UNROLL_N
    for (int c = 0; c < 4; ++c) {
      h3[c] = ld(DC::H + (2 * N + eh) * 4 + c);
      double v = hd[0][c];
      AQ_UNROLL_N
      for (int q = 1; q < N; ++q) v = (eh == q) ? hd[q][c] : v;
      h3d[c] = v;
    }
    cplx S1[LAP ? N : 1];
    cplx S1e = {0.0, 0.0}, S2 = {0.0, 0.0};
    if (LAP) for (int l = 0; l < N; ++l) S1[l] = {0.0, 0.0};
    AQ_UNROLL_N
    for (int j = 0; j < N; ++j) {
      J yo = Op::cst(0.0);
      AQ_UNROLL_N
      for (int m = 0; m < 6; ++m) yo = yo + ye[m] * P[L.y_w + m * N + j];
      const J E = enve * yo;
      cplx p = {Bv[2 * j], Bv[2 * j + 1]}, dp = {0.0, 0.0};
      AQ_UNROLL_N
      for (int c = 0; c < 4; ++c) {
        const double wr = W[c * 2 * N + 2 * j], wi = W[c * 2 * N + 2 * j + 1];
        p.re += h3[c] * wr; p.im += h3[c] * wi;
        dp.re += h3d[c] * wr; dp.im += h3d[c] * wi;
      }
      const cplx pdE = cscale(p, E.d[0]);
      const cplx mje = {ld(DC::MI + (j * N + e) * 2), ld(DC::MI + (j * N + e) * 2 + 1)};
      cfma(S1e, pdE, mje);
      if (LAP) {
        AQ_UNROLL_N
        for (int l = 0; l < N; ++l) {
          const cplx mjl = {ld(DC::MI + (j * N + l) * 2), ld(DC::MI + (j * N + l) * 2 + 1)};
          cfma(S1[l], pdE, mjl);
        }
        const cplx t2 = cadd(cscale(dp, 2.0 * E.d[0]), cscale(p, E.s[0]));
        cfma(S2, t2, mje);
      }
    }
    // gradient: Re tr(X) = Re [ sum_k sum_c dh[sigma_k,c] Gm[k,c] + S1[e] ]
    double gsum = S1e.re + dJ, l2 = S2.re + sJ;
    AQ_UNROLL_N
    for (int k = 0; k < N; ++k) {
      const int ek = sys.sigma[k];
      AQ_UNROLL_N
      for (int c = 0; c < 4; ++c) {
        double vd = hd[0][c], vs = LAP ? hs[0][c] : 0.0;
        AQ_UNROLL_N
        for (int q = 1; q < N; ++q) { vd = (ek == q) ? hd[q][c] : vd; if (LAP) vs = (ek == q) ? hs[q][c] : vs; }
        const double gm = ld(DC::GMAT + (k * 4 + c) * 2);
        gsum += vd * gm;
        if (LAP) l2 += vs * gm;
      }
    }
    g_out = gsum;
    if (LAP) {
      // X[k,l] = sum_c dh[sigma_k,c] T[k,c,l] + delta_ke S1[l];  subtract Re tr(X X)
      cplx X[N * N];
      AQ_UNROLL_N
      for (int k = 0; k < N; ++k) {
        const int ek = sys.sigma[k];
        double dh[4];
        AQ_UNROLL_N
        for (int c = 0; c < 4; ++c) {
          double v = hd[0][c];
          AQ_UNROLL_N
          for (int q = 1; q < N; ++q) v = (ek == q) ? hd[q][c] : v;
          dh[c] = v;
        }
        AQ_UNROLL_N
        for (int l = 0; l < N; ++l) {
          cplx acc = (k == e) ? S1[l] : cplx{0.0, 0.0};
          AQ_UNROLL_N
          for (int c = 0; c < 4; ++c) {
            acc.re += dh[c] * ld(DC::TT + ((k * 4 + c) * N + l) * 2);
            acc.im += dh[c] * ld(DC::TT + ((k * 4 + c) * N + l) * 2 + 1);
          }
          X[k * N + l] = acc;
        }
      }
      double trxx = 0.0;
      AQ_UNROLL_N
      for (int k = 0; k < N; ++k)
        AQ_UNROLL_N
        for (int l = 0; l < N; ++l) trxx += X[k * N + l].re * X[l * N + k].re - X[k * N + l].im * X[l * N + k].im;
      lap_out = l2 - trxx;
    }
    keep_alive(cr); keep_alive(cc); keep_alive(hd); keep_alive(hs); keep_alive(h0e); keep_alive(ye);
  }

  // tangent of one-electron layer l, row k (one_layer of psi_core.cuh with cached tanh outputs).
  // In/out: hd/hs = tangent of h_l[k] (ignored for l == 0), replaced by the tangent of h_{l+1}[k].
  template <bool LAP, int DIN, class J>
  static AQ_HD void row_layer(const double* __restrict__ P, int l, int k, int e, int se, const double inv_n[2],
                              const double* __restrict__ dc, int64_t stride, const J* __restrict__ h0e,
                              double hd[4], double hs[4], const double (*gmd)[4 * NA > 4 ? 4 * NA : 4],
                              const double (*gms)[4 * NA > 4 ? 4 * NA : 4], const Tan4& crk, const double (*sud)[4],
                              const double (*sus)[4]) {
    constexpr LayoutC<NE, NA> L{};
    constexpr int DTOT = 3 * DIN + 8, Q = DTOT / 4;
    auto ld = [&](int slot) -> double { return dc[(int64_t)slot * stride]; };
    const double* cw = P + L.conv_w[l] + k * DTOT;
    const double* sw = P + L.sing_w[l];
    const bool diag = (k == e);
    // tangent of the layer input [h_k (DIN), g_up (DIN), g_dn (DIN), G_up[k]/n_up (4), G_dn[k]/n_dn (4)]
    double xd[DTOT], xs[DTOT];
    AQ_UNROLL_N
    for (int q = 0; q < DIN; ++q) {
      if (l == 0) { xd[q] = diag ? h0e[q].d[0] : 0.0; xs[q] = (LAP && diag) ? h0e[q].s[0] : 0.0; }
      else { xd[q] = hd[q]; xs[q] = LAP ? hs[q] : 0.0; }
      xd[DIN + q] = gmd[0][q]; xd[2 * DIN + q] = gmd[1][q];
      xs[DIN + q] = LAP ? gms[0][q] : 0.0; xs[2 * DIN + q] = LAP ? gms[1][q] : 0.0;
    }
    AQ_UNROLL_N
    for (int c = 0; c < 4; ++c) {
      // k == e: column block sums; otherwise only block s_e moves, by the row chain (e,k)
      xd[3 * DIN + c] = (diag ? sud[0][c] : (se == 0 ? crk.d[c] : 0.0)) * inv_n[0];
      xd[3 * DIN + 4 + c] = (diag ? sud[1][c] : (se == 1 ? crk.d[c] : 0.0)) * inv_n[1];
      xs[3 * DIN + c] = LAP ? (diag ? sus[0][c] : (se == 0 ? crk.s[c] : 0.0)) * inv_n[0] : 0.0;
      xs[3 * DIN + 4 + c] = LAP ? (diag ? sus[1][c] : (se == 1 ? crk.s[c] : 0.0)) * inv_n[1] : 0.0;
    }
    double zd[4] = {0.0, 0.0, 0.0, 0.0}, zs[4] = {0.0, 0.0, 0.0, 0.0};
    AQ_UNROLL_N
    for (int q = 0; q < Q; ++q) {
      double pd = 0.0, ps = 0.0;
      AQ_UNROLL_N
      for (int c = 0; c < 4; ++c) { pd += xd[4 * q + c] * cw[4 * q + c]; if (LAP) ps += xs[4 * q + c] * cw[4 * q + c]; }
      pd *= 0.25; ps *= 0.25;
      const double t = ld(DC::T1 + (l * N + k) * QM + q);
      const double g = 1.0 - t * t;
      const double od = g * pd;
      const double os = LAP ? g * (ps - 2.0 * t * pd * pd) : 0.0;
      AQ_UNROLL_N
      for (int m = 0; m < 4; ++m) { zd[m] += od * sw[q * 4 + m]; if (LAP) zs[m] += os * sw[q * 4 + m]; }
    }
    AQ_UNROLL_N
    for (int m = 0; m < 4; ++m) {
      const double hn = ld(DC::H + (l * N + k) * 4 + m);                 // h_{l+1}[k][m]
      double t, ind, ins;
      if (DIN == 4) {                                                     // residual layer (quirk Q5)
        const double hp = (l == 0) ? ld(DC::H0 + k * 4 * NA + m) : ld(DC::H + ((l - 1) * N + k) * 4 + m);
        t = kSqrt2 * hn - hp;
        ind = xd[m]; ins = xs[m];
      } else {
        t = hn; ind = 0.0; ins = 0.0;
      }
      const double g = 1.0 - t * t;
      const double wd = g * zd[m];
      const double ws = LAP ? g * (zs[m] - 2.0 * t * zd[m] * zd[m]) : 0.0;
      hd[m] = (DIN == 4) ? (ind + wd) * kInvSqrt2 : wd;
      hs[m] = LAP ? ((DIN == 4) ? (ins + ws) * kInvSqrt2 : ws) : 0.0;
    }
  }

  // ---- gradient of log|psi| by ONE reverse (adjoint) sweep, fused with the forward pass in the same thread.
  //      3N forward tangents cost ~23 k FP64 instructions per configuration at N=4 (and grow like N); the adjoint
  //      costs ~7 k whatever N: every cached tanh output is used once, every weight twice.
  //        (1) d log|det| = Re tr(M^-1 dM) gives the adjoints of h_3, of the envelopes and of the Ynlm outputs;
  //        (2) the three one-electron layers are walked backwards (adjoints of the per-row inputs, of the block
  //            means and of the pair block sums G_l);
  //        (3) every ordered pair chain is walked backwards from its three levels to the inter-electron vector;
  //        (4) each electron's local part (features, Ynlm stream, envelope, e-n Jastrow) is contracted with its
  //            adjoints through a 3-direction jet (one transcendental evaluation per electron).
  static AQ_HD void grad_reverse(const AiqmcSystem& sys, const double* __restrict__ P, const double* __restrict__ x,
                                 double& phase, double& logabs, double* __restrict__ grad) {
    constexpr LayoutC<NE, NA> L{};
    typename PS::Primal pr;
    cplx Mi[N * N];
    double hp[3 * N * N * 4];                 // pair chains, all levels: hp[((l*N + i)*N + j)*4 + c]
    double t1[3 * N * QM];                    // first-stage tanh outputs of the one-electron layers
    PS::forward(sys, P, x, pr, Mi, hp, 1, t1, 1);
    double ld;
    PS::gj_inverse(Mi, ld, phase);
    logabs = ld + pr.jastrow;
    const double inv_n[2] = {1.0 / sys.n_up, 1.0 / sys.n_dn};

    // (1) determinant
    double h_bar[N][4], y_bar[N][6], env_bar[N];
    for (int k = 0; k < N; ++k) {
      env_bar[k] = 0.0;
      for (int c = 0; c < 4; ++c) h_bar[k][c] = 0.0;
      for (int m = 0; m < 6; ++m) y_bar[k][m] = 0.0;
    }
    for (int k = 0; k < N; ++k) {
      const int s = k < sys.n_up_rows ? 0 : 1;
      const double* W = P + L.orb_w[s];
      const double* Bv = P + L.orb_b[s];
      const int e = sys.sigma[k];
      for (int j = 0; j < N; ++j) {
        double yo = 0.0;
        for (int m = 0; m < 6; ++m) yo += pr.y[k][m] * P[L.y_w + m * N + j];
        double pre = Bv[2 * j], pim = Bv[2 * j + 1];
        for (int c = 0; c < 4; ++c) { pre += pr.h[3][e][c] * W[c * 2 * N + 2 * j]; pim += pr.h[3][e][c] * W[c * 2 * N + 2 * j + 1]; }
        const cplx mi = Mi[j * N + k];
        const double ev = pr.env[k] * yo;
        const double pre_bar = mi.re * ev, pim_bar = -mi.im * ev;
        const double ev_bar = mi.re * pre - mi.im * pim;
        for (int c = 0; c < 4; ++c) h_bar[e][c] += pre_bar * W[c * 2 * N + 2 * j] + pim_bar * W[c * 2 * N + 2 * j + 1];
        env_bar[k] += ev_bar * yo;
        const double yo_bar = ev_bar * pr.env[k];
        for (int m = 0; m < 6; ++m) y_bar[k][m] += yo_bar * P[L.y_w + m * N + j];
      }
    }

    // (2) one-electron layers, backwards
    double G_bar[3][2][N][4];
    double h0_bar[N][4 * A];
    for (int l = 2; l >= 0; --l) {
      if (l == 0) layer_reverse<4 * A>(sys, P, 0, pr, t1, inv_n, h_bar, G_bar[0], h0_bar);
      else layer_reverse<4>(sys, P, l, pr, t1, inv_n, h_bar, G_bar[l], h_bar);
    }

    // (3) pair chains, backwards
    for (int q = 0; q < 3 * N; ++q) grad[q] = 0.0;
    for (int i = 0; i < N; ++i) {
      const int s = i < sys.n_up ? 0 : 1;
      for (int j = 0; j < N; ++j) {
        if (i == j) continue;
        const double* a0 = hp + ((0 * N + i) * N + j) * 4;
        const double* a1 = hp + ((1 * N + i) * N + j) * 4;
        const double* a2 = hp + ((2 * N + i) * N + j) * 4;
        double b1[4], b0[4];
        chain_reverse(P + L.dbl_w[1], a1, a2, G_bar[2][s][j], G_bar[1][s][j], b1);
        chain_reverse(P + L.dbl_w[0], a0, a1, b1, G_bar[0][s][j], b0);
        const double r = a0[0];
        double r_bar = b0[0];
        if (i < j) {                                            // e-e Pade term (Jastrow.py:23-41)
          const double q = s_inv(1.0 + P[L.jas_alpha + i * N + j] * r);
          r_bar += P[L.jas_cusp + i * N + j] * q * q;
        }
        const double rs = r_bar * s_inv(r);
        for (int c = 0; c < 3; ++c) {
          const double db = b0[1 + c] + rs * a0[1 + c];           // d = x_j - x_i
          grad[3 * j + c] += db;
          grad[3 * i + c] -= db;
        }
      }
    }

    // (4) electron-local parts through a 3-direction jet
    using J = Jet<false, 3>;
    using Op = ScalarOps<J>;
    for (int e = 0; e < N; ++e) {
      J xj[3];
      for (int c = 0; c < 3; ++c) { xj[c] = Op::cst(x[3 * e + c]); xj[c].d[c] = 1.0; }
      J h0e[4 * A], ye[6], enve, jaee;
      PS::template electron_local<J>(P, e, xj, h0e, ye, enve, jaee);
      for (int c = 0; c < 3; ++c) {
        double g = jaee.d[c] + env_bar[e] * enve.d[c];
        for (int q = 0; q < 4 * A; ++q) g += h0_bar[e][q] * h0e[q].d[c];
        for (int m = 0; m < 6; ++m) g += y_bar[e][m] * ye[m].d[c];
        grad[3 * e + c] += g;
      }
    }
    // nvcc 12.9's stack colouring has overlapped live local arrays in this code base three times (DESIGN.md,
    // toolchain notes); keeping every tape / adjoint array observably alive to the end of the function forbids it.
    keep_alive(&pr); keep_alive(Mi); keep_alive(hp); keep_alive(t1); keep_alive(h_bar); keep_alive(y_bar);
    keep_alive(env_bar); keep_alive(G_bar); keep_alive(h0_bar);
  }

  // ---- the same reverse sweep, reading the primal quantities from the SoA derivative cache that primal<false>
  //      wrote (N > 16: the per-thread tape of grad_reverse would be 12 N^2 doubles; the 3N forward tangents of the
  //      two-pass path read the whole cache record 3N times -- 14 MB per C6H6 configuration, L2-bandwidth bound).
  //      One thread per configuration; its own state is the adjoints only (~20 kB at N = 30).
  static AQ_HD void grad_reverse_cached(const AiqmcSystem& sys, const double* __restrict__ P, const double* __restrict__ dc,
                                        int64_t stride, double* __restrict__ grad) {
    constexpr LayoutC<NE, NA> L{};
    auto ld = [&](int slot) -> double { return dc[(int64_t)slot * stride]; };
    const double inv_n[2] = {1.0 / sys.n_up, 1.0 / sys.n_dn};
    double x[3 * N];
    for (int q = 0; q < 3 * N; ++q) x[q] = ld(DC::X + q);

    // (1) determinant
    double h_bar[N][4], y_bar[N][6], env_bar[N];
    for (int k = 0; k < N; ++k) {
      env_bar[k] = 0.0;
      for (int c = 0; c < 4; ++c) h_bar[k][c] = 0.0;
      for (int m = 0; m < 6; ++m) y_bar[k][m] = 0.0;
    }
    for (int k = 0; k < N; ++k) {
      const int s = k < sys.n_up_rows ? 0 : 1;
      const double* W = P + L.orb_w[s];
      const double* Bv = P + L.orb_b[s];
      const int e = sys.sigma[k];
      double h3[4], yk[6];
      for (int c = 0; c < 4; ++c) h3[c] = ld(DC::H + (2 * N + e) * 4 + c);
      for (int m = 0; m < 6; ++m) yk[m] = ld(DC::YV + k * 6 + m);
      const double envk = ld(DC::ENVV + k);
      for (int j = 0; j < N; ++j) {
        double yo = 0.0;
        for (int m = 0; m < 6; ++m) yo += yk[m] * P[L.y_w + m * N + j];
        double pre = Bv[2 * j], pim = Bv[2 * j + 1];
        for (int c = 0; c < 4; ++c) { pre += h3[c] * W[c * 2 * N + 2 * j]; pim += h3[c] * W[c * 2 * N + 2 * j + 1]; }
        const double mre = ld(DC::MI + 2 * (j * N + k)), mim = ld(DC::MI + 2 * (j * N + k) + 1);
        const double ev = envk * yo;
        const double pre_bar = mre * ev, pim_bar = -mim * ev;
        const double ev_bar = mre * pre - mim * pim;
        for (int c = 0; c < 4; ++c) h_bar[e][c] += pre_bar * W[c * 2 * N + 2 * j] + pim_bar * W[c * 2 * N + 2 * j + 1];
        env_bar[k] += ev_bar * yo;
        const double yo_bar = ev_bar * envk;
        for (int m = 0; m < 6; ++m) y_bar[k][m] += yo_bar * P[L.y_w + m * N + j];
      }
    }

    // (2) one-electron layers, backwards (adjoints of the per-row inputs accumulate in place)
    double G_bar[3][2][N][4];
    double h0_bar[N][4 * A];
    for (int l = 2; l >= 0; --l) {
      if (l == 0) layer_reverse_cached<4 * A>(sys, P, 0, dc, stride, inv_n, h_bar, G_bar[0], h0_bar);
      else layer_reverse_cached<4>(sys, P, l, dc, stride, inv_n, h_bar, G_bar[l], h_bar);
    }

    // (3) pair chains, backwards
    for (int q = 0; q < 3 * N; ++q) grad[q] = 0.0;
    for (int i = 0; i < N; ++i) {
      const int s = i < sys.n_up ? 0 : 1;
      for (int j = 0; j < N; ++j) {
        if (i == j) continue;
        double a0[4], a1[4], a2[4];
        for (int c = 0; c < 4; ++c) {
          a0[c] = ld(DC::HP + ((0 * N + i) * N + j) * 4 + c);
          a1[c] = ld(DC::HP + ((1 * N + i) * N + j) * 4 + c);
          a2[c] = ld(DC::HP + ((2 * N + i) * N + j) * 4 + c);
        }
        double b1[4], b0[4];
        chain_reverse(P + L.dbl_w[1], a1, a2, G_bar[2][s][j], G_bar[1][s][j], b1);
        chain_reverse(P + L.dbl_w[0], a0, a1, b1, G_bar[0][s][j], b0);
        const double r = a0[0];
        double r_bar = b0[0];
        if (i < j) {                                            // e-e Pade term (Jastrow.py:23-41)
          const double q = s_inv(1.0 + P[L.jas_alpha + i * N + j] * r);
          r_bar += P[L.jas_cusp + i * N + j] * q * q;
        }
        const double rs = r_bar * s_inv(r);
        for (int c = 0; c < 3; ++c) {
          const double db = b0[1 + c] + rs * a0[1 + c];           // d = x_j - x_i
          grad[3 * j + c] += db;
          grad[3 * i + c] -= db;
        }
      }
    }

    // (4) electron-local parts through a 3-direction jet
    using J = Jet<false, 3>;
    using Op = ScalarOps<J>;
    for (int e = 0; e < N; ++e) {
      J xj[3];
      for (int c = 0; c < 3; ++c) { xj[c] = Op::cst(x[3 * e + c]); xj[c].d[c] = 1.0; }
      J h0e[4 * A], ye[6], enve, jaee;
      PS::template electron_local<J>(P, e, xj, h0e, ye, enve, jaee);
      for (int c = 0; c < 3; ++c) {
        double g = jaee.d[c] + env_bar[e] * enve.d[c];
        for (int q = 0; q < 4 * A; ++q) g += h0_bar[e][q] * h0e[q].d[c];
        for (int m = 0; m < 6; ++m) g += y_bar[e][m] * ye[m].d[c];
        grad[3 * e + c] += g;
      }
    }
    keep_alive(x); keep_alive(h_bar); keep_alive(y_bar); keep_alive(env_bar); keep_alive(G_bar); keep_alive(h0_bar);
  }

  // layer_reverse on the cache; `hin_bar` doubles as the per-row accumulator (it may alias h_bar for l >= 1: row k's
  // h_bar is consumed before row k's hin_bar is written)
  template <int DIN>
  static AQ_HD void layer_reverse_cached(const AiqmcSystem& sys, const double* __restrict__ P, int l,
                                         const double* __restrict__ dc, int64_t stride, const double inv_n[2],
                                         const double (*h_bar)[4], double (*G_bar_l)[N][4], double (*hin_bar)[DIN]) {
    constexpr LayoutC<NE, NA> L{};
    constexpr int DTOT = 3 * DIN + 8, Q = DTOT / 4;
    auto ld = [&](int slot) -> double { return dc[(int64_t)slot * stride]; };
    const double* sw = P + L.sing_w[l];
    double gup_bar[DIN], gdn_bar[DIN];
    for (int q = 0; q < DIN; ++q) { gup_bar[q] = 0.0; gdn_bar[q] = 0.0; }
    for (int k = 0; k < N; ++k) {
      const double* cw = P + L.conv_w[l] + k * DTOT;
      double zb[4], ob4[4];
      for (int m = 0; m < 4; ++m) {
        const double hn = ld(DC::H + (l * N + k) * 4 + m);           // h_{l+1}[k][m]
        double t, ob;
        if (DIN == 4) {                                               // residual layer (quirk Q5)
          const double hprev = (l == 0) ? ld(DC::H0 + k * 4 * A + m) : ld(DC::H + ((l - 1) * N + k) * 4 + m);
          t = kSqrt2 * hn - hprev;
          ob = h_bar[k][m] * kInvSqrt2;
        } else {
          t = hn;
          ob = h_bar[k][m];
        }
        zb[m] = ob * (1.0 - t * t);
        ob4[m] = ob;
      }
      for (int q = 0; q < DIN; ++q) hin_bar[k][q] = (DIN == 4) ? ob4[q < 4 ? q : 0] : 0.0;
      for (int q = 0; q < Q; ++q) {
        double ob = 0.0;
        for (int m = 0; m < 4; ++m) ob += zb[m] * sw[q * 4 + m];
        const double t = ld(DC::T1 + (l * N + k) * QM + q);
        const double pb = ob * (1.0 - t * t) * 0.25;
        for (int c = 0; c < 4; ++c) {
          const int idx = 4 * q + c;
          const double xb = pb * cw[idx];
          if (idx < DIN) hin_bar[k][idx] += xb;
          else if (idx < 2 * DIN) gup_bar[idx - DIN] += xb;
          else if (idx < 3 * DIN) gdn_bar[idx - 2 * DIN] += xb;
          else if (idx < 3 * DIN + 4) G_bar_l[0][k][idx - 3 * DIN] = xb * inv_n[0];
          else G_bar_l[1][k][idx - 3 * DIN - 4] = xb * inv_n[1];
        }
      }
    }
    for (int k = 0; k < N; ++k) {
      const bool up = k < sys.n_up;
      for (int q = 0; q < DIN; ++q) hin_bar[k][q] += (up ? gup_bar[q] * inv_n[0] : gdn_bar[q] * inv_n[1]);
    }
  }

  static AQ_HD void keep_alive(const void* p) {
#ifdef __CUDA_ARCH__
    asm volatile("" ::"l"(p) : "memory");
#else
    (void)p;
#endif
  }

  // adjoint of out = (in + tanh(in . W + b)) / sqrt2 ; t recovered from the cached levels: t = sqrt2 out - in.
  // in_bar = extra + [out_bar + W (g o out_bar)] / sqrt2
  static AQ_HD void chain_reverse(const double* __restrict__ W, const double* __restrict__ in, const double* __restrict__ out,
                                  const double* __restrict__ out_bar, const double* __restrict__ extra,
                                  double* __restrict__ in_bar) {
    double zb[4];
    for (int m = 0; m < 4; ++m) {
      const double t = kSqrt2 * out[m] - in[m];
      zb[m] = out_bar[m] * kInvSqrt2 * (1.0 - t * t);
    }
    for (int q = 0; q < 4; ++q) {
      double v = out_bar[q] * kInvSqrt2 + extra[q];
      for (int m = 0; m < 4; ++m) v += W[q * 4 + m] * zb[m];
      in_bar[q] = v;
    }
  }

  // adjoint of one-electron layer l for all rows.  In: h_bar = adjoint of h_{l+1}.  Out: G_bar_l[s][k][c] = adjoint
  // of the pair block sums G_l[s][k][c]; hin_bar = adjoint of the layer's per-row input (h_l, or h0 for l == 0).
  // h_bar and hin_bar may alias (they do for l >= 1).
  template <int DIN>
  static AQ_HD void layer_reverse(const AiqmcSystem& sys, const double* __restrict__ P, int l,
                                  const typename PS::Primal& pr, const double* __restrict__ t1, const double inv_n[2],
                                  const double (*h_bar)[4], double (*G_bar_l)[N][4], double (*hin_bar)[DIN]) {
    constexpr LayoutC<NE, NA> L{};
    constexpr int DTOT = 3 * DIN + 8, Q = DTOT / 4;
    const double* sw = P + L.sing_w[l];
    double gup_bar[DIN], gdn_bar[DIN];
    for (int q = 0; q < DIN; ++q) { gup_bar[q] = 0.0; gdn_bar[q] = 0.0; }
    double own[N][DIN];                                   // adjoint reaching row k's own input directly
    for (int k = 0; k < N; ++k) {
      const double* cw = P + L.conv_w[l] + k * DTOT;
      double zb[4];
      for (int m = 0; m < 4; ++m) {
        const double hn = pr.h[l + 1][k][m];
        double t, ob;
        if (DIN == 4) {                                   // residual layer (quirk Q5)
          const double hprev = (l == 0) ? pr.h0[k][m] : pr.h[l][k][m];
          t = kSqrt2 * hn - hprev;
          ob = h_bar[k][m] * kInvSqrt2;
        } else {
          t = hn;
          ob = h_bar[k][m];
        }
        zb[m] = ob * (1.0 - t * t);
        if (DIN == 4) own[k][m] = ob;
      }
      if (DIN != 4) for (int q = 0; q < DIN; ++q) own[k][q] = 0.0;
      for (int q = 0; q < Q; ++q) {
        double ob = 0.0;
        for (int m = 0; m < 4; ++m) ob += zb[m] * sw[q * 4 + m];
        const double t = t1[(l * N + k) * QM + q];
        const double pb = ob * (1.0 - t * t) * 0.25;
        for (int c = 0; c < 4; ++c) {
          const int idx = 4 * q + c;
          const double xb = pb * cw[idx];
          if (idx < DIN) own[k][idx] += xb;
          else if (idx < 2 * DIN) gup_bar[idx - DIN] += xb;
          else if (idx < 3 * DIN) gdn_bar[idx - 2 * DIN] += xb;
          else if (idx < 3 * DIN + 4) G_bar_l[0][k][idx - 3 * DIN] = xb * inv_n[0];
          else G_bar_l[1][k][idx - 3 * DIN - 4] = xb * inv_n[1];
        }
      }
    }
    for (int k = 0; k < N; ++k) {
      const bool up = k < sys.n_up;
      for (int q = 0; q < DIN; ++q) hin_bar[k][q] = own[k][q] + (up ? gup_bar[q] * inv_n[0] : gdn_bar[q] * inv_n[1]);
    }
  }

  // ---- whole gradient (+ Laplacian) of one configuration through the two passes (host checker, small jobs)
  template <bool LAP>
  static AQ_HD void eval(const AiqmcSystem& sys, const double* __restrict__ P, const double* __restrict__ x,
                         double* __restrict__ scratch /* DC::size(LAP) doubles */, double& phase, double& logabs,
                         double* __restrict__ grad, double& lap) {
    primal<LAP>(sys, P, x, scratch, 1, nullptr, phase, logabs);
    double acc = 0.0;
    for (int e = 0; e < N; ++e)
      for (int dir = 0; dir < 3; ++dir) {
        double g, l2 = 0.0;
        tangent<LAP>(sys, P, scratch, 1, e, dir, g, l2);
        grad[3 * e + dir] = g;
        acc += l2;
      }
    lap = acc;
  }
};

}  // namespace aiqmc
