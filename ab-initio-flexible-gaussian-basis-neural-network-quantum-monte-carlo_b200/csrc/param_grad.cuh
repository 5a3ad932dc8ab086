// param_grad.cuh -- d(alpha log|psi| + beta phase)/d(params): the parameter side of the loss gradient
// (Loss/pploss.py:186-223: `jax.jvp(batch_network, ...)` contracted with the clipped local energies).
//
// The reference obtains it as the transpose of a forward-mode JVP, one network evaluation per parameter direction
// in effect; here ONE reverse (adjoint) sweep per walker yields the whole gradient in the packed layout:
//   (1) determinant: Re[(alpha - i beta) tr(M^-1 dM)] seeds the adjoints of h_3, envelopes, Ynlm outputs and gives
//       the orbital / y_w weight gradients (column by column, so every shared weight is summed over the rows in
//       registers before it leaves the thread);
//   (2) the one-electron layers backwards (conv / single weights and biases);
//   (3) every ordered pair chain backwards, INCLUDING the diagonal i == j chains (position independent, but they
//       depend on the double-layer weights and feed the block sums), plus the e-e Pade parameters;
//   (4) each electron's local part: Ynlm stream weights, envelope and e-n Jastrow parameters.
// Contributions are handed to a `Sink` (add(offset, value)): on the device a warp-wide fixed-order reduction into a
// shared-memory accumulator (one walker per lane, deterministic), on the host checker a plain array.
// Entries of the packed layout that are not trainable (jas_cusp, jas_c34/c14, atoms, charges) stay zero; y_w and
// env_sx are gradients w.r.t. the PACKED quantities (row-normalised y weights, sigma * xi) -- the host mirror applies
// the chain rule back to the reference's pytree (system.py: unpack_param_grad).
#pragma once
#include "deriv_split.cuh"

namespace aiqmc {

template <int NE, int NA>
struct ParamGrad {
  using PS = Psi<NE, NA>;
  static constexpr int N = NE, A = NA, QM = PS::QM, K0 = 4 * NA + 2;
  static constexpr double kSqrt2 = 1.41421356237309504880;

  static AQ_HD void keep_alive(const void* p) {
#ifdef __CUDA_ARCH__
    asm volatile("" ::"l"(p) : "memory");
#else
    (void)p;
#endif
  }


  // Where the primal quantities of the reverse sweep come from: the thread's own forward pass (N <= 16), or the SoA
  // derivative cache written by DerivSplit::primal<false> (N > 16, where the per-thread tape would be 12 N^2 doubles).
  struct LocalView {
    const typename PS::Primal& pr;
    const cplx* Mi;
    const double* hp;
    const double* t1;
    AQ_HD double y(int k, int m) const { return pr.y[k][m]; }
    AQ_HD double env(int k) const { return pr.env[k]; }
    AQ_HD double hin(int l, int k, int q) const { return l == 0 ? pr.h0[k][q] : pr.h[l][k][q]; }
    AQ_HD double hout(int l, int k, int m) const { return pr.h[l + 1][k][m]; }
    AQ_HD double gmean(int l, int s, int q) const { return l == 0 ? pr.g0[s][q] : pr.g[l][s][q]; }
    AQ_HD double G(int l, int s, int k, int c) const { return pr.G[l][s][k][c]; }
    AQ_HD double t1v(int l, int k, int q) const { return t1[(l * N + k) * QM + q]; }
    AQ_HD double hpv(int l, int i, int j, int c) const { return hp[((l * N + i) * N + j) * 4 + c]; }
    AQ_HD cplx mi(int j, int k) const { return Mi[j * N + k]; }
  };
  struct CacheView {
    using DC = DerivCache<NE, NA>;
    const double* dc;
    int64_t stride;
    AQ_HD double ld(int slot) const { return dc[(int64_t)slot * stride]; }
    AQ_HD double y(int k, int m) const { return ld(DC::YV + k * 6 + m); }
    AQ_HD double env(int k) const { return ld(DC::ENVV + k); }
    AQ_HD double hin(int l, int k, int q) const { return l == 0 ? ld(DC::H0 + k * 4 * A + q) : ld(DC::H + ((l - 1) * N + k) * 4 + q); }
    AQ_HD double hout(int l, int k, int m) const { return ld(DC::H + (l * N + k) * 4 + m); }
    AQ_HD double gmean(int l, int s, int q) const { return l == 0 ? ld(DC::G0M + s * 4 * A + q) : ld(DC::GM + ((l - 1) * 2 + s) * 4 + q); }
    AQ_HD double G(int l, int s, int k, int c) const { return ld(DC::GS + ((l * 2 + s) * N + k) * 4 + c); }
    AQ_HD double t1v(int l, int k, int q) const { return ld(DC::T1 + (l * N + k) * QM + q); }
    AQ_HD double hpv(int l, int i, int j, int c) const { return ld(DC::HP + ((l * N + i) * N + j) * 4 + c); }
    AQ_HD cplx mi(int j, int k) const { return {ld(DC::MI + 2 * (j * N + k)), ld(DC::MI + 2 * (j * N + k) + 1)}; }
  };

  // adjoint of the one-electron layer l; mirrors DerivSplit::layer_reverse and adds the weight gradients
  template <int DIN, class View, class Sink>
  static AQ_HD void layer_reverse(const AiqmcSystem& sys, const double* __restrict__ P, int l, const View& v,
                                  const double inv_n[2], const double (*h_bar)[4], double (*G_bar_l)[N][4],
                                  double (*hin_bar)[DIN], Sink& sink) {
    constexpr LayoutC<NE, NA> L{};
    constexpr int DTOT = 3 * DIN + 8, Q = DTOT / 4;
    const double* sw = P + L.sing_w[l];
    double gup_bar[DIN], gdn_bar[DIN];
    for (int q = 0; q < DIN; ++q) { gup_bar[q] = 0.0; gdn_bar[q] = 0.0; }
    double own[N][DIN];
    double sw_bar[Q][4], sb_bar[4];
    for (int m = 0; m < 4; ++m) sb_bar[m] = 0.0;
    for (int q = 0; q < Q; ++q)
      for (int m = 0; m < 4; ++m) sw_bar[q][m] = 0.0;
    for (int k = 0; k < N; ++k) {
      const double* cw = P + L.conv_w[l] + k * DTOT;
      double zb[4];
      for (int m = 0; m < 4; ++m) {
        const double hn = v.hout(l, k, m);
        double t, ob;
        if (DIN == 4) {                                   // residual layer (quirk Q5)
          t = kSqrt2 * hn - v.hin(l, k, m);
          ob = h_bar[k][m] * kInvSqrt2;
        } else {
          t = hn;
          ob = h_bar[k][m];
        }
        zb[m] = ob * (1.0 - t * t);
        sb_bar[m] += zb[m];
        if (DIN == 4) own[k][m] = ob;
      }
      if (DIN != 4) for (int q = 0; q < DIN; ++q) own[k][q] = 0.0;
      for (int q = 0; q < Q; ++q) {
        double ob = 0.0;
        for (int m = 0; m < 4; ++m) ob += zb[m] * sw[q * 4 + m];
        const double t = v.t1v(l, k, q);
        for (int m = 0; m < 4; ++m) sw_bar[q][m] += t * zb[m];
        const double pre_bar = ob * (1.0 - t * t);
        sink.add(L.conv_b[l] + k * Q + q, pre_bar);
        const double pb = pre_bar * 0.25;
        for (int c = 0; c < 4; ++c) {
          const int idx = 4 * q + c;
          const double xin = idx < DIN ? v.hin(l, k, idx) : idx < 2 * DIN ? v.gmean(l, 0, idx - DIN)
                             : idx < 3 * DIN ? v.gmean(l, 1, idx - 2 * DIN)
                             : idx < 3 * DIN + 4 ? v.G(l, 0, k, idx - 3 * DIN) * inv_n[0]
                                                 : v.G(l, 1, k, idx - 3 * DIN - 4) * inv_n[1];
          sink.add(L.conv_w[l] + k * DTOT + idx, pb * xin);
          const double xb = pb * cw[idx];
          if (idx < DIN) own[k][idx] += xb;
          else if (idx < 2 * DIN) gup_bar[idx - DIN] += xb;
          else if (idx < 3 * DIN) gdn_bar[idx - 2 * DIN] += xb;
          else if (idx < 3 * DIN + 4) G_bar_l[0][k][idx - 3 * DIN] = xb * inv_n[0];
          else G_bar_l[1][k][idx - 3 * DIN - 4] = xb * inv_n[1];
        }
      }
    }
    for (int q = 0; q < Q; ++q)
      for (int m = 0; m < 4; ++m) sink.add(L.sing_w[l] + q * 4 + m, sw_bar[q][m]);
    for (int m = 0; m < 4; ++m) sink.add(L.sing_b[l] + m, sb_bar[m]);
    for (int k = 0; k < N; ++k) {
      const bool up = k < sys.n_up;
      for (int q = 0; q < DIN; ++q) hin_bar[k][q] = own[k][q] + (up ? gup_bar[q] * inv_n[0] : gdn_bar[q] * inv_n[1]);
    }
  }

  // adjoint of out = (in + tanh(in . W + b)) / sqrt2 with the weight gradients accumulated into wb[16], bb[4];
  // in_bar = extra + [out_bar + W (g o out_bar)] / sqrt2  (in_bar may be null: level 0 has nothing below it)
  static AQ_HD void chain_reverse(const double* __restrict__ W, const double* __restrict__ in, const double* __restrict__ out,
                                  const double* __restrict__ out_bar, const double* __restrict__ extra,
                                  double* __restrict__ in_bar, double* __restrict__ wb, double* __restrict__ bb) {
    double zb[4];
    for (int m = 0; m < 4; ++m) {
      const double t = kSqrt2 * out[m] - in[m];
      zb[m] = out_bar[m] * kInvSqrt2 * (1.0 - t * t);
      bb[m] += zb[m];
    }
    for (int q = 0; q < 4; ++q) {
      double v = in_bar ? out_bar[q] * kInvSqrt2 + extra[q] : 0.0;
      for (int m = 0; m < 4; ++m) {
        wb[q * 4 + m] += in[q] * zb[m];
        v += W[q * 4 + m] * zb[m];
      }
      if (in_bar) in_bar[q] = v;
    }
  }

  template <class Sink>
  static AQ_HD void run(const AiqmcSystem& sys, const double* __restrict__ P, const double* __restrict__ x, double alpha,
                        double beta, double& phase, double& logabs, Sink& sink) {
    typename PS::Primal pr;
    cplx Mi[N * N];
    double hp[3 * N * N * 4];
    double t1[3 * N * QM];
    PS::forward(sys, P, x, pr, Mi, hp, 1, t1, 1);
    double ld;
    PS::gj_inverse(Mi, ld, phase);
    logabs = ld + pr.jastrow;
    const LocalView v{pr, Mi, hp, t1};
    run_impl(sys, P, x, v, alpha, beta, sink);
    keep_alive(&pr); keep_alive(Mi); keep_alive(hp); keep_alive(t1);
  }

  // the sweep on the derivative cache of one configuration (written by DerivSplit::primal<false>)
  template <class Sink>
  static AQ_HD void run_cached(const AiqmcSystem& sys, const double* __restrict__ P, const double* __restrict__ dc,
                               int64_t stride, double alpha, double beta, Sink& sink) {
    double x[3 * N];
    for (int q = 0; q < 3 * N; ++q) x[q] = dc[(int64_t)(DerivCache<NE, NA>::X + q) * stride];
    const CacheView v{dc, stride};
    run_impl(sys, P, x, v, alpha, beta, sink);
    keep_alive(x);
  }

  template <class View, class Sink>
  static AQ_HD void run_impl(const AiqmcSystem& sys, const double* __restrict__ P, const double* __restrict__ x,
                             const View& v, double alpha, double beta, Sink& sink) {
    constexpr LayoutC<NE, NA> L{};
    const double inv_n[2] = {1.0 / sys.n_up, 1.0 / sys.n_dn};

    // (1) determinant, column by column
    double h_bar[N][4], y_bar[N][6], env_bar[N];
    for (int k = 0; k < N; ++k) {
      env_bar[k] = 0.0;
      for (int c = 0; c < 4; ++c) h_bar[k][c] = 0.0;
      for (int m = 0; m < 6; ++m) y_bar[k][m] = 0.0;
    }
    for (int j = 0; j < N; ++j) {
      double ow[2][4][2], ob[2][2], yw[6];
      for (int s = 0; s < 2; ++s) {
        ob[s][0] = ob[s][1] = 0.0;
        for (int c = 0; c < 4; ++c) ow[s][c][0] = ow[s][c][1] = 0.0;
      }
      for (int m = 0; m < 6; ++m) yw[m] = 0.0;
      for (int k = 0; k < N; ++k) {
        const int s = k < sys.n_up_rows ? 0 : 1;
        const double* W = P + L.orb_w[s];
        const double* Bv = P + L.orb_b[s];
        const int e = sys.sigma[k];
        double yo = 0.0;
        for (int m = 0; m < 6; ++m) yo += v.y(k, m) * P[L.y_w + m * N + j];
        double pre = Bv[2 * j], pim = Bv[2 * j + 1];
        double h3[4];
        for (int c = 0; c < 4; ++c) h3[c] = v.hout(2, e, c);
        for (int c = 0; c < 4; ++c) { pre += h3[c] * W[c * 2 * N + 2 * j]; pim += h3[c] * W[c * 2 * N + 2 * j + 1]; }
        const cplx mi = v.mi(j, k);
        const double cre = alpha * mi.re + beta * mi.im, cim = alpha * mi.im - beta * mi.re;   // (alpha - i beta) M^-1[j,k]
        const double envk = v.env(k);
        const double ev = envk * yo;
        const double pre_bar = cre * ev, pim_bar = -cim * ev;
        const double ev_bar = cre * pre - cim * pim;
        ob[s][0] += pre_bar; ob[s][1] += pim_bar;
        for (int c = 0; c < 4; ++c) {
          ow[s][c][0] += h3[c] * pre_bar;
          ow[s][c][1] += h3[c] * pim_bar;
          h_bar[e][c] += pre_bar * W[c * 2 * N + 2 * j] + pim_bar * W[c * 2 * N + 2 * j + 1];
        }
        env_bar[k] += ev_bar * yo;
        const double yo_bar = ev_bar * envk;
        for (int m = 0; m < 6; ++m) { y_bar[k][m] += yo_bar * P[L.y_w + m * N + j]; yw[m] += yo_bar * v.y(k, m); }
      }
      for (int s = 0; s < 2; ++s) {
        sink.add(L.orb_b[s] + 2 * j, ob[s][0]);
        sink.add(L.orb_b[s] + 2 * j + 1, ob[s][1]);
        for (int c = 0; c < 4; ++c) {
          sink.add(L.orb_w[s] + c * 2 * N + 2 * j, ow[s][c][0]);
          sink.add(L.orb_w[s] + c * 2 * N + 2 * j + 1, ow[s][c][1]);
        }
      }
      for (int m = 0; m < 6; ++m) sink.add(L.y_w + m * N + j, yw[m]);
    }

    // (2) one-electron layers, backwards
    double G_bar[3][2][N][4];
    double h0_bar[N][4 * A];
    for (int l = 2; l >= 0; --l) {
      if (l == 0) layer_reverse<4 * A>(sys, P, 0, v, inv_n, h_bar, G_bar[0], h0_bar, sink);
      else layer_reverse<4>(sys, P, l, v, inv_n, h_bar, G_bar[l], h_bar, sink);
    }

    // (3) pair chains, backwards (all ordered pairs, diagonal included); e-e Pade parameters
    double wb[2][16], bb[2][4];
    for (int l = 0; l < 2; ++l) {
      for (int q = 0; q < 16; ++q) wb[l][q] = 0.0;
      for (int m = 0; m < 4; ++m) bb[l][m] = 0.0;
    }
    for (int i = 0; i < N; ++i) {
      const int s = i < sys.n_up ? 0 : 1;
      for (int j = 0; j < N; ++j) {
        double a0[4], a1[4], a2[4];
        for (int c = 0; c < 4; ++c) { a0[c] = v.hpv(0, i, j, c); a1[c] = v.hpv(1, i, j, c); a2[c] = v.hpv(2, i, j, c); }
        double b1[4];
        chain_reverse(P + L.dbl_w[1], a1, a2, G_bar[2][s][j], G_bar[1][s][j], b1, wb[1], bb[1]);
        chain_reverse(P + L.dbl_w[0], a0, a1, b1, nullptr, nullptr, wb[0], bb[0]);
        if (i < j) {                                            // u = cusp r / (1 + a r)  (Jastrow.py:23-41)
          const double r = a0[0];
          const double q = s_inv(1.0 + P[L.jas_alpha + i * N + j] * r);
          sink.add(L.jas_alpha + i * N + j, -alpha * P[L.jas_cusp + i * N + j] * r * r * q * q);
        }
      }
    }
    for (int l = 0; l < 2; ++l) {
      for (int q = 0; q < 16; ++q) sink.add(L.dbl_w[l] + q, wb[l][q]);
      for (int m = 0; m < 4; ++m) sink.add(L.dbl_b[l] + m, bb[l][m]);
    }

    // (4) electron-local parts: Ynlm stream (recomputed with its tape), envelope, e-n Jastrow
    double ywb[2][36], ybb[3][6], yw0b[K0 * 6];
    for (int l = 0; l < 2; ++l) for (int q = 0; q < 36; ++q) ywb[l][q] = 0.0;
    for (int l = 0; l < 3; ++l) for (int m = 0; m < 6; ++m) ybb[l][m] = 0.0;
    for (int q = 0; q < K0 * 6; ++q) yw0b[q] = 0.0;
    for (int e = 0; e < N; ++e) {
      const double c0 = 0.28209479177387814, c1 = 0.48860251190291992, k15h = 1.0925484305920792,
                   k5q = 0.31539156525252005, k15q = 0.54627421529603959, k35 = 0.59004358992664352,
                   k105h = 2.8906114426405538, k21 = 0.45704579946446577, k7q = 0.37317633259011546,
                   k105q = 1.4453057213202769;
      double in0[K0];
      double sum_df = 0.0, sum_sp = 0.0, sum_E1 = 0.0;
      for (int a = 0; a < A; ++a) {
        double ae[3];
        for (int c = 0; c < 3; ++c) ae[c] = x[3 * e + c] - P[L.atoms + 3 * a + c];
        const double r2 = ae[0] * ae[0] + ae[1] * ae[1] + ae[2] * ae[2];
        double r, ri;
        s_sqrt_inv(r2, r, ri);
        const double t0 = ae[0] * ri, t1v = ae[1] * ri, t2 = ae[2] * ri;
        const double sp[4] = {c0, t0 * c1, t1v * c1, t2 * c1};
        for (int q = 0; q < 4; ++q) { in0[4 * a + q] = sp[q]; sum_sp += sp[q]; }
        const double ri2 = ri * ri, ri3 = ri2 * ri;
        const double t00 = t0 * t0, t11 = t1v * t1v, t22 = t2 * t2;
        const double d2 = (t0 * t1v) * k15h + (t1v * t2) * k15h + (t22 * 3.0 - r2) * k5q + (t0 * t2) * k15h + (t00 - t11) * k15q;
        const double f3 = (t1v * (t00 * 3.0 - t11)) * k35 + (t0 * t1v * t2) * k105h + (t1v * (t22 * 5.0 - r2)) * k21 +
                          (t22 * t2 * 5.0 - t2 * r2 * 3.0) * k7q + (t0 * (t22 * 5.0 - r2)) * k21 +
                          ((t00 - t11) * t2) * k105q + (t0 * (t00 - t11 * 3.0)) * k35;
        sum_df += d2 * ri2 + f3 * ri3;
        // envelope (envelope.py:26-30) and e-n Pade term (Jastrow.py:84) parameters of this (electron, atom)
        const double be = P[L.env_beta + e * A + a], al = P[L.env_alpha + e];
        const double E1 = s_exp(r2 * (-be));
        sink.add(L.env_beta + e * A + a, env_bar[e] * al * E1 * (-r2));
        for (int c = 0; c < 3; ++c) {
          const double pi = P[L.env_pi + (e * A + a) * 3 + c], sx = P[L.env_sx + (e * A + a) * 3 + c];
          const double E2 = s_exp(ae[c] * (-pi));
          sink.add(L.env_sx + (e * A + a) * 3 + c, env_bar[e] * E2);
          sink.add(L.env_pi + (e * A + a) * 3 + c, env_bar[e] * sx * E2 * (-ae[c]));
        }
        sum_E1 += E1;                                            // env_alpha[e] multiplies sum_a exp(-beta r^2)
        const double bj = P[L.jas_beta + e * A + a], c14 = P[L.jas_c14 + a], c34 = P[L.jas_c34 + a];
        const double ex = s_exp(r * (-c14 * bj));
        const double ib = s_inv(bj);
        sink.add(L.jas_beta + e * A + a, alpha * (ex * (-r * c14) * c34 * 0.5 * ib - (ex - 1.0) * c34 * 0.5 * ib * ib));
      }
      sink.add(L.env_alpha + e, env_bar[e] * sum_E1);
      in0[4 * A] = sum_df * (1.0 / (12.0 * A));
      in0[4 * A + 1] = sum_sp * (1.0 / (4.0 * A));
      // forward Ynlm stream with its levels
      double ylv[3][6];                                          // stream after layer 0, 1, 2
      {
        double z[6];
        for (int m = 0; m < 6; ++m) z[m] = P[L.yn_b[0] + m];
        for (int q = 0; q < K0; ++q)
          for (int m = 0; m < 6; ++m) z[m] += in0[q] * P[L.yn_w[0] + q * 6 + m];
        for (int m = 0; m < 6; ++m) {
          const double t = s_tanh(z[m]);
          ylv[0][m] = (A == 1) ? (in0[m] + t) * kInvSqrt2 : t;    // residual only if 4A+2 == 6 (quirk Q5)
        }
        for (int l = 1; l < 3; ++l) {
          double zz[6];
          for (int m = 0; m < 6; ++m) zz[m] = P[L.yn_b[l] + m];
          for (int q = 0; q < 6; ++q)
            for (int m = 0; m < 6; ++m) zz[m] += ylv[l - 1][q] * P[L.yn_w[l] + q * 6 + m];
          for (int m = 0; m < 6; ++m) ylv[l][m] = (ylv[l - 1][m] + s_tanh(zz[m])) * kInvSqrt2;
        }
      }
      // backwards
      double ob[6];
      for (int m = 0; m < 6; ++m) ob[m] = y_bar[e][m];
      for (int l = 2; l >= 1; --l) {
        double zb[6], ib[6];
        for (int m = 0; m < 6; ++m) {
          const double t = kSqrt2 * ylv[l][m] - ylv[l - 1][m];
          zb[m] = ob[m] * kInvSqrt2 * (1.0 - t * t);
          ybb[l][m] += zb[m];
        }
        for (int q = 0; q < 6; ++q) {
          double v = ob[q] * kInvSqrt2;
          for (int m = 0; m < 6; ++m) {
            ywb[l - 1][q * 6 + m] += ylv[l - 1][q] * zb[m];
            v += P[L.yn_w[l] + q * 6 + m] * zb[m];
          }
          ib[q] = v;
        }
        for (int m = 0; m < 6; ++m) ob[m] = ib[m];
      }
      for (int m = 0; m < 6; ++m) {
        const double t = (A == 1) ? kSqrt2 * ylv[0][m] - in0[m] : ylv[0][m];
        const double zb = ((A == 1) ? ob[m] * kInvSqrt2 : ob[m]) * (1.0 - t * t);
        ybb[0][m] += zb;
        for (int q = 0; q < K0; ++q) yw0b[q * 6 + m] += in0[q] * zb;
      }
    }
    for (int q = 0; q < K0 * 6; ++q) sink.add(L.yn_w[0] + q, yw0b[q]);
    for (int l = 1; l < 3; ++l)
      for (int q = 0; q < 36; ++q) sink.add(L.yn_w[l] + q, ywb[l - 1][q]);
    for (int l = 0; l < 3; ++l)
      for (int m = 0; m < 6; ++m) sink.add(L.yn_b[l] + m, ybb[l][m]);

    keep_alive(h_bar); keep_alive(y_bar);
    keep_alive(env_bar); keep_alive(G_bar); keep_alive(h0_bar); keep_alive(wb); keep_alive(bb); keep_alive(ywb);
    keep_alive(ybb); keep_alive(yw0b);
  }
};

}  // namespace aiqmc
