// ecp_coop.cuh -- the ccECP non-local quadrature (pseudopotential.py:272-318, pp_energy_test.py:
// 45-105), B200 version 2: sub-warp groups, one lane per electron, single-electron-move caching.
//
// Every quadrature point needs log psi at a configuration that differs from the walker's in ONE
// electron i.  Of the whole network only this changes:
//   * the 2(N-1) pair chains that touch i         (lane k recomputes (i,k) and (k,i))
//   * electron i's Ynlm stream / envelope / e-n Jastrow
//   * every row of the one-electron stream, but only through cached block sums:
//        G'_l[s][k] = G_l[s][k] + [s == s_i] (h'_l[i,k] - h_l[i,k])          (k != i)
//        G'_l[s][i] = sum_{k in s} h'_l[k,i]                                 (group reduction)
//   * the N x N complex determinant (no rank-1 shortcut: every row changes)
// A group of G = 4/8/16/32 lanes owns one point; lane k is electron k and row k of the LU
// (partial pivoting by group arg-max, pivot row broadcast by shuffles, permutation parity from an
// inversion count).  One CTA owns one walker: parameters, the walker's MoveCache and the exp table
// sit in shared memory; per-group partial sums are reduced in a fixed order (deterministic, no
// atomics).  Registers per lane drop from 255 (+2.7 kB local) to ~100, no local memory.
#pragma once
#include "psi_core.cuh"

namespace aiqmc {

template <int NE> struct GroupSize { static constexpr int G = NE <= 4 ? 4 : NE <= 8 ? 8 : NE <= 16 ? 16 : 32; };

template <int G>
__device__ __forceinline__ double gsum(double v, unsigned mask) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, G);
  return v;
}
template <int G>
__device__ __forceinline__ int gsum_i(int v, unsigned mask) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, G);
  return v;
}

// threads per CTA: G * groups with groups | E when possible (no idle tail pass), a multiple of 32, <= 256
template <int NE, int NA>
constexpr int coop_threads() {
  constexpr int G = GroupSize<NE>::G;
  constexpr int E = NE * NA * AIQMC_NQUAD;
  for (int m = 8; m >= 1; --m) {
    const int groups = 32 * m / G;
    if (groups > 0 && E % groups == 0) return 32 * m;
  }
  return 128;
}
#ifndef AIQMC_COOP_MINB
#define AIQMC_COOP_MINB 4          // resident CTAs/SM the register allocator must allow
#endif

template <int NE, int NA>
struct CoopSmem {
  static constexpr int kParams = (make_layout(NE, NA).total + 1) & ~1;
  static constexpr int kCache = (MoveCache<NE, NA>::SIZE + 1) & ~1;
  static constexpr int kPos = (3 * NE + 1) & ~1;
  static __host__ __device__ constexpr int doubles(int groups) { return kParams + kCache + kPos + 2 * groups + 16; }
};

template <int NE, int NA>
__global__ void __launch_bounds__((coop_threads<NE, NA>()), AIQMC_COOP_MINB) k_ecp_coop(AiqmcSystem sys, const double* __restrict__ params,
                                                  const double* __restrict__ pos, const double* __restrict__ rot,
                                                  int64_t B, const double* __restrict__ cache_all, EnergyWs w) {
  constexpr int G = GroupSize<NE>::G;
  constexpr int N = NE, A = NA;
  using MC = MoveCache<NE, NA>;
  using SM = CoopSmem<NE, NA>;
  constexpr LayoutC<NE, NA> L{};
  extern __shared__ double smem[];
  double* sP = smem;
  double* sC = sP + SM::kParams;
  double* sX = sC + SM::kCache;
  double* sAcc = sX + SM::kPos;
  const int64_t b = blockIdx.x;
  {
    constexpr int total = make_layout(NE, NA).total;
    for (int q = threadIdx.x; q < total; q += blockDim.x) sP[q] = params[q];
    const double* cb = cache_all + b * MC::SIZE;
    for (int q = threadIdx.x; q < MC::SIZE; q += blockDim.x) sC[q] = cb[q];
    for (int q = threadIdx.x; q < 3 * N; q += blockDim.x) sX[q] = pos[b * 3 * N + q];
    if (threadIdx.x < kExpTab) g_exp_tab[threadIdx.x] = exp2((double)threadIdx.x * (1.0 / kExpTab));
  }
  __syncthreads();
  const double* P = sP;
  const int k = threadIdx.x % G;
  const int grp = threadIdx.x / G, ngrp = blockDim.x / G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (((threadIdx.x & 31) / G) * G));
  const bool act = k < N;
  const int kk = act ? k : N - 1;
  const int sk = kk < sys.n_up ? 0 : 1;
  const double inv_n[2] = {1.0 / sys.n_up, 1.0 / sys.n_dn};
  const double xk[3] = {sX[3 * kk], sX[3 * kk + 1], sX[3 * kk + 2]};
  double R[9];
#pragma unroll
  for (int q = 0; q < 9; ++q) R[q] = rot[b * 9 + q];
  const double den_r = sC[MC::MISC + 1], den_i = sC[MC::MISC + 2];
  const double den_inv = 1.0 / (den_r * den_r + den_i * den_i);
  constexpr int E = N * A * AIQMC_NQUAD;
  double acc_re = 0.0, acc_im = 0.0;

  for (int ev = grp; ev < E; ev += ngrp) {
    const int i = ev / (A * AIQMC_NQUAD);
    const int a = (ev - i * A * AIQMC_NQUAD) / AIQMC_NQUAD;
    const int p = ev - (i * A + a) * AIQMC_NQUAD;
    const double* vl = w.vl + ((b * N + i) * A + a) * 4;
    const double v0 = vl[0], v1 = vl[1], v2 = vl[2], v3 = vl[3];
    if (v0 == 0.0 && v1 == 0.0 && v2 == 0.0 && v3 == 0.0) continue;   // exact zero channel: contributes 0
    // ---- rotated point, cos(theta) (quirks Q13, Q14) -- computed redundantly by every lane
    double ae[3], nh[3], xn[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) ae[c] = sX[3 * i + c] - P[L.atoms + 3 * a + c];
    const double r = sqrt(ae[0] * ae[0] + ae[1] * ae[1] + ae[2] * ae[2]);
#pragma unroll
    for (int l = 0; l < 3; ++l)
      nh[l] = c_ecp.quad_pts[p][0] * R[l] + c_ecp.quad_pts[p][1] * R[3 + l] + c_ecp.quad_pts[p][2] * R[6 + l];
    double dot = 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) { xn[c] = r * nh[c]; dot += ae[c] * xn[c]; }
    const double cs = dot / (r * (r * w.gnorm[4 * b + quad_group(p)]));
    const int si = i < sys.n_up ? 0 : 1;
    const bool diag = (kk == i);

    // ---- pair chains touching electron i: row (i,k): d = x_k - x_i', col (k,i): d = x_i' - x_k
    double cr[4], cc[4];
    {
      double d[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) d[c] = diag ? 0.0 : xk[c] - xn[c];
      const double rik = diag ? 0.0 : sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      cr[0] = rik; cc[0] = rik;
#pragma unroll
      for (int c = 0; c < 3; ++c) { cr[1 + c] = d[c]; cc[1 + c] = -d[c]; }
    }
    // e-e Jastrow change (Jastrow.py:23-41)
    double du = 0.0;
    if (act && !diag) {
      const int lo = i < k ? i : k, hi = i < k ? k : i;
      const double cu = P[L.jas_cusp + lo * N + hi], al = P[L.jas_alpha + lo * N + hi];
      const double rold = sC[MC::HP + ((0 * N + i) * N + k) * 4];
      du = cu * cr[0] / (1.0 + al * cr[0]) - cu * rold / (1.0 + al * rold);
    }
    const double dJee = gsum<G>(du, gmask);

    // ---- electron i at its new position: features, Ynlm stream, envelope, e-n Jastrow
    double h0n[4 * A], yn[6], envn, jaen;
    Psi<NE, NA>::template electron_local<double>(P, i, xn, h0n, yn, envn, jaen);

    // ---- one-electron stream, lane k = electron k
    double h[4];
#pragma unroll
    for (int l = 0; l < 3; ++l) {
      double Gu[4], Gd[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double su = gsum<G>((act && sk == 0) ? cc[c] : 0.0, gmask);
        const double sd = gsum<G>((act && sk == 1) ? cc[c] : 0.0, gmask);
        double gu = sC[MC::GS + ((l * 2 + 0) * N + kk) * 4 + c], gd = sC[MC::GS + ((l * 2 + 1) * N + kk) * 4 + c];
        const double delta = cr[c] - sC[MC::HP + ((l * N + i) * N + kk) * 4 + c];
        if (si == 0) gu += delta; else gd += delta;
        Gu[c] = (diag ? su : gu) * inv_n[0];
        Gd[c] = (diag ? sd : gd) * inv_n[1];
      }
      if (l == 0) {
        double hk[4 * A], g0u[4 * A], g0d[4 * A];
#pragma unroll
        for (int q = 0; q < 4 * A; ++q) {
          const double dh = (h0n[q] - sC[MC::H0 + i * 4 * A + q]);
          hk[q] = diag ? h0n[q] : sC[MC::H0 + kk * 4 * A + q];
          g0u[q] = sC[MC::G0M + q] + (si == 0 ? dh * inv_n[0] : 0.0);
          g0d[q] = sC[MC::G0M + 4 * A + q] + (si == 1 ? dh * inv_n[1] : 0.0);
        }
        Psi<NE, NA>::template one_layer<4 * A, double>(P, 0, kk, hk, g0u, g0d, Gu, Gd, h);
      } else {
        double gm[2][4], hn[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          gm[0][c] = gsum<G>((act && sk == 0) ? h[c] : 0.0, gmask) * inv_n[0];
          gm[1][c] = gsum<G>((act && sk == 1) ? h[c] : 0.0, gmask) * inv_n[1];
        }
        Psi<NE, NA>::template one_layer<4, double>(P, l, kk, h, gm[0], gm[1], Gu, Gd, hn);
#pragma unroll
        for (int c = 0; c < 4; ++c) h[c] = hn[c];
      }
      if (l < 2) {   // advance both pair chains through double-layer l (nn.py:305-309)
        const double* W = P + L.dbl_w[l];
        double zr[4], zc[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) { zr[m] = P[L.dbl_b[l] + m]; zc[m] = zr[m]; }
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int m = 0; m < 4; ++m) { zr[m] += cr[q] * W[q * 4 + m]; zc[m] += cc[q] * W[q * 4 + m]; }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          cr[m] = (cr[m] + s_tanh(zr[m])) * kInvSqrt2;
          cc[m] = (cc[m] + s_tanh(zc[m])) * kInvSqrt2;
        }
      }
    }

    // ---- orbital-matrix row of lane k: P[k,j] * env[k] * Yo[k,j]   (nn.py:432-504)
    cplx row[N];
    {
      const int e = sys.sigma[kk];
      double hs[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) hs[c] = __shfl_sync(gmask, h[c], e, G);
      const int srow = kk < sys.n_up_rows ? 0 : 1;
      const double* W = P + L.orb_w[srow];
      const double* Bv = P + L.orb_b[srow];
      double yr[6];
#pragma unroll
      for (int m = 0; m < 6; ++m) yr[m] = diag ? yn[m] : sC[MC::Y + kk * 6 + m];
      const double envr = diag ? envn : sC[MC::ENV + kk];
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double pre = Bv[2 * j], pim = Bv[2 * j + 1];
#pragma unroll
        for (int c = 0; c < 4; ++c) { pre += hs[c] * W[c * 2 * N + 2 * j]; pim += hs[c] * W[c * 2 * N + 2 * j + 1]; }
        double yo = 0.0;
#pragma unroll
        for (int m = 0; m < 6; ++m) yo += yr[m] * P[L.y_w + m * N + j];
        const double evv = envr * yo;
        row[j] = {pre * evv, pim * evv};
      }
    }
    // ---- complex LU across lanes: pivot = group arg-max of |row[c]|^2 among unused lanes
    bool used = !act;
    int mystep = act ? N : -1;
    double mant = 1.0;
    int ex = 0;
    cplx ph = {1.0, 0.0};
#pragma unroll
    for (int c = 0; c < N; ++c) {
      double bm = used ? -1.0 : row[c].re * row[c].re + row[c].im * row[c].im;
      int best = k;
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) {
        const double om = __shfl_xor_sync(gmask, bm, o, G);
        const int ol = __shfl_xor_sync(gmask, best, o, G);
        if (om > bm || (om == bm && ol < best)) { bm = om; best = ol; }
      }
      cplx piv = {__shfl_sync(gmask, row[c].re, best, G), __shfl_sync(gmask, row[c].im, best, G)};
      const double amag = sqrt(bm);
      int e2;
      mant = frexp(mant * amag, &e2);
      ex += e2;
      ph = cscale(cmul(ph, piv), 1.0 / amag);
      const cplx pinv = cinv(piv);
      if (k == best) { used = true; mystep = c; }
      const cplx f = cmul(row[c], pinv);
#pragma unroll
      for (int j = c + 1; j < N; ++j) {
        const cplx pj = {__shfl_sync(gmask, row[j].re, best, G), __shfl_sync(gmask, row[j].im, best, G)};
        if (!used) cfms(row[j], f, pj);
      }
    }
    // permutation parity: inversions of (lane -> step)
    int inv = 0;
#pragma unroll
    for (int q = 0; q < N; ++q) {
      const int st = __shfl_sync(gmask, mystep, q, G);
      if (act && q > k && st < mystep) ++inv;
    }
    if (gsum_i<G>(inv, gmask) & 1) { ph.re = -ph.re; ph.im = -ph.im; }

    const double la = log(mant) + ex * 0.69314718055994530942 + sC[MC::MISC + 0] + dJee + (jaen - sC[MC::JAE + i]);
    const double pha = atan2(ph.im, ph.re);
    // ratio = log psi(x') / log psi(x) * weight with complex logs (quirk Q12)
    const double wq = c_ecp.quad_wts[p] * den_inv;
    const double rr = (la * den_r + pha * den_i) * wq, ri = (pha * den_r - la * den_i) * wq;
    const double k4 = 0.07957747154594767;   // 1/(4 pi)
    const double f = v0 * k4 + v1 * (3.0 * k4 * cs) + v2 * (2.5 * k4 * (3.0 * cs * cs - 1.0)) +
                     v3 * (3.5 * k4 * (5.0 * cs * cs * cs - 3.0 * cs));
    acc_re += f * rr;
    acc_im += f * ri;
  }
  if (k == 0) { sAcc[2 * grp] = acc_re; sAcc[2 * grp + 1] = acc_im; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sr = 0.0, si2 = 0.0;
    for (int g = 0; g < ngrp; ++g) { sr += sAcc[2 * g]; si2 += sAcc[2 * g + 1]; }
    w.epp[2 * b] = sr;
    w.epp[2 * b + 1] = si2;
  }
}

}  // namespace aiqmc
