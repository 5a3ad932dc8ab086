// coop_lap.cuh -- gradient + per-coordinate second derivatives of log|psi| (the forward-Laplacian of the kinetic
// energy, Energy/pphamiltonian.py:74-106) as ONE kernel with the derivative cache in SHARED memory.
//
// Version 2 (deriv_split.cuh) ran two kernels: a one-thread-per-configuration primal pass (255 registers + 3.6-7.6 kB
// of local memory per thread) that wrote a 5-25 kB record per configuration to HBM, and a tangent pass that read it
// back through L2 with ~250 dependent loads per thread (0.05 of the FP64 roofline).  Here a CTA works on
// tiles of floor(32/N) configurations whose records never leave the SM, as a two-stage software pipeline over a
// double-buffered record area: while the tangent warps consume tile j, the producer warp builds tile j+1 (the two
// stages cost about the same per tile; run back to back behind one barrier the tangent warps idled a third of the time):
//   phase 1  (warp 0, one lane per electron -- the forward half of coop_grad.cuh): features, pair chains (lane k owns
//            column k, so the block sums are lane-local), the three one-electron layers, the orbital matrix, its
//            Gauss-Jordan inverse across the lanes, and the contractions Gm / T of M^-1 with the orbital weights; every
//            intermediate a derivative needs goes into the configuration's DerivCache record in shared memory;
//   phase 2  (all warps, one thread per (configuration, electron, direction)): DerivSplit::tangent -- the validated
//            tangent pass of version 2 -- pushed through the record with shared-memory latency instead of L2 latency.
// The single-electron-move cache of the quadrature kernels (MoveCache, HBM, array-of-structs) is copied out of the
// records by the whole CTA with coalesced stores.  Outputs as k_primal + k_tangent: grad (n_cfg,3N), lap_parts
// (3N,stride), phase / log|psi|.
#pragma once
#include "coop_grad.cuh"

#ifndef AIQMC_LAP_PIPE
#define AIQMC_LAP_PIPE 0           // 1: producer warp + tangent warps as a double-buffered pipeline (measured: 1.04 vs 0.97 ms
#endif                             //    on carbon -- the second record buffer costs two resident CTAs per SM)
#ifndef AIQMC_LAP_COOP_TANGENT
#define AIQMC_LAP_COOP_TANGENT 0   // 0: one thread per (cfg, electron, direction) = DerivSplit::tangent (1.3 kB of local arrays);
                                   // 1: lane-per-row tangent pass below: no local memory (176 B stack), but the moved electron's
                                   //    jets serialise inside each group and every lane repeats the block sums -- 2x the
                                   //    instructions: measured 1.66 ms against 0.97 ms for the kinetic stage of carbon
#endif
#ifndef AIQMC_LAP_MINB
#define AIQMC_LAP_MINB (AIQMC_LAP_PIPE ? 2 : 4)   // resident CTAs/SM the register allocator must allow
#endif

namespace aiqmc {

template <int NE, int NA>
struct CoopLapCfg {
  using DC = DerivCache<NE, NA>;
  using MC = MoveCache<NE, NA>;
  static constexpr int N = NE, A = NA;
  static constexpr int GPW = 32 / NE;                         // configurations per tile (phase 1 = one warp)
  static constexpr int NG = GPW;
  static constexpr int oJAE = DC::SIZE_LAP, oJEE = oJAE + NE, oMISC = oJEE + NE;
  static constexpr int kRecRaw = (oMISC + 4 + 1) & ~1;
  static constexpr int REC = kRecRaw + ((2 - kRecRaw % 16) + 16) % 16;        // even, = 2 mod 16: groups 4 banks apart
  // phase-1 scratch per group
  static constexpr int oMS = 0;                               // [N][N][2] transpose / inverse exchange
  static constexpr int oPIV = oMS + 2 * NE * NE;              // [2][N][2]
  static constexpr int kScrRaw = (oPIV + 4 * NE + 1) & ~1;
  static constexpr int SCR = kScrRaw + ((2 - kScrRaw % 16) + 16) % 16;
  // phase-2 scratch per (direction, configuration) group of the lane-per-row tangent pass
  static constexpr int o2H0 = 0;                              // [2][4A]   d, s of the moved electron's layer-0 features
  static constexpr int o2A = o2H0 + 8 * NA;                   // [N][8]    per-row deposits (column-chain tangents / row tangents)
  static constexpr int o2S1 = o2A + 8 * NE;                   // [N][2]    S1[l] of the moved electron's row
  static constexpr int o2X = o2S1 + 2 * NE;                   // [N][N][2] X = dM M^-1, row k from lane k
  static constexpr int o2R = o2X + 2 * NE * NE;               // [N][2]    per-lane partial results
  static constexpr int kScr2Raw = (o2R + 2 * NE + 1) & ~1;
  static constexpr int SCR2 = kScr2Raw + ((2 - kScr2Raw % 16) + 16) % 16;
  static constexpr bool kCoopTan = AIQMC_LAP_COOP_TANGENT != 0;
  static constexpr int kTan = NG * 3 * NE;                    // phase-2 threads that have work
  static constexpr int T2 = ((kTan > 32 ? kTan : 32) + 31) / 32 * 32;   // tangent (consumer) threads
  static constexpr bool kPipe = AIQMC_LAP_PIPE != 0;
  static constexpr int T = kPipe ? 32 + T2 : T2;              // pipeline: + a dedicated producer warp
  static constexpr int kPar = (make_layout(NE, NA).total + 1) & ~1;
  static constexpr int kDoubles = kPar + NG * ((kPipe ? 2 : 1) * REC + SCR) + (kCoopTan ? 3 * NG * SCR2 : 0);
  static_assert(T2 == 96, "phase 2 maps one warp to one Cartesian direction");
  static constexpr int kBytes = kDoubles * 8;
  static constexpr int fit_blocks(int want) { return want <= 1 ? 1 : (kBytes * want <= 224 * 1024 ? want : fit_blocks(want - 1)); }
  static constexpr int kMinBlocks = fit_blocks(AIQMC_LAP_MINB);
};


// ---- phase 2, lane-per-row: one group of N lanes = (configuration, Cartesian direction); lane k carries the tangent
//      (first and second derivative along x_{e,dir}) of ROW k of the one-electron stream and of the two pair chains
//      (e,k), (k,e) through electron e, for e = 0..N-1 in turn.  The per-thread version (DerivSplit::tangent) keeps
//      those for ALL rows in arrays indexed by a loop variable: 1.3 kB of local memory per thread, 1 GB of DRAM writes
//      per launch.  Here every array index is a compile-time constant; cross-lane traffic (block means, the gather
//      through sigma, X = dM M^-1 for tr(XX)) goes through a small per-group scratch.  Same mathematics, same record.
template <int NE, int NA>
__device__ __forceinline__ void lap_tangent_coop(const AiqmcSystem& sys, const double* __restrict__ P,
                                                 const double* __restrict__ rec, double* __restrict__ scr, int k, bool idle,
                                                 int dir, double& g_out, double& l_out, int e_sel) {
  using DC = DerivCache<NE, NA>;
  using DS = DerivSplit<NE, NA>;
  using PS = Psi<NE, NA>;
  using CF = CoopLapCfg<NE, NA>;
  using Tan4 = typename DS::Tan4;
  using J = Jet<true, 1>;
  using Op = ScalarOps<J>;
  constexpr int N = NE, A = NA, QM = DC::QM;
  constexpr LayoutC<NE, NA> L{};
  constexpr double kSqrt2 = 1.41421356237309504880;
  const int e = e_sel;
  const int n_up = sys.n_up;
  const double inv_n[2] = {1.0 / sys.n_up, 1.0 / sys.n_dn};
  const int se = e < n_up ? 0 : 1;
  const bool diag = (k == e);
  double* S_H0 = scr + CF::o2H0;
  double* S_A = scr + CF::o2A;
  double2* S_S1 = reinterpret_cast<double2*>(scr + CF::o2S1);
  double2* S_X = reinterpret_cast<double2*>(scr + CF::o2X);
  double* S_R = scr + CF::o2R;

  // ---- T0: the moved electron's local part as a jet (lane e only); its layer-0 feature tangents go to the scratch
  double xe[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) xe[c] = rec[DC::X + 3 * e + c];
  J h0e[4 * A], ye[6], enve, jaee;
#pragma unroll
  for (int q = 0; q < 4 * A; ++q) h0e[q] = Op::cst(0.0);
#pragma unroll
  for (int m = 0; m < 6; ++m) ye[m] = Op::cst(0.0);
  enve = Op::cst(0.0); jaee = Op::cst(0.0);
  __syncwarp();                                        // the previous electron's readers of the scratch are done
  if (diag && !idle) {
    J xj[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { xj[c] = Op::cst(xe[c]); xj[c].d[0] = (c == dir) ? 1.0 : 0.0; }
    PS::template electron_local<J>(P, e, xj, h0e, ye, enve, jaee);
#pragma unroll
    for (int q = 0; q < 4 * A; ++q) { S_H0[q] = h0e[q].d[0]; S_H0[4 * A + q] = h0e[q].s[0]; }
  }
  __syncwarp();
  double dJ = diag ? jaee.d[0] : 0.0, sJ = diag ? jaee.s[0] : 0.0;

  // ---- T1: level-0 tangents of the two pair chains through e that end in this lane: row (e,k): d = x_k - x_e,
  //      column (k,e): d = x_e - x_k
  Tan4 cr, cc;
#pragma unroll
  for (int c = 0; c < 4; ++c) { cr.d[c] = 0.0; cr.s[c] = 0.0; cc.d[c] = 0.0; cc.s[c] = 0.0; }
  if (!diag) {
    const double dd = rec[DC::X + 3 * k + dir] - xe[dir];
    const double r = rec[DC::HP + ((0 * N + e) * N + k) * 4];
    const double ri = s_inv(r);
    const double rt = -dd * ri;                                 // dr/dx_{e,dir}
    const double rs = (1.0 - rt * rt) * ri;                     // d2r/dx_{e,dir}^2
    cr.d[0] = rt; cc.d[0] = rt; cr.s[0] = rs; cc.s[0] = rs;
#pragma unroll
    for (int c = 0; c < 3; ++c) { cr.d[1 + c] = (c == dir) ? -1.0 : 0.0; cc.d[1 + c] = (c == dir) ? 1.0 : 0.0; }
    const int lo = e < k ? e : k, hi = e < k ? k : e;           // e-e Pade term (Jastrow.py:23-41)
    const double cu = P[L.jas_cusp + lo * N + hi], al = P[L.jas_alpha + lo * N + hi];
    const double q = s_inv(1.0 + al * r);
    const double u1 = cu * q * q, u2 = -2.0 * al * u1 * q;
    dJ += u1 * rt;
    sJ += u2 * rt * rt + u1 * rs;
  }

  // ---- the three one-electron layers of row k; the pair chains advance with them
  double hd[4] = {0.0, 0.0, 0.0, 0.0}, hs[4] = {0.0, 0.0, 0.0, 0.0};
  constexpr int GW = 4 * A > 4 ? 4 * A : 4;
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    // column block sums (only lane e uses them): every other lane deposits its column-chain tangent
    __syncwarp();
    if (!idle) {
#pragma unroll
      for (int c = 0; c < 4; ++c) { S_A[k * 8 + c] = cc.d[c]; S_A[k * 8 + 4 + c] = cc.s[c]; }
    }
    __syncwarp();
    double sud[2][4], sus[2][4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      double ud = 0.0, dd2 = 0.0, us = 0.0, ds = 0.0;
      if (diag) {
        for (int r = 0; r < N; ++r) {
          const double vd = S_A[r * 8 + c], vs = S_A[r * 8 + 4 + c];       // lane e deposited zeros
          if (r < n_up) { ud += vd; us += vs; } else { dd2 += vd; ds += vs; }
        }
      }
      sud[0][c] = ud; sud[1][c] = dd2; sus[0][c] = us; sus[1][c] = ds;
    }
    // tangents of the block means fed to every row
    double gmd[2][GW], gms[2][GW];
    if (l == 0) {
#pragma unroll
      for (int q = 0; q < 4 * A; ++q) {
        const double d0 = S_H0[q] * inv_n[se], s0 = S_H0[4 * A + q] * inv_n[se];
        gmd[0][q] = se == 0 ? d0 : 0.0; gmd[1][q] = se == 1 ? d0 : 0.0;
        gms[0][q] = se == 0 ? s0 : 0.0; gms[1][q] = se == 1 ? s0 : 0.0;
      }
    } else {
      __syncwarp();
      if (!idle) {
#pragma unroll
        for (int c = 0; c < 4; ++c) { S_A[k * 8 + c] = hd[c]; S_A[k * 8 + 4 + c] = hs[c]; }
      }
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        double ud = 0.0, dd2 = 0.0, us = 0.0, ds = 0.0;
        for (int r = 0; r < N; ++r) {
          const double vd = S_A[r * 8 + c], vs = S_A[r * 8 + 4 + c];
          if (r < n_up) { ud += vd; us += vs; } else { dd2 += vd; ds += vs; }
        }
        gmd[0][c] = ud * inv_n[0]; gmd[1][c] = dd2 * inv_n[1];
        gms[0][c] = us * inv_n[0]; gms[1][c] = ds * inv_n[1];
      }
    }
    if (l == 0) DS::template row_layer<true, 4 * A>(P, 0, k, e, se, inv_n, rec, 1, h0e, hd, hs, gmd, gms, cr, sud, sus);
    else DS::template row_layer<true, 4>(P, l, k, e, se, inv_n, rec, 1, h0e, hd, hs, gmd, gms, cr, sud, sus);
    if (l < 2 && !diag) {     // advance the tangents of the two pair chains through double-layer l
      const double* W = P + L.dbl_w[l];
      double tr[4], tc[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        tr[m] = kSqrt2 * rec[DC::HP + (((l + 1) * N + e) * N + k) * 4 + m] - rec[DC::HP + ((l * N + e) * N + k) * 4 + m];
        tc[m] = kSqrt2 * rec[DC::HP + (((l + 1) * N + k) * N + e) * 4 + m] - rec[DC::HP + ((l * N + k) * N + e) * 4 + m];
      }
      Tan4 nr, nc;
      DS::template chain_layer<true>(W, cr, tr, nr);
      DS::template chain_layer<true>(W, cc, tc, nc);
      cr = nr; cc = nc;
    }
  }

  // ---- determinant: row k of dM reads the tangent of h of electron sigma[k] (quirk Q4)
  __syncwarp();
  if (!idle) {
#pragma unroll
    for (int c = 0; c < 4; ++c) { S_A[k * 8 + c] = hd[c]; S_A[k * 8 + 4 + c] = hs[c]; }
  }
  __syncwarp();
  const int ek = sys.sigma[k];
  double dhd[4], dhs[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) { dhd[c] = S_A[ek * 8 + c]; dhs[c] = S_A[ek * 8 + 4 + c]; }
  double gsum = dJ, l2 = sJ;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const double gm = rec[DC::GMAT + (k * 4 + c) * 2];
    gsum += dhd[c] * gm;
    l2 += dhs[c] * gm;
  }
  // row e of E = env * Yo as jets (lane e): S1[l] = sum_j P[e,j] dE[e,j] Minv[j,l], S2 = sum_j (2 dP dE + P d2E) Minv[j,e]
  if (diag && !idle) {
    const int srow = e < sys.n_up_rows ? 0 : 1;
    const double* W = P + L.orb_w[srow];
    const double* Bv = P + L.orb_b[srow];
    const int eh = sys.sigma[e];
    double h3[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) h3[c] = rec[DC::H + (2 * N + eh) * 4 + c];
    cplx S1[N];
    cplx S2 = {0.0, 0.0};
#pragma unroll
    for (int l = 0; l < N; ++l) S1[l] = {0.0, 0.0};
#pragma unroll
    for (int j = 0; j < N; ++j) {
      J yo = Op::cst(0.0);
#pragma unroll
      for (int m = 0; m < 6; ++m) yo = yo + ye[m] * P[L.y_w + m * N + j];
      const J E = enve * yo;
      cplx p = {Bv[2 * j], Bv[2 * j + 1]}, dp = {0.0, 0.0};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double wr = W[c * 2 * N + 2 * j], wi = W[c * 2 * N + 2 * j + 1];
        p.re += h3[c] * wr; p.im += h3[c] * wi;
        dp.re += dhd[c] * wr; dp.im += dhd[c] * wi;          // lane e's dhd = tangent of h of electron sigma[e]
      }
      const cplx pdE = cscale(p, E.d[0]);
#pragma unroll
      for (int l = 0; l < N; ++l) {
        const cplx mjl = {rec[DC::MI + (j * N + l) * 2], rec[DC::MI + (j * N + l) * 2 + 1]};
        cfma(S1[l], pdE, mjl);
      }
      const cplx mje = {rec[DC::MI + (j * N + e) * 2], rec[DC::MI + (j * N + e) * 2 + 1]};
      const cplx t2 = cadd(cscale(dp, 2.0 * E.d[0]), cscale(p, E.s[0]));
      cfma(S2, t2, mje);
    }
#pragma unroll
    for (int l = 0; l < N; ++l) S_S1[l] = make_double2(S1[l].re, S1[l].im);
    l2 += S2.re;
  }
  __syncwarp();
  if (diag) gsum += S_S1[e].x;                               // S1[e].re
  // X[k][l] = sum_c dh[sigma_k,c] T[k,c,l] + delta_ke S1[l]; tr(X X) needs column k of X: exchange through the scratch
  cplx X[N];
#pragma unroll
  for (int l = 0; l < N; ++l) {
    cplx acc = {0.0, 0.0};
    if (diag) { const double2 s1 = S_S1[l]; acc = {s1.x, s1.y}; }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      acc.re += dhd[c] * rec[DC::TT + ((k * 4 + c) * N + l) * 2];
      acc.im += dhd[c] * rec[DC::TT + ((k * 4 + c) * N + l) * 2 + 1];
    }
    X[l] = acc;
    if (!idle) S_X[k * N + l] = make_double2(acc.re, acc.im);
  }
  __syncwarp();
  double trxx = 0.0;
#pragma unroll
  for (int l = 0; l < N; ++l) {
    const double2 xt = S_X[l * N + k];
    trxx += X[l].re * xt.x - X[l].im * xt.y;
  }
  // ---- fixed-order sum of the lanes' partials
  if (!idle) { S_R[2 * k] = gsum; S_R[2 * k + 1] = l2 - trxx; }
  __syncwarp();
  double g = 0.0, lp = 0.0;
  for (int r = 0; r < N; ++r) { g += S_R[2 * r]; lp += S_R[2 * r + 1]; }
  g_out = g;
  l_out = lp;
}

template <int NE, int NA>
__global__ void __launch_bounds__((CoopLapCfg<NE, NA>::T), (CoopLapCfg<NE, NA>::kMinBlocks)) k_lap_coop(
    AiqmcSystem sys, const double* __restrict__ params, const double* __restrict__ pos, int64_t n_cfg,
    double* __restrict__ mc_all, double* __restrict__ phase, double* __restrict__ logabs, double* __restrict__ gout,
    double* __restrict__ lap_parts, int64_t lap_stride) {
  using CF = CoopLapCfg<NE, NA>;
  using DC = DerivCache<NE, NA>;
  using MC = MoveCache<NE, NA>;
  using PS = Psi<NE, NA>;
  constexpr int N = NE, A = NA, GPW = CF::GPW, NG = CF::NG, QM = DC::QM, Q0 = 3 * NA + 2;
  constexpr LayoutC<NE, NA> L{};
  extern __shared__ __align__(16) double smem_cl[];
  const double* P = stage_params<NE, NA>(params, smem_cl);
  double* rec_buf = smem_cl + CF::kPar;                 // [1 or 2][NG][REC]
  double* scrs = rec_buf + (CF::kPipe ? 2 : 1) * NG * CF::REC;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_up = sys.n_up;
  const double inv_n[2] = {1.0 / sys.n_up, 1.0 / sys.n_dn};
  const int64_t tiles = (n_cfg + NG - 1) / NG;

  // =========================== phase 1: the producer warp, one lane per electron ===========================
  auto produce = [&](int64_t tile, double* recs) {
    const int64_t cfg0 = tile * NG;
    const int ncfg = (int)((n_cfg - cfg0) < NG ? (n_cfg - cfg0) : NG);
    {
      const bool idle = lane >= GPW * N;
      const int g = idle ? GPW - 1 : lane / N;
      const int k = idle ? N - 1 : lane - g * N;
      const bool act = !idle && g < ncfg;              // groups past the end of the batch redo the last configuration, silently
      const int gg = g < ncfg ? g : ncfg - 1;
      constexpr unsigned kLaneMask = GPW * N >= 32 ? 0xffffffffu : ((1u << ((GPW * N) % 32)) - 1u);
      unsigned gmask = (N >= 32 ? 0xffffffffu : ((1u << N) - 1u)) << (g * N);
      if (g == GPW - 1) gmask |= ~kLaneMask;                                       // idle lanes ride with the last group
      double* rec = recs + g * CF::REC;
      double* scr = scrs + g * CF::SCR;
      double2* MS = reinterpret_cast<double2*>(scr + CF::oMS);
      double2* PIV = reinterpret_cast<double2*>(scr + CF::oPIV);
      const int64_t cfg = cfg0 + gg;
      const int sig = sys.sigma[k];
      const int srow = k < sys.n_up_rows ? 0 : 1;

      double xk[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) xk[c] = pos[cfg * 3 * N + 3 * k + c];
      // ---- electron-local part
      double h0[4 * A], y[6], env, jae;
      PS::template electron_local<double>(P, k, xk, h0, y, env, jae);
      if (!idle) {
#pragma unroll
        for (int c = 0; c < 3; ++c) rec[DC::X + 3 * k + c] = xk[c];
#pragma unroll
        for (int q = 0; q < 4 * A; ++q) rec[DC::H0 + k * 4 * A + q] = h0[q];
#pragma unroll
        for (int m = 0; m < 6; ++m) rec[DC::YV + k * 6 + m] = y[m];
        rec[DC::ENVV + k] = env;
        rec[CF::oJAE + k] = jae;
      }
      __syncwarp();
      // ---- column k of the pair chains; block sums are lane-local
      double Gf[3][2][4];
#pragma unroll
      for (int l = 0; l < 3; ++l)
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
          for (int c = 0; c < 4; ++c) Gf[l][s][c] = 0.0;
      double jee = 0.0;
#pragma unroll 1
      for (int t = 0; t < N; ++t) {
        int i = k + t;
        i = i >= N ? i - N : i;
        double d[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) d[c] = xk[c] - rec[DC::X + 3 * i + c];       // pair (i, j = k): d = x_j - x_i
        double a0[4], a1[4], a2[4];
        PS::template pair_chain<double>(P, d, t == 0, a0, a1, a2);
        const bool iu = i < n_up;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          Gf[0][0][c] += iu ? a0[c] : 0.0; Gf[0][1][c] += iu ? 0.0 : a0[c];
          Gf[1][0][c] += iu ? a1[c] : 0.0; Gf[1][1][c] += iu ? 0.0 : a1[c];
          Gf[2][0][c] += iu ? a2[c] : 0.0; Gf[2][1][c] += iu ? 0.0 : a2[c];
        }
        if (!idle) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            rec[DC::HP + ((0 * N + i) * N + k) * 4 + c] = a0[c];
            rec[DC::HP + ((1 * N + i) * N + k) * 4 + c] = a1[c];
            rec[DC::HP + ((2 * N + i) * N + k) * 4 + c] = a2[c];
          }
        }
        if (t > 0) {                                                               // e-e Pade term (Jastrow.py:23-41)
          const int lo = i < k ? i : k, hi = i < k ? k : i;
          const double r = a0[0];
          jee += P[L.jas_cusp + lo * N + hi] * r * s_inv(1.0 + P[L.jas_alpha + lo * N + hi] * r);
        }
      }
      if (!idle) {
        rec[CF::oJEE + k] = jee;
#pragma unroll
        for (int l = 0; l < 3; ++l)
#pragma unroll
          for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int c = 0; c < 4; ++c) rec[DC::GS + ((l * 2 + s) * N + k) * 4 + c] = Gf[l][s][c];
      }
      // ---- block means of the layer-0 features (fixed order)
      double g0u[4 * A], g0d[4 * A];
#pragma unroll
      for (int q = 0; q < 4 * A; ++q) {
        double u = 0.0, dn = 0.0;
        for (int r = 0; r < N; ++r) { const double v = rec[DC::H0 + r * 4 * A + q]; if (r < n_up) u += v; else dn += v; }
        g0u[q] = u * inv_n[0];
        g0d[q] = dn * inv_n[1];
        if (!idle && k == 0) { rec[DC::G0M + q] = g0u[q]; rec[DC::G0M + 4 * A + q] = g0d[q]; }
      }
      // ---- the three one-electron layers of row k; layer outputs go straight into the record (they are also the
      //      exchange buffer of the block means)
      double hcur[4], hnext[4], gm[2][4];
      {
        double Gu[4], Gd[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) { Gu[c] = Gf[0][0][c] * inv_n[0]; Gd[c] = Gf[0][1][c] * inv_n[1]; }
        double t1[Q0];
        PS::template one_layer<4 * A, double>(P, 0, k, h0, g0u, g0d, Gu, Gd, hcur, t1, 1);
        if (!idle) {
#pragma unroll
          for (int q = 0; q < Q0; ++q) rec[DC::T1 + (0 * N + k) * QM + q] = t1[q];
        }
      }
#pragma unroll
      for (int l = 1; l < 3; ++l) {
        if (!idle) {
#pragma unroll
          for (int c = 0; c < 4; ++c) rec[DC::H + ((l - 1) * N + k) * 4 + c] = hcur[c];
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          double u = 0.0, dn = 0.0;
          for (int r = 0; r < N; ++r) { const double v = rec[DC::H + ((l - 1) * N + r) * 4 + c]; if (r < n_up) u += v; else dn += v; }
          gm[0][c] = u * inv_n[0];
          gm[1][c] = dn * inv_n[1];
          if (!idle && k == 0) { rec[DC::GM + ((l - 1) * 2 + 0) * 4 + c] = gm[0][c]; rec[DC::GM + ((l - 1) * 2 + 1) * 4 + c] = gm[1][c]; }
        }
        double Gu[4], Gd[4], t1[5];
#pragma unroll
        for (int c = 0; c < 4; ++c) { Gu[c] = Gf[l][0][c] * inv_n[0]; Gd[c] = Gf[l][1][c] * inv_n[1]; }
        PS::template one_layer<4, double>(P, l, k, hcur, gm[0], gm[1], Gu, Gd, hnext, t1, 1);
        if (!idle) {
#pragma unroll
          for (int q = 0; q < 5; ++q) rec[DC::T1 + (l * N + k) * QM + q] = t1[q];
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) hcur[c] = hnext[c];
      }
      if (!idle) {
#pragma unroll
        for (int c = 0; c < 4; ++c) rec[DC::H + (2 * N + k) * 4 + c] = hcur[c];
      }
      __syncwarp();
      // ---- orbital-matrix row k (quirk Q4: h of electron sigma[k], envelope / Ynlm of electron k)
      double hsg[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) hsg[c] = rec[DC::H + (2 * N + sig) * 4 + c];
      const double* Wo = P + L.orb_w[srow];
      const double* Bo = P + L.orb_b[srow];
      double evj[N];                                   // E[k][j] = env_k * (y_k . Yw[:, j])
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double pre = Bo[2 * j], pim = Bo[2 * j + 1];
#pragma unroll
        for (int c = 0; c < 4; ++c) { pre += hsg[c] * Wo[c * 2 * N + 2 * j]; pim += hsg[c] * Wo[c * 2 * N + 2 * j + 1]; }
        double yo = 0.0;
#pragma unroll
        for (int m = 0; m < 6; ++m) yo += y[m] * P[L.y_w + m * N + j];
        evj[j] = env * yo;
        if (!idle) MS[k * N + j] = make_double2(pre * evj[j], pim * evj[j]);
      }
      __syncwarp();
      // ---- A = M^T, Gauss-Jordan inverse with implicit row pivoting across the lanes (as coop_grad.cuh)
      double are[N], aim[N];
#pragma unroll
      for (int j = 0; j < N; ++j) { const double2 v = MS[j * N + k]; are[j] = v.x; aim[j] = v.y; }
      bool used = idle;
      unsigned unused = N >= 32 ? 0xffffffffu : ((1u << N) - 1u);
      int par = 0, ex = 0, mycol = 0, pbuf = 0;
      int piv_lane[N];
      cplx prod = {1.0, 0.0};
      StaticFor<0, N>::run([&](auto cc) {
        constexpr int c = decltype(cc)::value;
        const double m2 = are[c] * are[c] + aim[c] * aim[c];
        const unsigned key = used ? 0u : (((((unsigned)hi_word(m2)) >> 5) + 1u) << 5) | (31u - (unsigned)k);
        const unsigned kmax = __reduce_max_sync(gmask, key);
        const int best = 31 - (int)(kmax & 31u);
        piv_lane[c] = best;
        double2* pb = PIV + pbuf * N;
        pbuf ^= 1;
        const bool me = !idle && (k == best);
        if (me) {
#pragma unroll
          for (int j = 0; j < N; ++j) pb[j] = make_double2(are[j], aim[j]);
          used = true;
          mycol = c;
        }
        __syncwarp();
        const double2 pv2 = pb[c];
        const cplx pv = {pv2.x, pv2.y};
        par ^= __popc(unused & ((1u << best) - 1u));
        unused &= ~(1u << best);
        prod = cmul(prod, pv);
        {
          const double mag = fabs(prod.re) + fabs(prod.im);
          int e = ((hi_word(mag) >> 20) & 0x7ff) - 1023;
          e = e < -1000 ? -1000 : (e > 1000 ? 1000 : e);
          const double sc = make_double((1023 - e) << 20, 0);
          prod.re *= sc; prod.im *= sc;
          ex += e;
        }
        const double pn = s_inv(pv.re * pv.re + pv.im * pv.im);
        const cplx pinv = {pv.re * pn, -pv.im * pn};
        const cplx f = cmul(cplx{are[c], aim[c]}, pinv);
#pragma unroll
        for (int j = 0; j < N; ++j) {
          if (j != c) {
            const double2 pj = pb[j];
            const double nre = fma(f.im, pj.y, fma(-f.re, pj.x, are[j]));
            const double nim = fma(-f.im, pj.x, fma(-f.re, pj.y, aim[j]));
            const double qre = pj.x * pinv.re - pj.y * pinv.im, qim = pj.x * pinv.im + pj.y * pinv.re;
            are[j] = me ? qre : nre;
            aim[j] = me ? qim : nim;
          }
        }
        are[c] = me ? pinv.re : -f.re;
        aim[c] = me ? pinv.im : -f.im;
      });
      __syncwarp();
      if (!idle) {
#pragma unroll
        for (int c2 = 0; c2 < N; ++c2) MS[mycol * N + piv_lane[c2]] = make_double2(are[c2], aim[c2]);
      }
      __syncwarp();                                    // MS[l][j] = M^-1[j][l]
      // ---- M^-1 (column k), Gm[k][c] = sum_j w_kcj M^-1[j][k], T[k][c][l] = sum_j w_kcj M^-1[j][l]
      if (!idle) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
          const double2 v = MS[k * N + j];
          rec[DC::MI + (j * N + k) * 2] = v.x;
          rec[DC::MI + (j * N + k) * 2 + 1] = v.y;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll 1
          for (int l = 0; l < N; ++l) {
            cplx acc = {0.0, 0.0};
#pragma unroll
            for (int j = 0; j < N; ++j) {
              const cplx w = {Wo[c * 2 * N + 2 * j] * evj[j], Wo[c * 2 * N + 2 * j + 1] * evj[j]};
              const double2 mi = MS[l * N + j];
              cfma(acc, w, cplx{mi.x, mi.y});
            }
            rec[DC::TT + ((k * 4 + c) * N + l) * 2] = acc.re;
            rec[DC::TT + ((k * 4 + c) * N + l) * 2 + 1] = acc.im;
            if (l == k) { rec[DC::GMAT + (k * 4 + c) * 2] = acc.re; rec[DC::GMAT + (k * 4 + c) * 2 + 1] = acc.im; }
          }
        }
      }
      __syncwarp();
      // ---- log|psi|, phase, total Jastrow
      if (!idle && k == 0) {
        double jtot = 0.0;
        for (int r = 0; r < N; ++r) jtot += rec[CF::oJAE + r] + 0.5 * rec[CF::oJEE + r];
        const double ldet = 0.5 * log(prod.re * prod.re + prod.im * prod.im) + ex * 0.69314718055994530942;
        const double ph = atan2((par & 1) ? -prod.im : prod.im, (par & 1) ? -prod.re : prod.re);
        rec[CF::oMISC + 0] = jtot;
        rec[CF::oMISC + 1] = ldet + jtot;
        rec[CF::oMISC + 2] = ph;
        rec[CF::oMISC + 3] = 0.0;
        if (act) {
          if (phase) phase[cfg] = ph;
          logabs[cfg] = ldet + jtot;
        }
      }
    }
  };

  // =========================== phase 2: one thread per (configuration, electron, direction) ===========================
  auto consume = [&](int64_t tile, const double* recs) {
    const int64_t cfg0 = tile * NG;
    const int ncfg = (int)((n_cfg - cfg0) < NG ? (n_cfg - cfg0) : NG);
    const int tid = (int)threadIdx.x - (CF::kPipe ? 32 : 0);      // consumer index
    if constexpr (CF::kCoopTan) {
      // warp = Cartesian direction, lanes = (configuration g, row k) exactly as in phase 1
      const int dir = tid >> 5, ln = tid & 31;
      const bool idle = ln >= GPW * N;
      const int g = idle ? GPW - 1 : ln / N;
      const int k = idle ? N - 1 : ln - g * N;
      const int gg = g < ncfg ? g : ncfg - 1;           // groups past the end of the batch redo the last one, silently
      double* scr2 = scrs + NG * CF::SCR + (dir * NG + g) * CF::SCR2;
#pragma unroll 1
      for (int e = 0; e < N; ++e) {
        double gq, l2;
        lap_tangent_coop<NE, NA>(sys, P, recs + gg * CF::REC, scr2, k, idle, dir, gq, l2, e);
        if (!idle && g < ncfg && k == 0) {
          const int64_t cfg = cfg0 + g;
          const int ed = 3 * e + dir;
          gout[cfg * 3 * N + ed] = gq;
          lap_parts[(int64_t)ed * lap_stride + cfg] = l2;
        }
      }
    } else if (tid < ncfg * 3 * N) {
      const int cl = tid / (3 * N), ed = tid - cl * 3 * N;
      const int e = ed / 3, dir = ed - 3 * e;
      double gq, l2 = 0.0;
      DerivSplit<NE, NA>::template tangent<true>(sys, P, recs + cl * CF::REC, 1, e, dir, gq, l2);
      const int64_t cfg = cfg0 + cl;
      gout[cfg * 3 * N + ed] = gq;
      lap_parts[(int64_t)ed * lap_stride + cfg] = l2;
    }
    // ---- the single-electron-move cache of the quadrature kernels, coalesced out of the records
    if (mc_all) {
      constexpr int n1 = 12 * N * N + 24 * N + 4 * A * N + 8 * A;      // HP, GS, H0, G0M: same order in both layouts
      static_assert(MC::GS == MC::HP + 12 * N * N && DC::GS == DC::HP + 12 * N * N, "cache layouts diverged");
      for (int cl = 0; cl < ncfg; ++cl) {
        const double* rec = recs + cl * CF::REC;
        double* mc = mc_all + (cfg0 + cl) * MC::SIZE;
        for (int q = tid; q < n1; q += CF::T2) mc[MC::HP + q] = rec[DC::HP + q];
        for (int q = tid; q < 6 * N; q += CF::T2) mc[MC::Y + q] = rec[DC::YV + q];
        for (int q = tid; q < N; q += CF::T2) {
          mc[MC::ENV + q] = rec[DC::ENVV + q];
          mc[MC::JAE + q] = rec[CF::oJAE + q];
          mc[MC::JEE + q] = rec[CF::oJEE + q];
        }
        if (tid < 4) mc[MC::MISC + tid] = rec[CF::oMISC + tid];
      }
    }
  };

  // ---- this CTA's tiles are blockIdx.x, blockIdx.x + gridDim.x, ...
  const bool producer = warp == 0;
  if constexpr (!CF::kPipe) {          // two phases behind one barrier each: four resident CTAs per SM interleave them
#pragma unroll 1
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      __syncthreads();                 // the previous tile's records are no longer read
      if (producer) produce(tile, rec_buf);
      __syncthreads();
      consume(tile, rec_buf);
    }
    return;
  }
  if (producer && (int64_t)blockIdx.x < tiles) produce(blockIdx.x, rec_buf);
  __syncthreads();
  int buf = 0;
#pragma unroll 1
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t next = tile + gridDim.x;
    if (producer) {
      if (next < tiles) produce(next, rec_buf + (buf ^ 1) * NG * CF::REC);
    } else {
      consume(tile, rec_buf + buf * NG * CF::REC);
    }
    __syncthreads();
    buf ^= 1;
  }
}

}  // namespace aiqmc
