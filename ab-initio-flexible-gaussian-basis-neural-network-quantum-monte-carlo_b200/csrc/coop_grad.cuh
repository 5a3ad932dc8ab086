// coop_grad.cuh -- value + gradient of log|psi| by a fused forward + reverse (adjoint) sweep, B200 version 3 of the
// derivative path: ONE LANE PER ELECTRON, floor(32/N) configurations per warp, nothing in local memory.
//
// Version 2 (deriv_split.cuh: grad_reverse, one thread per configuration) needed 255 registers plus a 6-23 kB stack
// frame per thread: the tape of a configuration (all N^2 pair chains, the tanh outputs of every layer, the N x N
// complex inverse) lived in local memory and the kernel ran the FP64 pipe at ~6 %.  Here the work of a
// configuration is split across the N lanes of a group exactly as the network factorises:
//   lane k  owns  electron k's local part (features, Ynlm stream, envelope, e-n Jastrow),
//                 COLUMN k of the pair-chain matrix, h_two[i,k] for all i -- so the block sums G_l[s][k] = sum_{i in s}
//                 h_two^l[i,k] that feed row k are lane-local, no reduction,
//                 row k of the one-electron stream through the three layers,
//                 row k of the orbital matrix M, and after a transpose through shared memory row k of A = M^T for the
//                 Gauss-Jordan inverse, which leaves column k of M^-1 in the lane: what row k's adjoint needs.
//   cross-lane  = the per-spin block means of the one-electron stream (N x 4 values through a per-group scratch, summed
//                 in a fixed order), the gather h[sigma[k]] (quirk Q4), the pivot row of each elimination step (one
//                 REDUX arg-max + a double-buffered scratch row, 1 __syncwarp per step), the scatter of the pair
//                 adjoints onto the OTHER electron of each pair (rotated pair order i = (k+t) mod N: every lane writes a
//                 different row of the accumulator in every step -- no atomics, fixed order).
// The pair-chain tape (h_two^1, h_two^2 per off-diagonal pair: 8 doubles) sits in shared memory [slot][thread] (conflict-free),
// everything else in registers with compile-time indices.  The mathematics is grad_reverse's (deriv_split.cuh):
// d log|det| = Re tr(M^-1 dM), layers walked backwards, chains walked backwards, electron-local parts contracted
// through a 3-direction jet.  Reference semantics: jax.grad(logabs_f) of VMC/VMCmcstep.py:41,79 (A10-A12).
#pragma once
#include "psi_core.cuh"
#include "deriv_split.cuh"

namespace aiqmc {

template <int NE, int NA>
struct CoopGradCfg {
  static constexpr int N = NE, A = NA;
  static constexpr int GPW = 32 / NE;                        // configurations per warp
#ifndef AIQMC_GRADCOOP_W
#define AIQMC_GRADCOOP_W 0                                   // 0 = pick per system
#endif
  static constexpr int QM = Psi<NE, NA>::QM;
  static constexpr int RW = 8 * NA > 8 ? 8 * NA : 8;         // doubles a lane deposits for a block reduction
  // per-group scratch (doubles)
  static constexpr int oX = 0;                               // [3N] positions
  // Two pairs of buffers are never live at the same time (every hand-over is separated by a __syncwarp) and share
  // their storage -- the scratch and the tape decide how many warps fit an SM (N2: 514 -> 306 doubles per group):
  //   RED (layer-0 block sums in the forward pass, the block adjoints in the reverse pass) with MS (orbital matrix ->
  //   inverse, in between);  HS (4-vector exchanges before and after the inverse) with PIV (pivot rows, during it).
  static constexpr int oRED = (oX + 3 * NE + 1) & ~1;        // [N][RW]   |  [N][N][2] matrix transpose / inverse exchange
  static constexpr int oMS = oRED;
  static constexpr int kRedMs = NE * RW > 2 * NE * NE ? NE * RW : 2 * NE * NE;
  static constexpr int oHS = oRED + kRedMs;                  // [N][4]   4-vector exchange (h levels, h_bar)  |  [2][N][2] pivot rows
  static constexpr int oPIV = oHS;
  static constexpr int oGACC = oHS + 4 * NE;                 // [N][3]   gradient contributions to the other electron
  // group stride: even (double2 rows stay 16-byte aligned) and = 2 mod 16 doubles, so that the <= 8 groups of a warp
  // start 4 banks apart: the scratch accesses of different groups never collide (a stride of 120 doubles put groups
  // 0,2,4,6 on the same banks: 20 M conflicts per sweep in the first ncu capture)
  static constexpr int kScrRaw = (oGACC + 3 * NE + 1) & ~1;
  static constexpr int SCR = kScrRaw + ((2 - kScrRaw % 16) + 16) % 16;
  // per thread: [N-1 off-diagonal pairs][h1[4], h2[4] (, r)].  Where shared memory decides the number of resident warps
  // (N > 6) the pair distance is recomputed in the reverse pass instead of taped (carbon: taped 0.67 ms, recomputed 0.69)
  static constexpr bool kTapeR = NE <= 6;
  static constexpr int TSLOT = kTapeR ? 9 : 8;
  static constexpr int TAPE = TSLOT * (NE - 1);
  static constexpr int kPar = (make_layout(NE, NA).total + 1) & ~1;
  // warps per CTA.  N <= 6: 4 (two or more CTAs per SM).  Beyond that the pair tape (8 (N-1) doubles per thread) lets
  // only ONE CTA fit an SM, so the CTA takes as many warps as shared memory (and 254 registers x 256 threads) hold.
  // N2 history: 4 warps (156 kB) left the SM at 4 resident warps, FP64 pipe 21 %, half the stall samples on instruction
  // fetch: sweep 16.4 ms; 6 warps (227 kB): 11.1 ms; aliased scratch + slimmer tape: 8 warps.
  static constexpr int kWarpDoubles = GPW * SCR + 32 * TAPE;
  static constexpr int fit_warps(int w) { return w <= 2 ? 2 : ((kPar + w * kWarpDoubles) * 8 <= 227 * 1024 ? w : fit_warps(w - 1)); }
  static constexpr int W = AIQMC_GRADCOOP_W > 0 ? AIQMC_GRADCOOP_W : (NE <= 6 ? 4 : fit_warps(8));
  static constexpr int T = 32 * W;
  static constexpr int NG = W * GPW;                         // configurations per CTA pass
  static constexpr int kDoubles = kPar + NG * SCR + T * TAPE;
  static constexpr int kBytes = kDoubles * 8;
};

// adjoint of one-electron layer l for ROW k only (layer_reverse of deriv_split.cuh restricted to one row)
template <int NE, int NA, int DIN>
__device__ __forceinline__ void coop_row_reverse(const double* __restrict__ P, int l, int k, const double hn[4],
                                                 const double* __restrict__ hprev, const double* __restrict__ t1,
                                                 const double h_bar[4], const double inv_n[2], double own[DIN],
                                                 double gup_bar[DIN], double gdn_bar[DIN], double Gb0[4], double Gb1[4]) {
  constexpr LayoutC<NE, NA> L{};
  constexpr int DTOT = 3 * DIN + 8, Q = DTOT / 4;
  constexpr double kSqrt2 = 1.41421356237309504880;
  const double* sw = P + L.sing_w[l];
  const double* cw = P + L.conv_w[l] + k * DTOT;
  double zb[4];
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    double t, ob;
    if (DIN == 4) { t = kSqrt2 * hn[m] - hprev[m]; ob = h_bar[m] * kInvSqrt2; }     // residual layer (quirk Q5)
    else { t = hn[m]; ob = h_bar[m]; }
    zb[m] = ob * (1.0 - t * t);
    if (DIN == 4) own[m] = ob;
  }
  if (DIN != 4) {
#pragma unroll
    for (int q = 0; q < DIN; ++q) own[q] = 0.0;
  }
#pragma unroll
  for (int q = 0; q < DIN; ++q) { gup_bar[q] = 0.0; gdn_bar[q] = 0.0; }
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    double ob = 0.0;
#pragma unroll
    for (int m = 0; m < 4; ++m) ob += zb[m] * sw[q * 4 + m];
    const double t = t1[q];
    const double pb = ob * (1.0 - t * t) * 0.25;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int idx = 4 * q + c;
      const double xb = pb * cw[idx];
      if (idx < DIN) own[idx] += xb;
      else if (idx < 2 * DIN) gup_bar[idx - DIN] += xb;
      else if (idx < 3 * DIN) gdn_bar[idx - 2 * DIN] += xb;
      else if (idx < 3 * DIN + 4) Gb0[idx - 3 * DIN] = xb * inv_n[0];
      else Gb1[idx - 3 * DIN - 4] = xb * inv_n[1];
    }
  }
}

// SRC / OUT as k_grad_reverse (engine_impl.cuh): SRC == 1: configuration t = (walker b, moved electron i) built from
// x1, the drift at x1 and gauss1; OUT == 1: only electron i's components are written, gnew (n_cfg,3).
// Always: per-CTA partial of sum g^2 over ALL 3N components -> partials[blockIdx.x * 4 + pcol] (quirk Q6).
#ifndef AIQMC_GRADCOOP_MINB
#define AIQMC_GRADCOOP_MINB 1      // 3 (168 registers, ~100 B of spills, 12 warps/SM) measured: sweep 0.678 ms either way
#endif
template <int NE, int NA, int SRC, int OUT>
__global__ void __launch_bounds__((CoopGradCfg<NE, NA>::T), (NE <= 6 ? AIQMC_GRADCOOP_MINB : 1)) k_grad_coop(
    AiqmcSystem sys, const double* __restrict__ params, const double* __restrict__ pos, int64_t n_cfg, MovedSrc ms,
    double* __restrict__ phase, double* __restrict__ logabs, double* __restrict__ gout, double* __restrict__ partials,
    int pcol) {
  using CF = CoopGradCfg<NE, NA>;
  using PS = Psi<NE, NA>;
  constexpr int N = NE, A = NA, GPW = CF::GPW, NG = CF::NG, T = CF::T, Q0 = 3 * NA + 2;
  constexpr LayoutC<NE, NA> L{};
  constexpr unsigned kLaneMask = GPW * N >= 32 ? 0xffffffffu : ((1u << ((GPW * N) % 32)) - 1u);      // lanes that own an electron
  extern __shared__ __align__(16) double smem_cg[];
  __shared__ double red[CF::W];
  const double* P = stage_params<NE, NA>(params, smem_cg);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool idle = lane >= GPW * N;
  const int g = idle ? GPW - 1 : lane / N;
  const int k = idle ? N - 1 : lane - g * N;
  const bool act = !idle;
  unsigned gmask = (N >= 32 ? 0xffffffffu : ((1u << N) - 1u)) << (g * N);
  if (g == GPW - 1) gmask |= ~kLaneMask;      // idle lanes ride with the last group
  double* scr = smem_cg + CF::kPar + (warp * GPW + g) * CF::SCR;
  double* X = scr + CF::oX;
  double* RED = scr + CF::oRED;
  double* HS = scr + CF::oHS;
  double* GACC = scr + CF::oGACC;
  double2* MS = reinterpret_cast<double2*>(scr + CF::oMS);
  double2* PIV = reinterpret_cast<double2*>(scr + CF::oPIV);
  double* tape = smem_cg + CF::kPar + NG * CF::SCR + tid;                     // element (slot) at tape[slot * T]
  const int n_up = sys.n_up;
  const double inv_n[2] = {1.0 / sys.n_up, 1.0 / sys.n_dn};
  const bool k_up = k < n_up;
  const int sig = sys.sigma[k];
  const int srow = k < sys.n_up_rows ? 0 : 1;
  double g2 = 0.0;

  const int64_t tiles = (n_cfg + NG - 1) / NG;
#pragma unroll 1
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t cfg_raw = tile * NG + warp * GPW + g;
    const bool valid = cfg_raw < n_cfg;
    const int64_t cfg = valid ? cfg_raw : n_cfg - 1;
    // ---- this lane's electron position; the whole configuration goes to the group scratch
    double xk[3];
    int imov = 0;
    if (SRC == 0) {
#pragma unroll
      for (int c = 0; c < 3; ++c) xk[c] = pos[cfg * 3 * N + 3 * k + c];
    } else {
      const int64_t b = cfg / N;
      imov = (int)(cfg - b * N);
      const double te = taueff_of(ms.scal[0], ms.tau, ms.acyrus);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double x0 = pos[b * 3 * N + 3 * k + c];
        // g = grad_eff * tstep + gauss ; x2 = x1 + g on electron i only   (VMCmcstep.py:60-78)
        const double step = (ms.grad[b * 3 * N + 3 * k + c] * te) * ms.tau + ms.gauss1[b * 3 * N + 3 * k + c];
        const double xn = step + x0;
        xk[c] = (k == imov) ? xn : x0;
        if (k == imov && valid && act) ms.xprop[cfg * 3 + c] = xn;
      }
    }
    __syncwarp();                                    // previous tile's readers of the scratch are done
    if (act) {
#pragma unroll
      for (int c = 0; c < 3; ++c) { X[3 * k + c] = xk[c]; GACC[3 * k + c] = 0.0; }
    }

    // ---- F1: electron-local part
    double h0[4 * A], y[6], env, jae;
    PS::template electron_local<double>(P, k, xk, h0, y, env, jae);
    if (act) {
#pragma unroll
      for (int q = 0; q < 4 * A; ++q) RED[k * CF::RW + q] = h0[q];
    }
    __syncwarp();

    // ---- F2: column k of the pair chains, rotated order i = (k + t) mod N; block sums are lane-local
    double Gf[3][2][4];
#pragma unroll
    for (int l = 0; l < 3; ++l)
#pragma unroll
      for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int c = 0; c < 4; ++c) Gf[l][s][c] = 0.0;
    double jas = jae;
#pragma unroll 1
    for (int t = 0; t < N; ++t) {
      int i = k + t;
      i = i >= N ? i - N : i;
      double d[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) d[c] = xk[c] - X[3 * i + c];                 // pair (i, j = k): d = x_j - x_i
      double a0[4], a1[4], a2[4];
      PS::template pair_chain<double>(P, d, t == 0, a0, a1, a2);
      const bool iu = i < n_up;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        Gf[0][0][c] += iu ? a0[c] : 0.0; Gf[0][1][c] += iu ? 0.0 : a0[c];
        Gf[1][0][c] += iu ? a1[c] : 0.0; Gf[1][1][c] += iu ? 0.0 : a1[c];
        Gf[2][0][c] += iu ? a2[c] : 0.0; Gf[2][1][c] += iu ? 0.0 : a2[c];
      }
      if (t > 0) {                                                             // the diagonal chain carries no position adjoint
#pragma unroll
        for (int c = 0; c < 4; ++c) { tape[((t - 1) * CF::TSLOT + c) * T] = a1[c]; tape[((t - 1) * CF::TSLOT + 4 + c) * T] = a2[c]; }
        if (CF::kTapeR) tape[((t - 1) * CF::TSLOT + 8) * T] = a0[0];
      }
      if (t > 0 && i < k) {                                                    // e-e Pade term, pairs i < j (Jastrow.py:23-41)
        const double r = a0[0];
        jas += P[L.jas_cusp + i * N + k] * r * s_inv(1.0 + P[L.jas_alpha + i * N + k] * r);
      }
    }

    // ---- F3: block means of the layer-0 features (fixed order)
    double g0u[4 * A], g0d[4 * A];
#pragma unroll
    for (int q = 0; q < 4 * A; ++q) {
      double u = 0.0, dn = 0.0;
      for (int r = 0; r < N; ++r) { const double v = RED[r * CF::RW + q]; if (r < n_up) u += v; else dn += v; }
      g0u[q] = u * inv_n[0];
      g0d[q] = dn * inv_n[1];
    }

    // ---- F4: the three one-electron layers of row k
    double h1[4], h2[4], h3[4], t1a[Q0], t1b[5], t1c[5], gm1[2][4], gm2[2][4];
    {
      double Gu[4], Gd[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) { Gu[c] = Gf[0][0][c] * inv_n[0]; Gd[c] = Gf[0][1][c] * inv_n[1]; }
      PS::template one_layer<4 * A, double>(P, 0, k, h0, g0u, g0d, Gu, Gd, h1, t1a, 1);
    }
    auto block_means = [&](const double hv[4], double gm[2][4]) {
      __syncwarp();
      if (act) {
#pragma unroll
        for (int c = 0; c < 4; ++c) HS[k * 4 + c] = hv[c];
      }
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        double u = 0.0, dn = 0.0;
        for (int r = 0; r < N; ++r) { const double v = HS[r * 4 + c]; if (r < n_up) u += v; else dn += v; }
        gm[0][c] = u * inv_n[0];
        gm[1][c] = dn * inv_n[1];
      }
    };
    block_means(h1, gm1);
    {
      double Gu[4], Gd[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) { Gu[c] = Gf[1][0][c] * inv_n[0]; Gd[c] = Gf[1][1][c] * inv_n[1]; }
      PS::template one_layer<4, double>(P, 1, k, h1, gm1[0], gm1[1], Gu, Gd, h2, t1b, 1);
    }
    block_means(h2, gm2);
    {
      double Gu[4], Gd[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) { Gu[c] = Gf[2][0][c] * inv_n[0]; Gd[c] = Gf[2][1][c] * inv_n[1]; }
      PS::template one_layer<4, double>(P, 2, k, h2, gm2[0], gm2[1], Gu, Gd, h3, t1c, 1);
    }

    // ---- F5: orbital-matrix row k: h of electron sigma[k], envelope / Ynlm of electron k (quirk Q4)
    __syncwarp();
    if (act) {
#pragma unroll
      for (int c = 0; c < 4; ++c) HS[k * 4 + c] = h3[c];
    }
    __syncwarp();
    double hsg[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) hsg[c] = HS[sig * 4 + c];
    const double* Wo = P + L.orb_w[srow];
    const double* Bo = P + L.orb_b[srow];
    if (act) {
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double pre = Bo[2 * j], pim = Bo[2 * j + 1];
#pragma unroll
        for (int c = 0; c < 4; ++c) { pre += hsg[c] * Wo[c * 2 * N + 2 * j]; pim += hsg[c] * Wo[c * 2 * N + 2 * j + 1]; }
        double yo = 0.0;
#pragma unroll
        for (int m = 0; m < 6; ++m) yo += y[m] * P[L.y_w + m * N + j];
        const double ev = env * yo;
        MS[k * N + j] = make_double2(pre * ev, pim * ev);
      }
    }
    __syncwarp();

    // ---- F6: A = M^T, in-place Gauss-Jordan inverse with implicit row pivoting across the lanes
    double are[N], aim[N];
#pragma unroll
    for (int j = 0; j < N; ++j) { const double2 v = MS[j * N + k]; are[j] = v.x; aim[j] = v.y; }
    bool used = !act;
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
      const bool me = act && (k == best);
      if (me) {
#pragma unroll
        for (int j = 0; j < N; ++j) pb[j] = make_double2(are[j], aim[j]);
        used = true;
        mycol = c;
      }
      __syncwarp();
      const double2 pv2 = pb[c];
      const cplx pv = {pv2.x, pv2.y};
      par ^= __popc(unused & ((1u << best) - 1u));              // Lehmer code of the row permutation
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
      const cplx f = cmul(cplx{are[c], aim[c]}, pinv);          // multiplier of this row (unused on the pivot lane)
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
    // the lane that pivoted column c holds row c of A^-1 with element slot c2 <-> column piv_lane[c2]:
    // unscramble through shared memory; afterwards MS[k][j] = A^-1[k][j] = M^-1[j][k]
    __syncwarp();
    if (act) {
#pragma unroll
      for (int c2 = 0; c2 < N; ++c2) MS[mycol * N + piv_lane[c2]] = make_double2(are[c2], aim[c2]);
    }
    __syncwarp();
    const double ldet = 0.5 * log(prod.re * prod.re + prod.im * prod.im) + ex * 0.69314718055994530942;
    const double ph = atan2((par & 1) ? -prod.im : prod.im, (par & 1) ? -prod.re : prod.re);

    // ---- R1: adjoints from d log|det| = Re tr(M^-1 dM), row k
    double hb_sig[4] = {0.0, 0.0, 0.0, 0.0}, y_bar[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, env_bar = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double yo = 0.0;
#pragma unroll
      for (int m = 0; m < 6; ++m) yo += y[m] * P[L.y_w + m * N + j];
      double pre = Bo[2 * j], pim = Bo[2 * j + 1];
#pragma unroll
      for (int c = 0; c < 4; ++c) { pre += hsg[c] * Wo[c * 2 * N + 2 * j]; pim += hsg[c] * Wo[c * 2 * N + 2 * j + 1]; }
      const double2 mi = MS[k * N + j];
      const double ev = env * yo;
      const double pre_bar = mi.x * ev, pim_bar = -mi.y * ev;
      const double ev_bar = mi.x * pre - mi.y * pim;
#pragma unroll
      for (int c = 0; c < 4; ++c) hb_sig[c] += pre_bar * Wo[c * 2 * N + 2 * j] + pim_bar * Wo[c * 2 * N + 2 * j + 1];
      env_bar += ev_bar * yo;
      const double yo_bar = ev_bar * env;
#pragma unroll
      for (int m = 0; m < 6; ++m) y_bar[m] += yo_bar * P[L.y_w + m * N + j];
    }
    // total Jastrow and the adjoint of h3: both through the scratch (row sigma[k] of HS belongs to electron sigma[k])
    __syncwarp();
    if (act) {
#pragma unroll
      for (int c = 0; c < 4; ++c) HS[sig * 4 + c] = hb_sig[c];
      RED[k * CF::RW] = jas;
    }
    __syncwarp();
    double h_bar[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) h_bar[c] = HS[k * 4 + c];
    double jtot = 0.0;
    for (int r = 0; r < N; ++r) jtot += RED[r * CF::RW];
    const double la = ldet + jtot;

    // ---- R2: the one-electron layers, backwards; block-mean adjoints are summed over the rows through the scratch
    double Gb[3][2][4];
    double h0_bar[4 * A];
    auto spread = [&](auto dinc, const double* gup_bar, const double* gdn_bar, double* own) {
      constexpr int DIN = decltype(dinc)::value;
      __syncwarp();
      if (act) {
#pragma unroll
        for (int q = 0; q < DIN; ++q) { RED[k * CF::RW + q] = gup_bar[q]; RED[k * CF::RW + DIN + q] = gdn_bar[q]; }
      }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < DIN; ++q) {
        double s = 0.0;
        for (int r = 0; r < N; ++r) s += RED[r * CF::RW + (k_up ? 0 : DIN) + q];
        own[q] += s * (k_up ? inv_n[0] : inv_n[1]);
      }
    };
    {
      double own[4], gub[4], gdb[4];
      coop_row_reverse<NE, NA, 4>(P, 2, k, h3, h2, t1c, h_bar, inv_n, own, gub, gdb, Gb[2][0], Gb[2][1]);
      spread(std::integral_constant<int, 4>{}, gub, gdb, own);
#pragma unroll
      for (int c = 0; c < 4; ++c) h_bar[c] = own[c];
    }
    {
      double own[4], gub[4], gdb[4];
      coop_row_reverse<NE, NA, 4>(P, 1, k, h2, h1, t1b, h_bar, inv_n, own, gub, gdb, Gb[1][0], Gb[1][1]);
      spread(std::integral_constant<int, 4>{}, gub, gdb, own);
#pragma unroll
      for (int c = 0; c < 4; ++c) h_bar[c] = own[c];
    }
    {
      double gub[4 * A], gdb[4 * A];
      coop_row_reverse<NE, NA, 4 * A>(P, 0, k, h1, h0, t1a, h_bar, inv_n, h0_bar, gub, gdb, Gb[0][0], Gb[0][1]);
      spread(std::integral_constant<int, 4 * A>{}, gub, gdb, h0_bar);
    }

    // ---- R3: the pair chains of column k, backwards; d = x_k - x_i: +db on this electron, -db on electron i
    double gk[3] = {0.0, 0.0, 0.0};
    __syncwarp();
#pragma unroll 1
    for (int t = 1; t < N; ++t) {
      int i = k + t;
      i = i >= N ? i - N : i;
      const bool iu = i < n_up;
      double a0[4], a1[4], a2[4], ob2[4], ex1[4], ex0[4];
#pragma unroll
      for (int c = 0; c < 3; ++c) a0[1 + c] = xk[c] - X[3 * i + c];
      a0[0] = CF::kTapeR ? tape[((t - 1) * CF::TSLOT + 8) * T]
                         : s_sqrt(a0[1] * a0[1] + a0[2] * a0[2] + a0[3] * a0[3]);   // as pair_chain computed it
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        a1[c] = tape[((t - 1) * CF::TSLOT + c) * T];
        a2[c] = tape[((t - 1) * CF::TSLOT + 4 + c) * T];
        ob2[c] = iu ? Gb[2][0][c] : Gb[2][1][c];
        ex1[c] = iu ? Gb[1][0][c] : Gb[1][1][c];
        ex0[c] = iu ? Gb[0][0][c] : Gb[0][1][c];
      }
      double b1[4], b0[4];
      DerivSplit<NE, NA>::chain_reverse(P + L.dbl_w[1], a1, a2, ob2, ex1, b1);
      DerivSplit<NE, NA>::chain_reverse(P + L.dbl_w[0], a0, a1, b1, ex0, b0);
      const double r = a0[0];
      double r_bar = b0[0];
      if (i < k) {                                                             // e-e Pade term (Jastrow.py:23-41)
        const double q = s_inv(1.0 + P[L.jas_alpha + i * N + k] * r);
        r_bar += P[L.jas_cusp + i * N + k] * q * q;
      }
      const double rs = r_bar * s_inv(r);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double db = b0[1 + c] + rs * a0[1 + c];
        gk[c] += db;
        if (act) GACC[3 * i + c] -= db;                                        // lanes hit distinct rows i = (k+t) mod N
      }
      __syncwarp();
    }

    // ---- R4: electron-local part through a 3-direction jet
    {
      using J = Jet<false, 3>;
      using Op = ScalarOps<J>;
      J xj[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) { xj[c] = Op::cst(xk[c]); xj[c].d[c] = 1.0; }
      J h0e[4 * A], ye[6], enve, jaee;
      PS::template electron_local<J>(P, k, xj, h0e, ye, enve, jaee);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double gg = jaee.d[c] + env_bar * enve.d[c];
#pragma unroll
        for (int q = 0; q < 4 * A; ++q) gg += h0_bar[q] * h0e[q].d[c];
#pragma unroll
        for (int m = 0; m < 6; ++m) gg += y_bar[m] * ye[m].d[c];
        gk[c] += gg + GACC[3 * k + c];
      }
    }

    // ---- outputs
    if (valid && act) {
      g2 += gk[0] * gk[0] + gk[1] * gk[1] + gk[2] * gk[2];
      if (OUT == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) gout[cfg * 3 * N + 3 * k + c] = gk[c];
      } else if (k == imov) {
#pragma unroll
        for (int c = 0; c < 3; ++c) gout[cfg * 3 + c] = gk[c];
      }
      if (k == 0) {
        if (phase) phase[cfg] = ph;
        logabs[cfg] = la;
      }
    }
  }
  if (partials) {
    const double s = block_sum<T>(g2, red);
    if (threadIdx.x == 0) partials[blockIdx.x * 4 + pcol] = s;
  }
}

}  // namespace aiqmc
