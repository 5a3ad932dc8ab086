// ecp_grp.cuh -- the ccECP non-local quadrature (pseudopotential.py:272-318, pp_energy_test.py:45-105) and the
// T-move amplitudes (DMC/Tmoves.py:88-112) for systems with 5 <= N <= 32 electrons (N2, C6H6, ...):
// B200 version 4, "packed lane-per-electron groups".
//
// k_ecp_coop (version 2) ran the FP64 pipe far below its peak on N2: every weight of every DFMA was an LDS from
// a shared-memory copy of the parameters, every lane recomputed the moved electron's own Ynlm / envelope part,
// the 48 block sums per point were shuffle butterflies (384 SHFL), the pivot rows of the LU travelled by
// shuffles, groups were padded to a power of two (10 of 16 lanes busy at N = 10) and tanh was evaluated one at
// a time.  This version keeps its single-electron-move caching but:
//   * a group of exactly N lanes owns one quadrature point and floor(32/N) groups share a warp (30 of 32 lanes
//     at N = 10 and at N = 30);
//   * electron-independent weights (single/double layers, y_w) live in CONSTANT memory with compile-time
//     offsets -- operands of the DFMA, not loads; per-electron weights (the grouped "conv") sit in shared
//     memory TRANSPOSED [input][electron] so a warp's loads are contiguous;
//   * the moved electron's local part (features, 18 tanh Ynlm stream, envelope, e-n Jastrow, cos(theta)) is
//     computed ONCE per point by a thread-per-point phase and staged in shared memory;
//   * block sums: lanes deposit their terms in a per-group scratch, 8-9 lanes each add one column, everyone reads
//     the totals back (2 __syncwarp per layer, no shuffles);
//   * LU with partial pivoting: rows in registers, arg-max by two REDUX.SYNC on the high word of |.|^2,
//     pivot row broadcast through a double-buffered scratch (1 __syncwarp per step), permutation parity from a
//     Lehmer code (popc), determinant carried as a complex product with integer exponent renormalisation (one
//     log and one atan2 per point);
//   * tanh in interleaved batches of up to 8 (fastmath.cuh, ACC = 1: the value-only 9-op variant).
// One CTA = one walker; the CTA walks the walker's electrons i, and for each the A*50 points in chunks.
// Per-point contributions are staged and summed in a fixed order (deterministic, no atomics).
#pragma once
#include <type_traits>
#include <utility>
#include "psi_core.cuh"

namespace aiqmc {

constexpr int kUniMax = 1024;
static __constant__ double c_uni[kUniMax];   // compact copy of the electron-independent weights (UniLayout)

// The packed layout stores, per layer l, [sing_w, sing_b, (dbl_w, dbl_b,) yn_w, yn_b] contiguously (make_layout):
// the compact constant copy is those three segments followed by y_w.
template <int NE, int NA>
struct UniLayout {
  static constexpr AiqmcLayout K = make_layout(NE, NA);
  static constexpr int seg_src(int l) { return K.sing_w[l]; }
  static constexpr int seg_len(int l) { return K.yn_b[l] + 6 - K.sing_w[l]; }
  static constexpr int seg_dst(int l) { return l == 0 ? 0 : (l == 1 ? seg_len(0) : seg_len(0) + seg_len(1)); }
  static constexpr int at(int l, int packed_off) { return seg_dst(l) + packed_off - K.sing_w[l]; }
  static constexpr int y_w = seg_dst(2) + seg_len(2);
  static constexpr int total = y_w + 6 * NE;
  static_assert(total <= kUniMax, "electron-independent weights exceed the constant-memory copy");
};

// Compile-time loop: f(integral_constant<int, J>) for J in [J0, J1).  The LU nest at N = 30 is beyond what
// `#pragma unroll` will flatten (nvcc kept the inner loops rolled -> the register rows were demoted to local memory).
template <int J0, int J1>
struct StaticFor {
  template <class F>
  static __device__ __forceinline__ void run(F&& f) {
    if constexpr (J0 < J1) {
      f(std::integral_constant<int, J0>{});
      StaticFor<J0 + 1, J1>::run(f);
    }
  }
};

#ifndef AIQMC_GRP_ALIGN
#define AIQMC_GRP_ALIGN -1     // 1: all warps of a CTA walk the point loop in lock step (shared instruction fetches: pays off for the
#endif                         // 70 kB loop body of N > 16); 0: free-running warps (N2: 133.9 vs 136.1 ms); -1: pick by N
#ifndef AIQMC_GRP_MINB
#define AIQMC_GRP_MINB 2
#endif
#ifndef AIQMC_GRP_FLAT
#define AIQMC_GRP_FLAT 1
#endif
#ifndef AIQMC_GRP_PIVW
#define AIQMC_GRP_PIVW 0       // 1: per-warp interleaved pivot buffers (see GrpCfg); A/B on N2: 132.3 ms either way
#endif
#ifndef AIQMC_GRP_WARPS
#define AIQMC_GRP_WARPS 0      // 0 = pick per system
#endif
constexpr int grp_pick_warps(int n, int a) {
  if (AIQMC_GRP_WARPS > 0) return AIQMC_GRP_WARPS;
  const int gpw = 32 / n, pw = a * AIQMC_NQUAD;
  const int wmax = 12;
  double beff = 0.0;
  for (int w = wmax; w >= 6; --w) {        // tail efficiency of the per-electron point loop
    const int ng = w * gpw, it = (pw + ng - 1) / ng;
    const double eff = (double)pw / (double)(it * ng);
    if (eff > beff) beff = eff;
  }
  for (int w = wmax; w >= 6; --w) {        // the largest CTA within 3 % of the best (N2: 12 warps measured 3 % faster than 7)
    const int ng = w * gpw, it = (pw + ng - 1) / ng;
    const double eff = (double)pw / (double)(it * ng);
    if (eff >= beff - 0.03) return w;
  }
  return 8;
}

template <int NE, int NA>
struct GrpCfg {
  static constexpr int N = NE, A = NA;
  static constexpr int GPW = 32 / NE;                      // groups (= points in flight) per warp
  static constexpr int PW = NA * AIQMC_NQUAD;              // points per moved electron
  static constexpr int W = grp_pick_warps(NE, NA);
  static constexpr int T = 32 * W;
  static constexpr int NG = W * GPW;
  static constexpr int LSTR = 4 * NA + 16;                 // per-point record: xn[3] cs v[4] env jae yn[6] h0n[4A]
  static constexpr int kLocalBudget = 8192;                // doubles of per-point records (64 kB)
  static constexpr int pc_raw() {
    int pc = PW;
    if (pc * LSTR > kLocalBudget) pc = (kLocalBudget / LSTR) / NG * NG;
    return pc < NG ? NG : pc;
  }
  // Flat mode (N <= 16): the point loop runs over ALL (electron, atom, point) triples of the walker in chunks of 4 NG
  // points, with row i of the pair-chain cache staged for every i at once.  Per electron the N2 loop had 100 points on
  // 36 groups (3 passes, the last one 78 % full, three CTA barriers per electron); flat it is 1000 points in 7 chunks
  // of 144 (99 % full).  Beyond N = 16 a row set would not fit next to the point records and PW is large anyway.
  static constexpr bool kFlat = (AIQMC_GRP_FLAT != 0) && NE <= 16;
  static constexpr int kRows = kFlat ? NE : 1;             // rows i of the cache staged at a time
  static constexpr int PC = kFlat ? 4 * NG : pc_raw();     // points per chunk
  static constexpr int ETOT = NE * PW;                     // points per walker
  static constexpr int oPIV = (9 * NE + 18 + 8 * NA + 1) & ~1;            // pivot-row buffers inside the group scratch (16 B aligned)
  // per-group scratch doubles; the stride between the groups of a warp continues the 9-double row stride of red_in
  // across groups (SCR = 9 N mod 16, kept even for the double2 pivot rows), so that the GPW * N lanes of a warp spread
  // evenly over the banks: with the unpadded 164 doubles at N = 10 groups 0 and 1 shared 6 of their 10 bank pairs
  // Pivot rows live in a per-WARP buffer [2][N][GS] of (re, im) with the warp's groups interleaved (GS = GPW rounded up
  // to a power of two): element j of all groups shares one 16*GS-byte segment of a 128-byte line, so the broadcast read
  // of a pivot element by the whole warp and the write of it by one lane per group are single wavefronts (per-group
  // buffers put the 3 groups of N2 in 3 different lines: 2 wavefronts per read, 3 per write; the elimination was 44 %
  // of the kernel's shared-memory wavefronts).  Measured: no effect on the run time (N2 quadrature 132.3 ms either way),
  // so the per-group buffers (less shared memory) stay the default.
  static constexpr bool kPivW = (AIQMC_GRP_PIVW != 0) && GPW > 1;
  static constexpr int GS = !kPivW ? 1 : (GPW <= 2 ? 2 : (GPW <= 4 ? 4 : 8));
  static constexpr int kPivWarp = kPivW ? 4 * NE * GS : 0;                  // doubles per warp
  static constexpr int kScrRaw = oPIV + (kPivW ? 0 : 4 * NE);
  static constexpr int kScrWant = ((9 * NE) % 16) & ~1;
  static constexpr int SCR = kScrRaw + ((kScrWant - kScrRaw % 16) + 16) % 16;
  // shared-memory carve-up (doubles)
  static constexpr int D0 = 12 * NA + 8, Q0 = 3 * NA + 2;
  static constexpr int oCW0 = 0, oCW1 = oCW0 + D0 * NE, oCW2 = oCW1 + 20 * NE;
  static constexpr int oCB0 = oCW2 + 20 * NE, oCB1 = oCB0 + Q0 * NE, oCB2 = oCB1 + 5 * NE;
  static constexpr int oOW = (oCB2 + 5 * NE + 1) & ~1;              // orb_w[0] orb_b[0] orb_w[1] orb_b[1]  (20 N)
  static constexpr int oGS = oOW + 20 * NE;                // [3][2][N][4]
  static constexpr int oH0T = oGS + 24 * NE;               // [4A][N]
  static constexpr int oG0M = oH0T + 4 * NA * NE;          // [8A]
  static constexpr int oY = oG0M + 8 * NA;                 // [N][6]
  static constexpr int oENV = oY + 6 * NE, oJAE = oENV + NE, oJEE = oJAE + NE, oMISC = oJEE + NE;
  static constexpr int oHP = oMISC + 4;                    // [kRows][3][4][N] row(s) i of the pair-chain cache
  static constexpr int oJA = oHP + 12 * NE * kRows, oJC = oJA + NE * kRows;     // [kRows][N] each
  static constexpr int oX = oJC + NE * kRows;              // [3N]
  static constexpr int oL = (oX + 3 * NE + 1) & ~1;        // [PC][LSTR]
  static constexpr int oACC = oL + PC * LSTR;              // [PC][2]
  static constexpr int oSCR = oACC + 2 * PC;               // [NG][SCR]
  static constexpr int oPIVW = (oSCR + NG * SCR + 15) & ~15;        // [W][2][N][GS][2], 128-byte aligned
  static constexpr int kDoubles = oPIVW + W * kPivWarp;
  static constexpr int kBytes = kDoubles * 8;
  static constexpr int cw(int l) { return l == 0 ? oCW0 : (l == 1 ? oCW1 : oCW2); }
  static constexpr int cb(int l) { return l == 0 ? oCB0 : (l == 1 ? oCB1 : oCB2); }
};

// One one-electron layer for the lane's electron (nn.py:280-311): grouped conv (mean of 4 products + bias) ->
// tanh -> linear(->4) -> tanh (-> residual).  Inputs come through `in(idx)` (idx is a compile-time constant
// after unrolling), conv weights / biases from shared memory at stride NE, the linear layer from constant memory.
template <int NE, int NA, int LYR, int DIN, class In>
__device__ __forceinline__ void grp_one_layer(const double* __restrict__ cw, const double* __restrict__ cb, In in,
                                              double hout[4]) {
  using U = UniLayout<NE, NA>;
  constexpr int DTOT = 3 * DIN + 8, Q = DTOT / 4, CH = 8;
  constexpr int SW = U::at(LYR, U::K.sing_w[LYR]), SB = U::at(LYR, U::K.sing_b[LYR]);
  double z[4];
#pragma unroll
  for (int m = 0; m < 4; ++m) z[m] = c_uni[SB + m];
#pragma unroll
  for (int q0 = 0; q0 < Q; q0 += CH) {
    double pre[CH], t[CH];
#pragma unroll
    for (int qq = 0; qq < CH; ++qq) {
      const int q = q0 + qq;
      pre[qq] = 0.0;
      if (q < Q) {
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c < 4; ++c) acc += in(4 * q + c) * cw[(4 * q + c) * NE];
        pre[qq] = acc * 0.25 + cb[q * NE];
      }
    }
    if (Q - q0 >= CH) tanhv<CH, kAcc>(pre, t);
    else tanhv<(Q % CH ? Q % CH : CH), kAcc>(pre, t);
#pragma unroll
    for (int qq = 0; qq < CH; ++qq) {
      const int q = q0 + qq;
      if (q < Q) {
#pragma unroll
        for (int m = 0; m < 4; ++m) z[m] += t[qq] * c_uni[SW + q * 4 + m];
      }
    }
  }
  if (DIN == 4) {                                                    // residual only if shapes match (Q5)
    double hin[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) hin[m] = in(m);
    tanh_res<4, kAcc>(z, hin, hout);
  } else {
    tanhv<4, kAcc>(z, hout);
  }
}


// ---- LU across the lanes of a group: the lane holds its row in registers with the CURRENT column in element 0.
// One step = pivot search (one REDUX on a key that packs the high word of |a|^2 with the lane index), pivot row
// broadcast through the double-buffered scratch, multiplier, update of the WD-1 trailing elements written one slot
// to the left (the shift is free), determinant bookkeeping.  lu_run unrolls the steps for N <= 16; beyond that
// it runs rolled segments of kLuSeg steps at a fixed width (17 % more FMAs at N = 30, but 1 k instead of 4.6 k
// instructions: the straight-line triangular nest alone was 70 kB of SASS and the kernel stalled on fetch).
struct LuState {
  bool used;
  unsigned unused;
  int par, ex, parity_buf;
  cplx prod;
};
constexpr int kLuSeg = 6;

template <int NE, int WD, int GS>
__device__ __forceinline__ void lu_step(LuState& st, double (&rre)[NE], double (&rim)[NE], double2* __restrict__ pivb,
                                        int k, unsigned mask) {
  const double m2 = rre[0] * rre[0] + rim[0] * rim[0];
  // [high word of m2 >> 5, + 1][31 - lane]: monotone in m2 (m2 >= 0), ties go to the lowest lane, 0 = finished row
  const unsigned key = st.used ? 0u : (((((unsigned)hi_word(m2)) >> 5) + 1u) << 5) | (31u - (unsigned)k);
  const unsigned kmax = __reduce_max_sync(mask, key);
  const unsigned best = 31u - (kmax & 31u);
  double2* pb = pivb + st.parity_buf * NE * GS;               // element j of this group at pb[j * GS]
  st.parity_buf ^= 1;
  if ((unsigned)k == best) {
    StaticFor<0, WD>::run([&](auto jc) { constexpr int j = decltype(jc)::value; pb[j * GS] = make_double2(rre[j], rim[j]); });
    st.used = true;
  }
  __syncwarp();
  const double2 pv2 = pb[0];
  const cplx pv = {pv2.x, pv2.y};
  st.par ^= __popc(st.unused & ((1u << best) - 1u));       // Lehmer code of the row permutation
  st.unused &= ~(1u << best);
  st.prod = cmul(st.prod, pv);
  {
    const double mag = fabs(st.prod.re) + fabs(st.prod.im);
    int e = ((hi_word(mag) >> 20) & 0x7ff) - 1023;
    e = e < -1000 ? -1000 : (e > 1000 ? 1000 : e);
    const double sc = make_double((1023 - e) << 20, 0);
    st.prod.re *= sc; st.prod.im *= sc;
    st.ex += e;
  }
  const double pn = s_inv(pv.re * pv.re + pv.im * pv.im);
  const cplx pinv = {pv.re * pn, -pv.im * pn};
  const cplx f = cmul(cplx{rre[0], rim[0]}, pinv);
  StaticFor<1, WD>::run([&](auto jc) {                      // finished rows compute garbage nobody reads
    constexpr int j = decltype(jc)::value;
    const double2 pj = pb[j * GS];
    rre[j - 1] = fma(f.im, pj.y, fma(-f.re, pj.x, rre[j]));        // 4 DFMA per complex update
    rim[j - 1] = fma(-f.im, pj.x, fma(-f.re, pj.y, rim[j]));
  });
  rre[WD - 1] = 0.0; rim[WD - 1] = 0.0;
}

template <int NE, int C0, int GS>
__device__ __forceinline__ void lu_run(LuState& st, double (&rre)[NE], double (&rim)[NE], double2* __restrict__ pivb,
                                       int k, unsigned mask) {
  if constexpr (C0 < NE) {
    if constexpr (NE <= 16) {
      lu_step<NE, NE - C0, GS>(st, rre, rim, pivb, k, mask);
      lu_run<NE, C0 + 1, GS>(st, rre, rim, pivb, k, mask);
    } else {
      constexpr int WD = NE - C0, ST = WD < kLuSeg ? WD : kLuSeg;
#pragma unroll 1
      for (int s = 0; s < ST; ++s) lu_step<NE, WD, GS>(st, rre, rim, pivb, k, mask);
      lu_run<NE, C0 + ST, GS>(st, rre, rim, pivb, k, mask);
    }
  }
}

template <int NE, int NA>
__global__ void __launch_bounds__((GrpCfg<NE, NA>::T), (NE <= 16 ? AIQMC_GRP_MINB : 1))
k_ecp_grp(AiqmcSystem sys, const double* __restrict__ params, const double* __restrict__ pos,
          const double* __restrict__ rot, int64_t B, const double* __restrict__ cache_all, EnergyWs w,
          double* __restrict__ tm_out, double tm_tau) {
  using CF = GrpCfg<NE, NA>;
  using MC = MoveCache<NE, NA>;
  using U = UniLayout<NE, NA>;
  constexpr int N = NE, A = NA, GPW = CF::GPW, NG = CF::NG, PW = CF::PW, PC = CF::PC, LSTR = CF::LSTR;
  constexpr LayoutC<NE, NA> L{};
  extern __shared__ __align__(16) double smem_grp[];
  double* const smem = smem_grp;
  const int tid = threadIdx.x;
  const int64_t b = blockIdx.x;
  const double* cache = cache_all + b * MC::SIZE;

  // ---- stage the walker-constant data
  for (int l = 0; l < 3; ++l) {
    const int dtot = l == 0 ? CF::D0 : 20, q = dtot / 4;
    for (int t = tid; t < dtot * N; t += CF::T) {         // conv_w[l][k][idx] -> [idx][k]
      const int kk = t / dtot, idx = t - kk * dtot;
      smem[CF::cw(l) + idx * N + kk] = params[L.conv_w[l] + t];
    }
    for (int t = tid; t < q * N; t += CF::T) {
      const int kk = t / q, qq = t - kk * q;
      smem[CF::cb(l) + qq * N + kk] = params[L.conv_b[l] + t];
    }
  }
  for (int t = tid; t < 20 * N; t += CF::T) smem[CF::oOW + t] = params[L.orb_w[0] + t];
  // per-lane rows are stored lane-contiguous ([component][electron]): lane k of a group reads element k, so a group's
  // access is one contiguous run of N doubles (the [electron][4] layout of the cache put lanes k and k+4 on the same
  // banks: 16 % of the kernel's excess shared-memory wavefronts in the round-2 ncu source page)
  for (int t = tid; t < 24 * N; t += CF::T) {             // GS[l][s][j][c] -> [l][s][c][j]
    const int ls = t / (4 * N), r = t - ls * 4 * N, j = r >> 2, c = r & 3;
    smem[CF::oGS + (ls * 4 + c) * N + j] = cache[MC::GS + t];
  }
  for (int t = tid; t < 4 * A * N; t += CF::T) {          // H0[k][q] -> [q][k]
    const int kk = t / (4 * A), q = t - kk * 4 * A;
    smem[CF::oH0T + q * N + kk] = cache[MC::H0 + t];
  }
  for (int t = tid; t < 8 * A + 9 * N + 4; t += CF::T) {   // G0M Y ENV JAE JEE MISC; Y[k][m] -> [m][k]
    const int ty = t - 8 * A;
    if (ty >= 0 && ty < 6 * N) smem[CF::oY + (ty % 6) * N + ty / 6] = cache[MC::G0M + t];
    else smem[CF::oG0M + t] = cache[MC::G0M + t];
  }
  for (int t = tid; t < 3 * N; t += CF::T) smem[CF::oX + t] = pos[b * 3 * N + t];
  if (tid < kExpTab) g_exp_tab[tid] = exp2((double)tid * (1.0 / kExpTab));
  if (kAcc == 1 && AIQMC_TANH_TAB64) fill_tanh_table(tid, (int)blockDim.x);

  // ---- lane roles
  const int lane = tid & 31, warp = tid >> 5;
  const bool idle = lane >= GPW * N;
  const int g = idle ? GPW - 1 : lane / N;                 // group within the warp
  const int k = idle ? N + (lane - GPW * N) : lane - g * N;
  const bool act = !idle;
  const int kk = act ? k : N - 1;
  unsigned gmask = (N >= 32 ? 0xffffffffu : ((1u << N) - 1u)) << (g * N);
  if (g == GPW - 1 && GPW * N < 32) gmask |= ~((GPW * N >= 32) ? 0xffffffffu : ((1u << (GPW * N)) - 1u));
  const int n_up = sys.n_up;
  const double inv_n[2] = {1.0 / sys.n_up, 1.0 / sys.n_dn};
  const int sig = sys.sigma[kk];
  const int srow = kk < sys.n_up_rows ? 0 : 1;
  double* scr = smem + CF::oSCR + (warp * GPW + g) * CF::SCR;
  double* red_in = scr;                                    // [N][9]
  double* red_out = scr + 9 * N;                           // [18]: up sums (9), down sums (9)
  double* g0 = red_out + 18;                               // [8A]
  double2* pivb = CF::kPivW ? reinterpret_cast<double2*>(smem + CF::oPIVW + warp * CF::kPivWarp) + g      // [2][N][GS] (re, im)
                            : reinterpret_cast<double2*>(scr + CF::oPIV);                               // [2][N]
  __syncthreads();
  const double xk[3] = {smem[CF::oX + 3 * kk], smem[CF::oX + 3 * kk + 1], smem[CF::oX + 3 * kk + 2]};
  const double den_r = smem[CF::oMISC + 1], den_i = smem[CF::oMISC + 2];
  const double den_inv = 1.0 / (den_r * den_r + den_i * den_i);
  double acc_re = 0.0, acc_im = 0.0;                       // warp 0: fixed-order accumulation of the contributions

  // stage row `i` of the pair-chain cache and the (i,k) Jastrow parameters into row slot `slot`
  auto stage_row = [&](int i, int slot) {
    for (int t = tid; t < 12 * N; t += CF::T) {           // HP[l][i][k][c] -> [l][c][k]
      const int l = t / (4 * N), r = t - l * 4 * N, kq = r >> 2, c = r & 3;
      smem[CF::oHP + slot * 12 * N + (l * 4 + c) * N + kq] = cache[MC::HP + (l * N + i) * N * 4 + r];
    }
    for (int t = tid; t < N; t += CF::T) {
      const int lo = i < t ? i : t, hi = i < t ? t : i;
      smem[CF::oJA + slot * N + t] = params[L.jas_alpha + lo * N + hi];
      smem[CF::oJC + slot * N + t] = (t == i) ? 0.0 : params[L.jas_cusp + lo * N + hi];
    }
  };
  static_assert(CF::kFlat, "k_ecp_grp is the flat point loop (N <= 16); k_ecp_grp_rows handles the rest");
  for (int i = 0; i < N; ++i) stage_row(i, i);            // published by the barrier after the first phase 0

  constexpr int kChunks = (CF::ETOT + PC - 1) / PC;
#pragma unroll 1
  for (int ch = 0; ch < kChunks; ++ch) {
    const int e0 = ch * PC;
    const int npt = (CF::ETOT - e0 < PC) ? CF::ETOT - e0 : PC;
    {
      // ---- phase 0: thread per point -- rotated point, cos(theta) (quirks Q13, Q14), electron i's local part
      for (int t = tid; t < npt; t += CF::T) {
        const int i = (e0 + t) / PW, ew = e0 + t - i * PW;
        const int a = ew / AIQMC_NQUAD, p = ew - a * AIQMC_NQUAD;
        double* Lp = smem + CF::oL + t * LSTR;
        double ae[3], xn[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) ae[c] = smem[CF::oX + 3 * i + c] - params[L.atoms + 3 * a + c];
        const double r = sqrt(ae[0] * ae[0] + ae[1] * ae[1] + ae[2] * ae[2]);
        double dot = 0.0;
#pragma unroll
        for (int l = 0; l < 3; ++l) {
          const double nh = c_ecp.quad_pts[p][0] * rot[b * 9 + l] + c_ecp.quad_pts[p][1] * rot[b * 9 + 3 + l] +
                            c_ecp.quad_pts[p][2] * rot[b * 9 + 6 + l];
          xn[l] = r * nh;
          dot += ae[l] * xn[l];
        }
        const double cs = dot / (r * (r * w.gnorm[4 * b + quad_group(p)]));
        double h0n[4 * A], yn[6], envn, jaen;
        Psi<NE, NA>::template electron_local<double, kAcc>(params, i, xn, h0n, yn, envn, jaen);
        const double* vl = w.vl + ((b * N + i) * A + a) * 4;
        Lp[0] = xn[0]; Lp[1] = xn[1]; Lp[2] = xn[2]; Lp[3] = cs;
        Lp[4] = vl[0]; Lp[5] = vl[1]; Lp[6] = vl[2]; Lp[7] = vl[3];
        Lp[8] = envn; Lp[9] = jaen;
#pragma unroll
        for (int m = 0; m < 6; ++m) Lp[10 + m] = yn[m];
#pragma unroll
        for (int q = 0; q < 4 * A; ++q) Lp[16 + q] = h0n[q];
      }
      __syncthreads();

      // ---- main phase: one group per point, lane = electron
#pragma unroll 1
      for (int it = 0;; ++it) {
        const int t0 = warp * GPW + it * NG;
        constexpr bool kAlign = AIQMC_GRP_ALIGN < 0 ? (NE > 16) : (AIQMC_GRP_ALIGN != 0);
        if constexpr (kAlign) {
          // all warps of the CTA walk the (large, straight-line) loop body together so that they share instruction
          // fetches; warps past the end of the chunk run a dummy pass instead of leaving early
          if (it * NG >= npt) break;                                      // CTA-uniform
          __syncthreads();
        } else {
          if (t0 >= npt) break;                                           // warp-uniform
        }
        const bool valid = t0 + g < npt;
        const int t = valid ? t0 + g : npt - 1;
        const int i = (e0 + t) / PW;                                      // the point's displaced electron (group-uniform)
        const int si = i < n_up ? 0 : 1;
        const bool diag = act && (k == i);
        const int row_off = i;                                            // row i of the pair-chain cache [3][4][N]
#define AQ_HP_I(idx) smem[CF::oHP + row_off * 12 * N + (idx)]
#define AQ_JA_I(idx) smem[CF::oJA + row_off * N + (idx)]
#define AQ_JC_I(idx) smem[CF::oJC + row_off * N + (idx)]
        const double* Lp = smem + CF::oL + t * LSTR;
        // level-0 pair features through i: row (i,k): d = x_k - x_i', column (k,i): -d; e-e Jastrow term
        double cr[4], cc[4], h[4];
        double ju;
        {
          double d[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) d[c] = diag ? 0.0 : xk[c] - Lp[c];
          const double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
          const double rik = diag ? 0.0 : r2 * s_rsqrt(diag ? 1.0 : r2);
          cr[0] = rik; cc[0] = rik;
#pragma unroll
          for (int c = 0; c < 3; ++c) { cr[1 + c] = d[c]; cc[1 + c] = -d[c]; }
          ju = AQ_JC_I(kk) * rik * s_inv(1.0 + AQ_JA_I(kk) * rik);     // 0 on the diagonal
        }
        double jee_tot = 0.0;
#pragma unroll
        for (int l = 0; l < 3; ++l) {
          // ---- deposit the terms of the block sums
          if (act) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              red_in[k * 9 + c] = diag ? AQ_HP_I((l * 4 + c) * N + k) : cc[c];   // G'_l[s][i] terms
              if (l > 0) red_in[k * 9 + 4 + c] = h[c];
            }
            if (l == 0) {
              red_in[k * 9 + 8] = ju;
              for (int q = k; q < 8 * A; q += N) {                       // block means of the layer-0 features
                const int s = q >= 4 * A, qq = q - s * 4 * A;
                const double dh = Lp[16 + qq] - smem[CF::oH0T + qq * N + i];
                g0[q] = smem[CF::oG0M + q] + (s == si ? dh * inv_n[s] : 0.0);
              }
            }
          }
          __syncwarp();
          if (act) {
            constexpr int kCols = 9;
            for (int col = k; col < (l == 0 ? kCols : 8); col += N) {
              if (l == 0 && col >= 4 && col < 8) continue;               // no h sums before the first layer
              double u = 0.0, dsum = 0.0;
#pragma unroll
              for (int q = 0; q < N; ++q) {                               // fixed order: deterministic
                const double v = red_in[q * 9 + col];
                u += q < n_up ? v : 0.0;
                dsum += q < n_up ? 0.0 : v;
              }
              red_out[col] = u;
              red_out[9 + col] = dsum;
            }
          }
          __syncwarp();
          double Gu[4], Gd[4], gm0[4], gm1[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            double gu = smem[CF::oGS + ((l * 2 + 0) * 4 + c) * N + kk], gd = smem[CF::oGS + ((l * 2 + 1) * 4 + c) * N + kk];
            const double delta = cr[c] - AQ_HP_I((l * 4 + c) * N + kk);
            if (si == 0) gu += delta; else gd += delta;
            Gu[c] = (diag ? red_out[c] : gu) * inv_n[0];
            Gd[c] = (diag ? red_out[9 + c] : gd) * inv_n[1];
            if (l > 0) { gm0[c] = red_out[4 + c] * inv_n[0]; gm1[c] = red_out[13 + c] * inv_n[1]; }
          }
          if (l == 0) jee_tot = red_out[8] + red_out[17];
          const double* cw = smem + CF::cw(l) + kk;
          const double* cb = smem + CF::cb(l) + kk;
          if (l == 0) {
            const double* hp = diag ? Lp + 16 : smem + CF::oH0T + kk;
            const int hstr = diag ? 1 : N;
            auto in0 = [&](int idx) -> double {
              return idx < 4 * A ? hp[idx * hstr] : idx < 12 * A ? g0[idx - 4 * A]
                     : idx < 12 * A + 4 ? Gu[idx - 12 * A] : Gd[idx - 12 * A - 4];
            };
            grp_one_layer<NE, NA, 0, 4 * A>(cw, cb, in0, h);
          } else {
            auto inl = [&](int idx) -> double {
              return idx < 4 ? h[idx] : idx < 8 ? gm0[idx - 4] : idx < 12 ? gm1[idx - 8] : idx < 16 ? Gu[idx - 12] : Gd[idx - 16];
            };
            double hn[4];
            if (l == 1) grp_one_layer<NE, NA, 1, 4>(cw, cb, inl, hn);
            else grp_one_layer<NE, NA, 2, 4>(cw, cb, inl, hn);
#pragma unroll
            for (int c = 0; c < 4; ++c) h[c] = hn[c];
          }
          if (l < 2) {   // advance both pair chains through double-layer l (nn.py:305-309)
            const int WO = l == 0 ? U::at(0, U::K.dbl_w[0]) : U::at(1, U::K.dbl_w[1]);
            const int BO = l == 0 ? U::at(0, U::K.dbl_b[0]) : U::at(1, U::K.dbl_b[1]);
            double z[8], tt[8];
#pragma unroll
            for (int m = 0; m < 4; ++m) { z[m] = c_uni[BO + m]; z[4 + m] = z[m]; }
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
              for (int m = 0; m < 4; ++m) { z[m] += cr[q] * c_uni[WO + q * 4 + m]; z[4 + m] += cc[q] * c_uni[WO + q * 4 + m]; }
#pragma unroll
            for (int m = 0; m < 4; ++m) { tt[m] = cr[m]; tt[4 + m] = cc[m]; }
            tanh_res<8, kAcc>(z, tt, tt);
#pragma unroll
            for (int m = 0; m < 4; ++m) { cr[m] = tt[m]; cc[m] = tt[4 + m]; }
          }
        }

        // ---- orbital-matrix row of lane k: reads h of electron sigma[k], envelope / Ynlm of electron k (quirk Q4)
        if (act) {
#pragma unroll
          for (int c = 0; c < 4; ++c) red_in[k * 9 + 4 + c] = h[c];
        }
        __syncwarp();
        double rre[N], rim[N];
        {
          double hs[4], yr[6];
#pragma unroll
          for (int c = 0; c < 4; ++c) hs[c] = red_in[sig * 9 + 4 + c];
#pragma unroll
          for (int m = 0; m < 6; ++m) yr[m] = diag ? Lp[10 + m] : smem[CF::oY + m * N + kk];
          const double envr = diag ? Lp[8] : smem[CF::oENV + kk];
          const double* Wt = smem + CF::oOW + srow * 10 * N;
          const double* Bv = Wt + 8 * N;
          StaticFor<0, N>::run([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            const double2* W2 = reinterpret_cast<const double2*>(Wt);        // (re, im) pairs: one LDS.128 each
            const double2 b2 = reinterpret_cast<const double2*>(Bv)[j];
            double pre = b2.x, pim = b2.y;
#pragma unroll
            for (int c = 0; c < 4; ++c) { const double2 w2 = W2[c * N + j]; pre += hs[c] * w2.x; pim += hs[c] * w2.y; }
            double yo = 0.0;
#pragma unroll
            for (int m = 0; m < 6; ++m) yo += yr[m] * c_uni[U::y_w + m * N + j];
            const double evv = envr * yo;
            rre[j] = pre * evv; rim[j] = pim * evv;
          });
        }

        // ---- complex LU across the group's lanes with partial pivoting (lu_step / lu_run above)
        LuState lu;
        lu.used = !act;
        lu.unused = N >= 32 ? 0xffffffffu : ((1u << N) - 1u);
        lu.par = 0; lu.ex = 0; lu.parity_buf = 0;
        lu.prod = {1.0, 0.0};
        lu_run<NE, 0, CF::GS>(lu, rre, rim, pivb, k, GPW == 1 ? 0xffffffffu : gmask);
        int par = lu.par, ex = lu.ex;
        cplx prod = lu.prod;
        if (valid && act && k == 0) {
          // leave the determinant (mantissa product, exponent, parity) and the Jastrow change in the point's record;
          // the log / atan2 / quadrature weight are applied by the thread-per-point epilogue below (one lane per
          // point instead of a whole warp walking through libm for GPW results)
          double* Lw = smem + CF::oL + t * LSTR;
          Lw[0] = (par & 1) ? -prod.re : prod.re;
          Lw[1] = (par & 1) ? -prod.im : prod.im;
          Lw[2] = (double)ex;
          Lw[10] = (jee_tot - smem[CF::oJEE + i]) + (Lp[9] - smem[CF::oJAE + i]);
        }
        __syncwarp();
      }
      __syncthreads();
      // ---- epilogue: thread per point -- log|det|, phase, the ratio of complex logs (quirk Q12), angular weights
      for (int t = tid; t < npt; t += CF::T) {
        const double* Lp = smem + CF::oL + t * LSTR;
        const double pr = Lp[0], pi = Lp[1];
        const double la = 0.5 * log(pr * pr + pi * pi) + Lp[2] * 0.69314718055994530942 + smem[CF::oMISC + 0] + Lp[10];
        const double pha = atan2(pi, pr);
        const int i = (e0 + t) / PW, ew = e0 + t - i * PW;
        const int a = ew / AIQMC_NQUAD, p = ew - a * AIQMC_NQUAD;
        const double wq = c_ecp.quad_wts[p] * den_inv;
        const double rr = (la * den_r + pha * den_i) * wq, ri = (pha * den_r - la * den_i) * wq;
        const double v0 = Lp[4], v1 = Lp[5], v2 = Lp[6], v3 = Lp[7], cs = Lp[3];
        const double k4 = 0.07957747154594767;   // 1/(4 pi)
        const double f = v0 * k4 + v1 * (3.0 * k4 * cs) + v2 * (2.5 * k4 * (3.0 * cs * cs - 1.0)) +
                         v3 * (3.5 * k4 * (5.0 * cs * cs * cs - 3.0 * cs));
        smem[CF::oACC + 2 * t] = f * rr;
        smem[CF::oACC + 2 * t + 1] = f * ri;
        if (tm_out) tmove_point_out(tm_out + (((b * N + i) * A + a) * AIQMC_NQUAD + p) * 4, v0, v1, v2, v3, cs, rr, ri, tm_tau);
      }
      __syncthreads();                                   // the records are rewritten by the next chunk's phase 0
      if (warp == 0 && !tm_out) {
#pragma unroll 1
        for (int q = lane; q < npt; q += 32) { acc_re += smem[CF::oACC + 2 * q]; acc_im += smem[CF::oACC + 2 * q + 1]; }
      }
    }
  }
  if (warp == 0 && !tm_out) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      acc_re += __shfl_xor_sync(0xffffffffu, acc_re, o);
      acc_im += __shfl_xor_sync(0xffffffffu, acc_im, o);
    }
    if (lane == 0) { w.epp[2 * b] = acc_re; w.epp[2 * b + 1] = acc_im; }
  }
}

// ---- N > 16 (C6H6): the per-electron form of the point loop -- electron i's cache row staged, its A * 50 points walked in
//      chunks, next electron.  The flat loop above is the same arithmetic; this copy is kept because the 70 kB loop body
//      of N = 30 is fetch-bound and lost 3.5 % (849 -> 880 ms per 2,368 walkers) when it was compiled through the
//      generalised loop nest, for no gain (600 points per electron already fill the 12 groups evenly).
template <int NE, int NA>
__global__ void __launch_bounds__((GrpCfg<NE, NA>::T), (NE <= 16 ? AIQMC_GRP_MINB : 1))
k_ecp_grp_rows(AiqmcSystem sys, const double* __restrict__ params, const double* __restrict__ pos,
          const double* __restrict__ rot, int64_t B, const double* __restrict__ cache_all, EnergyWs w,
          double* __restrict__ tm_out, double tm_tau) {
  using CF = GrpCfg<NE, NA>;
  using MC = MoveCache<NE, NA>;
  using U = UniLayout<NE, NA>;
  constexpr int N = NE, A = NA, GPW = CF::GPW, NG = CF::NG, PW = CF::PW, PC = CF::PC, LSTR = CF::LSTR;
  constexpr LayoutC<NE, NA> L{};
  extern __shared__ __align__(16) double smem_grp[];
  double* const smem = smem_grp;
  const int tid = threadIdx.x;
  const int64_t b = blockIdx.x;
  const double* cache = cache_all + b * MC::SIZE;

  // ---- stage the walker-constant data
  for (int l = 0; l < 3; ++l) {
    const int dtot = l == 0 ? CF::D0 : 20, q = dtot / 4;
    for (int t = tid; t < dtot * N; t += CF::T) {         // conv_w[l][k][idx] -> [idx][k]
      const int kk = t / dtot, idx = t - kk * dtot;
      smem[CF::cw(l) + idx * N + kk] = params[L.conv_w[l] + t];
    }
    for (int t = tid; t < q * N; t += CF::T) {
      const int kk = t / q, qq = t - kk * q;
      smem[CF::cb(l) + qq * N + kk] = params[L.conv_b[l] + t];
    }
  }
  for (int t = tid; t < 20 * N; t += CF::T) smem[CF::oOW + t] = params[L.orb_w[0] + t];
  // per-lane rows are stored lane-contiguous ([component][electron]): lane k of a group reads element k, so a group's
  // access is one contiguous run of N doubles (the [electron][4] layout of the cache put lanes k and k+4 on the same
  // banks: 16 % of the kernel's excess shared-memory wavefronts in the round-2 ncu source page)
  for (int t = tid; t < 24 * N; t += CF::T) {             // GS[l][s][j][c] -> [l][s][c][j]
    const int ls = t / (4 * N), r = t - ls * 4 * N, j = r >> 2, c = r & 3;
    smem[CF::oGS + (ls * 4 + c) * N + j] = cache[MC::GS + t];
  }
  for (int t = tid; t < 4 * A * N; t += CF::T) {          // H0[k][q] -> [q][k]
    const int kk = t / (4 * A), q = t - kk * 4 * A;
    smem[CF::oH0T + q * N + kk] = cache[MC::H0 + t];
  }
  for (int t = tid; t < 8 * A + 9 * N + 4; t += CF::T) {   // G0M Y ENV JAE JEE MISC; Y[k][m] -> [m][k]
    const int ty = t - 8 * A;
    if (ty >= 0 && ty < 6 * N) smem[CF::oY + (ty % 6) * N + ty / 6] = cache[MC::G0M + t];
    else smem[CF::oG0M + t] = cache[MC::G0M + t];
  }
  for (int t = tid; t < 3 * N; t += CF::T) smem[CF::oX + t] = pos[b * 3 * N + t];
  if (tid < kExpTab) g_exp_tab[tid] = exp2((double)tid * (1.0 / kExpTab));
  if (kAcc == 1 && AIQMC_TANH_TAB64) fill_tanh_table(tid, (int)blockDim.x);

  // ---- lane roles
  const int lane = tid & 31, warp = tid >> 5;
  const bool idle = lane >= GPW * N;
  const int g = idle ? GPW - 1 : lane / N;                 // group within the warp
  const int k = idle ? N + (lane - GPW * N) : lane - g * N;
  const bool act = !idle;
  const int kk = act ? k : N - 1;
  unsigned gmask = (N >= 32 ? 0xffffffffu : ((1u << N) - 1u)) << (g * N);
  if (g == GPW - 1 && GPW * N < 32) gmask |= ~((GPW * N >= 32) ? 0xffffffffu : ((1u << (GPW * N)) - 1u));
  const int n_up = sys.n_up;
  const double inv_n[2] = {1.0 / sys.n_up, 1.0 / sys.n_dn};
  const int sig = sys.sigma[kk];
  const int srow = kk < sys.n_up_rows ? 0 : 1;
  double* scr = smem + CF::oSCR + (warp * GPW + g) * CF::SCR;
  double* red_in = scr;                                    // [N][9]
  double* red_out = scr + 9 * N;                           // [18]: up sums (9), down sums (9)
  double* g0 = red_out + 18;                               // [8A]
  double2* pivb = CF::kPivW ? reinterpret_cast<double2*>(smem + CF::oPIVW + warp * CF::kPivWarp) + g      // [2][N][GS] (re, im)
                            : reinterpret_cast<double2*>(scr + CF::oPIV);                               // [2][N]
  __syncthreads();
  const double xk[3] = {smem[CF::oX + 3 * kk], smem[CF::oX + 3 * kk + 1], smem[CF::oX + 3 * kk + 2]};
  const double den_r = smem[CF::oMISC + 1], den_i = smem[CF::oMISC + 2];
  const double den_inv = 1.0 / (den_r * den_r + den_i * den_i);
  double acc_re = 0.0, acc_im = 0.0;                       // warp 0: fixed-order accumulation of the contributions

#pragma unroll 1
  for (int i = 0; i < N; ++i) {
    const int si = i < n_up ? 0 : 1;
    // row i of the pair-chain cache and the (i,k) Jastrow parameters
    for (int t = tid; t < 12 * N; t += CF::T) {           // HP[l][i][k][c] -> [l][c][k]
      const int l = t / (4 * N), r = t - l * 4 * N, kq = r >> 2, c = r & 3;
      smem[CF::oHP + (l * 4 + c) * N + kq] = cache[MC::HP + (l * N + i) * N * 4 + r];
    }
    for (int t = tid; t < N; t += CF::T) {
      const int lo = i < t ? i : t, hi = i < t ? t : i;
      smem[CF::oJA + t] = params[L.jas_alpha + lo * N + hi];
      smem[CF::oJC + t] = (t == i) ? 0.0 : params[L.jas_cusp + lo * N + hi];
    }
#pragma unroll 1
    for (int c0 = 0; c0 < PW; c0 += PC) {
      const int npt = (PW - c0 < PC) ? PW - c0 : PC;
      // ---- phase 0: thread per point -- rotated point, cos(theta) (quirks Q13, Q14), electron i's local part
      for (int t = tid; t < npt; t += CF::T) {
        const int e = c0 + t, a = e / AIQMC_NQUAD, p = e - a * AIQMC_NQUAD;
        double* Lp = smem + CF::oL + t * LSTR;
        double ae[3], xn[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) ae[c] = smem[CF::oX + 3 * i + c] - params[L.atoms + 3 * a + c];
        const double r = sqrt(ae[0] * ae[0] + ae[1] * ae[1] + ae[2] * ae[2]);
        double dot = 0.0;
#pragma unroll
        for (int l = 0; l < 3; ++l) {
          const double nh = c_ecp.quad_pts[p][0] * rot[b * 9 + l] + c_ecp.quad_pts[p][1] * rot[b * 9 + 3 + l] +
                            c_ecp.quad_pts[p][2] * rot[b * 9 + 6 + l];
          xn[l] = r * nh;
          dot += ae[l] * xn[l];
        }
        const double cs = dot / (r * (r * w.gnorm[4 * b + quad_group(p)]));
        double h0n[4 * A], yn[6], envn, jaen;
        Psi<NE, NA>::template electron_local<double, kAcc>(params, i, xn, h0n, yn, envn, jaen);
        const double* vl = w.vl + ((b * N + i) * A + a) * 4;
        Lp[0] = xn[0]; Lp[1] = xn[1]; Lp[2] = xn[2]; Lp[3] = cs;
        Lp[4] = vl[0]; Lp[5] = vl[1]; Lp[6] = vl[2]; Lp[7] = vl[3];
        Lp[8] = envn; Lp[9] = jaen;
#pragma unroll
        for (int m = 0; m < 6; ++m) Lp[10 + m] = yn[m];
#pragma unroll
        for (int q = 0; q < 4 * A; ++q) Lp[16 + q] = h0n[q];
      }
      __syncthreads();

      // ---- main phase: one group per point, lane = electron
      const bool diag = act && (k == i);
#pragma unroll 1
      for (int it = 0;; ++it) {
        const int t0 = warp * GPW + it * NG;
        constexpr bool kAlign = AIQMC_GRP_ALIGN < 0 ? (NE > 16) : (AIQMC_GRP_ALIGN != 0);
        if constexpr (kAlign) {
          // all warps of the CTA walk the (large, straight-line) loop body together so that they share instruction
          // fetches; warps past the end of the chunk run a dummy pass instead of leaving early
          if (it * NG >= npt) break;                                      // CTA-uniform
          __syncthreads();
        } else {
          if (t0 >= npt) break;                                           // warp-uniform
        }
        const bool valid = t0 + g < npt;
        const int t = valid ? t0 + g : npt - 1;
        const double* Lp = smem + CF::oL + t * LSTR;
        // level-0 pair features through i: row (i,k): d = x_k - x_i', column (k,i): -d; e-e Jastrow term
        double cr[4], cc[4], h[4];
        double ju;
        {
          double d[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) d[c] = diag ? 0.0 : xk[c] - Lp[c];
          const double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
          const double rik = diag ? 0.0 : r2 * s_rsqrt(diag ? 1.0 : r2);
          cr[0] = rik; cc[0] = rik;
#pragma unroll
          for (int c = 0; c < 3; ++c) { cr[1 + c] = d[c]; cc[1 + c] = -d[c]; }
          ju = smem[CF::oJC + kk] * rik * s_inv(1.0 + smem[CF::oJA + kk] * rik);     // 0 on the diagonal
        }
        double jee_tot = 0.0;
#pragma unroll
        for (int l = 0; l < 3; ++l) {
          // ---- deposit the terms of the block sums
          if (act) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              red_in[k * 9 + c] = diag ? smem[CF::oHP + (l * 4 + c) * N + k] : cc[c];   // G'_l[s][i] terms
              if (l > 0) red_in[k * 9 + 4 + c] = h[c];
            }
            if (l == 0) {
              red_in[k * 9 + 8] = ju;
              for (int q = k; q < 8 * A; q += N) {                       // block means of the layer-0 features
                const int s = q >= 4 * A, qq = q - s * 4 * A;
                const double dh = Lp[16 + qq] - smem[CF::oH0T + qq * N + i];
                g0[q] = smem[CF::oG0M + q] + (s == si ? dh * inv_n[s] : 0.0);
              }
            }
          }
          __syncwarp();
          if (act) {
            constexpr int kCols = 9;
            for (int col = k; col < (l == 0 ? kCols : 8); col += N) {
              if (l == 0 && col >= 4 && col < 8) continue;               // no h sums before the first layer
              double u = 0.0, dsum = 0.0;
#pragma unroll
              for (int q = 0; q < N; ++q) {                               // fixed order: deterministic
                const double v = red_in[q * 9 + col];
                u += q < n_up ? v : 0.0;
                dsum += q < n_up ? 0.0 : v;
              }
              red_out[col] = u;
              red_out[9 + col] = dsum;
            }
          }
          __syncwarp();
          double Gu[4], Gd[4], gm0[4], gm1[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            double gu = smem[CF::oGS + ((l * 2 + 0) * 4 + c) * N + kk], gd = smem[CF::oGS + ((l * 2 + 1) * 4 + c) * N + kk];
            const double delta = cr[c] - smem[CF::oHP + (l * 4 + c) * N + kk];
            if (si == 0) gu += delta; else gd += delta;
            Gu[c] = (diag ? red_out[c] : gu) * inv_n[0];
            Gd[c] = (diag ? red_out[9 + c] : gd) * inv_n[1];
            if (l > 0) { gm0[c] = red_out[4 + c] * inv_n[0]; gm1[c] = red_out[13 + c] * inv_n[1]; }
          }
          if (l == 0) jee_tot = red_out[8] + red_out[17];
          const double* cw = smem + CF::cw(l) + kk;
          const double* cb = smem + CF::cb(l) + kk;
          if (l == 0) {
            const double* hp = diag ? Lp + 16 : smem + CF::oH0T + kk;
            const int hstr = diag ? 1 : N;
            auto in0 = [&](int idx) -> double {
              return idx < 4 * A ? hp[idx * hstr] : idx < 12 * A ? g0[idx - 4 * A]
                     : idx < 12 * A + 4 ? Gu[idx - 12 * A] : Gd[idx - 12 * A - 4];
            };
            grp_one_layer<NE, NA, 0, 4 * A>(cw, cb, in0, h);
          } else {
            auto inl = [&](int idx) -> double {
              return idx < 4 ? h[idx] : idx < 8 ? gm0[idx - 4] : idx < 12 ? gm1[idx - 8] : idx < 16 ? Gu[idx - 12] : Gd[idx - 16];
            };
            double hn[4];
            if (l == 1) grp_one_layer<NE, NA, 1, 4>(cw, cb, inl, hn);
            else grp_one_layer<NE, NA, 2, 4>(cw, cb, inl, hn);
#pragma unroll
            for (int c = 0; c < 4; ++c) h[c] = hn[c];
          }
          if (l < 2) {   // advance both pair chains through double-layer l (nn.py:305-309)
            const int WO = l == 0 ? U::at(0, U::K.dbl_w[0]) : U::at(1, U::K.dbl_w[1]);
            const int BO = l == 0 ? U::at(0, U::K.dbl_b[0]) : U::at(1, U::K.dbl_b[1]);
            double z[8], tt[8];
#pragma unroll
            for (int m = 0; m < 4; ++m) { z[m] = c_uni[BO + m]; z[4 + m] = z[m]; }
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
              for (int m = 0; m < 4; ++m) { z[m] += cr[q] * c_uni[WO + q * 4 + m]; z[4 + m] += cc[q] * c_uni[WO + q * 4 + m]; }
#pragma unroll
            for (int m = 0; m < 4; ++m) { tt[m] = cr[m]; tt[4 + m] = cc[m]; }
            tanh_res<8, kAcc>(z, tt, tt);
#pragma unroll
            for (int m = 0; m < 4; ++m) { cr[m] = tt[m]; cc[m] = tt[4 + m]; }
          }
        }

        // ---- orbital-matrix row of lane k: reads h of electron sigma[k], envelope / Ynlm of electron k (quirk Q4)
        if (act) {
#pragma unroll
          for (int c = 0; c < 4; ++c) red_in[k * 9 + 4 + c] = h[c];
        }
        __syncwarp();
        double rre[N], rim[N];
        {
          double hs[4], yr[6];
#pragma unroll
          for (int c = 0; c < 4; ++c) hs[c] = red_in[sig * 9 + 4 + c];
#pragma unroll
          for (int m = 0; m < 6; ++m) yr[m] = diag ? Lp[10 + m] : smem[CF::oY + m * N + kk];
          const double envr = diag ? Lp[8] : smem[CF::oENV + kk];
          const double* Wt = smem + CF::oOW + srow * 10 * N;
          const double* Bv = Wt + 8 * N;
          StaticFor<0, N>::run([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            const double2* W2 = reinterpret_cast<const double2*>(Wt);        // (re, im) pairs: one LDS.128 each
            const double2 b2 = reinterpret_cast<const double2*>(Bv)[j];
            double pre = b2.x, pim = b2.y;
#pragma unroll
            for (int c = 0; c < 4; ++c) { const double2 w2 = W2[c * N + j]; pre += hs[c] * w2.x; pim += hs[c] * w2.y; }
            double yo = 0.0;
#pragma unroll
            for (int m = 0; m < 6; ++m) yo += yr[m] * c_uni[U::y_w + m * N + j];
            const double evv = envr * yo;
            rre[j] = pre * evv; rim[j] = pim * evv;
          });
        }

        // ---- complex LU across the group's lanes with partial pivoting (lu_step / lu_run above)
        LuState lu;
        lu.used = !act;
        lu.unused = N >= 32 ? 0xffffffffu : ((1u << N) - 1u);
        lu.par = 0; lu.ex = 0; lu.parity_buf = 0;
        lu.prod = {1.0, 0.0};
        lu_run<NE, 0, CF::GS>(lu, rre, rim, pivb, k, GPW == 1 ? 0xffffffffu : gmask);
        int par = lu.par, ex = lu.ex;
        cplx prod = lu.prod;
        if (valid && act && k == 0) {
          // leave the determinant (mantissa product, exponent, parity) and the Jastrow change in the point's record;
          // the log / atan2 / quadrature weight are applied by the thread-per-point epilogue below (one lane per
          // point instead of a whole warp walking through libm for GPW results)
          double* Lw = smem + CF::oL + t * LSTR;
          Lw[0] = (par & 1) ? -prod.re : prod.re;
          Lw[1] = (par & 1) ? -prod.im : prod.im;
          Lw[2] = (double)ex;
          Lw[10] = (jee_tot - smem[CF::oJEE + i]) + (Lp[9] - smem[CF::oJAE + i]);
        }
        __syncwarp();
      }
      __syncthreads();
      // ---- epilogue: thread per point -- log|det|, phase, the ratio of complex logs (quirk Q12), angular weights
      for (int t = tid; t < npt; t += CF::T) {
        const double* Lp = smem + CF::oL + t * LSTR;
        const double pr = Lp[0], pi = Lp[1];
        const double la = 0.5 * log(pr * pr + pi * pi) + Lp[2] * 0.69314718055994530942 + smem[CF::oMISC + 0] + Lp[10];
        const double pha = atan2(pi, pr);
        const int e = c0 + t, a = e / AIQMC_NQUAD, p = e - a * AIQMC_NQUAD;
        const double wq = c_ecp.quad_wts[p] * den_inv;
        const double rr = (la * den_r + pha * den_i) * wq, ri = (pha * den_r - la * den_i) * wq;
        const double v0 = Lp[4], v1 = Lp[5], v2 = Lp[6], v3 = Lp[7], cs = Lp[3];
        const double k4 = 0.07957747154594767;   // 1/(4 pi)
        const double f = v0 * k4 + v1 * (3.0 * k4 * cs) + v2 * (2.5 * k4 * (3.0 * cs * cs - 1.0)) +
                         v3 * (3.5 * k4 * (5.0 * cs * cs * cs - 3.0 * cs));
        smem[CF::oACC + 2 * t] = f * rr;
        smem[CF::oACC + 2 * t + 1] = f * ri;
        if (tm_out) tmove_point_out(tm_out + (((b * N + i) * A + a) * AIQMC_NQUAD + p) * 4, v0, v1, v2, v3, cs, rr, ri, tm_tau);
      }
      __syncthreads();                                   // the records are rewritten by the next chunk's phase 0
      if (warp == 0 && !tm_out)
        for (int q = lane; q < npt; q += 32) { acc_re += smem[CF::oACC + 2 * q]; acc_im += smem[CF::oACC + 2 * q + 1]; }
    }
  }
  if (warp == 0 && !tm_out) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      acc_re += __shfl_xor_sync(0xffffffffu, acc_re, o);
      acc_im += __shfl_xor_sync(0xffffffffu, acc_im, o);
    }
    if (lane == 0) { w.epp[2 * b] = acc_re; w.epp[2 * b + 1] = acc_im; }
  }
}

#undef AQ_HP_I
#undef AQ_JA_I
#undef AQ_JC_I
}  // namespace aiqmc
