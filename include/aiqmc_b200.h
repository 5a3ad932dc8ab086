/* aiqmc_b200.h -- C ABI of the B200-native AIQMC walker engine.
 *
 * The reference (Yongda1/...-Quantum-Monte-Carlo, AIQMCrelease3) has no native interface:
 * its boundary is a set of Python closures over jax.numpy.  Every entry
 * point below replaces one of those closures, is natively batched over walkers, takes raw
 * DEVICE pointers (float64 unless noted) + a cudaStream_t passed as void*, never allocates
 * behind the caller's back (workspace size queries + caller-provided workspace) and returns
 * an int status: 0 = ok, <0 = AIQMC_E_* below.  No torch / jax types appear here; the XLA
 * FFI shim and the torch/ctypes harness both sit on top of these symbols.
 *
 * Citations are file:line under /root/reference/AIQMCrelease3/.
 */
#ifndef AIQMC_B200_H_
#define AIQMC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AIQMC_MAX_ELEC 32
#define AIQMC_MAX_ATOMS 16
#define AIQMC_ECP_MAX_L 4    /* l = 0..3, P_l table of pseudopotential.py:250-269 */
#define AIQMC_ECP_MAX_K 4    /* gaussians per channel */
#define AIQMC_NQUAD 50       /* 6+12+8+24 octahedral points, pseudopotential.py:181-225 */

enum {
  AIQMC_OK = 0,
  AIQMC_E_UNSUPPORTED = -1,  /* (n_elec, n_atoms) pair has no compiled instantiation */
  AIQMC_E_BADARG = -2,
  AIQMC_E_CUDA = -3,         /* a CUDA runtime call failed; aiqmc_last_cuda_error() has the code */
  AIQMC_E_WORKSPACE = -4,    /* workspace too small */
  AIQMC_E_NCCL = -5          /* NCCL could not be bound at run time (libnccl.so.2), or an NCCL call failed */
};

/* Static description of the molecule + spin layout.  Mirrors the arguments of
 * make_ai_net (wavefunction_Ynlm/nn.py:511-526) and the index tables of spin_indices.py:5-46. */
typedef struct AiqmcSystem {
  int32_t n_elec;       /* N, nelectrons */
  int32_t n_atoms;      /* A, natoms */
  int32_t n_up;         /* nspins[0]: electrons [0,n_up) form the first symmetric-feature block */
  int32_t n_dn;         /* nspins[1] (nn.py:142-153) */
  int32_t n_up_rows;    /* len(spin_up_indices): rows [0,n_up_rows) use params['orbitals'][0] */
  int32_t sigma[AIQMC_MAX_ELEC]; /* orbital-matrix row k reads h_one of electron sigma[k]
                                    (spin_up_indices then spin_down_indices, nn.py:432-474) */
} AiqmcSystem;

/* ccECP tables, shapes as in example/single_atom_C/single_atom_C.py:13-23 (padded). */
typedef struct AiqmcEcp {
  int32_t k_loc;                 /* gaussians in the local channel */
  int32_t n_l;                   /* list_l + 1 angular channels actually summed */
  int32_t k_nl;                  /* gaussians per non-local channel */
  int32_t pad_;
  double rn_local[AIQMC_MAX_ATOMS][AIQMC_ECP_MAX_K];      /* Rn_local (used as r^(n-2), pseudopotential.py:95) */
  double local_coes[AIQMC_MAX_ATOMS][AIQMC_ECP_MAX_K];
  double local_exps[AIQMC_MAX_ATOMS][AIQMC_ECP_MAX_K];
  double rn_non_local[AIQMC_MAX_ATOMS][AIQMC_ECP_MAX_L][AIQMC_ECP_MAX_K];  /* used as r^n, :150 */
  double non_local_coes[AIQMC_MAX_ATOMS][AIQMC_ECP_MAX_L][AIQMC_ECP_MAX_K];
  double non_local_exps[AIQMC_MAX_ATOMS][AIQMC_ECP_MAX_L][AIQMC_ECP_MAX_K];
  double quad_pts[AIQMC_NQUAD][3];   /* unrotated points, order OA,OB,OC,OD */
  double quad_wts[AIQMC_NQUAD];      /* 4/315, 64/2835, 27/1280, 14641/725760 per group */
} AiqmcEcp;

/* ---- parameter packing --------------------------------------------------------------- */
/* Offsets (in doubles) of every leaf of the reference parameter pytree (nn.py:203-278,
 * 370-407) inside the single packed buffer the kernels read.  Filled by aiqmc_param_layout. */
typedef struct AiqmcLayout {
  int32_t conv_w[3], conv_b[3];      /* streams[l].convolutional  w (N,d_l)  b (N,d_l/4) */
  int32_t sing_w[3], sing_b[3];      /* streams[l].single         w (d_l/4,4) b (4)      */
  int32_t dbl_w[2], dbl_b[2];        /* streams[l].double         w (4,4) b (4), l=0,1   */
  int32_t yn_w[3], yn_b[3];          /* streams_y[l].single_Ynlm  w (k_l,6) b (6)        */
  int32_t orb_w[2], orb_b[2];        /* orbitals[s]               w (4,2N) b (2N)        */
  int32_t y_w;                       /* y[0].w (6,N), row-normalised as at nn.py:449-451  */
  int32_t jas_alpha, jas_cusp;       /* (N,N) i<j tables built from ee_par/ee_anti + spin_indices.py */
  int32_t jas_beta;                  /* jastrow_ae.ae (N,A) */
  int32_t jas_c34, jas_c14;          /* (2Z)^(3/4), (2Z)^(1/4) per atom (Jastrow.py:84) */
  int32_t env_pi, env_sx;            /* envelope[i].pi (A,3); sigma*xi (A,3) */
  int32_t env_alpha, env_beta;       /* envelope[i].alpha (1) ; beta (A) */
  int32_t atoms, charges;            /* (A,3), (A) */
  int32_t total;                     /* total number of doubles */
} AiqmcLayout;

int aiqmc_param_layout(int32_t n_elec, int32_t n_atoms, AiqmcLayout* out);
/* 1 if the kernels of (n_elec,n_atoms) can be bound, else 0.  Every system is its own shared object
 * (libaiqmc_sys_<N>_<A>.so beside this library, or in $AIQMC_PLUGIN_DIR), loaded on first use; any n_elec <= 32,
 * n_atoms <= 16 can be built (aiqmc_b200.build.ensure_system).  aiqmc_rescan_systems() forgets failed look-ups after
 * a plugin has been built at run time. */
int aiqmc_supported(int32_t n_elec, int32_t n_atoms);
void aiqmc_rescan_systems(void);
int aiqmc_last_cuda_error(void);
/* Number of CUDA kernels this library has launched so far in this process (all entry points). */
int64_t aiqmc_launch_count(void);
const char* aiqmc_version(void);

/* ---- wavefunction: replaces Network.apply == signed_network (nn.py:545-551) ------------ */
/* pos (n_cfg,3N) -> phase (n_cfg) [angle of the determinant], logabs (n_cfg) [log|psi|]. */
int aiqmc_psi_fwd(const AiqmcSystem* sys, const double* params, const double* pos, int64_t n_cfg,
                  double* phase, double* logabs, void* stream);
/* + grad (n_cfg,3N) = d log|psi| / d pos: replaces jax.grad(logabs_f, argnums=1)
 * (VMC/VMCmcstep.py:41, Energy/hamiltonian.py:104). */
/* Scratch for the two derivative entry points below (structure-of-arrays derivative cache of one chunk of
 * configurations; bounded by 1 GiB whatever n_cfg).  with_lap = 0 for aiqmc_psi_grad, 1 for aiqmc_psi_fwdlap. */
int64_t aiqmc_psi_workspace_bytes(const AiqmcSystem* sys, int64_t n_cfg, int32_t with_lap);
int aiqmc_psi_grad(const AiqmcSystem* sys, const double* params, const double* pos, int64_t n_cfg,
                   double* phase, double* logabs, double* grad, void* workspace, int64_t workspace_bytes,
                   void* stream);
/* + lap (n_cfg) = sum_i d^2 log|psi| / d pos_i^2 by forward Laplacian: replaces
 * jax.linearize(grad) + the fori_loop of 3N jvps (Energy/pphamiltonian.py:74-106). */
int aiqmc_psi_fwdlap(const AiqmcSystem* sys, const double* params, const double* pos, int64_t n_cfg,
                     double* phase, double* logabs, double* grad, double* lap, void* workspace,
                     int64_t workspace_bytes, void* stream);

/* ---- VMC: replaces walkers_update (VMC/VMCmcstep.py:28-111) ---------------------------- */
int64_t aiqmc_vmc_workspace_bytes(const AiqmcSystem* sys, int64_t n_walkers);
/* One drift-diffusion Metropolis sweep over the whole per-device batch, in place.
 *  gauss1 (B,3N), gauss2 (B,N,3N): sqrt(tstep)*N(0,1); rnd (B,N): U[0,1)  (parity mode inputs,
 *  the arrays the reference draws at VMCmcstep.py:58,83 and :19-20).
 *  accept (B,N) uint8 out (may be NULL); signed_ratio!=0 gives the DMC variant
 *  (DMC/drift_diffusion.py:87-89).  aux_out (may be NULL): [0]=sum(x_new), [1]=sum(x_proposed)
 *  for tdamp (drift_diffusion.py:21), [2]=v2 of grad(x1), [3]=v2 of grad(x2) (limdrift sums, quirk Q6).
 *  grad_eff_old (B,3N) out may be NULL (drift_diffusion.py:45). */
int aiqmc_vmc_sweep(const AiqmcSystem* sys, const double* params, double* pos, const double* gauss1,
                    const double* gauss2, const double* rnd, int64_t n_walkers, double tstep,
                    double acyrus, int32_t signed_ratio, uint8_t* accept, double* grad_eff_old,
                    double* aux_out, void* workspace, int64_t workspace_bytes, void* stream);

/* Same sweep with the second Gaussian array in COMPACT form gauss2c (B,N,3) = gauss2[b,i,3i:3i+3]: the only entries
 * of the reference's (B,N,3N) draw that walkers_update reads (VMCmcstep.py:86-94 indexes [b,i,i,d]).  12N instead of
 * 12N^2 bytes per walker cross the boundary. */
int aiqmc_vmc_sweep_compact(const AiqmcSystem* sys, const double* params, double* pos, const double* gauss1,
                            const double* gauss2c, const double* rnd, int64_t n_walkers, double tstep,
                            double acyrus, int32_t signed_ratio, uint8_t* accept, double* grad_eff_old,
                            double* aux_out, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- device-side random inputs ("throughput mode"): counter-based Philox4x32-10 keyed by `seed`, counter =
 * (global walker id, step, slot), so a walker's numbers do not depend on the batch size or on the sharding over GPUs.
 * They stand in for the draws the reference makes inside its jitted graph (jax.random.normal / uniform / orthogonal,
 * VMC/VMCmcstep.py:19-20,58,83; pseudopotential/pseudopotential.py:233-235; DMC/Tmoves.py:146,216-217); jax's threefry
 * streams themselves are not reproducible without jax.  walker0 = global index of this batch's first walker.
 * aiqmc_rng_sweep: gauss1 (B,3N), gauss2c (B,N,3) ~ sqrt(tstep) N(0,1), rnd (B,N) ~ U[0,1).
 * aiqmc_rng_rotations: rot (B,3,3) Haar-random orthogonal (QR of a Gaussian matrix, R's diagonal positive).
 * aiqmc_rng_uniform: out (B,cols) ~ U[0,1); tag 0..15 selects an independent stream (T-move u, rnd, comb u ...). */
int aiqmc_rng_sweep(uint64_t seed, uint32_t step, int64_t walker0, int64_t n_walkers, int32_t n_elec, double tstep,
                    double* gauss1, double* gauss2c, double* rnd, void* stream);
int aiqmc_rng_rotations(uint64_t seed, uint32_t step, int64_t walker0, int64_t n_walkers, double* rot, void* stream);
int aiqmc_rng_uniform(uint64_t seed, uint32_t step, int64_t walker0, int64_t n_walkers, int32_t cols, uint32_t tag,
                      double* out, void* stream);

/* ---- local energy: replaces _e_l (Energy/hamiltonian.py:248-258, pphamiltonian.py:177-188)
 *      wrapped in jax.vmap over walkers (Loss/pploss.py:145-153) ---------------------------- */
int64_t aiqmc_energy_workspace_bytes(const AiqmcSystem* sys, int64_t n_walkers, int32_t with_ecp);
/* All-electron: E_L = V_ee + V_en + V_nn + KE (real).  e_l (B). */
int aiqmc_local_energy_ae(const AiqmcSystem* sys, const double* params, const double* pos,
                          int64_t n_walkers, double* e_l, void* workspace, int64_t workspace_bytes,
                          void* stream);
/* ccECP: E_L = V_ee + V_nn + KE + local channel + non-local 50-point quadrature (complex,
 * quirk Q25).  rot (B,3,3) is the per-walker rotation jax.random.orthogonal would have drawn
 * (pseudopotential.py:233-235).  e_l (B,2) interleaved (re,im). */
int aiqmc_local_energy_ecp(const AiqmcSystem* sys, const AiqmcEcp* ecp, const double* params,
                           const double* pos, const double* rot, int64_t n_walkers, double* e_l,
                           void* workspace, int64_t workspace_bytes, void* stream);
/* Same, running only the stages in stage_mask: 1 = KE + Coulomb + local channel + v_l tables
 * (k_energy_base), 2 = non-local quadrature, 4 = assembly.  The quadrature kernel is picked by system
 * size (thread-per-point on the single-electron-move cache for N <= 4, lane-per-electron up to N = 16,
 * full evaluation per point beyond); +8 forces the full-evaluation kernel, +16 the lane-per-electron
 * one (cross-checks).  Stages must be issued in
 * order over the same workspace; used by bench.py to time the dominant kernel on its own stream. */
int aiqmc_local_energy_ecp_stages(const AiqmcSystem* sys, const AiqmcEcp* ecp, const double* params,
                                  const double* pos, const double* rot, int64_t n_walkers, double* e_l,
                                  void* workspace, int64_t workspace_bytes, int32_t stage_mask,
                                  void* stream);
/* ---- contracted Gaussian basis: replaces primitive_Gaussian_basis (AIQMC/Gaussian_orbitals.py:11-13) and
 *      follows ferminet/utils/gto.py:117-135,338-389 (real solid harmonics, AO = radial * angular, m = -l..l) ----
 * AO_k(r) = [sum_p coef_p exp(-alpha_p |r-R|^2)] * |r-R|^l Y_lm(r-R), l <= 3, with analytic gradient and Laplacian.
 * shells / centres are HOST arrays (uploaded to constant memory); points (n,3) device.  val (n,nao), grad (n,nao,3)
 * and lap (n,nao) are device outputs; grad and lap may be NULL. */
#define AIQMC_GTO_MAX_PRIM 16
#define AIQMC_GTO_MAX_SHELLS 48
#define AIQMC_GTO_MAX_CENTRES 16
typedef struct AiqmcGtoShell {
  int32_t l;          /* angular momentum 0..3 */
  int32_t n_prim;     /* primitives in the contraction */
  int32_t centre;     /* index into centres */
  int32_t ao_offset;  /* first AO column of this shell (its 2l+1 functions are consecutive, m = -l..l) */
  double alpha[AIQMC_GTO_MAX_PRIM];
  double coef[AIQMC_GTO_MAX_PRIM];
} AiqmcGtoShell;
int aiqmc_gto_eval(const AiqmcGtoShell* shells, int32_t n_shells, const double* centres, int32_t n_centres,
                   const double* points, int64_t n_points, int32_t nao, double* val, double* grad, double* lap,
                   void* stream);

/* ---- DMC T-moves: replaces compute_tmoves / calculate_ratio_weight_tmoves (DMC/Tmoves.py:32-225) ------
 * One non-local move attempt per electron of every walker: amplitudes (exp(-tstep v_l) - 1) P_l(cos) * ratio on
 * the 50-point quadrature of every (electron, atom), clipped at 0 in jnp's lexicographic complex order, the cdf /
 * searchsorted selection driven by u (B) (the jax.random.uniform(key) of :146), the back-amplitude norm with the
 * reference's electron-axis indexing and hard-coded 1:19:55:79:151 slices (quirk Q18), and the accept test against
 * rnd (B,N) (:216-217).  rot (B,3,3) as in aiqmc_local_energy_ecp.  pos_out (B,3N), acceptance (B,N),
 * selected (B,N) int32 move index (0 = stay; may be NULL). */
int64_t aiqmc_dmc_tmove_workspace_bytes(const AiqmcSystem* sys, int64_t n_walkers);
int aiqmc_dmc_tmove(const AiqmcSystem* sys, const AiqmcEcp* ecp, const double* params, const double* pos,
                    const double* rot, const double* u, const double* rnd, int64_t n_walkers, double tstep,
                    double* pos_out, double* acceptance, int32_t* selected, void* workspace,
                    int64_t workspace_bytes, void* stream);
/* Block-reduced [sum Re E, sum Im E, sum |E|^2, count] -> stats[4] (device), the partials of
 * pmean(mean(e_l)) and the variance at Loss/pploss.py:165-167; all-reduced over GPUs by the host
 * (NCCL sum of 4 doubles).  e_l_stride = 1 (real) or 2 (complex interleaved). */
int aiqmc_energy_stats(const double* e_l, int32_t e_l_stride, int64_t n_walkers, double* stats,
                       void* stream);
/* The same statistics for batches beyond 2^18 walkers (one 8-CTA cluster cannot stream them at HBM rate): partial sums of
 * fixed 4096-walker chunks in the caller's workspace, added in a fixed order -- the result depends on the batch size
 * only.  Up to 2^18 walkers it forwards to aiqmc_energy_stats (workspace may be NULL: the size query returns 0). */
int64_t aiqmc_energy_stats_workspace_bytes(int64_t n_walkers);
int aiqmc_energy_stats_ws(const double* e_l, int32_t e_l_stride, int64_t n_walkers, double* stats, void* workspace,
                          int64_t workspace_bytes, void* stream);

/* ---- parameter side of the loss gradient: replaces the `jax.jvp(batch_network, primals, tangents)`
 * of make_loss.total_energy_jvp (Loss/pploss.py:186-223) as seen through jax.grad.
 * grad_out (aiqmc_param_layout(...).total doubles, packed layout) = sum over walkers of
 *     alpha[w] * d log|psi_w| / d params  +  beta[w] * d phase_w / d params,
 * one reverse (adjoint) sweep per walker.  With diff = clipped E_L - centre (clip_local_values, pploss.py:73-135):
 * alpha = 2 Re(diff) / B, beta = 2 (Im(diff) + Im(clipped E_L)) / B reproduces tangents_out of :215-218 for complex
 * outputs, alpha = diff / B, beta = 0 the real branch (:222).  Entries for quantities pack_params derives on the host
 * (row-normalised y weights, sigma * xi) are gradients w.r.t. the PACKED values; non-trainable slots are 0.
 * phase / logabs (n_walkers) may be NULL.  Sums are fixed-order (bit reproducible).  N <= 16: one fused forward +
 * reverse pass per walker; N > 16: the primal pass fills the derivative cache (part of the workspace) and the sweep
 * runs on it. */
int64_t aiqmc_param_grad_workspace_bytes(const AiqmcSystem* sys, int64_t n_walkers);
int aiqmc_psi_param_grad(const AiqmcSystem* sys, const double* params, const double* pos, int64_t n_walkers,
                         const double* alpha, const double* beta, double* grad_out, double* phase, double* logabs,
                         void* workspace, int64_t workspace_bytes, void* stream);

/* ---- all-electron Metropolis-Hastings: replaces mh_update / mh_accept of
 * AIQMCrelease2/MonteCarloSample/mcstep.py:26-68 (= ferminet/mcmc.py:67-150), one all-electron move per call:
 * x2 = x1 + stddev * hmean(x1) * noise with hmean the per-electron harmonic mean distance to the nuclei (:12-16),
 * ratio = 2 log|psi(x2)| + log q(x1|x2) - lp - log q(x2|x1) (:59-63), accepted where ratio > log(u) (:29-30).
 * pos (B,3N) and lp (B) = 2 log|psi(pos)| are updated in place; noise (B,3N) ~ N(0,1) and u (B) ~ U(0,1) are the two
 * draws of :49-56,:28-29 (parity mode inputs); *num_accepts (device uint64) is incremented by the number of accepted
 * walkers; accept (B) uint8 may be NULL.  The caller loops `steps` times and adapts the width (update_mcmc_width). */
int64_t aiqmc_mh_workspace_bytes(const AiqmcSystem* sys, int64_t n_walkers);
int aiqmc_mh_step(const AiqmcSystem* sys, const double* params, double* pos, double* lp, const double* noise,
                  const double* u, int64_t n_walkers, double stddev, uint8_t* accept, uint64_t* num_accepts,
                  void* workspace, int64_t workspace_bytes, void* stream);

/* ---- correlated sampling under a nuclear displacement -------------------------
 * aiqmc_correlated_samples replaces correlated_samples (correlatedsamples/corrsamples.py:23-47): the space-warp
 * move x_i += sum_a w_ia (R'_a - R_a), w_ia = r_ia^-4 / sum_b r_ib^-4, for every electron of every walker.
 * aiqmc_weights_jacobian replaces weights_jacobian (correlatedsamples/jacobianWeights.py:22-51) as written there
 * (including the use of the x displacement for all three directions): jacobian (B).
 * atoms / new_atoms: HOST arrays (n_atoms,3); pos, pos_out, jacobian: device. */
int aiqmc_correlated_samples(const double* atoms, const double* new_atoms, int32_t n_atoms, const double* pos,
                             int64_t n_walkers, int32_t n_elec, double* pos_out, void* stream);
int aiqmc_weights_jacobian(const double* atoms, const double* new_atoms, int32_t n_atoms, const double* pos,
                           int64_t n_walkers, int32_t n_elec, double* jacobian, void* stream);

/* ---- DMC: replaces DMC/drift_diffusion.py, S_matrix.py, dmc.py:86-92, branch.py ---------- */
/* Step 1 of comput_S (S_matrix.py:22-23): min over this device's walkers of
 * min(|E_est - Re E_L[b]|, branchcut[b]) -> ecut_min (device scalar).  The reference takes this
 * min over ALL walkers and devices (quirk Q20): with several GPUs the host MIN-all-reduces
 * the scalar (NCCL) before step 2. */
int aiqmc_dmc_ecut_min(const double* e_l, int32_t e_l_stride, int64_t n_walkers, double e_est,
                       const double* branchcut, double* ecut_min, void* stream);
/* Step 2: S[b] = E_T - E_est + ecut_min*sign(E_est - Re E_L[b]) / (1 + (v2[b]*tau/N)^2) with
 * v2[b] = sum_k drift[b,k]^2; drift (B,3N) is the limited drift the reference squares at dmc.py:86-89. */
int aiqmc_dmc_s(const double* e_l, int32_t e_l_stride, const double* drift, int64_t n_walkers,
                int32_t n_elec, double e_trial, double e_est, const double* ecut_min, double tau,
                double* s_out, void* stream);
/* weights *= exp(tau * tdamp * 0.5 * (S_new + S_old))   (dmc.py:91-92). */
int aiqmc_dmc_weights(double* weights, const double* s_old, const double* s_new, int64_t n_walkers,
                      double tau, double tdamp, void* stream);
/* Systematic comb (branch.py:17-23): newinds (B) int32, new_weight (device scalar) = wtot/B.
 * u in [0,1) is the uniform the reference draws at branch.py:21.  The prefix sum is a blocked scan (2048 walkers per
 * block, one CTA each; block totals scanned sequentially): its association order does not depend on B, which is what
 * lets aiqmc_rebalance_nccl reproduce it across GPUs.  weights must be 16-byte aligned. */
int64_t aiqmc_branch_workspace_bytes(int64_t n_walkers);
int aiqmc_branch_comb(const double* weights, int64_t n_walkers, double u, int32_t* newinds,
                      double* new_weight, void* workspace, int64_t workspace_bytes, void* stream);
/* Gather walkers by index: pos_out[b] = pos_in[newinds[b]]  (HBM-bound, 24N+4 B/walker). */
int aiqmc_gather_walkers(const double* pos_in, const int32_t* newinds, int64_t n_walkers,
                         int32_t row_doubles, double* pos_out, void* stream);

/* ---- multi-GPU: the three exchanges of the walker path behind the C ABI ------------------------
 * NCCL is bound at run time (dlopen of libnccl.so.2 or $AIQMC_NCCL_LIB); the communicator is created once by the
 * host from a 128-byte unique id that rank 0 makes and the host distributes (any transport), and is passed in as an
 * opaque handle (an ncclComm_t). */
int aiqmc_nccl_available(void);
int aiqmc_nccl_unique_id(void* id128);
int aiqmc_nccl_comm_init(int32_t world, int32_t rank, const void* id128, void** comm_out);
int aiqmc_nccl_comm_destroy(void* comm);
/* (1) energy mean / variance: in-place SUM all-reduce of the 4-double statistics of aiqmc_energy_stats
 *     (the two pmean's of Loss/pploss.py:165-167). */
int aiqmc_energy_allreduce(double* stats, void* comm, void* stream);
/* (2) quirk Q20: in-place MIN all-reduce of the e_cut scalar of aiqmc_dmc_ecut_min (DMC/S_matrix.py:23). */
int aiqmc_ecut_allreduce_min(double* ecut_min, void* comm, void* stream);
/* (3) cross-GPU DMC population control (new capability; the reference combs per device, DMC/branch.py:10-34 under
 *     pmap): the systematic comb over the weights of ALL ranks + migration of the selected walkers.  Exchanged: the
 *     block totals of each rank's blocked scan (n_walkers/2048 doubles per rank) by all-gather, then ONE grouped
 *     ncclSend/ncclRecv of only the walkers that change rank.  Equal n_walkers on every rank.  The result is identical,
 *     bit for bit, to aiqmc_branch_comb + aiqmc_gather_walkers run on one GPU over the concatenated batch.
 *     pos (B,row) -> pos_out (B,row): slot k of rank r receives the walker tooth r*B + k selects; new_weight (device
 *     scalar) = total weight / (world*B); src_rank_out (B) int32 (may be NULL) = rank each new walker came from;
 *     *moved_bytes_out (host, may be NULL) = bytes this rank sent to other ranks.  Synchronises the stream once (the
 *     message sizes must reach the host).  world == 1 needs no communicator.
 *     mode 0 ("ordered"): the layout above, identical to the single-GPU comb; because the comb's base offset u*wtot
 *     rotates the teeth by a fraction u of the population, almost every walker changes rank.
 *     mode 1 ("balanced"): the SAME multiset of walkers (same survivors, same multiplicities), but every rank keeps the
 *     walkers it already holds (in tooth order) and only the surplus of ranks owning more than B teeth moves, to the
 *     ranks owning fewer (a deterministic rank-order matching computed identically everywhere): the wire carries the
 *     population imbalance only.  Walkers are exchangeable, so the DMC estimators are unaffected. */
int64_t aiqmc_rebalance_workspace_bytes(int64_t n_walkers, int32_t row_doubles, int32_t world);
int aiqmc_rebalance_nccl(const double* weights, const double* pos, int64_t n_walkers, int32_t row_doubles, double u,
                         int32_t world, int32_t rank, void* comm, int32_t mode, double* pos_out, double* new_weight,
                         int32_t* src_rank_out, int64_t* moved_bytes_out, void* workspace, int64_t workspace_bytes,
                         void* stream);

/* ---- measurement aid ------------------------------------------------------------------ */
/* Dependent-free FP64 FMA loop over the whole chip (148*8 CTAs x 256 threads x 8 chains):
 * performs *flops_out = grid*256*8*2*iters flops; time it with CUDA events to get the DFMA peak
 * that bench.py uses as the roofline denominator (MEASURED_PEAKS.json has no FP64 figure). */
int aiqmc_bench_dfma(int64_t iters, double* sink, double* flops_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AIQMC_B200_H_ */
