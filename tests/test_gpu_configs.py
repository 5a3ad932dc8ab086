"""GPU parity tests that close the round-1 gaps on the BASELINE configurations: C6H6 local energy against the ORACLE
(not against another CUDA kernel), N2 dmc_propagate, a full-size (65,536-walker) bit-exact accept mask with the same
per-device batch on both sides (quirk Q6 couples every accept to a batch-wide sum), the device-side Philox inputs,
two engines on two streams, and the complex e_est / e_trial feed-back of a DMC loop."""
import math
import threading

import numpy as np
import pytest
import torch

from common import CASES, Case, O, benzene_case, build_bench_case, ecp_tables, make_rand
from oracle import philox as PH

import aiqmc_b200

pytestmark = pytest.mark.gpu
TSTEP = 0.05


def engine(case, ecp=None):
    return aiqmc_b200.WalkerEngine(case.spec(), case.params, ecp=ecp)


def test_benzene_local_energy_matches_oracle():
    """BASELINE configs[4] (N=30, A=12): E_L of 4 walkers -- kinetic (forward Laplacian), Coulomb, local channel and the
    18,000-point non-local quadrature per walker -- against the float64 oracle."""
    case = benzene_case(4, width=0.8)
    tabs = ecp_tables(case.a, rich=True)
    eng = engine(case, ecp=aiqmc_b200.make_ecp(case.a, list_l=2, **tabs))
    rot = torch.tensor(O.random_rotations(case.rng, case.B))
    e = eng.local_energy(torch.tensor(case.pos), rot).cpu().numpy()
    le = O.local_energy_ecp(case.net.apply, O.make_log_network(case.net.apply), case.charges, None, tabs['rn_local'],
                            tabs['local_coes'], tabs['local_exps'], tabs['rn_non_local'], tabs['non_local_coes'],
                            tabs['non_local_exps'], case.a, case.n, 3, 2)
    ref, _ = le(case.params, rot, case.oracle_data(batched_static=False))
    np.testing.assert_allclose(e, ref.numpy(), atol=1e-5, rtol=0)            # north_star: 1e-5 Ha
    np.testing.assert_allclose(e, ref.numpy(), atol=1e-7, rtol=1e-8)


def test_n2_dmc_propagate_matches_oracle():
    """dmc_propagate_run (DMC/dmc.py:72-93) on N2 (N=10, A=2): positions, E_L and weights after one full DMC step."""
    case = Case(**CASES["N2_ecp"], nwalkers=12, width=0.7)
    tabs = ecp_tables(case.a, rich=True)
    rng = case.rng
    key = dict(tmove=dict(rot=torch.tensor(O.random_rotations(rng, case.B)), u=torch.tensor(rng.uniform(size=case.B)),
                          rnd=torch.tensor(rng.uniform(size=(case.B, case.n)))),
               sweep=case.sweep_rand(TSTEP), rot=torch.tensor(O.random_rotations(rng, case.B)))
    weights = torch.tensor(rng.uniform(0.5, 1.5, size=case.B))
    branchcut = torch.full((case.B,), 3.0)
    e_trial, e_est = -19.9, -20.0
    ref = O.dmc_propagate(case.net.apply, TSTEP, case.n, case.a, 3, case.B, case.charges, **tabs)
    e_ref, w_ref, d_ref = ref(case.params, key, case.oracle_data(), weights, branchcut, e_trial, e_est)
    net = aiqmc_b200.make_ai_net(**case.kw)
    run = aiqmc_b200.dmc_propagate(net.apply, net.apply, TSTEP, case.n, case.a, 3, case.B, case.charges, **tabs)
    data = aiqmc_b200.AINetData(positions=torch.tensor(case.pos), spins=case.t_spins, atoms=case.t_atoms,
                                charges=torch.tensor(case.charges))
    e_gpu, w_gpu, d_gpu = run(case.params, key, data, weights, branchcut, e_trial, e_est)
    np.testing.assert_allclose(d_gpu.positions.cpu().numpy(), d_ref.positions.numpy(), rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(e_gpu.cpu().numpy(), e_ref.numpy(), atol=1e-5, rtol=0)        # north_star: 1e-5 Ha
    np.testing.assert_allclose(e_gpu.cpu().numpy(), e_ref.numpy(), atol=1e-7, rtol=1e-8)
    np.testing.assert_allclose(w_gpu.cpu().numpy(), w_ref.numpy(), rtol=1e-8)


def test_full_size_accept_mask_is_bit_exact_against_the_oracle():
    """BASELINE configs[1] at its stated size: ONE sweep of 65,536 walkers; the oracle runs the same 65,536-walker
    batch (limdrift's v2 is a sum over the whole per-device batch, quirk Q6), 262,144 accept decisions compared bit
    for bit, positions to 1e-10.  (The oracle side is 5 x 65,536 autograd evaluations on the host: about a minute.)"""
    B = 65536
    case, _ = build_bench_case(B)
    eng = engine(case)
    rng = np.random.default_rng(2024)
    rand = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in make_rand(rng, B, case.n, TSTEP).items()}
    pos = torch.from_numpy(case.pos.copy()).cuda()
    out = eng.vmc_sweep(pos, rand["gauss1"].cuda(), rand["gauss2"].cuda(), rand["rnd"].cuda(), TSTEP, want_aux=True)
    new_data, aux = O.walkers_update(O.select_output(case.net.apply, 1), case.params, case.oracle_data(), rand, TSTEP, 3,
                                     case.n, B, return_aux=True)
    got = out["accept"].cpu().numpy().astype(bool)
    ref = aux["accept"].numpy()
    assert got.shape == ref.shape == (B, case.n)
    assert int((got != ref).sum()) == 0, f"{int((got != ref).sum())} of {got.size} accept decisions differ"
    np.testing.assert_allclose(pos.cpu().numpy(), new_data.positions.numpy(), rtol=1e-10, atol=1e-10)
    assert 0.2 < got.mean() < 0.999


def test_compact_gauss2_is_the_diagonal_of_the_reference_array():
    case = Case(**CASES["N2_ecp"], nwalkers=40, width=0.8)
    eng = engine(case)
    rand = case.sweep_rand(TSTEP)
    g2c = torch.stack([rand["gauss2"][:, i, 3 * i:3 * i + 3] for i in range(case.n)], dim=1).contiguous()
    p1, p2 = torch.tensor(case.pos).cuda(), torch.tensor(case.pos).cuda()
    o1 = eng.vmc_sweep(p1, rand["gauss1"].cuda(), rand["gauss2"].cuda(), rand["rnd"].cuda(), TSTEP)
    o2 = eng.vmc_sweep(p2, rand["gauss1"].cuda(), g2c.cuda(), rand["rnd"].cuda(), TSTEP)
    assert torch.equal(p1, p2) and torch.equal(o1["accept"], o2["accept"])


def test_device_philox_inputs_match_the_numpy_restatement():
    case = Case(**CASES["C_ecp"], nwalkers=2)
    eng = engine(case)
    seed, step, w0, B = 0x1234567890ABCDEF, 7, 5_000_000_000, 1000          # walker ids beyond 32 bits
    g1, g2c, u = eng.rng_sweep(seed, step, w0, B, TSTEP)
    r1, r2, ru = PH.rng_sweep(seed, step, w0, B, case.n, TSTEP)
    assert np.array_equal(u.cpu().numpy(), ru)                               # uniforms: bit-exact
    np.testing.assert_allclose(g1.cpu().numpy(), r1, rtol=1e-13, atol=1e-15)  # normals: to libm rounding
    np.testing.assert_allclose(g2c.cpu().numpy(), r2, rtol=1e-13, atol=1e-15)
    rot = eng.rng_rotations(seed, step, w0, B).cpu().numpy()
    np.testing.assert_allclose(rot, PH.rng_rotations(seed, step, w0, B), rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(rot @ np.swapaxes(rot, -1, -2), np.broadcast_to(np.eye(3), (B, 3, 3)), atol=1e-13)
    uu = eng.rng_uniform(seed, step, w0, B, 5, 3).cpu().numpy()
    assert np.array_equal(uu, PH.rng_uniform(seed, step, w0, B, 5, 3))
    # a walker's numbers do not depend on the batch it is generated in (sharding independence)
    h1, _, hu = eng.rng_sweep(seed, step, w0 + 600, 400, TSTEP)
    assert torch.equal(h1, g1[600:]) and torch.equal(hu, u[600:])
    # distribution sanity on a larger draw
    big, _, bu = eng.rng_sweep(3, 0, 0, 200_000, TSTEP)
    z = big / math.sqrt(TSTEP)
    assert abs(float(z.mean())) < 5e-3 and abs(float(z.std()) - 1.0) < 5e-3 and abs(float(bu.mean()) - 0.5) < 3e-3


def test_seeded_sweep_equals_parity_sweep_on_the_same_numbers():
    """mc_step with an int key (device Philox) == mc_step with the explicit arrays the numpy restatement produces."""
    case = Case(**CASES["C_ecp"], nwalkers=64, width=0.9)
    net = aiqmc_b200.make_ai_net(**case.kw)
    mc = aiqmc_b200.main_monte_carlo(net.apply, TSTEP, 3, case.n, 2, case.B)
    data = aiqmc_b200.AINetData(positions=torch.tensor(case.pos), spins=case.t_spins, atoms=case.t_atoms,
                                charges=torch.tensor(case.charges))
    out_seed = mc(case.params, data, 99).positions
    keys = []
    for s in range(2):
        g1, g2c, u = PH.rng_sweep(99, s, 0, case.B, case.n, TSTEP)
        keys.append(dict(gauss1=g1, gauss2=PH.expand_gauss2(g2c), rnd=u))
    out_par = mc(case.params, data, keys).positions
    np.testing.assert_allclose(out_seed.cpu().numpy(), out_par.cpu().numpy(), rtol=1e-12, atol=1e-13)


def test_two_engines_on_two_streams_do_not_race_on_constant_memory():
    """Two engines of the SAME system with DIFFERENT parameters and ECP tables, driven from two host threads on two
    streams: the per-instantiation __constant__ copies (parameters of k_ecp_pt, ECP table) are shared by both; the
    library orders their use (ConstScope, engine_impl.cuh).  Results must equal the serial ones bit for bit."""
    case_a = Case(**CASES["C_ecp"], nwalkers=3000, width=0.9)
    case_b = Case(**{**CASES["C_ecp"], "seed": 77}, nwalkers=3000, width=0.9)
    tabs_a, tabs_b = ecp_tables(1, rich=True), ecp_tables(1, rich=False)
    eng_a = engine(case_a, ecp=aiqmc_b200.make_ecp(1, list_l=2, **tabs_a))
    eng_b = engine(case_b, ecp=aiqmc_b200.make_ecp(1, list_l=2, **tabs_b))
    rot = torch.tensor(O.random_rotations(np.random.default_rng(5), 3000)).cuda()
    pa, pb = torch.tensor(case_a.pos).cuda(), torch.tensor(case_b.pos).cuda()
    ref_a, ref_b = eng_a.local_energy(pa, rot).clone(), eng_b.local_energy(pb, rot).clone()
    torch.cuda.synchronize()
    results, errors = {}, []

    def worker(tag, eng, p, ref):
        try:
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                for it in range(40):
                    e = eng.local_energy(p, rot)
                    if not torch.equal(e, ref):
                        errors.append((tag, it, float((e - ref).abs().max())))
            st.synchronize()
            results[tag] = True
        except Exception as exc:                                   # surfaced below
            errors.append((tag, repr(exc)))
    ta = threading.Thread(target=worker, args=("a", eng_a, pa, ref_a))
    tb = threading.Thread(target=worker, args=("b", eng_b, pb, ref_b))
    ta.start(); tb.start(); ta.join(); tb.join()
    assert not errors, errors[:5]
    assert results == {"a": True, "b": True}


def test_dmc_loop_feeds_complex_estimates_back():
    """estimate_energy / trial_energy of ccECP (complex) energies go back into dmc_propagate_run on the
    next block, as DMC/main_dmc.py:188-242 does (the reference takes the real parts, S_matrix.py:18-20)."""
    case = Case(**CASES["C_ecp"], nwalkers=16, width=0.7)
    tabs = ecp_tables(1, rich=True)
    rng = case.rng
    net = aiqmc_b200.make_ai_net(**case.kw)
    run = aiqmc_b200.dmc_propagate(net.apply, net.apply, TSTEP, case.n, 1, 3, case.B, case.charges, **tabs)
    data = aiqmc_b200.AINetData(positions=torch.tensor(case.pos), spins=case.t_spins, atoms=case.t_atoms,
                                charges=torch.tensor(case.charges))
    w = torch.ones(case.B, dtype=torch.float64)
    e_trial, e_est = -5.39, -5.41
    for block in range(2):
        key = dict(tmove=dict(rot=torch.tensor(O.random_rotations(rng, case.B)), u=torch.tensor(rng.uniform(size=case.B)),
                              rnd=torch.tensor(rng.uniform(size=(case.B, case.n)))),
                   sweep=case.sweep_rand(TSTEP), rot=torch.tensor(O.random_rotations(rng, case.B)))
        e, w, data = run(case.params, key, data, w, torch.full((case.B,), 3.0), e_trial, e_est)
        e_est = aiqmc_b200.estimate_energy(e, w)                    # complex tensor
        e_trial = aiqmc_b200.trial_energy(e_est, w, 1.0)            # complex tensor
        assert torch.is_complex(e_est)
    assert torch.isfinite(w).all()


def test_rebalance_on_one_rank_equals_comb_plus_gather():
    """aiqmc_rebalance_nccl with world = 1 (no communicator): 'ordered' is the single-GPU comb + gather bit for bit,
    'balanced' holds the same walkers with the same multiplicities (the multi-rank paths are checked on real GPUs by
    tools/multigpu_check.py, profiles/r2_multigpu_check_*.txt, and on CPU by tests/test_distributed_gloo.py)."""
    case = Case(**CASES["C_ecp"], nwalkers=2)
    eng = engine(case)
    rng = np.random.default_rng(5)
    for B in (1000, 65536):
        w = torch.tensor(rng.uniform(0.05, 3.0, size=B)).cuda()
        p = torch.tensor(rng.normal(size=(B, 12))).cuda()
        neww, inds = eng.branch_comb(w, 0.61)
        ref = eng.gather_walkers(p, inds)
        n0, p0, s0, m0 = eng.rebalance(w, p, 0.61, None, mode="ordered")
        assert float(n0) == float(neww) and torch.equal(p0, ref) and m0 == 0 and int(s0.abs().sum()) == 0
        n1, p1, s1, m1 = eng.rebalance(w, p, 0.61, None, mode="balanced")
        assert float(n1) == float(neww) and m1 == 0
        assert torch.equal(torch.sort(p1[:, 0]).values, torch.sort(ref[:, 0]).values)
        # oracle: the comb of DMC/branch.py on the same weights selects the same walkers (ulp ties aside)
        _, io = O.branch(w.cpu(), 0.61)
        assert float((inds.cpu() != io).double().mean()) <= 2e-3       # a tooth within an ulp of a boundary may flip: the scans associate differently


def test_n2_large_batch_accept_mask_is_bit_exact():
    """BASELINE configs[2] (N2, N=10, A=2): one sweep of 4,096 walkers against the oracle run on the same 4,096-walker
    batch (quirk Q6: the drift limiter sums over the whole per-device batch): 40,960 accept decisions bit for bit."""
    B = 4096
    case = Case(**CASES["N2_ecp"], nwalkers=B, width=0.9)
    eng = engine(case)
    rand = case.sweep_rand(TSTEP)
    pos = torch.tensor(case.pos).cuda()
    out = eng.vmc_sweep(pos, rand["gauss1"].cuda(), rand["gauss2"].cuda(), rand["rnd"].cuda(), TSTEP)
    new_data, aux = O.walkers_update(O.select_output(case.net.apply, 1), case.params, case.oracle_data(), rand, TSTEP, 3,
                                     case.n, B, return_aux=True)
    got, ref = out["accept"].cpu().numpy().astype(bool), aux["accept"].numpy()
    assert int((got != ref).sum()) == 0, f"{int((got != ref).sum())} of {got.size} accept decisions differ"
    np.testing.assert_allclose(pos.cpu().numpy(), new_data.positions.numpy(), rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("n,a", [(8, 2), (12, 2), (2, 1), (3, 1)])
def test_lane_per_electron_derivative_kernels_on_every_group_shape(n, a):
    """coop_grad.cuh / coop_lap.cuh pack floor(32/N) configurations per warp: N = 8 (4 groups, no idle lanes), N = 12
    (2 groups, 8 idle lanes), N = 2 (16 groups), N = 3 (10 groups, 2 idle lanes) -- C2 with ccECP / all-electron are the
    reference's own examples.  log|psi|, gradient and Laplacian against autograd on the oracle, ragged batch sizes."""
    spins = [1.] * ((n + 1) // 2) + [-1.] * (n // 2)
    for nw in (1, 7):
        case = Case(n=n, natoms=a, spins=spins, seed=40 + n, nwalkers=nw, width=0.8)
        eng = engine(case)
        ph, la, g, lap = eng.psi(torch.tensor(case.pos), mode=2)
        _, la1, g1 = eng.psi(torch.tensor(case.pos), mode=1)
        f = lambda x: case.net.apply(case.params, x, case.t_spins, case.t_atoms)[1]
        lat, gt, d2 = O.grad_and_hess_diag(f, torch.tensor(case.pos))
        np.testing.assert_allclose(la.cpu().numpy(), lat.detach().numpy(), rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(la1.cpu().numpy(), lat.detach().numpy(), rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(g.cpu().numpy(), gt.numpy(), rtol=1e-7, atol=1e-8)
        np.testing.assert_allclose(g1.cpu().numpy(), gt.numpy(), rtol=1e-7, atol=1e-8)
        np.testing.assert_allclose(lap.cpu().numpy(), d2.sum(-1).numpy(), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("n,a", [(5, 2), (6, 1), (8, 2), (12, 2)])
def test_group_quadrature_kernel_on_every_group_shape(n, a):
    """k_ecp_grp packs floor(32/N) quadrature points per warp and, for N <= 16, walks ALL N*A*50 points of a walker in
    chunks of 4 NG: N = 5 (6 groups, 2 idle lanes), 6 (5 groups), 8 (4 groups, none idle), 12 (2 groups, 8 idle lanes);
    chunk sizes 4 NG that do and do not divide N*A*50.  ccECP local energy against the oracle, ragged batch."""
    spins = [1.] * ((n + 1) // 2) + [-1.] * (n // 2)
    case = Case(n=n, natoms=a, spins=spins, seed=70 + n, nwalkers=3, width=0.8)
    tabs = ecp_tables(a, rich=True)
    eng = engine(case, ecp=aiqmc_b200.make_ecp(a, list_l=2, **tabs))
    rot = torch.tensor(O.random_rotations(case.rng, case.B))
    e = eng.local_energy(torch.tensor(case.pos), rot).cpu().numpy()
    le = O.local_energy_ecp(case.net.apply, O.make_log_network(case.net.apply), case.charges, None,
                            tabs['rn_local'], tabs['local_coes'], tabs['local_exps'], tabs['rn_non_local'],
                            tabs['non_local_coes'], tabs['non_local_exps'], a, n, 3, 2)
    ref, _ = le(case.params, rot, case.oracle_data(batched_static=False))
    np.testing.assert_allclose(e, ref.numpy(), atol=1e-7, rtol=1e-8)


@pytest.mark.gpu
def test_energy_stats_large_batch_workspace_path():
    """Beyond 2^18 walkers the statistics go through chunk partials in a workspace (aiqmc_energy_stats_ws): same numbers
    as a float64 reference sum to round-off, deterministic, and identical to the single-cluster kernel at its size limit."""
    import aiqmc_b200
    from aiqmc_b200 import workloads as W
    eng = W.build("c_ecp", 2).engine()
    rng = np.random.default_rng(3)
    for B in (1 << 18, (1 << 18) + 1, 1_000_003):
        e = torch.tensor(rng.normal(size=(B, 2)) * 3.0 - 5.0).cuda()
        ec = torch.view_as_complex(e)
        got = eng.energy_stats(ec).cpu().numpy()
        ref = np.array([float(e[:, 0].sum()), float(e[:, 1].sum()), float((e * e).sum()), B])
        np.testing.assert_allclose(got, ref, rtol=1e-12)
        assert np.array_equal(got, eng.energy_stats(ec).cpu().numpy())                 # run-to-run identical
        gr = eng.energy_stats(e[:, 0].contiguous()).cpu().numpy()                      # real input (stride 1)
        np.testing.assert_allclose(gr, [ref[0], 0.0, float((e[:, 0] ** 2).sum()), B], rtol=1e-12, atol=1e-9)
