"""GPU parity: the CUDA path (through the C ABI / aiqmc_b200 host mirror) against the oracle on the
same seeded inputs.  Tolerances are north_star's: log|psi| 1e-6 relative, E_L 1e-5 Ha, accept
masks and comb indices bit-exact."""
import math

import numpy as np
import pytest
import torch

from common import CASES, Case, O, ecp_tables

import aiqmc_b200

pytestmark = pytest.mark.gpu


def engine(case, ecp=None):
    return aiqmc_b200.WalkerEngine(case.spec(), case.params, ecp=ecp)


@pytest.mark.parametrize("name", list(CASES))
def test_signed_network_value_grad_laplacian(name):
    case = Case(**CASES[name], nwalkers=64)
    eng = engine(case)
    ph, la, g, lp = (t.cpu().numpy() for t in eng.psi(case.pos, mode=2))
    f = lambda x: case.net.apply(case.params, x, case.t_spins, case.t_atoms)[1]
    pht, lat = case.net.apply(case.params, torch.tensor(case.pos), case.t_spins, case.t_atoms)
    _, gt, dt = O.grad_and_hess_diag(f, torch.tensor(case.pos))
    np.testing.assert_allclose(la, lat.numpy(), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(np.angle(np.exp(1j * (ph - pht.numpy()))), 0.0, atol=1e-6)
    np.testing.assert_allclose(g, gt.numpy(), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(lp, dt.sum(-1).numpy(), rtol=1e-6, atol=1e-6)
    ph0, la0 = (t.cpu().numpy() for t in eng.psi(case.pos, mode=0))
    np.testing.assert_allclose(la0, la, rtol=1e-12, atol=1e-12)
    ph1, la1, g1 = (t.cpu().numpy() for t in eng.psi(case.pos, mode=1))
    np.testing.assert_allclose(g1, g, rtol=1e-10, atol=1e-10)


def test_signed_network_dropin_signature():
    case = Case(**CASES["C_ecp"], nwalkers=16)
    net = aiqmc_b200.make_ai_net(**case.kw)
    phase, logabs = net.apply(case.params, torch.tensor(case.pos), case.t_spins, case.t_atoms,
                              torch.tensor(case.charges))
    pht, lat = case.net.apply(case.params, torch.tensor(case.pos), case.t_spins, case.t_atoms)
    np.testing.assert_allclose(logabs.cpu().numpy(), lat.numpy(), rtol=1e-6)
    # single configuration (no batch axis), float32 input as the reference passes
    p1, l1 = net.apply(case.params, torch.tensor(case.pos[0], dtype=torch.float32), case.t_spins, case.t_atoms, None)
    assert p1.shape == () and l1.shape == ()
    np.testing.assert_allclose(float(l1), float(lat[0]), rtol=1e-5)


def test_empty_and_ragged_batches():
    case = Case(**CASES["C_ecp"], nwalkers=130)      # not a multiple of the CTA size
    eng = engine(case)
    ph, la = eng.psi(case.pos[:0], mode=0)
    assert ph.shape == (0,) and la.shape == (0,)
    _, la = eng.psi(case.pos, mode=0)
    _, lat = case.net.apply(case.params, torch.tensor(case.pos), case.t_spins, case.t_atoms)
    np.testing.assert_allclose(la.cpu().numpy(), lat.numpy(), rtol=1e-6)


@pytest.mark.parametrize("name,signed", [("C_ecp", False), ("C_ae", False), ("N2_ecp", False), ("odd", True)])
def test_vmc_sweep_accept_mask_bit_exact(name, signed):
    tstep = 0.05
    case = Case(**CASES[name], nwalkers=48)
    eng = engine(case)
    rand = case.sweep_rand(tstep)
    new_data, aux = O.walkers_update(O.select_output(case.net.apply, 1), case.params, case.oracle_data(), rand, tstep,
                                     3, case.n, case.B, signed=signed, return_aux=True)
    pos = torch.tensor(case.pos).cuda()
    out = eng.vmc_sweep(pos, rand['gauss1'].cuda(), rand['gauss2'].cuda().contiguous(), rand['rnd'].cuda(), tstep,
                        signed_ratio=signed, want_drift=True, want_aux=True)
    torch.cuda.synchronize()
    assert np.array_equal(out['accept'].cpu().numpy().astype(bool), aux['accept'].numpy())      # bit-exact
    np.testing.assert_allclose(pos.cpu().numpy(), new_data.positions.numpy(), rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(out['grad_eff_old'].cpu().numpy(), aux['grad_eff'].numpy(), rtol=1e-7, atol=1e-9)
    auxv = out['aux'].cpu().numpy()
    np.testing.assert_allclose(auxv[2], float((aux['grad'] ** 2).sum()), rtol=1e-8)
    assert 0 < aux['accept'].float().mean() < 1 or case.n <= 2


def test_mc_step_dropin_three_sweeps():
    tstep, nsteps = 0.05, 3
    case = Case(**CASES["C_ecp"], nwalkers=32)
    keys = [case.sweep_rand(tstep) for _ in range(nsteps)]
    ref = O.main_monte_carlo(case.net.apply, tstep, 3, case.n, nsteps, case.B)(case.params, case.oracle_data(), keys)
    net = aiqmc_b200.make_ai_net(**case.kw)
    mc_step = aiqmc_b200.main_monte_carlo(net.apply, tstep, 3, case.n, nsteps, case.B)
    data = aiqmc_b200.AINetData(positions=torch.tensor(case.pos), spins=case.t_spins, atoms=case.t_atoms,
                                charges=torch.tensor(case.charges))
    out = mc_step(case.params, data, keys)
    np.testing.assert_allclose(out.positions.cpu().numpy(), ref.positions.numpy(), rtol=1e-8, atol=1e-9)


@pytest.mark.parametrize("name", ["C_ae", "odd", "h2like"])
def test_local_energy_all_electron(name):
    case = Case(**CASES[name], nwalkers=40)
    eng = engine(case)
    e = eng.local_energy(torch.tensor(case.pos)).cpu().numpy()
    le = O.local_energy_ae(case.net.apply, case.charges)
    ref, _ = le(case.params, None, case.oracle_data(batched_static=False))
    np.testing.assert_allclose(e, ref.numpy(), atol=1e-5, rtol=0)          # north_star: 1e-5 Ha
    np.testing.assert_allclose(e, ref.numpy(), atol=1e-8, rtol=1e-9)


@pytest.mark.parametrize("name,rich", [("C_ecp", False), ("C_ecp", True), ("N2_ecp", True)])
def test_local_energy_ecp(name, rich):
    case = Case(**CASES[name], nwalkers=24, width=0.7)
    tabs = ecp_tables(case.a, rich=rich)
    ecp = aiqmc_b200.make_ecp(case.a, list_l=2, **tabs)
    eng = engine(case, ecp=ecp)
    rot = torch.tensor(O.random_rotations(case.rng, case.B))
    e = eng.local_energy(torch.tensor(case.pos), rot).cpu().numpy()
    le = O.local_energy_ecp(case.net.apply, O.make_log_network(case.net.apply), case.charges, None,
                            tabs['rn_local'], tabs['local_coes'], tabs['local_exps'], tabs['rn_non_local'],
                            tabs['non_local_coes'], tabs['non_local_exps'], case.a, case.n, 3, 2)
    ref, _ = le(case.params, rot, case.oracle_data(batched_static=False))
    np.testing.assert_allclose(e.real, ref.real.numpy(), atol=1e-5, rtol=0)
    np.testing.assert_allclose(e.imag, ref.imag.numpy(), atol=1e-5, rtol=0)
    np.testing.assert_allclose(e, ref.numpy(), atol=1e-8, rtol=1e-9)
    stats = eng.energy_stats(torch.tensor(e).cuda()).cpu().numpy()
    np.testing.assert_allclose(stats, [e.real.sum(), e.imag.sum(), (np.abs(e) ** 2).sum(), len(e)], rtol=1e-12)


@pytest.mark.parametrize("name,scale", [("C_ecp", 30.0), ("N2_ecp", 12.0)])
def test_saturated_tanh_network_value_and_ecp_energy(name, scale):
    """Weights scaled until most tanh arguments sit far in the saturated tails (|z| up to a few hundred): the in-house
    tanh bodies (sign-free 1 - 2/(1 + exp(2z)); the quadrature kernels' 512-entry-table variant with the clamped, biased
    exponent) must stay finite and agree with torch.tanh through log|psi| and the ccECP local energy."""
    case = Case(**CASES[name], nwalkers=16, width=0.9)

    def blow(t):
        for lp in t['layers']['streams']:
            for key in ('convolutional', 'single', 'double'):
                if key in lp:
                    lp[key]['w'] = lp[key]['w'] * scale
        for lp in t['layers']['streams_y']:
            lp['single_Ynlm']['w'] = lp['single_Ynlm']['w'] * scale
        return t
    case.params = blow(case.params)
    tabs = ecp_tables(case.a, rich=True)
    eng = engine(case, ecp=aiqmc_b200.make_ecp(case.a, list_l=2, **tabs))
    phase, la = eng.psi(torch.tensor(case.pos), mode=0)
    p0, l0 = case.net.apply(case.params, torch.tensor(case.pos), case.t_spins, case.t_atoms)
    assert np.isfinite(la.cpu().numpy()).all()
    np.testing.assert_allclose(la.cpu().numpy(), l0.numpy(), rtol=1e-9, atol=1e-9)
    rot = torch.tensor(O.random_rotations(case.rng, case.B))
    e = eng.local_energy(torch.tensor(case.pos), rot).cpu().numpy()
    le = O.local_energy_ecp(case.net.apply, O.make_log_network(case.net.apply), case.charges, None,
                            tabs['rn_local'], tabs['local_coes'], tabs['local_exps'], tabs['rn_non_local'],
                            tabs['non_local_coes'], tabs['non_local_exps'], case.a, case.n, 3, 2)
    ref, _ = le(case.params, rot, case.oracle_data(batched_static=False))
    assert np.isfinite(e).all()
    scale_e = np.maximum(1.0, np.abs(ref.numpy()))
    assert (np.abs(e - ref.numpy()) / scale_e).max() < 1e-6              # derivatives of a saturated network are large


def test_local_energy_dropin_factory():
    case = Case(**CASES["C_ecp"], nwalkers=8)
    tabs = ecp_tables(1)
    net = aiqmc_b200.make_ai_net(**case.kw)
    le = aiqmc_b200.local_energy(net.apply, case.charges, lognetwork=None, natoms=1, nelectrons=4, ndim=3, list_l=2,
                                 **tabs)
    rot = torch.tensor(O.random_rotations(case.rng, case.B))
    data = aiqmc_b200.AINetData(positions=torch.tensor(case.pos), spins=case.t_spins, atoms=case.t_atoms,
                                charges=torch.tensor(case.charges))
    e, aux = le(case.params, rot, data)
    assert aux is None and e.is_complex() and e.shape == (case.B,)
    ref, _ = O.local_energy_ecp(case.net.apply, O.make_log_network(case.net.apply), case.charges, None,
                                tabs['rn_local'], tabs['local_coes'], tabs['local_exps'], tabs['rn_non_local'],
                                tabs['non_local_coes'], tabs['non_local_exps'], 1, 4, 3, 2)(
        case.params, rot, case.oracle_data(batched_static=False))
    np.testing.assert_allclose(e.cpu().numpy(), ref.numpy(), atol=1e-5)


def test_dmc_drift_S_weights_branch():
    tstep = 0.05
    case = Case(**CASES["C_ecp"], nwalkers=64)
    eng = engine(case)
    net = aiqmc_b200.make_ai_net(**case.kw)
    packed = net.pack(case.params, case.atoms)
    rand = case.sweep_rand(tstep)
    # oracle
    dd = O.propose_drift_diffusion(O.select_output(case.net.apply, 1), tstep, 3, case.n, case.B)
    new_data, tdamp, g_old, g_new, aux = dd(case.params, rand, case.oracle_data())
    # engine
    data = aiqmc_b200.AINetData(positions=torch.tensor(case.pos), spins=case.t_spins, atoms=case.t_atoms,
                                charges=torch.tensor(case.charges))
    nd, tdamp_g, go_g, gn_g, acc = aiqmc_b200.propose_drift_diffusion(net.apply, tstep, 3, case.n, case.B)(packed, rand, data)
    assert np.array_equal(acc.cpu().numpy().astype(bool), aux['accept'].numpy())
    np.testing.assert_allclose(float(tdamp_g), float(tdamp), rtol=1e-10)
    np.testing.assert_allclose(go_g.cpu().numpy(), g_old.numpy(), rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(gn_g.cpu().numpy(), g_new.numpy(), rtol=1e-7, atol=1e-9)
    # S and weights on synthetic energies
    rng = case.rng
    eloc = torch.tensor(rng.normal(-5.4, 0.5, size=case.B) + 1j * rng.normal(0, 0.01, size=case.B))
    branchcut = torch.full((case.B,), 10 * 0.5)
    s_ref = O.comput_S(-5.41, -5.40, branchcut, g_old ** 2, tstep, eloc, case.n)
    s_gpu = aiqmc_b200.comput_S(eng, -5.41, -5.40, branchcut.cuda(), go_g, tstep, eloc.cuda())
    np.testing.assert_allclose(s_gpu.cpu().numpy(), s_ref.numpy(), rtol=1e-10, atol=1e-12)
    s2_ref = O.comput_S(-5.41, -5.40, branchcut, g_new ** 2, tstep, eloc * 1.01, case.n)
    s2_gpu = aiqmc_b200.comput_S(eng, -5.41, -5.40, branchcut.cuda(), gn_g, tstep, (eloc * 1.01).cuda())
    w = torch.ones(case.B).cuda()
    eng.dmc_weights(w, s_gpu, s2_gpu, tstep, float(tdamp))
    w_ref = torch.exp(tstep * tdamp * (0.5 * s2_ref + 0.5 * s_ref))
    np.testing.assert_allclose(w.cpu().numpy(), w_ref.numpy(), rtol=1e-12)


@pytest.mark.parametrize("B", [8, 1000, 65536])
def test_branch_comb_bit_exact_on_dyadic_weights_and_gather(B):
    rng = np.random.default_rng(B)
    w = torch.tensor(rng.integers(0, 9, size=B) / 4.0)          # dyadic rationals: cumsum exact in any order
    w[0] = 1.0
    case = Case(**CASES["C_ecp"], nwalkers=2)
    eng = engine(case)
    u = 0.3125
    neww_ref, inds_ref = O.branch(w, u)
    neww, inds = eng.branch_comb(w.cuda(), u)
    assert np.array_equal(inds.cpu().numpy(), inds_ref.numpy().astype(np.int32))                   # bit-exact
    assert float(neww) == float(neww_ref)
    counts = np.bincount(inds.cpu().numpy(), minlength=B)
    assert counts.sum() == B and np.all(np.abs(counts - w.numpy() / float(neww_ref)) < 1 + 1e-9)
    pos = torch.tensor(rng.normal(size=(B, 12))).cuda()
    np.testing.assert_array_equal(eng.gather_walkers(pos, inds).cpu().numpy(), pos.cpu().numpy()[inds.cpu().numpy()])
    # random weights: statistical agreement (indices may differ only where cumsum rounding differs)
    wr = torch.tensor(rng.uniform(0.2, 1.8, size=B))
    _, ir = O.branch(wr, 0.77)
    _, ig = eng.branch_comb(wr.cuda(), 0.77)
    assert np.mean(ig.cpu().numpy() != ir.numpy()) < 1e-3


@pytest.mark.parametrize("rows,width,picks", [(1, 9, 1), (1025, 9, 1023), (4099, 27, 5000), (3000, 90, 1024 * 4 + 1), (777, 12, 0)])
def test_gather_walkers_ragged_sizes_and_odd_row_widths(rows, width, picks):
    """The gather moves 4 elements per thread 256 apart: sizes that are not multiples of 1024 elements, odd row widths
    (8-byte path), fewer / more picks than rows, and an empty selection."""
    rng = np.random.default_rng(rows * 31 + width)
    eng = engine(Case(**CASES["C_ecp"], nwalkers=2))
    pos = torch.tensor(rng.normal(size=(rows, width))).cuda()
    inds = torch.tensor(rng.integers(0, rows, size=picks), dtype=torch.int32).cuda()
    out = eng.gather_walkers(pos, inds)
    assert out.shape == (picks, width)
    np.testing.assert_array_equal(out.cpu().numpy(), pos.cpu().numpy()[inds.cpu().numpy()])


def test_unsupported_system_and_bad_args_fail_loudly(monkeypatch):
    monkeypatch.setenv("AIQMC_NO_AUTOBUILD", "1")          # (7,3) is not in the prebuilt set; do not compile it here
    case = Case(n=7, natoms=3, spins=[1.] * 4 + [-1.] * 3, seed=3)
    with pytest.raises(aiqmc_b200.lib.AiqmcError):
        engine(case)
    monkeypatch.delenv("AIQMC_NO_AUTOBUILD")
    case = Case(**CASES["C_ecp"], nwalkers=4)
    eng = engine(case)
    with pytest.raises(ValueError):
        eng.ecp = aiqmc_b200.make_ecp(1, list_l=2, **ecp_tables(1))
        eng.local_energy(torch.tensor(case.pos))           # rotation missing


def test_system_outside_the_prebuilt_set_runs_through_its_own_plugin():
    """(N, A) = (3, 2) is not named in csrc/dispatch.h: its kernels come from libaiqmc_sys_3_2.so, built on demand by
    aiqmc_b200.build.ensure_system (tests/test_plugins.py builds it on the CPU box; here it is bound and checked)."""
    case = Case(n=3, natoms=2, spins=[1., 1., -1.], seed=31, nwalkers=9, width=0.8)
    eng = engine(case)
    ph, la, g, lap = eng.psi(torch.tensor(case.pos), mode=2)
    f = lambda x: case.net.apply(case.params, x, case.t_spins, case.t_atoms)[1]
    lat, gt, d2 = O.grad_and_hess_diag(f, torch.tensor(case.pos))
    np.testing.assert_allclose(la.cpu().numpy(), lat.detach().numpy(), rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(g.cpu().numpy(), gt.numpy(), rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(lap.cpu().numpy(), d2.sum(-1).numpy(), rtol=1e-7, atol=1e-7)


@pytest.mark.parametrize("name,rich", [("C_ecp", True), ("C_ecp", False), ("N2_ecp", True), ("odd", True), ("h2like", False)])
def test_cached_quadrature_kernels_match_full_evaluation_kernel(name, rich):
    """The three quadrature kernels against each other: default (thread-per-point on the single-electron-move
    cache, ecp_pt.cuh, N <= 4; else lane-per-electron), forced lane-per-electron (ecp_coop.cuh, stage bit 16)
    and the plain full-evaluation one-thread-per-point kernel (stage bit 8)."""
    case = Case(**CASES[name], nwalkers=33, width=0.8)
    tabs = ecp_tables(case.a, rich=rich)
    eng = engine(case, ecp=aiqmc_b200.make_ecp(case.a, list_l=2, **tabs))
    rot = torch.tensor(O.random_rotations(case.rng, case.B))
    e_def = eng.local_energy(torch.tensor(case.pos), rot, stages=7).cpu().numpy()
    e_coop = eng.local_energy(torch.tensor(case.pos), rot, stages=7 | 16).cpu().numpy()
    e_ref = eng.local_energy(torch.tensor(case.pos), rot, stages=7 | 8).cpu().numpy()
    np.testing.assert_allclose(e_def, e_ref, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(e_coop, e_ref, rtol=1e-10, atol=1e-10)
    # run-to-run bit reproducibility of the deterministic reductions
    assert np.array_equal(e_def, eng.local_energy(torch.tensor(case.pos), rot, stages=7).cpu().numpy())
    assert np.array_equal(e_coop, eng.local_energy(torch.tensor(case.pos), rot, stages=7 | 16).cpu().numpy())


def test_packed_group_quadrature_benzene_matches_full_evaluation_kernel():
    """C6H6 (N=30, A=12, BASELINE configs[4]): the packed lane-per-electron kernel (ecp_grp.cuh, one 30-lane group per
    warp, rows of the 30x30 complex LU in registers) against the plain full-evaluation kernel (stage bit 8)."""
    ring = lambda r, n: [[r * math.cos(2 * math.pi * k / n), r * math.sin(2 * math.pi * k / n), 0.0] for k in range(n)]
    case = Case(n=30, natoms=12, spins=[1.] * 15 + [-1.] * 15, seed=21, atoms=ring(2.640, 6) + ring(4.689, 6),
                charges=[4.0] * 6 + [1.0] * 6, nwalkers=5, width=0.8)
    tabs = ecp_tables(case.a, rich=True)
    eng = engine(case, ecp=aiqmc_b200.make_ecp(case.a, list_l=2, **tabs))
    rot = torch.tensor(O.random_rotations(case.rng, case.B))
    e_def = eng.local_energy(torch.tensor(case.pos), rot, stages=7).cpu().numpy()
    e_ref = eng.local_energy(torch.tensor(case.pos), rot, stages=7 | 8).cpu().numpy()
    assert np.all(np.isfinite(e_ref))
    np.testing.assert_allclose(e_def, e_ref, rtol=1e-9, atol=1e-9)
    assert np.array_equal(e_def, eng.local_energy(torch.tensor(case.pos), rot, stages=7).cpu().numpy())


def test_benzene_gradient_and_sweep_use_the_cached_reverse_pass():
    """N = 30 > 16: gradients come from the primal pass + ONE reverse sweep on the derivative cache
    (deriv_split.cuh: grad_reverse_cached).  Gradient vs torch autograd on the oracle; the sweep's accept mask vs
    the oracle's walkers_update, bit for bit."""
    ring = lambda r, n: [[r * math.cos(2 * math.pi * k / n), r * math.sin(2 * math.pi * k / n), 0.0] for k in range(n)]
    case = Case(n=30, natoms=12, spins=[1.] * 15 + [-1.] * 15, seed=21, atoms=ring(2.640, 6) + ring(4.689, 6),
                charges=[4.0] * 6 + [1.0] * 6, nwalkers=3, width=0.8)
    eng = engine(case)
    ph, la, g = eng.psi(torch.tensor(case.pos), mode=1)[:3]
    f = lambda x: case.net.apply(case.params, x, case.t_spins, case.t_atoms)[1]
    lat, gt, _ = O.value_and_grad(f, torch.tensor(case.pos))
    np.testing.assert_allclose(la.cpu().numpy(), lat.detach().numpy(), rtol=1e-10)
    np.testing.assert_allclose(g.cpu().numpy(), gt.numpy(), rtol=1e-8, atol=1e-9)
    tstep = 0.05
    rand = case.sweep_rand(tstep)
    pos = torch.tensor(case.pos).cuda()
    out = eng.vmc_sweep(pos, rand['gauss1'].cuda(), rand['gauss2'].cuda(), rand['rnd'].cuda(), tstep)
    new_data, aux = O.walkers_update(O.select_output(case.net.apply, 1), case.params, case.oracle_data(), rand, tstep,
                                     3, case.n, case.B, return_aux=True)
    assert np.array_equal(out['accept'].cpu().numpy().astype(bool), aux['accept'].numpy())
    np.testing.assert_allclose(pos.cpu().numpy(), new_data.positions.numpy(), rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("name,rich,tstep,scale", [("C_ecp", False, 0.05, 1.0), ("C_ecp", True, 1.0, -3.0),
                                                   ("N2_ecp", True, 1.0, -3.0), ("h2like", True, 1.0, -5.0)])
def test_dmc_tmoves_match_oracle(name, rich, tstep, scale):
    """compute_tmoves (DMC/Tmoves.py:32-225): selected move per electron bit-exact, acceptance to 1e-8, final
    configuration identical.  The physical carbon table at tstep 0.05 never leaves move 0 ("stay"); the rich
    tables with attractive (negative) coefficients and a long step make non-trivial moves common."""
    case = Case(**CASES[name], nwalkers=24, width=0.6)
    tabs = ecp_tables(case.a, rich=rich)
    tabs['non_local_coes'] = tabs['non_local_coes'] * scale
    rng = case.rng
    rot = O.random_rotations(rng, case.B)
    u = rng.uniform(size=case.B)
    rnd = rng.uniform(size=(case.B, case.n))
    net = aiqmc_b200.make_ai_net(**case.kw)
    tm = aiqmc_b200.compute_tmoves(2, tstep, case.n, case.a, 3, net.apply, tabs['rn_non_local'], tabs['non_local_coes'],
                                   tabs['non_local_exps'])
    data = aiqmc_b200.AINetData(positions=torch.tensor(case.pos), spins=case.t_spins, atoms=case.t_atoms,
                                charges=torch.tensor(case.charges))
    new_pos, acc = tm(data, case.params, dict(rot=torch.tensor(rot), u=torch.tensor(u), rnd=torch.tensor(rnd)))
    eng = net.apply.bind(case.params, case.t_atoms)        # the engine `tm` ran on; the table travels with the closure
    _, _, sel = eng.dmc_tmove(torch.tensor(case.pos), torch.tensor(rot), torch.tensor(u), torch.tensor(rnd), tstep, ecp=tm.ecp)
    ref = O.compute_tmoves(2, tstep, case.n, case.a, 3, O.make_log_network(case.net.apply), tabs['rn_non_local'],
                           tabs['non_local_coes'], tabs['non_local_exps'])
    moved = 0
    for b in range(case.B):
        d1 = O.AINetData(positions=torch.tensor(case.pos[b]), spins=case.t_spins, atoms=case.t_atoms,
                         charges=torch.tensor(case.charges))
        final, acceptance, aux = ref(d1, case.params, dict(rot=torch.tensor(rot[b]), u=u[b], rnd=torch.tensor(rnd[b])))
        assert np.array_equal(sel[b].cpu().numpy(), aux['selected'].numpy()), b          # bit-exact selection
        np.testing.assert_allclose(acc[b].cpu().numpy(), acceptance.numpy().ravel(), rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(new_pos[b].cpu().numpy(), final.numpy(), rtol=1e-10, atol=1e-10)
        moved += int((aux['selected'] > 0).sum())
    if rich:
        assert moved > 0, "test inputs never selected a non-trivial T-move"


def test_dmc_propagate_one_full_step():
    """dmc_propagate_run (DMC/dmc.py:72-93): T-move -> drift-diffusion -> E_L(old), E_L(new) -> S -> weights."""
    tstep = 0.05
    case = Case(**CASES["C_ecp"], nwalkers=20, width=0.7)
    tabs = ecp_tables(1, rich=True)
    rng = case.rng
    key = dict(tmove=dict(rot=torch.tensor(O.random_rotations(rng, case.B)), u=torch.tensor(rng.uniform(size=case.B)),
                          rnd=torch.tensor(rng.uniform(size=(case.B, case.n)))),
               sweep=case.sweep_rand(tstep), rot=torch.tensor(O.random_rotations(rng, case.B)))
    weights = torch.tensor(rng.uniform(0.5, 1.5, size=case.B))
    branchcut = torch.full((case.B,), 10 * 0.3)
    e_trial, e_est = -5.39, -5.41
    ref = O.dmc_propagate(case.net.apply, tstep, case.n, 1, 3, case.B, case.charges, **tabs)
    e_ref, w_ref, d_ref = ref(case.params, key, case.oracle_data(), weights, branchcut, e_trial, e_est)
    net = aiqmc_b200.make_ai_net(**case.kw)
    run = aiqmc_b200.dmc_propagate(net.apply, net.apply, tstep, case.n, 1, 3, case.B, case.charges, **tabs)
    data = aiqmc_b200.AINetData(positions=torch.tensor(case.pos), spins=case.t_spins, atoms=case.t_atoms,
                                charges=torch.tensor(case.charges))
    e_gpu, w_gpu, d_gpu = run(case.params, key, data, weights, branchcut, e_trial, e_est)
    np.testing.assert_allclose(d_gpu.positions.cpu().numpy(), d_ref.positions.numpy(), rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(e_gpu.cpu().numpy(), e_ref.numpy(), atol=1e-5, rtol=0)        # north_star: 1e-5 Ha
    np.testing.assert_allclose(e_gpu.cpu().numpy(), e_ref.numpy(), atol=1e-8, rtol=1e-9)
    np.testing.assert_allclose(w_gpu.cpu().numpy(), w_ref.numpy(), rtol=1e-9)


def test_reconfigure_and_energy_estimators():
    """main_dmc.py:218-242 + estimate_energy.py: unique survivors + noise padding, weighted average, E_trial."""
    rng = np.random.default_rng(9)
    B = 300
    case = Case(**CASES["C_ecp"], nwalkers=2)
    eng = engine(case)
    pos = torch.tensor(rng.normal(size=(B, 12)))
    w = torch.tensor(rng.uniform(0.0, 2.5, size=B))
    noise = torch.tensor(rng.uniform(size=(B, 12)))
    _, inds_ref = O.branch(w, 0.61)
    ref, nuniq_ref = O.reconfigure(pos, inds_ref, noise)
    _, inds = eng.branch_comb(w.cuda(), 0.61)
    got, nuniq = aiqmc_b200.reconfigure(eng, pos.cuda(), inds, noise)
    assert nuniq == nuniq_ref and nuniq < B
    np.testing.assert_array_equal(got.cpu().numpy(), ref.numpy())
    e = torch.tensor(rng.normal(-5.4, 0.2, size=(3, 4, B)))
    ww = torch.tensor(rng.uniform(0.5, 1.5, size=(3, 4, B)))
    np.testing.assert_allclose(float(aiqmc_b200.estimate_energy(e.cuda(), ww.cuda())), np.average(e.numpy(), weights=ww.numpy()), rtol=1e-13)
    np.testing.assert_allclose(float(aiqmc_b200.trial_energy(-5.4, w.cuda(), 1.0)), -5.4 - math.log(float(w.mean())), rtol=1e-13)
