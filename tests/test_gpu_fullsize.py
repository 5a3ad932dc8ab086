"""GPU tests at BASELINE.json's full size (carbon ccECP, 65,536 walkers per GPU) through size-independent
properties, plus an oracle spot check on a random subset of the same batch."""
import numpy as np
import pytest
import torch

from common import O, ecp_tables

import aiqmc_b200
import common as bench

pytestmark = pytest.mark.gpu
B = 65536


@pytest.fixture(scope="module")
def setup():
    case, tabs = bench.build_bench_case(B)
    eng = aiqmc_b200.WalkerEngine(case.spec(), case.params, ecp=aiqmc_b200.make_ecp(1, list_l=2, **tabs))
    rng = np.random.default_rng(77)
    rot = torch.from_numpy(bench.random_rot(rng, B)).cuda()
    pos = torch.from_numpy(case.pos.copy()).cuda()
    return case, tabs, eng, rng, pos, rot


def test_local_energy_walkers_are_independent_and_reduction_is_deterministic(setup):
    case, tabs, eng, rng, pos, rot = setup
    e = eng.local_energy(pos, rot).clone()
    assert torch.isfinite(torch.view_as_real(e)).all()
    assert torch.equal(e, eng.local_energy(pos, rot))                      # bit-reproducible
    idx = torch.from_numpy(rng.choice(B, size=777, replace=False)).cuda()   # ragged subset, other CTA packing
    e_sub = eng.local_energy(pos[idx].contiguous(), rot[idx].contiguous())
    assert torch.equal(e_sub, e[idx])                                       # a walker's energy does not depend on its batch
    perm = torch.from_numpy(rng.permutation(B)).cuda()
    assert torch.equal(eng.local_energy(pos[perm].contiguous(), rot[perm].contiguous()), e[perm])
    # statistics are additive over any split of the batch (what the multi-GPU all-reduce relies on)
    s_all = eng.energy_stats(e).cpu().numpy()
    s_parts = sum(eng.energy_stats(e[lo:hi].contiguous()).cpu().numpy() for lo, hi in [(0, 1000), (1000, 40001), (40001, B)])
    np.testing.assert_allclose(s_all, s_parts, rtol=1e-12)
    assert s_all[3] == B
    # oracle spot check on 48 walkers of the batch
    pick = idx[:48].cpu().numpy()
    le = O.local_energy_ecp(case.net.apply, O.make_log_network(case.net.apply), case.charges, None, tabs['rn_local'],
                            tabs['local_coes'], tabs['local_exps'], tabs['rn_non_local'], tabs['non_local_coes'],
                            tabs['non_local_exps'], 1, case.n, 3, 2)
    ref, _ = le(case.params, rot[idx[:48]].cpu(), case.oracle_data(pos=case.pos[pick], batched_static=False))
    np.testing.assert_allclose(e[idx[:48]].cpu().numpy(), ref.numpy(), atol=1e-5, rtol=0)      # north_star: 1e-5 Ha
    np.testing.assert_allclose(e[idx[:48]].cpu().numpy(), ref.numpy(), atol=1e-7, rtol=1e-9)


def test_sweep_full_size_properties(setup):
    case, tabs, eng, rng, pos, rot = setup
    tstep = bench.TSTEP
    r = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in bench.make_rand(rng, B, case.n, tstep).items()}
    p1, p2 = pos.clone(), pos.clone()
    o1 = eng.vmc_sweep(p1, r["gauss1"], r["gauss2"], r["rnd"], tstep, want_drift=True, want_aux=True)
    o2 = eng.vmc_sweep(p2, r["gauss1"], r["gauss2"], r["rnd"], tstep, want_drift=True, want_aux=True)
    assert torch.equal(p1, p2) and torch.equal(o1["accept"], o2["accept"]) and torch.equal(o1["aux"], o2["aux"])
    acc = o1["accept"].bool()
    assert 0.2 < float(acc.float().mean()) < 0.999
    # accepted electrons moved, rejected ones kept their coordinates bit for bit
    moved = (p1.reshape(B, case.n, 3) != pos.reshape(B, case.n, 3)).any(-1)
    assert torch.equal(moved, acc)
    # the batch-global limdrift sum (quirk Q6) equals the sum of squares of the gradients the psi entry point returns
    _, _, g = eng.psi(pos, mode=1)
    np.testing.assert_allclose(float(o1["aux"][2]), float((g ** 2).sum()), rtol=1e-11)
    # drift written out = grad * taueff(v2), VMCmcstep.py:11-14
    v2 = float((g ** 2).sum())
    te = (np.sqrt(1 + 2 * tstep * 0.25 * v2) - 1) / (0.25 * v2)
    np.testing.assert_allclose(o1["grad_eff_old"].cpu().numpy(), (g * te).cpu().numpy(), rtol=1e-10, atol=1e-14)
    # sum of the new coordinates (the numerator of tdamp, quirk Q19)
    np.testing.assert_allclose(float(o1["aux"][0]), float(p1.sum()), rtol=1e-9)


def test_gradient_and_laplacian_paths_agree_at_full_size(setup):
    case, tabs, eng, rng, pos, rot = setup
    ph2, la2, g2, lp2 = eng.psi(pos, mode=2)            # two-pass tangent path
    ph1, la1, g1 = eng.psi(pos, mode=1)                 # fused reverse path
    ph0, la0 = eng.psi(pos, mode=0)                     # value-only LU
    assert torch.isfinite(lp2).all()
    np.testing.assert_allclose(g1.cpu().numpy(), g2.cpu().numpy(), rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(la1.cpu().numpy(), la2.cpu().numpy(), rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(la0.cpu().numpy(), la2.cpu().numpy(), rtol=1e-12, atol=1e-12)


def test_host_step_pipeline_matches_device_resident_steps():
    """aiqmc_b200.HostStepPipeline (host buffers in / out every step, copies overlapped with the kernels) gives exactly
    the positions and statistics of the same steps run on device-resident data."""
    import common as bench
    B, nsteps = 4096, 3
    case, tabs = bench.build_bench_case(B)
    eng = aiqmc_b200.WalkerEngine(case.spec(), case.params, ecp=aiqmc_b200.make_ecp(1, list_l=2, **tabs))
    rng = np.random.default_rng(3)
    host_sets = []
    for _ in range(nsteps):
        r = bench.make_rand(rng, B, case.n, bench.TSTEP)
        r["rot"] = bench.random_rot(rng, B)
        host_sets.append({k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in r.items()})
    pos_host = torch.from_numpy(case.pos.copy()).pin_memory()
    stats_host = torch.empty(4, dtype=torch.float64).pin_memory()
    aiqmc_b200.HostStepPipeline(eng, bench.TSTEP).run(pos_host, host_sets, stats_host)
    pos = torch.from_numpy(case.pos.copy()).cuda()
    for s in host_sets:
        d = {k: v.cuda() for k, v in s.items()}
        eng.vmc_sweep(pos, d["gauss1"], d["gauss2"], d["rnd"], bench.TSTEP, want_accept=False)
        stats = eng.energy_stats(eng.local_energy(pos, d["rot"]))
    assert torch.equal(pos.cpu(), pos_host)
    assert torch.equal(stats.cpu(), stats_host)
