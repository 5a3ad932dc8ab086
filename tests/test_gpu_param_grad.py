"""GPU parity for SURVEY 8f N1: aiqmc_psi_param_grad (csrc/param_grad.cuh) and the make_loss mirror of
Loss/pploss.py:137-223 against torch.autograd on the oracle."""
import numpy as np
import pytest
import torch

from common import CASES, Case, O, ecp_tables
from test_param_grad_host import oracle_param_grad, tree_leaves

import aiqmc_b200

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,B", [("C_ecp", 70), ("C_ae", 33), ("N2_ecp", 40), ("odd", 5), ("h2like", 64)])
def test_param_gradient_kernel_matches_autograd(name, B):
    """Weighted sum over walkers of d(alpha log|psi| + beta phase)/d params; B is not a multiple of 32 so the tail
    lanes (zero seed) are exercised; bit-reproducible run to run (fixed-order reductions)."""
    case = Case(**CASES[name], nwalkers=B)
    eng = aiqmc_b200.WalkerEngine(case.spec(), case.params)
    rng = np.random.default_rng(7)
    alpha, beta = rng.normal(size=B), rng.normal(size=B)
    g, ph, la = eng.param_grad(torch.tensor(case.pos), alpha, beta)
    g2, _, _ = eng.param_grad(torch.tensor(case.pos), alpha, beta)
    assert torch.equal(g, g2)
    pht, lat = case.net.apply(case.params, torch.tensor(case.pos), case.t_spins, case.t_atoms)
    np.testing.assert_allclose(la.cpu().numpy(), lat.numpy(), rtol=1e-6, atol=1e-12)      # north_star tolerance
    np.testing.assert_allclose(la.cpu().numpy(), lat.numpy(), rtol=1e-10, atol=1e-11)
    got = dict(tree_leaves(aiqmc_b200.unpack_param_grad(eng.layout, g.cpu().numpy(), case.params, case.spec())))
    ref = oracle_param_grad(case, case.pos, alpha, beta)
    assert set(got) == set(ref)
    for key, r in ref.items():
        if r.size:
            np.testing.assert_allclose(np.asarray(got[key]).reshape(r.shape), r, rtol=1e-7,
                                       atol=1e-8 * max(1.0, float(np.abs(r).max())), err_msg=key)


def test_param_gradient_large_batch_linearity_and_empty():
    """65,536 walkers (BASELINE configs[1] size): the gradient is linear in the seeds, so grad(alpha1 + alpha2) must
    equal grad(alpha1) + grad(alpha2) to rounding, and a zero seed gives exactly zero; an empty batch returns zeros."""
    case = Case(**CASES["C_ecp"], nwalkers=65536)
    eng = aiqmc_b200.WalkerEngine(case.spec(), case.params)
    rng = np.random.default_rng(9)
    pos = torch.tensor(case.pos).cuda()
    a1, a2, b1 = (torch.tensor(rng.normal(size=case.B)).cuda() for _ in range(3))
    zero = torch.zeros(case.B, dtype=torch.float64, device="cuda")
    g1, _, _ = eng.param_grad(pos, a1, b1)
    g2, _, _ = eng.param_grad(pos, a2, zero)
    g12, _, _ = eng.param_grad(pos, a1 + a2, b1)
    scale = float(g12.abs().max())
    assert scale > 0 and torch.isfinite(g12).all()
    np.testing.assert_allclose((g1 + g2).cpu().numpy(), g12.cpu().numpy(), rtol=0, atol=1e-9 * scale)
    g0, _, _ = eng.param_grad(pos, zero, zero)
    assert float(g0.abs().max()) == 0.0
    ge, _, _ = eng.param_grad(pos[:0], zero[:0], zero[:0])
    assert ge.shape == (eng.layout.total,) and float(ge.abs().max()) == 0.0


def test_benzene_param_gradient_runs_on_the_derivative_cache():
    """N = 30 > 16 (BASELINE configs[4]): primal pass into the derivative cache + the sweep on the cache; every leaf
    against torch autograd, and bit-reproducible."""
    case = Case(n=30, natoms=12, spins=[1.] * 15 + [-1.] * 15, seed=21, nwalkers=35, charges=[4.0] * 6 + [1.0] * 6)
    eng = aiqmc_b200.WalkerEngine(case.spec(), case.params)
    rng = np.random.default_rng(2)
    alpha, beta = rng.normal(size=case.B), rng.normal(size=case.B)
    g, ph, la = eng.param_grad(torch.tensor(case.pos), alpha, beta)
    g2, _, _ = eng.param_grad(torch.tensor(case.pos), alpha, beta)
    assert torch.equal(g, g2)
    _, lat = case.net.apply(case.params, torch.tensor(case.pos), case.t_spins, case.t_atoms)
    np.testing.assert_allclose(la.cpu().numpy(), lat.numpy(), rtol=1e-10)
    got = dict(tree_leaves(aiqmc_b200.unpack_param_grad(eng.layout, g.cpu().numpy(), case.params, case.spec())))
    ref = oracle_param_grad(case, case.pos, alpha, beta)
    for key, r in ref.items():
        if r.size:
            np.testing.assert_allclose(np.asarray(got[key]).reshape(r.shape), r, rtol=1e-7,
                                       atol=1e-8 * max(1.0, float(np.abs(r).max())), err_msg=key)


@pytest.mark.parametrize("clip,median", [(0.0, True), (1.0, True), (1.0, False)])
def test_make_loss_value_and_grad_matches_oracle(clip, median):
    """Drop-in for make_loss (pploss.py:137-223): loss, variance, clipped centre and the parameter gradient that
    jax.value_and_grad obtains through the custom JVP, ccECP carbon, complex local energies."""
    case = Case(**CASES["C_ecp"], nwalkers=48, width=0.7)
    tabs = ecp_tables(1, rich=True)
    net = aiqmc_b200.make_ai_net(**case.kw)
    le = aiqmc_b200.local_energy(net.apply, case.charges, lognetwork=None, natoms=1, nelectrons=4, ndim=3, list_l=2, **tabs)
    loss_fn = aiqmc_b200.make_loss(net.apply, le, clip_local_energy=clip, clip_from_median=median,
                                   center_at_clipped_energy=True, complex_output=True)
    rot = torch.tensor(O.random_rotations(case.rng, case.B))
    data = aiqmc_b200.AINetData(positions=torch.tensor(case.pos), spins=case.t_spins, atoms=case.t_atoms,
                                charges=torch.tensor(case.charges))
    (loss, aux), grads = loss_fn.value_and_grad(case.params, rot, data)
    le_o = O.local_energy_ecp(case.net.apply, O.make_log_network(case.net.apply), case.charges, None, tabs['rn_local'],
                              tabs['local_coes'], tabs['local_exps'], tabs['rn_non_local'], tabs['non_local_coes'],
                              tabs['non_local_exps'], 1, 4, 3, 2)
    loss_o = O.make_loss(case.net.apply, lambda p, k, d: le_o(p, k, case.oracle_data(batched_static=False)),
                         clip_local_energy=clip, clip_from_median=median, center_at_clipped_energy=True, complex_output=True)
    (lo, auxo), go = loss_o.value_and_grad(case.params, rot, case.oracle_data())
    np.testing.assert_allclose(float(loss), float(lo), atol=1e-5)                       # north_star: 1e-5 Ha
    np.testing.assert_allclose(float(aux.variance), float(auxo['variance']), rtol=1e-7, atol=1e-8)
    got, ref = dict(tree_leaves(grads)), dict(tree_leaves(go))
    assert set(got) == set(ref)
    for key, r in ref.items():
        r = r.numpy()
        if r.size:
            np.testing.assert_allclose(np.asarray(got[key]).reshape(r.shape), r, rtol=1e-6,
                                       atol=1e-7 * max(1.0, float(np.abs(r).max())), err_msg=key)
