"""CPU check of csrc/param_grad.cuh (the d log psi / d params side of Loss/pploss.py:186-223) compiled
for the host (tests/hostcore, test-only): per-walker d(alpha log|psi| + beta phase)/d(params), mapped back to the
reference's pytree by system.unpack_param_grad, against torch.autograd on the oracle network."""
import ctypes as C

import numpy as np
import pytest
import torch

from common import CASES, Case
import hostcore_util as H

import aiqmc_b200


@pytest.fixture(scope="module")
def hostlib():
    return H.load()


def tree_leaves(t, prefix=""):
    if isinstance(t, dict):
        for k in sorted(t):
            yield from tree_leaves(t[k], prefix + "/" + str(k))
    elif isinstance(t, (list, tuple)):
        for i, v in enumerate(t):
            yield from tree_leaves(v, prefix + "/" + str(i))
    else:
        yield prefix, t


def oracle_param_grad(case, pos, alpha, beta):
    """sum_w alpha_w d log|psi_w| + beta_w d phase_w by autograd through the oracle (float64)."""
    leaves = list(tree_leaves(case.params))
    req = [v.clone().requires_grad_(True) for _, v in leaves]
    it = iter(req)

    def rebuild(t):
        if isinstance(t, dict):
            return {k: rebuild(t[k]) for k in sorted(t)}
        if isinstance(t, (list, tuple)):
            return [rebuild(v) for v in t]
        return next(it)

    params = rebuild(case.params)
    ph, la = case.net.apply(params, torch.tensor(pos), case.t_spins, case.t_atoms)
    obj = (torch.tensor(alpha) * la + torch.tensor(beta) * ph).sum()
    grads = torch.autograd.grad(obj, req, allow_unused=True)
    return {name: (g.numpy() if g is not None else np.zeros(tuple(v.shape))) for (name, v), g in zip(leaves, grads)}


@pytest.mark.parametrize("name", list(CASES))
def test_param_gradient_matches_autograd(hostlib, name):
    case = Case(**CASES[name], nwalkers=6)
    lay = aiqmc_b200.system.AiqmcLayout()
    hostlib.hc_layout(case.n, case.a, C.byref(lay))
    packed = aiqmc_b200.pack_params(lay, case.params, case.spec())
    rng = np.random.default_rng(3)
    alpha, beta = rng.normal(size=case.B), rng.normal(size=case.B)
    per_walker = H.host_param_grad(hostlib, case.spec().c_struct(), packed, case.pos, alpha, beta)
    got = aiqmc_b200.unpack_param_grad(lay, per_walker.sum(0), case.params, case.spec())
    ref = oracle_param_grad(case, case.pos, alpha, beta)
    got_leaves = dict(tree_leaves(got))
    assert set(got_leaves) == set(ref)
    for key, r in ref.items():
        g = np.asarray(got_leaves[key]).reshape(r.shape)
        if r.size == 0:
            continue
        scale = max(1.0, float(np.abs(r).max()))
        np.testing.assert_allclose(g, r, rtol=1e-8, atol=1e-9 * scale, err_msg=key)
    # the same sweep run ON the derivative cache (the kernels' path for N > 16): identical to rounding
    pw_cached = H.host_param_grad(hostlib, case.spec().c_struct(), packed, case.pos, alpha, beta, cached=True)
    np.testing.assert_allclose(pw_cached, per_walker, rtol=1e-11, atol=1e-12 * max(1.0, float(np.abs(per_walker).max())))
    # the phase has a branch cut but a smooth gradient: a pure-phase seed must work on its own
    pw_phase = H.host_param_grad(hostlib, case.spec().c_struct(), packed, case.pos, np.zeros(case.B), np.ones(case.B))
    ref_phase = oracle_param_grad(case, case.pos, np.zeros(case.B), np.ones(case.B))
    got_phase = dict(tree_leaves(aiqmc_b200.unpack_param_grad(lay, pw_phase.sum(0), case.params, case.spec())))
    for key, r in ref_phase.items():
        if r.size == 0:
            continue
        np.testing.assert_allclose(np.asarray(got_phase[key]).reshape(r.shape), r, rtol=1e-8,
                                   atol=1e-9 * max(1.0, float(np.abs(r).max())), err_msg=key)


def test_benzene_param_gradient_on_the_cache_matches_autograd(hostlib):
    """N = 30, A = 12 (BASELINE configs[4]): the cached sweep against torch autograd, every leaf."""
    case = Case(n=30, natoms=12, spins=[1.] * 15 + [-1.] * 15, seed=21, nwalkers=2, charges=[4.0] * 6 + [1.0] * 6)
    lay = aiqmc_b200.system.AiqmcLayout()
    hostlib.hc_layout(30, 12, C.byref(lay))
    packed = aiqmc_b200.pack_params(lay, case.params, case.spec())
    alpha, beta = np.array([0.7, -1.3]), np.array([0.4, 0.9])
    pw = H.host_param_grad(hostlib, case.spec().c_struct(), packed, case.pos, alpha, beta, cached=True)
    got = dict(tree_leaves(aiqmc_b200.unpack_param_grad(lay, pw.sum(0), case.params, case.spec())))
    ref = oracle_param_grad(case, case.pos, alpha, beta)
    assert set(got) == set(ref)
    for key, r in ref.items():
        if r.size:
            np.testing.assert_allclose(np.asarray(got[key]).reshape(r.shape), r, rtol=1e-7,
                                       atol=1e-8 * max(1.0, float(np.abs(r).max())), err_msg=key)
