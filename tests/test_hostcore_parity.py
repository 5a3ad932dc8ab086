"""CPU checks of the kernel arithmetic: csrc/psi_core.cuh compiled for the host (tests/hostcore,
test-only) against the oracle.  log|psi| to 1e-6 relative, gradient / Laplacian far tighter."""
import ctypes as C

import numpy as np
import pytest
import torch

from common import CASES, Case, O
import hostcore_util as H

import aiqmc_b200


@pytest.fixture(scope="module")
def hostlib():
    return H.load()


@pytest.mark.parametrize("name", list(CASES))
def test_value_grad_laplacian_match_oracle(hostlib, name):
    case = Case(**CASES[name], nwalkers=5)
    lay = aiqmc_b200.system.AiqmcLayout()
    hostlib.hc_layout(case.n, case.a, C.byref(lay))
    packed = aiqmc_b200.pack_params(lay, case.params, case.spec())
    ph, la, g, lp = H.host_psi(hostlib, case.spec().c_struct(), packed, case.pos, 2)
    f = lambda x: case.net.apply(case.params, x, case.t_spins, case.t_atoms)[1]
    pht, lat = case.net.apply(case.params, torch.tensor(case.pos), case.t_spins, case.t_atoms)
    _, gt, dt = O.grad_and_hess_diag(f, torch.tensor(case.pos))
    np.testing.assert_allclose(la, lat.numpy(), rtol=1e-6, atol=1e-12)           # north_star tolerance
    np.testing.assert_allclose(la, lat.numpy(), rtol=1e-11, atol=1e-12)          # what fp64 actually gives
    np.testing.assert_allclose(np.angle(np.exp(1j * (ph - pht.numpy()))), 0.0, atol=1e-11)
    np.testing.assert_allclose(g, gt.numpy(), rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(lp, dt.sum(-1).numpy(), rtol=1e-9, atol=1e-9)
    # value-only and gradient-only code paths agree with the Laplacian path
    ph0, la0, _, _ = H.host_psi(hostlib, case.spec().c_struct(), packed, case.pos, 0)
    np.testing.assert_allclose(la0, la, rtol=1e-13, atol=1e-13)
    _, _, g1, _ = H.host_psi(hostlib, case.spec().c_struct(), packed, case.pos, 1)
    np.testing.assert_allclose(g1, g, rtol=1e-12, atol=1e-12)
    # the two-pass derivative path the kernels use (deriv_split.cuh: primal cache + tangent pass)
    ph3, la3, g3, _ = H.host_psi(hostlib, case.spec().c_struct(), packed, case.pos, 3)
    ph4, la4, g4, lp4 = H.host_psi(hostlib, case.spec().c_struct(), packed, case.pos, 4)
    np.testing.assert_allclose(la3, la, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(g3, gt.numpy(), rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(g4, g3, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(lp4, dt.sum(-1).numpy(), rtol=1e-9, atol=1e-9)
    # the fused forward + reverse (adjoint) gradient the sweep kernels use
    ph5, la5, g5, _ = H.host_psi(hostlib, case.spec().c_struct(), packed, case.pos, 5)
    np.testing.assert_allclose(la5, la, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(g5, gt.numpy(), rtol=1e-9, atol=1e-10)
    # the reverse sweep on the derivative cache (systems with N > 16 in the kernels; any system here)
    ph6, la6, g6, _ = H.host_psi(hostlib, case.spec().c_struct(), packed, case.pos, 6)
    np.testing.assert_allclose(la6, la, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(g6, gt.numpy(), rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(g6, g5, rtol=1e-12, atol=1e-12)


def test_benzene_sized_system_value(hostlib):
    n, a = 30, 12
    case = Case(n=n, natoms=a, spins=[1.] * 15 + [-1.] * 15, seed=21, nwalkers=2,
                charges=[4.0] * 6 + [1.0] * 6)
    lay = aiqmc_b200.system.AiqmcLayout()
    hostlib.hc_layout(n, a, C.byref(lay))
    packed = aiqmc_b200.pack_params(lay, case.params, case.spec())
    ph, la, g, lp = H.host_psi(hostlib, case.spec().c_struct(), packed, case.pos, 1)
    pht, lat = case.net.apply(case.params, torch.tensor(case.pos), case.t_spins, case.t_atoms)
    np.testing.assert_allclose(la, lat.numpy(), rtol=1e-10)
    f = lambda x: case.net.apply(case.params, x, case.t_spins, case.t_atoms)[1]
    _, gt, _ = O.value_and_grad(f, torch.tensor(case.pos))
    np.testing.assert_allclose(g, gt.numpy(), rtol=1e-8, atol=1e-9)
    _, la6, g6, _ = H.host_psi(hostlib, case.spec().c_struct(), packed, case.pos, 6)     # the path the kernels take at N = 30
    np.testing.assert_allclose(la6, lat.numpy(), rtol=1e-10)
    np.testing.assert_allclose(g6, gt.numpy(), rtol=1e-8, atol=1e-9)


def test_layout_matches_library_and_is_dense(hostlib):
    for n, a in [(4, 1), (10, 2), (30, 12)]:
        l1, l2 = aiqmc_b200.system.AiqmcLayout(), aiqmc_b200.lib.param_layout(n, a)
        hostlib.hc_layout(n, a, C.byref(l1))
        assert bytes(l1) == bytes(l2)
        nparams = (n * (12 * a + 8) + n * (3 * a + 2) + (3 * a + 2) * 4 + 4 + 2 * (n * 20 + n * 5 + 24) + 2 * 20 +
                   (4 * a + 2) * 6 + 6 + 2 * 42 + 2 * 10 * n + 6 * n)
        extras = 2 * n * n + n * a + 2 * a + 2 * n * a * 3 + n + n * a + 3 * a + a
        assert l2.total == nparams + extras
