"""Frozen vectors (tests/golden/*.npz, written by tests/golden/make_golden.py from the float64 oracle -- the reference
itself cannot run here, see the script's header).  CPU: the oracle still reproduces them (guards the checker against
drift); GPU: the CUDA path reproduces them through the C ABI."""
import os

import numpy as np
import pytest
import torch

from common import CASES, Case, O, ecp_tables

import aiqmc_b200

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["C_ecp", "C_ae", "N2_ecp"]
TSTEP = 0.05


def load(name):
    g = dict(np.load(os.path.join(GOLD, name + ".npz")))
    case = Case(**CASES[name], nwalkers=g["pos"].shape[0], width=0.8)
    assert int(g["seed"]) == CASES[name]["seed"]
    np.testing.assert_array_equal(case.pos, g["pos"])          # same seeded inputs as when the file was written
    return g, case


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden_vectors(name):
    g, case = load(name)
    ph, la = case.net.apply(case.params, torch.tensor(g["pos"]), case.t_spins, case.t_atoms)
    np.testing.assert_allclose(la.numpy(), g["logabs"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(np.angle(np.exp(1j * (ph.numpy() - g["phase"]))), 0.0, atol=1e-12)
    rand = {k: torch.tensor(g[k]) for k in ("gauss1", "gauss2", "rnd")}
    new_data, aux = O.walkers_update(O.select_output(case.net.apply, 1), case.params, case.oracle_data(), rand, TSTEP, 3,
                                     case.n, case.B, return_aux=True)
    assert np.array_equal(aux["accept"].numpy(), g["accept"])
    np.testing.assert_allclose(new_data.positions.numpy(), g["pos_after_sweep"], rtol=1e-12, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_path_reproduces_golden_vectors(name):
    g, case = load(name)
    with_ecp = "rot" in g
    ecp = aiqmc_b200.make_ecp(case.a, list_l=2, **ecp_tables(case.a, rich=True)) if with_ecp else None
    eng = aiqmc_b200.WalkerEngine(case.spec(), case.params, ecp=ecp)
    pos = torch.tensor(g["pos"]).cuda()
    ph, la, grad, lap = eng.psi(pos, mode=2)
    np.testing.assert_allclose(la.cpu().numpy(), g["logabs"], rtol=1e-6, atol=0)             # north_star: 1e-6 relative
    np.testing.assert_allclose(la.cpu().numpy(), g["logabs"], rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(np.angle(np.exp(1j * (ph.cpu().numpy() - g["phase"]))), 0.0, atol=1e-10)
    np.testing.assert_allclose(grad.cpu().numpy(), g["grad"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(lap.cpu().numpy(), g["lap"], rtol=1e-8, atol=1e-8)
    e = eng.local_energy(pos, torch.tensor(g["rot"])) if with_ecp else eng.local_energy(pos)
    np.testing.assert_allclose(e.cpu().numpy(), g["e_l"], atol=1e-5, rtol=0)                  # north_star: 1e-5 Ha
    np.testing.assert_allclose(e.cpu().numpy(), g["e_l"], atol=1e-8, rtol=1e-9)
    p2 = pos.clone()
    out = eng.vmc_sweep(p2, torch.tensor(g["gauss1"]).cuda(), torch.tensor(g["gauss2"]).cuda(),
                        torch.tensor(g["rnd"]).cuda(), TSTEP)
    assert np.array_equal(out["accept"].cpu().numpy().astype(bool), g["accept"])             # bit-exact accept mask
    np.testing.assert_allclose(p2.cpu().numpy(), g["pos_after_sweep"], rtol=1e-10, atol=1e-10)
