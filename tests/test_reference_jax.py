"""Parity pin against the GENUINE reference (unmodified AIQMCrelease3 modules under JAX), whenever JAX is importable.

In the image this repository was built in `import jax` fails, so these tests SKIP -- loudly, with the probe's reason --
and parity stays "unpinned" (the oracle is then pinned only by the reference's known answers, tests/test_oracle_pins.py).
On any box with jax + the reference sources (baseline/_ref, /root/reference or $AIQMC_REFERENCE_ROOT) they run and
assert oracle == reference on the committed golden inputs: log|psi| / phase, the post-sweep positions and accept
pattern on identical gauss/uniform arrays, and the ccECP / all-electron local energy on identical rotations.
tests/golden/make_golden_reference.py regenerates tests/golden/*.npz from the genuine reference in the same situation.
"""
import os
import warnings

import numpy as np
import pytest
import torch

from common import CASES, Case, O, ecp_tables
from oracle import reference_jax as RJ

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
OK, WHY, REF_ROOT = RJ.probe()
needs_reference = pytest.mark.skipif(not OK, reason=f"GENUINE-REFERENCE PARITY NOT RUN: {WHY}")
TSTEP = 0.05


def test_probe_is_explicit():
    """The probe never fails silently: either the reference is usable or the reason says what is missing."""
    assert isinstance(OK, bool) and isinstance(WHY, str) and WHY
    if not OK:
        warnings.warn(f"parity is pinned by known answers only: {WHY}")


def _load(name):
    g = dict(np.load(os.path.join(GOLD, name + ".npz")))
    case = Case(**CASES[name], nwalkers=g["pos"].shape[0], width=0.8)
    h = RJ.ReferenceHarness(case.kw, case.params, case.atoms, case.charges, case.spins, x64=True)
    return g, case, h


@needs_reference
@pytest.mark.parametrize("name", ["C_ecp", "C_ae", "N2_ecp"])
def test_oracle_equals_reference_signed_network(name):
    g, case, h = _load(name)
    ph, la = h.psi(g["pos"])
    np.testing.assert_allclose(la, g["logabs"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(np.angle(np.exp(1j * (ph - g["phase"]))), 0.0, atol=1e-8)


@needs_reference
@pytest.mark.parametrize("name", ["C_ecp", "C_ae", "N2_ecp"])
def test_oracle_equals_reference_sweep(name):
    g, case, h = _load(name)
    new_pos = h.walkers_update(g["pos"], {k: g[k] for k in ("gauss1", "gauss2", "rnd")}, TSTEP)
    np.testing.assert_allclose(new_pos, g["pos_after_sweep"], rtol=1e-9, atol=1e-9)
    moved = np.any(np.abs(new_pos.reshape(case.B, case.n, 3) - g["pos"].reshape(case.B, case.n, 3)) > 0, axis=-1)
    assert np.array_equal(moved, g["accept"])                                  # accept mask bit-exact


@needs_reference
@pytest.mark.parametrize("name", ["C_ecp", "N2_ecp"])
def test_oracle_equals_reference_ecp_energy(name):
    g, case, h = _load(name)
    e = h.local_energy_ecp(g["pos"], g["rot"], ecp_tables(case.a, rich=True))
    np.testing.assert_allclose(e, g["e_l"], atol=1e-7, rtol=1e-8)


@needs_reference
def test_oracle_equals_reference_ae_energy():
    g, case, h = _load("C_ae")
    np.testing.assert_allclose(h.local_energy_ae(g["pos"]), g["e_l"], atol=1e-7, rtol=1e-8)
