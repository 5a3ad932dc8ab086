#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the GENUINE reference (unmodified AIQMCrelease3 under JAX, float64) instead of
the oracle.  Usable only where `import jax` works (oracle/reference_jax.py: probe()); in the image this repository was
built in it prints the reason and exits 2, and the committed files remain the oracle-generated ones of make_golden.py.
Same seeded inputs and the same keys as make_golden.py, so tests/test_golden.py runs unchanged on the result and then
pins BOTH the oracle and the CUDA path to the reference's own numbers."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from common import CASES, Case, O, ecp_tables  # noqa: E402
from oracle import reference_jax as RJ  # noqa: E402

TSTEP = 0.05


def make(name, nwalkers, with_ecp):
    case = Case(**CASES[name], nwalkers=nwalkers, width=0.8)
    h = RJ.ReferenceHarness(case.kw, case.params, case.atoms, case.charges, case.spins, x64=True)
    old = dict(np.load(os.path.join(HERE, f"{name}.npz")))          # derivative entries have no reference closure: kept
    rng = np.random.default_rng(1234)
    out = dict(pos=case.pos, seed=np.int64(CASES[name]["seed"]), grad=old["grad"], lap=old["lap"])
    ph, la = h.psi(case.pos)
    out.update(phase=ph, logabs=la)
    rand = dict(gauss1=rng.standard_normal((nwalkers, 3 * case.n)) * TSTEP ** 0.5,
                gauss2=rng.standard_normal((nwalkers, case.n, 3 * case.n)) * TSTEP ** 0.5,
                rnd=rng.uniform(size=(nwalkers, case.n)))
    new_pos = h.walkers_update(case.pos, rand, TSTEP)
    moved = np.any(np.abs(new_pos.reshape(nwalkers, case.n, 3) - case.pos.reshape(nwalkers, case.n, 3)) > 0, axis=-1)
    out.update(**rand, accept=moved, pos_after_sweep=new_pos)
    if with_ecp:
        rot = O.random_rotations(rng, nwalkers)
        out.update(rot=rot, e_l=h.local_energy_ecp(case.pos, rot, ecp_tables(case.a, rich=True)))
    else:
        out.update(e_l=h.local_energy_ae(case.pos))
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print(name, "written from the genuine reference")


if __name__ == "__main__":
    ok, why, _ = RJ.probe()
    if not ok:
        print("cannot regenerate from the reference:", why)
        sys.exit(2)
    make("C_ecp", 6, True)
    make("C_ae", 5, False)
    make("N2_ecp", 3, True)
