#!/usr/bin/env python
"""Generates tests/golden/*.npz: frozen input/output vectors of the walker hot path.

The reference (AIQMCrelease3) cannot run in this image (no jax), so these vectors come from the float64 CPU oracle
(oracle/aiqmc_oracle.py), which is itself pinned by the reference's known answers (tests/test_oracle_pins.py).  They
freeze the oracle: a later edit that changes any number below fails tests/test_golden.py on the CPU, and the CUDA path
is checked against the same files on the GPU.  Re-run only on purpose:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from common import CASES, Case, O, ecp_tables  # noqa: E402

TSTEP = 0.05


def make(name, nwalkers, with_ecp):
    case = Case(**CASES[name], nwalkers=nwalkers, width=0.8)
    rng = np.random.default_rng(1234)
    out = dict(pos=case.pos, seed=np.int64(CASES[name]["seed"]))
    ph, la = case.net.apply(case.params, torch.tensor(case.pos), case.t_spins, case.t_atoms)
    f = lambda x: case.net.apply(case.params, x, case.t_spins, case.t_atoms)[1]
    _, g, d2 = O.grad_and_hess_diag(f, torch.tensor(case.pos))
    out.update(phase=ph.numpy(), logabs=la.numpy(), grad=g.numpy(), lap=d2.sum(-1).numpy())
    rand = dict(gauss1=rng.standard_normal((nwalkers, 3 * case.n)) * TSTEP ** 0.5,
                gauss2=rng.standard_normal((nwalkers, case.n, 3 * case.n)) * TSTEP ** 0.5,
                rnd=rng.uniform(size=(nwalkers, case.n)))
    new_data, aux = O.walkers_update(O.select_output(case.net.apply, 1), case.params, case.oracle_data(),
                                     {k: torch.tensor(v) for k, v in rand.items()}, TSTEP, 3, case.n, nwalkers, return_aux=True)
    out.update(**rand, accept=aux['accept'].numpy(), pos_after_sweep=new_data.positions.numpy())
    if with_ecp:
        tabs = ecp_tables(case.a, rich=True)
        rot = O.random_rotations(rng, nwalkers)
        le = O.local_energy_ecp(case.net.apply, O.make_log_network(case.net.apply), case.charges, None, tabs['rn_local'],
                                tabs['local_coes'], tabs['local_exps'], tabs['rn_non_local'], tabs['non_local_coes'],
                                tabs['non_local_exps'], case.a, case.n, 3, 2)
        e, _ = le(case.params, torch.tensor(rot), case.oracle_data(batched_static=False))
        out.update(rot=rot, e_l=e.numpy())
    else:
        e, _ = O.local_energy_ae(case.net.apply, case.charges)(case.params, None, case.oracle_data(batched_static=False))
        out.update(e_l=e.numpy())
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(path, {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    make("C_ecp", 6, True)
    make("C_ae", 5, False)
    make("N2_ecp", 3, True)
