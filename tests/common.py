"""Shared builders for the parity tests: one synthetic system + oracle network + packed params."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import aiqmc_oracle as O  # noqa: E402

# carbon ccECP tables, verbatim from AIQMCrelease3/example/single_atom_C/single_atom_C.py:13-23
C_ECP = dict(rn_local=np.array([[1.0, 3.0, 2.0]]), local_coes=np.array([[4.00000, 57.74008, -25.81955]]),
             local_exps=np.array([[14.43502, 8.39889, 7.38188]]),
             rn_non_local=np.array([[[2.0, 2.0], [2.0, 2.0], [2.0, 2.0]]]),
             non_local_coes=np.array([[[52.13345, 0], [0, 0], [0, 0]]]),
             non_local_exps=np.array([[[7.76079, 0], [0, 0], [0, 0]]]))


def ecp_tables(natoms, rich=False):
    """Same-shape tables for `natoms` atoms.  rich=True fills l=1,2 and the second gaussian with
    non-zero (made-up) values so the P_1/P_2 branches are exercised."""
    t = {k: np.repeat(v, natoms, axis=0).copy() for k, v in C_ECP.items()}
    if rich:
        t['non_local_coes'][:, 1, :] = [3.1, -0.7]
        t['non_local_coes'][:, 2, :] = [-1.3, 0.4]
        t['non_local_coes'][:, 0, 1] = 2.2
        t['non_local_exps'][:, 1, :] = [1.9, 0.8]
        t['non_local_exps'][:, 2, :] = [1.1, 2.5]
        t['non_local_exps'][:, 0, 1] = 0.9
        t['rn_non_local'][:, 1, 1] = 0.0
        t['rn_non_local'][:, 2, 0] = 1.0
    return t


class Case:
    def __init__(self, n, natoms, spins, seed, nspins=None, atoms=None, charges=None, nwalkers=8, width=1.0):
        torch.set_default_dtype(torch.float64)
        rng = np.random.default_rng(seed)
        self.rng = rng
        self.n, self.a = n, natoms
        self.spins = np.asarray(spins, dtype=np.float64)
        self.atoms = np.asarray(atoms, dtype=np.float64) if atoms is not None else rng.normal(size=(natoms, 3))
        self.charges = np.asarray(charges, dtype=np.float64) if charges is not None else np.full(natoms, float(n) / natoms)
        par, anti, npar, nanti = O.jastrow_indices_ee(self.spins, n)
        up, dn = O.spin_indices_h(self.spins)
        self.nspins = tuple(nspins) if nspins is not None else (len(up), len(dn))
        self.kw = dict(nspins=self.nspins, charges=self.charges, parallel_indices=par, antiparallel_indices=anti,
                       spin_up_indices=up, spin_down_indices=dn, n_parallel=npar, n_antiparallel=nanti, ndim=3,
                       natoms=natoms, nelectrons=n)
        self.net = O.make_ai_net(**self.kw)
        self.params = self.net.init(rng, randomize_all=True)
        centres = np.concatenate([np.tile(self.atoms[i % natoms], 1) for i in range(n)])
        self.pos = centres[None, :] + width * rng.normal(size=(nwalkers, 3 * n))
        self.B = nwalkers

    # oracle-side views
    @property
    def t_atoms(self):
        return torch.tensor(self.atoms)

    @property
    def t_spins(self):
        return torch.tensor(self.spins)

    def oracle_data(self, pos=None, batched_static=True):
        pos = torch.tensor(self.pos if pos is None else pos)
        B = pos.shape[0]
        if batched_static:
            return O.AINetData(positions=pos, spins=self.t_spins.expand(B, self.n),
                               atoms=self.t_atoms.expand(B, self.a, 3), charges=torch.tensor(self.charges).expand(B, self.a))
        return O.AINetData(positions=pos, spins=self.t_spins, atoms=self.t_atoms, charges=torch.tensor(self.charges))

    def sweep_rand(self, tstep, B=None):
        B = B or self.B
        n = self.n
        return dict(gauss1=torch.tensor(self.rng.standard_normal((B, 3 * n))) * tstep ** 0.5,
                    gauss2=torch.tensor(self.rng.standard_normal((B, n, 3 * n))) * tstep ** 0.5,
                    rnd=torch.tensor(self.rng.uniform(size=(B, n))))

    def spec(self):
        import aiqmc_b200
        return aiqmc_b200.SystemSpec(self.n, self.a, self.nspins, self.atoms, self.charges, self.kw['spin_up_indices'],
                                     self.kw['spin_down_indices'], self.kw['parallel_indices'],
                                     self.kw['antiparallel_indices'])


CASES = {
    "C_ecp": dict(n=4, natoms=1, spins=[1., -1., 1., -1.], seed=11, atoms=[[0., 0., 0.]], charges=[4.0]),
    "C_ae": dict(n=6, natoms=1, spins=[1., 1., 1., -1., -1., -1.], seed=12, atoms=[[0., 0., 0.]], charges=[6.0]),
    "N2_ecp": dict(n=10, natoms=2, spins=[1.] * 5 + [-1.] * 5, seed=13, atoms=[[0, 0, -1.034], [0, 0, 1.034]],
                   charges=[5.0, 5.0]),
    "odd": dict(n=5, natoms=2, spins=[1., 1., 1., -1., -1.], seed=14),
    "h2like": dict(n=2, natoms=2, spins=[1., -1.], seed=15),
}


# ---- the bench.py headline system (carbon ccECP, BASELINE configs[1]) as an oracle-backed Case ----
TSTEP = 0.05
BENCH_SEED = 20260101


def build_bench_case(nwalkers, seed=BENCH_SEED):
    case = Case(n=4, natoms=1, spins=[1., -1., 1., -1.], seed=seed, atoms=[[0., 0., 0.]], charges=[4.0],
                nwalkers=nwalkers, width=1.0)
    # "random-init params": the reference's init scales (weights N(0,1)/sqrt(fan_in), biases N(0,1),
    # Jastrow/envelope = 1), SURVEY 8(d)
    case.params = case.net.init(np.random.default_rng(seed), randomize_all=False)
    return case, ecp_tables(1)


def make_rand(rng, B, n, tstep):
    return dict(gauss1=(rng.standard_normal((B, 3 * n)) * np.sqrt(tstep)),
                gauss2=(rng.standard_normal((B, n, 3 * n)) * np.sqrt(tstep)),
                rnd=rng.uniform(size=(B, n)))


def random_rot(rng, B):
    q, r = np.linalg.qr(rng.standard_normal((B, 3, 3)))
    return q * np.sign(np.diagonal(r, axis1=-2, axis2=-1))[:, None, :]


def benzene_case(nwalkers, seed=21, width=0.8):
    import math
    ring = lambda r, n: [[r * math.cos(2 * math.pi * k / n), r * math.sin(2 * math.pi * k / n), 0.0] for k in range(n)]
    return Case(n=30, natoms=12, spins=[1.] * 15 + [-1.] * 15, seed=seed, atoms=ring(2.640, 6) + ring(4.689, 6),
                charges=[4.0] * 6 + [1.0] * 6, nwalkers=nwalkers, width=width)
