"""The BASELINE.json configurations as synthetic workloads of the walker engine.

Product-side construction of systems, random-init parameters and starting walkers: bench.py, the multi-GPU tools and
the tests build their cases from here (nothing under oracle/ or tests/ is needed to run the engine).

  c_ae   configs[0]  C atom all-electron, N=6, A=1, 4,096 walkers
  c_ecp  configs[1]  C atom ccECP, N=4, A=1, 65,536 walkers per GPU           (the bench.py headline)
  n2     configs[2]  N2, R = 2.068 bohr, ccECP, N=10, A=2, 65,536 walkers per GPU
  dmc    configs[3]  C atom ccECP fixed-node DMC (dmc_propagate_run + cross-GPU comb), atom at (0,0,-1) as
                     example/single_atom_C/C2testDMC.py:7-42, 65,536 walkers per GPU
  c6h6   configs[4]  benzene, ccECP, N=30, A=12, 262,144 walkers in total (32,768 per GPU at 8 GPUs)

ccECP tables: carbon verbatim from AIQMCrelease3/example/single_atom_C/single_atom_C.py:13-23; the reference ships no
nitrogen / hydrogen tables and no N2 / benzene geometry, so every atom re-uses the carbon table (same shapes; the
throughput is value independent) and benzene is a D6h ring with R_C = 2.640, R_H = 4.689 bohr -- builder-defined and
declared as such in the bench line.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np

from .system import SystemSpec, make_ecp

SEED = 20260101
TSTEP = 0.05

C_ECP_TABLES = dict(rn_local=np.array([[1.0, 3.0, 2.0]]), local_coes=np.array([[4.00000, 57.74008, -25.81955]]),
                    local_exps=np.array([[14.43502, 8.39889, 7.38188]]),
                    rn_non_local=np.array([[[2.0, 2.0], [2.0, 2.0], [2.0, 2.0]]]),
                    non_local_coes=np.array([[[52.13345, 0], [0, 0], [0, 0]]]),
                    non_local_exps=np.array([[[7.76079, 0], [0, 0], [0, 0]]]))


# the reference's own contracted-Gaussian basis file, AIQMC/C.cc-pVDZ.nwchem:1-27 verbatim (2s + 2p + 1d = 13 AOs from
# 21 primitives): the A0 micro-benchmark of SURVEY 8(d) evaluates it at walkers x electrons points
C_CC_PVDZ = """C s
13.073594 0.0051583
6.541187 0.0603424
4.573411 -0.1978471
1.637494 -0.0810340
0.819297 0.2321726
0.409924 0.2914643
0.231300 0.4336405
0.102619 0.2131940
0.051344 0.0049848
C s
0.127852 1.000000
C p
9.934169 0.0209076
3.886955 0.0572698
1.871016 0.1122682
0.935757 0.2130082
0.468003 0.2835815
0.239473 0.3011207
0.117063 0.2016934
0.058547 0.0453575
0.029281 0.0029775
C p
0.149161 1.000000
C d
0.561160 1.000000
"""


def ecp_tables(natoms: int) -> Dict[str, np.ndarray]:
    return {k: np.repeat(v, natoms, axis=0).copy() for k, v in C_ECP_TABLES.items()}


def _ring(r, n=6):
    return [[r * math.cos(2 * math.pi * k / n), r * math.sin(2 * math.pi * k / n), 0.0] for k in range(n)]


SYSTEMS = {
    "c_ae": dict(label="C atom all-electron (N=6, A=1), BASELINE configs[0]", spins=[1.] * 3 + [-1.] * 3,
                 atoms=[[0., 0., 0.]], charges=[6.0], ecp=False, walkers=4096),
    "c_ecp": dict(label="C atom ccECP (N=4, A=1): VMC sweep + ccECP local energy, BASELINE configs[1]",
                  spins=[1., -1., 1., -1.], atoms=[[0., 0., 0.]], charges=[4.0], ecp=True, walkers=65536),
    "n2": dict(label="N2 ccECP (N=10, A=2, R=2.068 bohr), BASELINE configs[2]", spins=[1.] * 5 + [-1.] * 5,
               atoms=[[0., 0., -1.034], [0., 0., 1.034]], charges=[5.0, 5.0], ecp=True, walkers=65536),
    "dmc": dict(label="C atom ccECP fixed-node DMC: dmc_propagate_run + cross-GPU comb (N=4, A=1), BASELINE configs[3]",
                spins=[1., -1., 1., -1.], atoms=[[0., 0., -1.0]], charges=[4.0], ecp=True, walkers=65536),
    "c6h6": dict(label="C6H6 ccECP (N=30, A=12), BASELINE configs[4]", spins=[1.] * 15 + [-1.] * 15,
                 atoms=_ring(2.640) + _ring(4.689), charges=[4.0] * 6 + [1.0] * 6, ecp=True, walkers=32768),
}


def flops_psi(n: int, a: int) -> float:
    """F(N,A) = (8/3)N^3 + 130N^2 + 40NA flop per psi value (fixed algorithmic count)."""
    return (8.0 / 3.0) * n ** 3 + 130.0 * n ** 2 + 40.0 * n * a


def flops_walker_step(name_or_n, a: Optional[int] = None, ecp: bool = True, dmc: bool = False) -> float:
    """Algorithmic flops of one walker-step: sweep 3(N+1)F + kinetic (3N+2)F [+ quadrature 50NA F];
    DMC step (150NA + 9N + 11)F."""
    if isinstance(name_or_n, str):
        s = SYSTEMS[name_or_n]
        n, a, ecp, dmc = len(s["spins"]), len(s["atoms"]), s["ecp"], name_or_n == "dmc"
    else:
        n = int(name_or_n)
    f = flops_psi(n, a)
    if dmc:
        return (150 * n * a + 9 * n + 11) * f
    return (6 * n + 5 + (50 * n * a if ecp else 0)) * f


def stage_flops(n: int, a: int, ecp: bool = True) -> Dict[str, float]:
    f = flops_psi(n, a)
    out = {"sweep": 3 * (n + 1) * f, "kinetic": (3 * n + 2) * f}
    if ecp:
        out["quadrature"] = 50.0 * n * a * f
    return out


def init_walkers(rng: np.random.Generator, atoms: np.ndarray, charges: np.ndarray, n: int, nwalkers: int,
                 width: float = 1.0) -> np.ndarray:
    """Atom-centred Gaussians (initial_electrons_positions/init.py:16-25): electrons are handed to the atoms in
    proportion to their (effective) charge, any remainder round-robin."""
    charges = np.asarray(charges, dtype=np.float64)
    quota = np.floor(charges * n / charges.sum()).astype(int)
    owner = [a for a, q in enumerate(quota) for _ in range(q)]
    a = 0
    while len(owner) < n:
        owner.append(a % len(charges))
        a += 1
    centres = np.concatenate([atoms[o] for o in owner[:n]])
    return centres[None, :] + width * rng.normal(size=(nwalkers, 3 * n))


@dataclass
class Workload:
    name: str
    label: str
    spec: SystemSpec
    params: dict
    pos: np.ndarray              # (B,3N) float64 starting walkers of this rank
    ecp: object                  # AiqmcEcp or None
    tables: Optional[dict]
    n: int
    a: int
    spins: np.ndarray = None

    def engine(self, device=None):
        from .engine import WalkerEngine
        return WalkerEngine(self.spec, self.params, ecp=self.ecp, device=device)


def build(name: str, nwalkers: Optional[int] = None, seed: int = SEED, rank: int = 0) -> Workload:
    """Random-init parameters with the reference's init scales (weights N(0,1)/sqrt(fan_in), biases N(0,1), Jastrow /
    envelope = 1; nn.py:203-278,370-407) -- identical on every rank -- and this rank's own walkers."""
    from .api import make_ai_net
    s = SYSTEMS[name]
    spins = np.asarray(s["spins"], dtype=np.float64)
    atoms = np.asarray(s["atoms"], dtype=np.float64).reshape(-1, 3)
    charges = np.asarray(s["charges"], dtype=np.float64)
    spec = SystemSpec.from_spins(atoms, charges, spins)
    n, a = spec.nelectrons, spec.natoms
    npar, nanti = spec.parallel_indices.shape[1], spec.antiparallel_indices.shape[1]
    net = make_ai_net(spec.nspins, charges, spec.parallel_indices, spec.antiparallel_indices, spec.spin_up_indices,
                      spec.spin_down_indices, npar, nanti, 3, a, n)
    params = net.init(np.random.default_rng(seed))
    B = int(nwalkers if nwalkers is not None else s["walkers"])
    pos = init_walkers(np.random.default_rng(seed + 1 + rank), atoms, charges, n, B)
    tabs = ecp_tables(a) if s["ecp"] else None
    ecp = make_ecp(a, list_l=2, **tabs) if tabs is not None else None
    return Workload(name, s["label"], spec, params, pos, ecp, tabs, n, a, spins)
