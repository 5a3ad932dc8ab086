"""reference_jax.py -- TEST INFRASTRUCTURE: drives the UNMODIFIED AIQMCrelease3 modules when JAX is importable.

The reference is pure Python/JAX.  JAX is not installed in the image this repository was built
in, so everything here is written against the reference's source (file:line cited per function) and has never been
executed by the builder: on a box that has jax (+ chex, kfac_jax, optax, which the reference imports at module import
time) `probe()` succeeds, tests/test_reference_jax.py compares the oracle (oracle/aiqmc_oracle.py) with the genuine
reference on the golden inputs, and `bench.py --impl reference` times the genuine reference (kind "reference").
Without jax every entry point reports why and the callers skip loudly / fall back to the oracle port.

Only tests/, __graft_entry__.smoke() and bench.py's reference / cpu_baseline legs may import this file.

How the reference's own random draws are made equal to the explicit arrays the oracle / CUDA path consume (north_star:
"identical ... proposal/uniform arrays"): `jax.random.normal / uniform / orthogonal` are replaced, for the duration of
one call, by functions that return the supplied arrays (selected by shape) -- the reference's arithmetic is untouched.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
from typing import Any, Dict, Optional, Tuple

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# places where the reference package directory `AIQMCrelease3/` may live (its parent goes on sys.path)
CANDIDATES = [os.environ.get("AIQMC_REFERENCE_ROOT", ""), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"]


def probe() -> Tuple[bool, str, Optional[str]]:
    """(usable, reason, reference root).  Usable = jax imports AND the AIQMCrelease3 sources are reachable AND the
    reference's own import chain (chex, kfac_jax, optax ...) resolves."""
    try:
        import jax  # noqa: F401
    except Exception as exc:                                   # ModuleNotFoundError in this image
        return False, f"jax is not importable ({exc!r}); the genuine reference cannot run here", None
    root = next((c for c in CANDIDATES if c and os.path.isfile(os.path.join(c, "AIQMCrelease3", "wavefunction_Ynlm", "nn.py"))), None)
    if root is None:
        return False, "AIQMCrelease3 sources not found (looked in $AIQMC_REFERENCE_ROOT, baseline/_ref, /root/reference)", None
    if root not in sys.path:
        sys.path.insert(0, root)
    try:
        importlib.import_module("AIQMCrelease3.wavefunction_Ynlm.nn")      # runs nn.py:557-599's import-time test (quirk Q2)
        importlib.import_module("AIQMCrelease3.VMC.VMCmcstep")
        importlib.import_module("AIQMCrelease3.Energy.pphamiltonian")
    except Exception as exc:
        return False, f"the reference's modules do not import ({exc!r})", root
    return True, "ok", root


def _jnp_tree(tree, dtype):
    import jax.numpy as jnp
    if isinstance(tree, dict):
        return {k: _jnp_tree(v, dtype) for k, v in tree.items()}
    if isinstance(tree, (list, tuple)):
        return type(tree)(_jnp_tree(v, dtype) for v in tree)
    a = tree.detach().cpu().numpy() if hasattr(tree, "detach") else np.asarray(tree)
    return jnp.asarray(a, dtype=dtype)


@contextlib.contextmanager
def _fixed_randoms(normal_by_shape: Dict[tuple, Any] = None, uniform_by_shape: Dict[tuple, Any] = None, orthogonal=None):
    """Replaces jax.random.{normal,uniform,orthogonal} by look-ups of explicit arrays (by requested shape)."""
    import jax
    import jax.numpy as jnp
    saved = (jax.random.normal, jax.random.uniform, jax.random.orthogonal)

    def normal(key, shape=(), dtype=None):
        if normal_by_shape and tuple(shape) in normal_by_shape:
            return jnp.asarray(normal_by_shape[tuple(shape)])
        return saved[0](key, shape=shape) if dtype is None else saved[0](key, shape=shape, dtype=dtype)

    def uniform(key, shape=(), dtype=None, minval=0.0, maxval=1.0):
        if uniform_by_shape and tuple(shape) in uniform_by_shape:
            return jnp.asarray(uniform_by_shape[tuple(shape)])
        return saved[1](key, shape=shape, minval=minval, maxval=maxval)

    def orth(key, n, shape=(), dtype=None):
        if orthogonal is not None:
            return jnp.asarray(orthogonal).reshape(tuple(shape) + (n, n))
        return saved[2](key, n, shape)

    jax.random.normal, jax.random.uniform, jax.random.orthogonal = normal, uniform, orth
    try:
        yield
    finally:
        jax.random.normal, jax.random.uniform, jax.random.orthogonal = saved


class ReferenceHarness:
    """The reference's closures for one system, built exactly as main/main_pp_adam_muti_GPU.py:97-148 builds them.
    `kw` are make_ai_net's keyword arguments (tests/common.py: Case.kw), `params` the parameter pytree (the oracle's
    tree has the reference's structure, nn.py:203-278,370-407), float64 requested through jax_enable_x64 so that the
    comparison with the float64 oracle is not limited by the reference's float32 default (quirk Q1)."""

    def __init__(self, kw: Dict[str, Any], params, atoms, charges, spins, x64: bool = True):
        ok, why, _ = probe()
        if not ok:
            raise RuntimeError(why)
        import jax
        import jax.numpy as jnp
        if x64:
            jax.config.update("jax_enable_x64", True)
        self.jax, self.jnp = jax, jnp
        self.dtype = jnp.float64 if x64 else jnp.float32
        from AIQMCrelease3.wavefunction_Ynlm import nn
        from AIQMCrelease3.VMC import VMCmcstep
        from AIQMCrelease3.Energy import pphamiltonian, hamiltonian
        from AIQMCrelease3.utils import utils
        self.nn, self.VMCmcstep, self.pph, self.ham, self.utils = nn, VMCmcstep, pphamiltonian, hamiltonian, utils
        k = dict(kw)
        for name in ("parallel_indices", "antiparallel_indices", "spin_up_indices", "spin_down_indices"):
            k[name] = jnp.asarray(np.asarray(k[name]))
        k["charges"] = jnp.asarray(np.asarray(k["charges"]), dtype=self.dtype)
        self.kw = k
        self.n, self.a = int(k["nelectrons"]), int(k["natoms"])
        self.network = nn.make_ai_net(**k)                                   # nn.py:511-553
        self.signed_network = self.network.apply
        self.params = _jnp_tree(params, self.dtype)
        self.atoms = jnp.asarray(np.asarray(atoms), dtype=self.dtype).reshape(self.a, 3)
        self.charges = k["charges"]
        self.spins = jnp.asarray(np.asarray(spins), dtype=self.dtype)

        def log_network(*args, **kwargs):                                    # main_pp_adam_muti_GPU.py:119-121
            phase, mag = self.signed_network(*args, **kwargs)
            return mag + 1.j * phase
        self.log_network = log_network

    # ---- signed_network(params, pos, spins, atoms, charges) -> (phase, log|psi|), nn.py:545-551
    def psi(self, pos: np.ndarray):
        jax, jnp = self.jax, self.jnp
        f = jax.vmap(lambda x: self.signed_network(self.params, x, self.spins, self.atoms, self.charges))
        ph, la = f(jnp.asarray(pos, dtype=self.dtype))
        return np.asarray(ph), np.asarray(la)

    def _data(self, pos):
        jnp = self.jnp
        pos = jnp.asarray(pos, dtype=self.dtype)
        B = pos.shape[0]
        return self.nn.AINetData(positions=pos, spins=jnp.broadcast_to(self.spins, (B, self.n)),
                                 atoms=jnp.broadcast_to(self.atoms, (B, self.a, 3)),
                                 charges=jnp.broadcast_to(self.charges, (B, self.a)))

    # ---- one walkers_update sweep, VMC/VMCmcstep.py:28-111, on explicit random arrays
    def walkers_update(self, pos: np.ndarray, rand: Dict[str, Any], tstep: float):
        jax = self.jax
        B, n = pos.shape[0], self.n
        s = float(np.sqrt(tstep))
        g1 = np.asarray(rand["gauss1"], dtype=np.float64) / s             # the reference multiplies by sqrt(tstep) itself
        g2 = np.asarray(rand["gauss2"], dtype=np.float64) / s
        u = np.asarray(rand["rnd"], dtype=np.float64)
        logabs_f = self.utils.select_output(self.signed_network, 1)
        with _fixed_randoms({(B, 3 * n): g1, (B, n, 3 * n): g2}, {(B, n): u}):
            new_data, _ = self.VMCmcstep.walkers_update(logabs_f, self.params, self._data(pos), jax.random.PRNGKey(0),
                                                        tstep=tstep, ndim=3, nelectrons=n, batch_size=B)
        return np.asarray(new_data.positions).reshape(B, 3 * n)

    # ---- ccECP local energy of every walker, Energy/pphamiltonian.py:130-190 under jax.vmap (Loss/pploss.py:145-153)
    def local_energy_ecp(self, pos: np.ndarray, rot: np.ndarray, tabs: Dict[str, Any], list_l: int = 2):
        jax, jnp = self.jax, self.jnp
        t = {k: jnp.asarray(np.asarray(v), dtype=self.dtype) for k, v in tabs.items()}
        le = self.pph.local_energy(f=self.signed_network, lognetwork=self.log_network, charges=self.charges, nspins=self.kw["nspins"],
                                   rn_local=t["rn_local"], local_coes=t["local_coes"], local_exps=t["local_exps"],
                                   rn_non_local=t["rn_non_local"], non_local_coes=t["non_local_coes"],
                                   non_local_exps=t["non_local_exps"], natoms=self.a, nelectrons=self.n, ndim=3, list_l=list_l)
        out = []
        for b in range(pos.shape[0]):                                     # one rotation per walker key (quirk Q17)
            d = self.nn.AINetData(positions=jnp.asarray(pos[b], dtype=self.dtype), spins=self.spins, atoms=self.atoms,
                                  charges=self.charges)
            with _fixed_randoms(orthogonal=np.asarray(rot[b]).reshape(1, 3, 3)):
                e, _ = le(self.params, jax.random.PRNGKey(b), d)
            out.append(complex(np.asarray(e)))
        return np.asarray(out)

    # ---- all-electron local energy, Energy/hamiltonian.py:236-260
    def local_energy_ae(self, pos: np.ndarray):
        jnp = self.jnp
        le = self.ham.local_energy(f=self.signed_network, charges=self.charges, nspins=self.kw["nspins"], use_scan=False)
        out = []
        for b in range(pos.shape[0]):
            d = self.nn.AINetData(positions=jnp.asarray(pos[b], dtype=self.dtype), spins=self.spins, atoms=self.atoms,
                                  charges=self.charges)
            e, _ = le(self.params, self.jax.random.PRNGKey(b), d)
            out.append(float(np.asarray(e).real))
        return np.asarray(out)

    # ---- the timed unit of bench.py --impl reference: mc_step (1 sweep, jitted) + vmapped local energy, drawing its
    #      own randoms exactly as main_pp_adam_muti_GPU.py:123-156,183-190 does (no patching in the timed path)
    def make_timed_step(self, tabs: Dict[str, Any], batch: int, tstep: float, list_l: int = 2):
        jax, jnp = self.jax, self.jnp
        t = {k: jnp.asarray(np.asarray(v), dtype=self.dtype) for k, v in tabs.items()}
        mc_step = self.VMCmcstep.main_monte_carlo(f=self.signed_network, tstep=tstep, ndim=3, nelectrons=self.n, nsteps=1,
                                                  batch_size=batch)
        le = self.pph.local_energy(f=self.signed_network, lognetwork=self.log_network, charges=self.charges, nspins=self.kw["nspins"],
                                   rn_local=t["rn_local"], local_coes=t["local_coes"], local_exps=t["local_exps"],
                                   rn_non_local=t["rn_non_local"], non_local_coes=t["non_local_coes"],
                                   non_local_exps=t["non_local_exps"], natoms=self.a, nelectrons=self.n, ndim=3, list_l=list_l)
        batch_le = jax.jit(jax.vmap(le, in_axes=(None, 0, self.nn.AINetData(positions=0, spins=0, atoms=0, charges=0)),
                                    out_axes=(0, None)))                    # pploss.py:145-153

        def step(pos, seed: int):
            key = jax.random.PRNGKey(seed)
            data = mc_step(self.params, self._data(pos), key)
            keys = jax.random.split(key, batch)
            e, _ = batch_le(self.params, keys, data)
            e.block_until_ready()
            return np.asarray(data.positions).reshape(batch, -1), np.asarray(e)
        return step
