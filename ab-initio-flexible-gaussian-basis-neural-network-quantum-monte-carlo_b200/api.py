"""Drop-in mirrors of the reference's Python closures for the walker hot path.

Same factory names, argument meaning and return conventions as
  AIQMCrelease3/wavefunction_Ynlm/nn.py:511-553      make_ai_net -> Network(init, apply, orbitals)
  AIQMCrelease3/VMC/VMCmcstep.py:121-140             main_monte_carlo -> mc_step(params, data, key)
  AIQMCrelease3/Energy/hamiltonian.py:236-260        local_energy    -> _e_l(params, key, data)
  AIQMCrelease3/Energy/pphamiltonian.py:130-190      local_energy (ccECP)
  AIQMCrelease3/DMC/{drift_diffusion,S_matrix,branch,dmc}.py
but natively batched over walkers (the reference wraps them in jax.vmap / jax.pmap) and backed by
the sm_100a kernels.  Differences a caller sees:
  * arrays are torch CUDA float64 tensors (numpy / float32 inputs are converted);
  * `key` is either a dict of explicit random arrays (parity mode: the arrays the reference draws
    from its PRNGKey) or an int seed (throughput mode: torch's device Philox generator);
  * `params` may be the reference pytree (packed on every call) or a PackedParams handle.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, replace
from typing import Any, Callable, Dict, Optional, Sequence

import numpy as np
import torch

from . import parallel
from .engine import WalkerEngine
from .system import SystemSpec, make_ecp, unpack_param_grad


@dataclass
class AINetData:
    """nn.py:20-25.  positions (B,3N); spins/atoms/charges may be per-walker replicas (B,...) as in
    main_pp_adam_muti_GPU.py:80-94 -- only row 0 is read, like VMCmcstep.py:42-44."""
    positions: Any
    spins: Any
    atoms: Any
    charges: Any


class PackedParams:
    """Parameters already flattened and resident on the device."""

    def __init__(self, engine: WalkerEngine, tree):
        engine.set_params(tree)
        self.engine = engine
        self.tree = tree


@dataclass
class Network:
    init: Callable
    apply: Callable
    orbitals: Optional[Callable]
    pack: Callable
    engine_for: Callable


def _first(x, ndim):
    x = torch.as_tensor(x) if not isinstance(x, torch.Tensor) else x
    return x.reshape(-1, *x.shape[-ndim:])[0] if x.dim() > ndim else x


def make_ai_net(nspins, charges, parallel_indices, antiparallel_indices, spin_up_indices, spin_down_indices,
                n_parallel: int, n_antiparallel: int, ndim: int, natoms: int, nelectrons: int,
                determinants: int = 1, bias_orbitals: bool = True, rescale_inputs: bool = False,
                hidden_dims=((4, 4), (4, 4), (4, 4)), hidden_dims_Ynlm=(6, 6, 6), device=None) -> Network:
    """nn.py:511-553.  Only the configuration every reference driver uses is compiled:
    ndim=3, determinants=1, rescale_inputs=False, hidden_dims=((4,4),)*3, hidden_dims_Ynlm=(6,6,6)."""
    if ndim != 3 or determinants != 1 or rescale_inputs or tuple(map(tuple, hidden_dims)) != ((4, 4),) * 3 \
            or tuple(hidden_dims_Ynlm) != (6, 6, 6):
        raise NotImplementedError("aiqmc_b200 compiles the reference's default architecture only")
    charges_np = np.asarray(charges.cpu() if hasattr(charges, "cpu") else charges, dtype=np.float64).reshape(-1)
    engines: Dict[bytes, WalkerEngine] = {}

    def engine_for(atoms) -> WalkerEngine:
        at = np.asarray(atoms.detach().cpu() if hasattr(atoms, "detach") else atoms, dtype=np.float64).reshape(-1, 3)
        key = at.tobytes()
        if key not in engines:
            spec = SystemSpec(nelectrons, natoms, tuple(nspins), at, charges_np,
                              np.asarray(spin_up_indices).reshape(-1), np.asarray(spin_down_indices).reshape(-1),
                              np.asarray(parallel_indices).reshape(2, -1), np.asarray(antiparallel_indices).reshape(2, -1))
            engines[key] = WalkerEngine(spec, device=device)
        return engines[key]

    def init(key):
        """Same leaf shapes / scales as nn.py:203-278,370-407; `key` is an int seed or numpy Generator."""
        rng = key if isinstance(key, np.random.Generator) else np.random.default_rng(int(key))

        def lin(i, o, bias=True):
            p = {'w': rng.standard_normal((i, o)) / math.sqrt(float(i))}
            if bias:
                p['b'] = rng.standard_normal((o,))
            return p
        layers, layers_y = [], []
        d_one, d_y = 4 * natoms, 4 * natoms + 2
        for i in range(3):
            d_in = 3 * d_one + 8
            lp = {'convolutional': {'w': rng.standard_normal((nelectrons, d_in)) / math.sqrt(float(nelectrons)),
                                    'b': rng.standard_normal((nelectrons, d_in // 4))},
                  'single': lin(d_in // 4, 4)}
            if i < 2:
                lp['double'] = lin(4, 4)
            layers.append(lp)
            layers_y.append({'single_Ynlm': lin(d_y, 6)})
            d_one, d_y = 4, 6
        one = np.ones
        return {'layers': {'input': {}, 'streams': layers, 'streams_y': layers_y},
                'orbitals': [lin(4, 2 * nelectrons) for _ in range(2)],
                'y': [lin(6, nelectrons, bias=False)],
                'jastrow_ee': {'ee_par': one(n_parallel), 'ee_anti': one(n_antiparallel)},
                'jastrow_ae': {'ae': one((nelectrons, natoms))},
                'envelope': [{'pi': one((natoms, 3)), 'sigma': one((natoms, 3)), 'alpha': one(1), 'beta': one(natoms),
                              'xi': one(1), 'eplion': one((natoms, 3)), 'mu': one(natoms), 'nu': one(natoms)}
                             for _ in range(nelectrons)]}

    def pack(params, atoms) -> PackedParams:
        return PackedParams(engine_for(atoms), params)

    def _bind(params, atoms) -> WalkerEngine:
        if isinstance(params, PackedParams):
            return params.engine
        eng = engine_for(_first(atoms, 2))
        eng.set_params(params)
        return eng

    def apply(params, pos, spins=None, atoms=None, charges=None):
        """signed_network(params, pos, spins, atoms, charges) -> (phase, log|psi|); pos (..., 3N)."""
        eng = _bind(params, atoms)
        return eng.psi(pos, mode=0)

    net = Network(init=init, apply=apply, orbitals=None, pack=pack, engine_for=engine_for)
    apply.network = net
    apply.bind = _bind
    return net


def _engine_of(f, params, data) -> WalkerEngine:
    if not hasattr(f, "bind"):
        raise TypeError("f must be the `apply` of aiqmc_b200.make_ai_net (the kernels are not a generic autodiff)")
    return f.bind(params, data.atoms)


def _positions(eng: WalkerEngine, data: AINetData) -> torch.Tensor:
    p = torch.as_tensor(data.positions).to(device=eng.device, dtype=torch.float64)
    return p.reshape(-1, 3 * eng.n).contiguous()


def _sweep_rand(eng: WalkerEngine, key, B: int, tstep: float, step: int):
    n = eng.n
    if isinstance(key, (list, tuple)):
        key = key[step]
    if isinstance(key, dict):
        cv = lambda a: torch.as_tensor(a).to(device=eng.device, dtype=torch.float64).contiguous()
        return cv(key['gauss1']), cv(key['gauss2']), cv(key['rnd'])
    # throughput mode: counter-based Philox in the library's own kernels (csrc/rng.cu), keyed by (seed, walker, step);
    # gauss2 comes in the compact (B,N,3) form -- only its diagonal blocks are ever read (VMCmcstep.py:86-94)
    return eng.rng_sweep(int(key), step, 0, B, tstep)


def main_monte_carlo(f, tstep: float, ndim: int, nelectrons: int, nsteps: int, batch_size: int):
    """VMCmcstep.py:121-140 -> mc_step(params, data, key) -> data (nsteps sweeps, in place on a copy)."""
    def mc_step(params, data: AINetData, key):
        eng = _engine_of(f, params, data)
        pos = _positions(eng, data).clone()
        for i in range(nsteps):
            g1, g2, u = _sweep_rand(eng, key, pos.shape[0], tstep, i)
            eng.vmc_sweep(pos, g1, g2, u, tstep, want_accept=False)
        return replace(data, positions=pos)
    return mc_step


def local_energy(f, charges, nspins=None, use_scan: bool = False, complex_output: bool = False, lognetwork=None,
                 rn_local=None, local_coes=None, local_exps=None, rn_non_local=None, non_local_coes=None,
                 non_local_exps=None, natoms: Optional[int] = None, nelectrons: Optional[int] = None,
                 ndim: int = 3, list_l: Optional[int] = None):
    """hamiltonian.py:236-260 (no ECP tables) / pphamiltonian.py:130-190 (with ECP tables).

    Returns _e_l(params, key, data) -> (E_L (B,), None).  For the ECP flavour `key` is the per-walker
    rotation (B,3,3) the reference draws at pp_energy_test.py:75, or an int seed.
    """
    if complex_output:
        raise NotImplementedError("no reference caller sets complex_output=True (quirk Q11)")
    ecp = None
    if rn_local is not None:
        ecp = make_ecp(natoms, rn_local, local_coes, local_exps, rn_non_local, non_local_coes, non_local_exps, list_l)

    def _e_l(params, key, data: AINetData):
        eng = _engine_of(f, params, data)
        pos = _positions(eng, data)
        rot = None
        if ecp is not None:
            rot = key if not isinstance(key, (int, np.integer)) else random_rotations(pos.shape[0], int(key), eng.device)
        return eng.local_energy(pos, rot, ecp=ecp), None       # this closure's table; the shared engine is not mutated
    _e_l.engine_of = lambda params, data: _engine_of(f, params, data)
    return _e_l


def random_rotations(n: int, seed: int, device, step: int = 0, walker0: int = 0) -> torch.Tensor:
    """Haar-random 3x3 orthogonal matrices (stand-in for jax.random.orthogonal, pseudopotential.py:234), from the
    library's Philox kernel (csrc/rng.cu: aiqmc_rng_rotations)."""
    import ctypes as C
    from . import lib as _lib
    rot = torch.empty((n, 3, 3), dtype=torch.float64, device=device)
    with torch.cuda.device(rot.device):
        _lib.check(_lib.load().aiqmc_rng_rotations(int(seed), int(step), int(walker0), n, C.c_void_p(rot.data_ptr()),
                                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)), "aiqmc_rng_rotations")
    return rot


def total_energy(local_energy_fn, process_group=None):
    """Loss/pploss.py:157-167 / DMC/total_energy.py:9-32 forward part: E_L per walker, mean and variance;
    the two pmean's become ONE 4-double all-reduce (NCCL) when a process group is given."""
    def _total(params, key, data: AINetData):
        e_l, _ = local_energy_fn(params, key, data)
        eng = local_energy_fn.engine_of(params, data)
        stats = eng.energy_stats(e_l)
        mean, variance, _ = parallel.allreduce_energy_stats(stats, process_group)
        return e_l, mean, variance
    return _total


# ---- loss: Loss/pploss.py:73-223 ----------------------------------------------
@dataclass
class AuxiliaryLossData:
    """Loss/pploss.py:20-33."""
    variance: Any
    local_energy: Any
    clipped_energy: Any
    grad_local_energy: Any = None
    local_energy_mat: Any = None


def clip_local_values(local_values: torch.Tensor, mean_local_values, clip_scale: float, clip_from_median: bool,
                      center_at_clipped_value: bool, complex_output: bool = False, process_group=None):
    """pploss.py:73-135 on device tensors; the pmean's are all-reduces over `process_group`."""
    batch_mean = lambda v: parallel.allreduce_mean(torch.mean(v), process_group)

    def clip_at_total_variation(values, center, scale):
        tv = batch_mean(torch.abs(values - center))
        return torch.minimum(torch.maximum(values, center - scale * tv), center + scale * tv)

    if clip_from_median:
        allv = parallel.all_gather_cat(local_values.real.contiguous(), process_group)
        clip_center = torch.quantile(allv, 0.5, interpolation='midpoint')      # jnp.median
    else:
        clip_center = mean_local_values
    if complex_output:
        cr = clip_center.real if torch.is_complex(clip_center) else clip_center
        ci = clip_center.imag if torch.is_complex(clip_center) else torch.zeros_like(cr)
        clipped = torch.complex(clip_at_total_variation(local_values.real, cr, clip_scale),
                                clip_at_total_variation(local_values.imag, ci, clip_scale))
    else:
        clipped = clip_at_total_variation(local_values, clip_center, clip_scale)
    diff_center = batch_mean(clipped) if center_at_clipped_value else mean_local_values
    return diff_center, clipped - diff_center


def make_loss(network, local_energy_fn, clip_local_energy: float = 0.0, clip_from_median: bool = True,
              center_at_clipped_energy: bool = True, complex_output: bool = True, process_group=None):
    """pploss.py:137-223.  Returns total_energy(params, key, data) -> (loss, AuxiliaryLossData) and, as the stand-in
    for jax.value_and_grad(total_energy, has_aux=True) through the custom JVP, total_energy.value_and_grad(params,
    key, data) -> ((loss, aux), grads) with `grads` a pytree shaped like the reference's params, already pmean'ed
    over `process_group` (Optimizer/adam.py:49-59).  `network` is make_ai_net(...).apply, `local_energy_fn` the
    closure of aiqmc_b200.local_energy."""
    def total_energy(params, key, data: AINetData):
        e_l, e_l_mat = local_energy_fn(params, key, data)
        eng = local_energy_fn.engine_of(params, data)
        mean, variance, _ = parallel.allreduce_energy_stats(eng.energy_stats(e_l), process_group)
        loss = mean if torch.is_complex(e_l) else mean.real
        return loss, AuxiliaryLossData(variance=variance, local_energy=e_l, clipped_energy=e_l, local_energy_mat=e_l_mat)

    def value_and_grad(params, key, data: AINetData):
        loss, aux = total_energy(params, key, data)
        e_l = aux.local_energy
        if clip_local_energy > 0.0:
            aux.clipped_energy, diff = clip_local_values(e_l, loss, clip_local_energy, clip_from_median,
                                                         center_at_clipped_energy, complex_output, process_group)
        else:
            diff = e_l - loss
        eng = _engine_of(network, params, data)
        B = e_l.shape[0]
        if complex_output:
            # (term1 - 2 term2).real / B of pploss.py:208-218 with psi_tangent = d(log|psi| + i phase)
            ce = aux.clipped_energy if torch.is_complex(aux.clipped_energy) else torch.complex(aux.clipped_energy, torch.zeros_like(aux.clipped_energy))
            d = diff if torch.is_complex(diff) else torch.complex(diff, torch.zeros_like(diff))
            ce = ce.expand(B) if ce.ndim == 0 else ce
            alpha, beta = 2.0 * d.real / B, 2.0 * (d.imag + ce.imag) / B
            out_loss = loss.real
        else:
            alpha, beta = diff.real / B, torch.zeros(B, dtype=torch.float64, device=eng.device)
            out_loss = loss.real if torch.is_complex(loss) else loss
        g, _, _ = eng.param_grad(_positions(eng, data), alpha, beta)
        g = parallel.allreduce_mean(g, process_group)
        tree = params.tree if isinstance(params, PackedParams) else params
        grads = unpack_param_grad(eng.layout, g.cpu().numpy(), tree, eng.spec)
        return (out_loss, aux), grads

    total_energy.value_and_grad = value_and_grad
    return total_energy


# ---- all-electron Metropolis-Hastings: AIQMCrelease2/MonteCarloSample/mcstep.py -------------
def make_mcmc_step(batch_network, batch_per_device: int, steps: int = 10, atoms=None, ndim: int = 3, blocks: int = 1,
                   process_group=None):
    """mcstep.py:71-104 -> mcmc_step(params, data, key, width) -> (new_data, pmove).  `batch_network` is the `apply`
    of make_ai_net; key = dict(noise (steps,B,3N) ~ N(0,1), u (steps,B) ~ U(0,1)) in parity mode (the draws of
    :49-56 and :28-29) or an int seed; pmove is pmean'ed over `process_group` as at :101."""
    if ndim != 3 or blocks != 1:
        raise NotImplementedError("the reference only ever uses ndim=3, blocks=1 (mcstep.py:47)")
    nsteps = steps * blocks

    def mcmc_step(params, data: AINetData, key, width):
        eng = _engine_of(batch_network, params, data)
        pos = _positions(eng, data).clone()
        B = pos.shape[0]
        _, la = eng.psi(pos, mode=0)
        lp = (2.0 * la).contiguous()
        count = torch.zeros((), dtype=torch.int64, device=eng.device)
        gen = None
        if isinstance(key, (int, np.integer)):
            gen = torch.Generator(device=eng.device)
            gen.manual_seed(int(key))
        for s in range(nsteps):
            if gen is None:
                noise, u = key['noise'][s], key['u'][s]
            else:
                noise = torch.randn((B, 3 * eng.n), dtype=torch.float64, device=eng.device, generator=gen)
                u = torch.rand((B,), dtype=torch.float64, device=eng.device, generator=gen)
            eng.mh_step(pos, lp, noise, u, float(width), count)
        pmove = count.to(torch.float64) / (nsteps * batch_per_device)
        return replace(data, positions=pos), parallel.allreduce_mean(pmove, process_group)
    return mcmc_step


def update_mcmc_width(t: int, width, adapt_frequency: int, pmove, pmoves: np.ndarray, pmove_max: float = 0.55,
                      pmove_min: float = 0.5):
    """mcstep.py:107-124 verbatim semantics (host side)."""
    t_since = t % adapt_frequency
    pmoves[t_since] = float(torch.as_tensor(pmove).reshape(-1)[0])
    if t > 0 and t_since == 0:
        if np.mean(pmoves) > pmove_max:
            width = width * 1.1
        elif np.mean(pmoves) < pmove_min:
            width = width / 1.1
    return width, pmoves


# ---- correlated sampling (correlatedsamples/*.py) and the ccECP file reader ------------------
def _corr_call(fn_name, atoms, new_atoms, pos, out_cols):
    from . import lib as _lib
    from .engine import _ptr, _stream
    import ctypes as C
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise _lib.AiqmcError("correlated sampling needs a CUDA device (no CPU fallback)")
    at = np.ascontiguousarray(np.asarray(atoms, dtype=np.float64).reshape(-1, 3))
    nat = np.ascontiguousarray(np.asarray(new_atoms, dtype=np.float64).reshape(-1, 3))
    if at.shape != nat.shape:
        raise ValueError("atoms and new_atoms must have the same shape")
    p = torch.as_tensor(pos).to(device="cuda", dtype=torch.float64)
    lead = p.shape[:-1]
    p2 = p.reshape(-1, p.shape[-1]).contiguous()
    B, n = p2.shape[0], p2.shape[1] // 3
    out = torch.empty((B, 3 * n) if out_cols else (B,), dtype=torch.float64, device="cuda")
    rc = getattr(lib, fn_name)(at.ctypes.data_as(C.c_void_p), nat.ctypes.data_as(C.c_void_p), at.shape[0], _ptr(p2), B, n,
                               _ptr(out), _stream())
    _lib.check(rc, fn_name)
    return out.reshape(*lead, 3 * n) if out_cols else out.reshape(lead)


def correlated_samples(atoms, new_atoms, pos):
    """corrsamples.py:23-47, natively batched: pos (..., 3N) -> space-warped positions for the displaced nuclei."""
    return _corr_call("aiqmc_correlated_samples", atoms, new_atoms, pos, True)


def weights_jacobian(pos, atoms, new_atoms):
    """jacobianWeights.py:22-51, natively batched: pos (..., 3N) -> the reference's Jacobian weight per walker."""
    return _corr_call("aiqmc_weights_jacobian", atoms, new_atoms, pos, False)


def read_ecp_nwchem(text: str, symbols: Sequence[str], n_channels: int = 3, n_gauss: int = 2) -> Dict[str, np.ndarray]:
    """What pseudopotential/readpp.py:1-47 set out to do (it stops half way): an NWChem-format ccECP block
    (`X nelec k`, `X ul` + rows `n exponent coefficient`, then `X S`, `X P`, ...) -> the tables the reference
    hard-codes (example/single_atom_C/single_atom_C.py:13-23): Rn_local / Local_exps / Local_coes (A, K_loc) and
    Rn_non_local / Non_local_exps / Non_local_coes (A, n_channels, n_gauss) zero padded, one row per atom of `symbols`,
    plus nelec_core (A,) and list_l = n_channels - 1.  Keys follow aiqmc_b200.make_ecp / local_energy."""
    blocks: Dict[str, Dict[str, list]] = {}
    core: Dict[str, int] = {}
    elem, chan = None, None
    for raw in text.splitlines():
        tok = raw.split('#')[0].split()
        if not tok or tok[0].upper() in ("ECP", "END"):
            continue
        if len(tok) == 3 and tok[1].lower() == "nelec":
            core[tok[0]] = int(tok[2])
        elif len(tok) == 2 and not tok[0][0].isdigit():
            elem, chan = tok[0], tok[1].lower()
            blocks.setdefault(elem, {})[chan] = []
        elif len(tok) == 3 and elem is not None:
            blocks[elem][chan].append((float(tok[0]), float(tok[1]), float(tok[2])))
        else:
            raise ValueError(f"cannot parse ECP line: {raw!r}")
    order = ["s", "p", "d", "f", "g"][:n_channels]
    for sname in symbols:
        if sname not in blocks or "ul" not in blocks[sname]:
            raise ValueError(f"no ECP block for element {sname}")
    k_loc = max(len(blocks[s]["ul"]) for s in symbols)
    A = len(symbols)
    out = {"rn_local": np.zeros((A, k_loc)), "local_exps": np.zeros((A, k_loc)), "local_coes": np.zeros((A, k_loc)),
           "rn_non_local": np.zeros((A, n_channels, n_gauss)), "non_local_exps": np.zeros((A, n_channels, n_gauss)),
           "non_local_coes": np.zeros((A, n_channels, n_gauss)), "nelec_core": np.zeros(A, dtype=np.int64),
           "list_l": n_channels - 1}
    out["rn_non_local"][:] = 2.0          # padding convention of the reference's tables (single_atom_C.py:15)
    for ia, s in enumerate(symbols):
        if s not in blocks or "ul" not in blocks[s]:
            raise ValueError(f"no ECP block for element {s}")
        out["nelec_core"][ia] = core.get(s, 0)
        for k, (nn, ex, co) in enumerate(blocks[s]["ul"]):
            out["rn_local"][ia, k], out["local_exps"][ia, k], out["local_coes"][ia, k] = nn, ex, co
        for l, name in enumerate(order):
            rows = blocks[s].get(name, [])
            if len(rows) > n_gauss:
                raise ValueError(f"{s} {name.upper()} channel has {len(rows)} gaussians, table holds {n_gauss}")
            for k, (nn, ex, co) in enumerate(rows):
                out["rn_non_local"][ia, l, k], out["non_local_exps"][ia, l, k], out["non_local_coes"][ia, l, k] = nn, ex, co
        extra = [c for c in blocks[s] if c != "ul" and c not in order]
        if extra:
            raise ValueError(f"{s}: channels {extra} do not fit n_channels={n_channels}")
    return out


# ---- DMC -------------------------------------------------------------------------------
def propose_drift_diffusion(f, tstep: float, ndim: int, nelectrons: int, batch_size: int):
    """DMC/drift_diffusion.py:25-107 -> (new_data, tdamp, grad_eff_old, grad_new_eff_s).
    `f` is signed_network's apply; key as in main_monte_carlo (one sweep)."""
    def drift_diffusion(params, key, data: AINetData):
        eng = _engine_of(f, params, data)
        pos = _positions(eng, data).clone()
        g1, g2, u = _sweep_rand(eng, key, pos.shape[0], tstep, 0)
        out = eng.vmc_sweep(pos, g1, g2, u, tstep, signed_ratio=True, want_accept=True, want_drift=True,
                            want_aux=True)
        tdamp = out['aux'][0] / out['aux'][1]                                 # quirk Q19
        _, _, g_s = eng.psi(pos, mode=1)
        v2 = torch.sum(g_s ** 2)
        taueff = (torch.sqrt(1 + 2 * tstep * 0.25 * v2) - 1) / (0.25 * v2)     # limdrift, batch-global (Q6)
        return replace(data, positions=pos), tdamp, out['grad_eff_old'], g_s * taueff, out['accept']
    return drift_diffusion


def compute_tmoves(list_l, tstep: float, nelectrons: int, natoms: int, ndim: int, lognetwork, Rn_non_local,
                   Non_local_coes, Non_local_exps, _ecp=None):
    """DMC/Tmoves.py:32-225 -> calculate_ratio_weight_tmoves(data, params, key) -> (final_configuration, acceptance),
    natively batched over walkers.  `lognetwork` is the `apply` of make_ai_net (its complex log is taken inside the
    kernels); key = dict(rot (B,3,3), u (B,), rnd (B,N)): the three draws the reference makes from its key."""
    zeros = np.zeros((natoms, 3))
    # the T-move kernels read only the non-local channels and the quadrature grid of the table; dmc_propagate hands in
    # its full table (_ecp) so that T-moves and local energies share ONE constant-memory upload per step
    ecp = _ecp if _ecp is not None else make_ecp(natoms, zeros, zeros, zeros, Rn_non_local, Non_local_coes, Non_local_exps, list_l)

    def calculate_ratio_weight_tmoves(data: AINetData, params, key):
        eng = _engine_of(lognetwork, params, data)
        pos = _positions(eng, data)
        new_pos, acceptance, _ = eng.dmc_tmove(pos, key['rot'], key['u'], key['rnd'], tstep, ecp=ecp)
        return new_pos, acceptance
    calculate_ratio_weight_tmoves.ecp = ecp            # the table of this closure (local channel zeroed)
    return calculate_ratio_weight_tmoves


def _real_scalar(x) -> float:
    """jnp.real(x) of a python / numpy / torch scalar (ccECP energies are complex, quirk Q25)."""
    if isinstance(x, torch.Tensor):
        x = x.detach().reshape(-1)[0]
        return float(x.real if x.is_complex() else x)
    return float(np.real(x))


def comput_S(engine: WalkerEngine, e_trial, e_est, branchcut, drift, tau, eloc, process_group=None):
    """DMC/S_matrix.py:4-25; `drift` is the limited drift whose square the reference passes as v2."""
    e_est, e_trial = _real_scalar(e_est), _real_scalar(e_trial)          # S_matrix.py:18-20 takes jnp.real of all three
    m = parallel.allreduce_min(engine.dmc_ecut_min(eloc, e_est, branchcut), process_group)   # quirk Q20
    return engine.dmc_s(eloc, drift, e_trial, e_est, m, tau)


def branch(engine: WalkerEngine, weights: torch.Tensor, key):
    """DMC/branch.py:10-34 -> (new weight scalar, newinds); `key` is the uniform u in [0,1)."""
    return engine.branch_comb(weights, float(key))


def dmc_propagate(signed_network, lognetwork, tstep: float, nelectrons: int, natoms: int, ndim: int, batch_size: int,
                  charges, rn_local, local_coes, local_exps, rn_non_local, non_local_coes, non_local_exps,
                  list_l: int = 2, process_group=None):
    """DMC/dmc.py:13-93 -> dmc_propagate_run(params, key, data, weights, branchcut_start, e_trial, e_est)
    -> (eloc_new, weights, new_data): T-move, drift-diffusion sweep, local energy at the old and new configurations,
    the two S values and the weight update w *= exp(tstep * tdamp * (S_new + S_old) / 2).
    key = dict(tmove=dict(rot, u, rnd), sweep=dict(gauss1, gauss2, rnd), rot=(B,3,3)): the draws the reference
    makes from its single key, as explicit arrays (parity mode)."""
    tm = compute_tmoves(list_l, tstep, nelectrons, natoms, ndim, lognetwork, rn_non_local, non_local_coes,
                        non_local_exps, _ecp=make_ecp(natoms, rn_local, local_coes, local_exps, rn_non_local, non_local_coes,
                                                      non_local_exps, list_l))
    dd = propose_drift_diffusion(signed_network, tstep, ndim, nelectrons, batch_size)
    le = local_energy(signed_network, charges, lognetwork=lognetwork, rn_local=rn_local, local_coes=local_coes,
                      local_exps=local_exps, rn_non_local=rn_non_local, non_local_coes=non_local_coes,
                      non_local_exps=non_local_exps, natoms=natoms, nelectrons=nelectrons, ndim=ndim, list_l=list_l)

    def dmc_propagate_run(params, key, data: AINetData, weights, branchcut_start, e_trial, e_est):
        eng = _engine_of(signed_network, params, data)
        pos, _ = tm(data, params, key['tmove'])
        t_move_data = replace(data, positions=pos)
        new_data, tdamp, grad_eff_old, grad_new_eff, _ = dd(params, key['sweep'], t_move_data)
        eloc_old, _ = le(params, key['rot'], data)            # dmc.py:81: the configuration BEFORE the T-move
        eloc_new, _ = le(params, key['rot'], new_data)
        bc = torch.as_tensor(branchcut_start).to(device=eng.device, dtype=torch.float64).reshape(-1).contiguous()
        s_old = comput_S(eng, e_trial, e_est, bc, grad_eff_old, tstep, eloc_old, process_group)
        s_new = comput_S(eng, e_trial, e_est, bc, grad_new_eff, tstep, eloc_new, process_group)
        w = torch.as_tensor(weights).to(device=eng.device, dtype=torch.float64).clone().contiguous()
        eng.dmc_weights(w, s_old, s_new, tstep, float(tdamp))
        return eloc_new, w, new_data
    return dmc_propagate_run


def reconfigure(engine: WalkerEngine, positions: torch.Tensor, newinds: torch.Tensor, noise: torch.Tensor):
    """DMC/main_dmc.py:218-231, on the device: keep the UNIQUE survivors of the comb (ascending index), then pad
    the batch back to B rows with copies of the last survivor + `noise` rows (the jax.random.uniform(key, (n, 3N))
    draw of :226).  Returns (new positions (B,3N), number of distinct survivors)."""
    uniq = torch.unique(newinds.to(torch.int64))                      # sorted, as jnp.unique
    kept = engine.gather_walkers(positions, uniq)
    nmiss = positions.shape[0] - uniq.shape[0]
    if nmiss > 0:
        extra = kept[-1][None, :] + torch.as_tensor(noise).to(device=kept.device, dtype=kept.dtype)[:nmiss]
        kept = torch.cat([kept, extra], dim=0)
    return kept, int(uniq.shape[0])


def estimate_energy(energy: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    """DMC/estimate_energy.py:4-5: weighted average over the whole (block, iteration, walker) history."""
    e, w = torch.as_tensor(energy), torch.as_tensor(weights).to(torch.as_tensor(energy).device)
    return (e * w).sum() / w.sum()


def trial_energy(e_est, weights: torch.Tensor, feedback: float):
    """DMC/main_dmc.py:242: E_T = E_est - feedback * log(mean w)."""
    return e_est - feedback * torch.log(torch.as_tensor(weights).mean()).real


def branch_global(engine: WalkerEngine, weights: torch.Tensor, positions: torch.Tensor, key, process_group=None,
                  return_bytes: bool = False, mode: str = "balanced"):
    """Population control across all GPUs of the job (the reference combs per device only): the
    systematic comb of DMC/branch.py:10-34 over the weights of ALL ranks + migration of the selected walkers, through
    the C ABI (aiqmc_rebalance_nccl): exchanged are the block totals of each rank's weight scan (all-gather) and, in one
    grouped NCCL send/recv, only the walkers that change rank.  Returns (new weight scalar, new positions (B,3N),
    source rank of every new walker (B,) int32, walkers imported from other ranks[, bytes this rank sent]).
    mode "balanced" (default): same survivors and multiplicities as the global comb, each rank keeps its own walkers
    and only the population imbalance travels; "ordered": the exact slot order of a single-GPU comb over the
    concatenated batch (moves nearly every walker: the comb's base offset rotates the teeth)."""
    from .engine import NcclComm
    import torch.distributed as dist
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1
    comm = NcclComm.for_group(process_group, engine.device) if multi else None
    neww, new_pos, src, moved = engine.rebalance(weights, positions, float(key), comm, mode=mode)
    rank = comm.rank if comm is not None else 0
    imported = int((src != rank).sum())
    return (neww, new_pos, src, imported, moved) if return_bytes else (neww, new_pos, src, imported)
