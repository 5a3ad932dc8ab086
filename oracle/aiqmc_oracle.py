"""CPU oracle for the AIQMCrelease3 walker hot path.  TEST INFRASTRUCTURE ONLY.

This file is a PyTorch-CPU restatement (float64 by default, float32 on request) of
the reference's JAX algorithm for SURVEY.md section 8(a) rows A1..A31.  It is used
only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs, and only as the checker or the timed CPU baseline -- the product
path (aiqmc_b200 + csrc/) never imports it.

PARITY STATUS: JAX is not installable in this image (no jax/jaxlib wheels, no
network), so the reference itself cannot be executed here.  The oracle is pinned
against every known answer the reference tree holds for this path
(ferminet/tests/hamiltonian_test.py:62-149, ferminet/tests/network_blocks_test.py:
27-46, quadrature self-checks of pseudopotential.py:181-225) -- see
tests/test_oracle_pins.py.  What stays "parity unpinned": jaxlib's complex64
`slogdet` rounding, the threefry RNG streams and `jax.random.orthogonal`
(side-stepped: every random array is an explicit input), and XLA's float32
summation order.

All functions are natively batched over arbitrary leading dimensions (the
reference wraps per-walker functions in jax.vmap); `pos` is (..., 3N).
Reference citations are relative to /root/reference/AIQMCrelease3/.
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass, replace
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

PI = math.pi


# --------------------------------------------------------------------------
# containers / static tables
# --------------------------------------------------------------------------
@dataclass
class AINetData:
    """wavefunction_Ynlm/nn.py:20-25."""
    positions: Any
    spins: Any
    atoms: Any
    charges: Any


def array_partitions(sizes: Sequence[int]) -> List[int]:
    """wavefunction_Ynlm/network_blocks.py:25-37."""
    return list(itertools.accumulate(sizes))[:-1]


def jastrow_indices_ee(spins, nelectrons: int):
    """spin_indices.py:5-19: (i<j) pair index lists in row-major nonzero order."""
    s = np.asarray(spins, dtype=np.float64).reshape(nelectrons)
    prod = np.triu(s[None, :] * s[:, None], k=1)
    par = np.array(np.nonzero(np.where(prod > 0, prod, 0.0)))
    anti = np.array(np.nonzero(np.where(prod < 0, prod, 0.0)))
    return par, anti, par.shape[1], anti.shape[1]


def spin_indices_h(spins):
    """spin_indices.py:38-46."""
    s = np.asarray(spins, dtype=np.float64)
    return np.nonzero(s > 0)[0], np.nonzero(s < 0)[0]


def init_electrons(rng: np.random.Generator, atoms, charges, spins, batch_size: int,
                   init_width: float):
    """initial_electrons_positions/init.py:7-30 (numpy RNG instead of threefry)."""
    atoms = np.asarray(atoms, dtype=np.float64)
    centres = np.concatenate([np.tile(atoms[i], int(charges[i])) for i in range(len(atoms))])
    pos = np.tile(centres[None, :], (batch_size, 1))
    pos = pos + rng.standard_normal(pos.shape) * init_width
    return pos, np.asarray(spins)


# --------------------------------------------------------------------------
# wavefunction  (wavefunction_Ynlm/nn.py, network_blocks.py, Jastrow.py, envelope.py)
# --------------------------------------------------------------------------
def construct_input_features(pos: torch.Tensor, atoms: torch.Tensor, ndim: int = 3):
    """nn.py:106-116.  ee[i,j] = r_j - r_i; r_ee has the diag-safe norm."""
    p = pos.reshape(*pos.shape[:-1], -1, ndim)
    ae = p[..., :, None, :] - atoms
    ee = p[..., None, :, :] - p[..., :, None, :]
    r_ae = torch.linalg.norm(ae, dim=-1, keepdim=True)
    n = p.shape[-2]
    eye = torch.eye(n, dtype=pos.dtype)
    r_ee = torch.linalg.norm(ee + eye[..., None], dim=-1) * (1.0 - eye)
    return ae, ee, r_ae, r_ee[..., None]


def ainet_features(ae, r_ae, ee, r_ee):
    """nn.py:125-137 (rescale_inputs=False, the only mode any driver uses)."""
    ae_features = torch.cat((r_ae, ae), dim=-1)
    ee_features = torch.cat((r_ee, ee), dim=-1)
    ae_features = ae_features.reshape(*ae_features.shape[:-2], -1)
    return ae_features, ee_features


def construct_symmetric_features(h_one, h_two, nspins):
    """nn.py:142-153.  Contiguous spin blocks along the electron axis (quirk Q4)."""
    parts = array_partitions(nspins)
    n = h_one.shape[-2]
    bounds = [0] + list(parts) + [n]
    g_one, g_two = [], []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        if hi > lo:
            g = h_one[..., lo:hi, :].mean(dim=-2, keepdim=True)
            g_one.append(g.expand(*h_one.shape))
            g_two.append(h_two[..., lo:hi, :, :].mean(dim=-3))
    return torch.cat([h_one] + g_one + g_two, dim=-1)


def y_l_real(x):
    """nn.py:156-167; x (...,3) unit vector -> (...,4)."""
    c0 = 0.5 * math.sqrt(1.0 / PI)
    c1 = math.sqrt(3.0 / (4.0 * PI))
    return torch.stack([torch.full_like(x[..., 0], c0), c1 * x[..., 0], c1 * x[..., 1],
                        c1 * x[..., 2]], dim=-1)


def y_l_real_high(x, y):
    """nn.py:169-193; x (...,3), y (...,1) -> (...,12,1).

    Quirk Q3: the reference reads x[3] of a length-3 vector; JAX clamps
    out-of-bounds reads, so x[3] == x[2].
    """
    x0, x1, x2 = x[..., 0:1], x[..., 1:2], x[..., 2:3]
    x3 = x2  # clamped read
    y2, y3 = y ** 2, y ** 3
    out = [0.5 * math.sqrt(15 / PI) * (x0 * x1 / y2),
           0.5 * math.sqrt(15 / PI) * (x1 * x2 / y2),
           0.25 * math.sqrt(5 / PI) * ((3 * x2 ** 2 - y2) / y2),
           0.5 * math.sqrt(15 / PI) * (x0 * x2 / y2),
           0.25 * math.sqrt(15 / PI) * ((x0 ** 2 - x1 ** 2) / y2),
           0.25 * math.sqrt(35 / (2 * PI)) * ((x1 * (3 * x0 ** 2 - x1 ** 2)) / y3),
           0.5 * math.sqrt(105 / PI) * (x0 * x1 * x2 / y3),
           0.25 * math.sqrt(21 / (2 * PI)) * ((x1 * (5 * x2 ** 2 - y2)) / y3),
           0.25 * math.sqrt(7 / PI) * ((5 * x2 ** 3 - 3 * x2 * y2) / y3),
           0.25 * math.sqrt(21 / (2 * PI)) * ((x0 * (5 * x2 ** 2 - y2)) / y3),
           0.25 * math.sqrt(105 / PI) * (((x0 ** 2 - x1 ** 2) * x3) / y3),
           0.25 * math.sqrt(35 / (2 * PI)) * ((x0 * (x0 ** 2 - 3 * x1 ** 2)) / y3)]
    return torch.stack(out, dim=-2)


def linear_layer(x, w, b=None):
    """network_blocks.py:119-133."""
    y = x @ w
    return y + b if b is not None else y


def convolu_layer(nelectrons, x, w, b=None):
    """network_blocks.py:106-116: per-electron grouped mean of 4-wide products."""
    xr = x.reshape(*x.shape[:-2], nelectrons, -1, 4)
    wr = w.reshape(nelectrons, -1, 4)
    y = (xr * wr).mean(dim=-1)
    return y + b


def _residual(x, y):
    """nn.py:284 -- residual only if shapes match (quirk Q5)."""
    return (x + y) / math.sqrt(2.0) if x.shape == y.shape else y


def slogdet(x):
    """network_blocks.py:138-158."""
    if x.shape[-1] == 1:
        v = x[..., 0, 0]
        sign = v / torch.abs(v) if torch.is_complex(v) else torch.sign(v)
        return sign, torch.log(torch.abs(v))
    return torch.linalg.slogdet(x)


def logdet_matmul(xs: Sequence[torch.Tensor]):
    """network_blocks.py:161-206 for w=None; each x is (..., n, n) (one determinant)."""
    det1d = 1.0
    phase_in, logdet = 1.0, 0.0
    for x in xs:
        if x.shape[-1] == 1:
            det1d = det1d * x[..., 0, 0]
        else:
            s, l = slogdet(x)
            phase_in, logdet = phase_in * s, logdet + l
    if not torch.is_tensor(logdet):
        logdet = torch.zeros(xs[0].shape[:-2], dtype=xs[0].real.dtype)
    maxlogdet = logdet  # max over the (single) determinant axis
    result = phase_in * det1d * torch.exp(logdet - maxlogdet)
    if torch.is_complex(result):
        phase_out = torch.angle(result)
    else:
        phase_out = torch.sign(result)
    return phase_out, torch.log(torch.abs(result)) + maxlogdet


@dataclass
class Network:
    """nn.py:99-103."""
    init: Callable
    apply: Callable
    orbitals: Callable
    config: Dict[str, Any]


def make_ai_net(nspins, charges, parallel_indices, antiparallel_indices, spin_up_indices,
                spin_down_indices, n_parallel: int, n_antiparallel: int, ndim: int, natoms: int,
                nelectrons: int, determinants: int = 1,
                hidden_dims=((4, 4), (4, 4), (4, 4)), hidden_dims_Ynlm=(6, 6, 6),
                dtype=torch.float64) -> Network:
    """nn.py:511-553.  `init(rng)` takes a numpy Generator (values need not equal JAX's
    threefry initialisation; parity is on *given* parameters)."""
    assert ndim == 3 and determinants == 1
    charges_t = torch.as_tensor(np.asarray(charges, dtype=np.float64)).to(dtype)
    par = np.asarray(parallel_indices).reshape(2, -1)
    anti = np.asarray(antiparallel_indices).reshape(2, -1)
    up = torch.as_tensor(np.asarray(spin_up_indices).reshape(-1), dtype=torch.long)
    dn = torch.as_tensor(np.asarray(spin_down_indices).reshape(-1), dtype=torch.long)
    nchannels = len([s for s in nspins if s > 0])
    nlayers = len(hidden_dims)

    def init(rng: np.random.Generator, randomize_all: bool = False):
        """nn.py:203-278,370-407; shapes identical to the reference pytree."""
        def lin(i, o, bias=True):
            p = {'w': rng.standard_normal((i, o)) / math.sqrt(float(i))}
            if bias:
                p['b'] = rng.standard_normal((o,))
            return p

        def ones(shape):
            if randomize_all:
                return 1.0 + 0.3 * rng.uniform(-1.0, 1.0, size=shape)
            return np.ones(shape)

        layers, layers_y = [], []
        d_one, d_two = natoms * 4, 4
        d_y = 4 * natoms + 2
        for i in range(nlayers):
            d_in = (nchannels + 1) * d_one + nchannels * d_two
            o_one, o_two = hidden_dims[i]
            lp = {'convolutional': {'w': rng.standard_normal((nelectrons, d_in)) / math.sqrt(float(nelectrons)),
                                    'b': rng.standard_normal((nelectrons, d_in // 4))},
                  'single': lin(d_in // 4, o_one)}
            ly = {'single_Ynlm': lin(d_y, hidden_dims_Ynlm[i])}
            if i < nlayers - 1:
                lp['double'] = lin(d_two, o_two)
            layers.append(lp)
            layers_y.append(ly)
            d_one, d_two, d_y = o_one, o_two, hidden_dims_Ynlm[i]
        params = {'layers': {'input': {}, 'streams': layers, 'streams_y': layers_y}}
        params['orbitals'] = [lin(d_one, 2 * nelectrons) for s in nspins if s > 0]
        params['y'] = [lin(d_y, nelectrons, bias=False)]
        params['jastrow_ee'] = {'ee_par': ones((n_parallel,)), 'ee_anti': ones((n_antiparallel,))}
        params['jastrow_ae'] = {'ae': ones((nelectrons, natoms))}
        params['envelope'] = [{'pi': ones((natoms, 3)), 'sigma': ones((natoms, 3)), 'alpha': ones((1,)),
                               'beta': ones((natoms,)), 'xi': ones((1,)), 'eplion': ones((natoms, 3)),
                               'mu': ones((natoms,)), 'nu': ones((natoms,))} for _ in range(nelectrons)]
        return tree_map(lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64)).to(dtype), params)

    def layers_apply(params, ae, r_ae, ee, r_ee):
        """nn.py:321-352."""
        ae_features, ee_features = ainet_features(ae, r_ae, ee, r_ee)
        temp = ae / r_ae
        y_sp = y_l_real(temp)                       # (...,N,A,4)
        y_df = y_l_real_high(temp, r_ae)            # (...,N,A,12,1)
        y_sp = y_sp.reshape(*y_sp.shape[:-2], -1)   # (...,N,4A)
        y_df = y_df.reshape(*y_df.shape[:-3], -1)   # (...,N,12A)
        y_one = torch.cat([y_sp, y_df.mean(dim=-1, keepdim=True), y_sp.mean(dim=-1, keepdim=True)], dim=-1)
        for i in range(len(hidden_dims_Ynlm)):
            p = params['streams_y'][i]['single_Ynlm']
            y_one = _residual(y_one, torch.tanh(linear_layer(y_one, p['w'], p['b'])))
        h_one, h_two = ae_features, ee_features
        for i in range(nlayers):
            p = params['streams'][i]
            h_in = construct_symmetric_features(h_one, h_two, nspins)
            h_con = torch.tanh(convolu_layer(nelectrons, h_in, p['convolutional']['w'], p['convolutional']['b']))
            h_next = torch.tanh(linear_layer(h_con, p['single']['w'], p['single']['b']))
            h_one = _residual(h_one, h_next)
            if 'double' in p:
                h_two = _residual(h_two, torch.tanh(linear_layer(h_two, p['double']['w'], p['double']['b'])))
        return h_one, y_one

    def jastrow_ee_apply(r_ee, p):
        """Jastrow.py:23-52: sum of cusp*r/(1+alpha*r) over parallel (1/4) and antiparallel (1/2) pairs."""
        total = 0.0
        for idx, cusp, alpha in ((par, 0.25, p['ee_par']), (anti, 0.5, p['ee_anti'])):
            if idx.shape[1] == 0:
                continue
            r = r_ee[..., idx[0], idx[1]]
            total = total + ((r * cusp) / (1.0 + alpha * r)).sum(dim=-1)
        return total

    def jastrow_ae_apply(r_ae, p):
        """Jastrow.py:74-93."""
        z2 = 2.0 * charges_t
        beta = p['ae']
        val = -1.0 * z2 ** 0.75 * (1.0 - torch.exp(-1.0 * z2 ** 0.25 * r_ae * beta)) / (2.0 * beta)
        return val.sum(dim=(-1, -2))

    def envelope_apply(r_ae_i, ae_i, p):
        """envelope.py:26-30 -> scalar per electron; r_ae_i (...,A), ae_i (...,A,3)."""
        return (torch.sum(torch.exp(-p['beta'] * r_ae_i ** 2) * p['alpha'], dim=-1) +
                torch.sum(torch.exp(-ae_i * p['pi']) * p['sigma'] * p['xi'], dim=(-1, -2)))

    def orbitals_apply(params, pos, spins, atoms, charges_unused=None):
        """nn.py:409-506 -> [M (...,N,N) complex]."""
        ae, ee, r_ae, r_ee = construct_input_features(pos, atoms, ndim=3)
        h, y = layers_apply(params['layers'], ae, r_ae, ee, r_ee)
        h_spin = [h[..., up, :], h[..., dn, :]]
        orbs = [linear_layer(hs, p['w'], p['b']) for hs, p in zip(h_spin, params['orbitals'])]
        wy = params['y'][0]['w']
        wy = wy / torch.linalg.norm(wy, dim=-1, keepdim=True)
        y_orbitals = linear_layer(y, wy)
        orbs = [torch.complex(o[..., ::2], o[..., 1::2]) for o in orbs]
        m = torch.cat(orbs, dim=-2)                                  # up rows then down rows
        env = torch.stack([envelope_apply(r_ae[..., i, :, 0], ae[..., i, :, :], params['envelope'][i])
                           for i in range(nelectrons)], dim=-1)     # (...,N) original electron order
        total = m * env[..., :, None] * y_orbitals
        jee = torch.exp(jastrow_ee_apply(r_ee[..., 0], params['jastrow_ee']) / nelectrons)
        jae = torch.exp(jastrow_ae_apply(r_ae[..., 0], params['jastrow_ae']) / nelectrons)
        return [total * jee[..., None, None] * jae[..., None, None]]

    def apply(params, pos, spins, atoms, charges_unused=None):
        """nn.py:545-551 -> (phase angle, log|psi|)."""
        return logdet_matmul(orbitals_apply(params, pos, spins, atoms))

    config = dict(nspins=tuple(nspins), charges=np.asarray(charges, dtype=np.float64), parallel_indices=par,
                  antiparallel_indices=anti, spin_up_indices=up.numpy(), spin_down_indices=dn.numpy(),
                  natoms=natoms, nelectrons=nelectrons)
    return Network(init=init, apply=apply, orbitals=orbitals_apply, config=config)


def tree_map(fn, tree):
    if isinstance(tree, dict):
        return {k: tree_map(fn, v) for k, v in tree.items()}
    if isinstance(tree, (list, tuple)):
        return [tree_map(fn, v) for v in tree]
    return fn(tree)


def select_output(f, argnum):
    """utils/utils.py:3-7."""
    return lambda *a, **k: f(*a, **k)[argnum]


def make_log_network(signed_network):
    """main/main_pp_adam_muti_GPU.py:119-121: log psi = log|psi| + i*phase."""
    def log_network(*a, **k):
        phase, mag = signed_network(*a, **k)
        return torch.complex(mag, phase)
    return log_network


# --------------------------------------------------------------------------
# autodiff helpers (mirror jax.grad / jax.linearize on a batch)
# --------------------------------------------------------------------------
def value_and_grad(fn, x, create_graph=False):
    x = x.detach().clone().requires_grad_(True)
    y = fn(x)
    g, = torch.autograd.grad(y.sum(), x, create_graph=create_graph)
    return y, g, x


def grad_and_hess_diag(fn, x):
    """gradient and the diagonal of the Hessian (3N jvp-of-grad in the reference)."""
    y, g, xr = value_and_grad(fn, x, create_graph=True)
    n = x.shape[-1]
    diag = []
    for i in range(n):
        gi, = torch.autograd.grad(g[..., i].sum(), xr, retain_graph=True)
        diag.append(gi[..., i])
    return y.detach(), g.detach(), torch.stack(diag, dim=-1).detach()


# --------------------------------------------------------------------------
# VMC sweep  (VMC/VMCmcstep.py)
# --------------------------------------------------------------------------
def limdrift(g, tau, acyrus):
    """VMCmcstep.py:11-14 -- v2 summed over the WHOLE array (quirk Q6)."""
    v2 = torch.sum(g ** 2)
    taueff = (torch.sqrt(1 + 2 * tau * acyrus * v2) - 1) / (acyrus * v2)
    return g * taueff


def walkers_update(logabs_f, params, data: AINetData, rand: Dict[str, torch.Tensor], tstep: float,
                   ndim: int, nelectrons: int, batch_size: int, signed: bool = False,
                   return_aux: bool = False):
    """VMCmcstep.py:28-111 (signed=False) and DMC/drift_diffusion.py:30-106 (signed=True).

    rand['gauss1'] (B,3N) and rand['gauss2'] (B,N,3N) are sqrt(tstep)*N(0,1); rand['rnd'] (B,N)
    is U[0,1) -- the reference draws them from one key with two shapes (quirk Q7).
    """
    B, N = batch_size, nelectrons
    x1 = data.positions
    atoms, spins = data.atoms[0], data.spins[0]
    f = lambda x: logabs_f(params, x, spins, atoms, None)
    wave_x1, grad, _ = value_and_grad(f, x1)
    wave_x1 = wave_x1.detach()
    grad_eff = limdrift(grad, tstep, 0.25)
    g = (grad_eff * tstep + rand['gauss1']).reshape(B, N, ndim)
    initial = x1.reshape(B, N, ndim)
    z = torch.zeros(B, N, N, ndim, dtype=x1.dtype)
    idx = torch.arange(N)
    z[:, idx, idx, :] = g
    x2 = (initial[:, None, :, :] + z).reshape(B, N, N * ndim)
    changed = g + initial
    wave_x2, grad_new, _ = value_and_grad(f, x2)
    wave_x2 = wave_x2.detach()
    grad_new_eff = limdrift(grad_new, tstep, 0.25)
    grad_eff_rep = grad_eff[:, None, :].expand(B, N, N * ndim)
    gauss2 = rand['gauss2']
    forward = gauss2 ** 2
    backward = (gauss2 + (grad_eff_rep + grad_new_eff) * tstep) ** 2
    t_prob = torch.exp((forward - backward) / (2 * tstep)).reshape(B, N, N, ndim).sum(dim=-1)
    t_pro = torch.diagonal(t_prob, dim1=-2, dim2=-1)
    wfratio = torch.exp(wave_x2 - wave_x1[:, None])
    acceptance = torch.abs(wfratio) ** 2 * t_pro
    if signed:
        acceptance = acceptance * torch.sign(wfratio)   # drift_diffusion.py:87-89 (sign of a positive number)
    cond = acceptance > rand['rnd']
    x_new = torch.where(cond[..., None], changed, initial)
    new_data = replace(data, positions=x_new.reshape(B, -1))
    if return_aux:
        aux = dict(accept=cond, acceptance=acceptance, grad=grad, grad_eff=grad_eff, grad_new_eff=grad_new_eff,
                   wave_x1=wave_x1, wave_x2=wave_x2, t_pro=t_pro, changed=changed)
        return new_data, aux
    return new_data


def main_monte_carlo(f, tstep: float, ndim: int, nelectrons: int, nsteps: int, batch_size: int):
    """VMCmcstep.py:121-140.  `key` is a list of `nsteps` random-array dicts."""
    logabs_f = select_output(f, 1)

    def mc_step(params, data, key: Sequence[Dict[str, torch.Tensor]]):
        for i in range(nsteps):
            data = walkers_update(logabs_f, params, data, key[i], tstep, ndim, nelectrons, batch_size)
        return data
    return mc_step


# --------------------------------------------------------------------------
# energies  (Energy/hamiltonian.py, Energy/pphamiltonian.py)
# --------------------------------------------------------------------------
def local_kinetic_energy(f, complex_output: bool = False):
    """hamiltonian.py:77-132 / pphamiltonian.py:67-106; batched over leading dims."""
    phase_f, logabs_f = select_output(f, 0), select_output(f, 1)

    def _lapl_over_f(params, data: AINetData):
        fn = lambda x: logabs_f(params, x, data.spins, data.atoms, data.charges)
        _, primal, diag = grad_and_hess_diag(fn, data.positions)
        result = -0.5 * diag.sum(dim=-1) - 0.5 * (primal ** 2).sum(dim=-1)
        if complex_output:
            fp = lambda x: phase_f(params, x, data.spins, data.atoms, data.charges)
            _, pprimal, pdiag = grad_and_hess_diag(fp, data.positions)
            result = torch.complex(result + 0.5 * (pprimal ** 2).sum(-1),
                                   -0.5 * pdiag.sum(-1) - (primal * pprimal).sum(-1))
        return result
    return _lapl_over_f


def potential_electron_electron(r_ee):
    """hamiltonian.py:177-189."""
    n = r_ee.shape[-2]
    iu = torch.triu_indices(n, n, 1)
    return (1.0 / r_ee[..., iu[0], iu[1], 0]).sum(dim=-1)


def potential_electron_nuclear(charges, r_ae):
    """hamiltonian.py:192-200."""
    return -torch.sum(charges / r_ae[..., 0], dim=(-1, -2))


def potential_nuclear_nuclear(charges, atoms):
    """hamiltonian.py:203-213."""
    r_aa = torch.linalg.norm(atoms[None, ...] - atoms[:, None], dim=-1)
    a = atoms.shape[0]
    iu = torch.triu_indices(a, a, 1)
    zz = charges[None, :] * charges[:, None]
    return (zz[iu[0], iu[1]] / r_aa[iu[0], iu[1]]).sum()


def potential_energy(r_ae, r_ee, atoms, charges):
    """hamiltonian.py:216-233."""
    return (potential_electron_electron(r_ee) + potential_electron_nuclear(charges, r_ae) +
            potential_nuclear_nuclear(charges, atoms))


def local_energy_ae(f, charges, nspins=None, use_scan=False, complex_output=False):
    """hamiltonian.py:236-260 -> _e_l(params, key, data) -> (E_L, None)."""
    ke = local_kinetic_energy(f, complex_output=complex_output)
    charges = torch.as_tensor(charges)

    def _e_l(params, key, data: AINetData):
        _, _, r_ae, r_ee = construct_input_features(data.positions, data.atoms)
        potential = potential_energy(r_ae, r_ee, data.atoms, charges.to(r_ae.dtype))
        return potential + ke(params, data), None
    return _e_l


# --------------------------------------------------------------------------
# ccECP  (pseudopotential/pseudopotential.py, pp_energy_test.py)
# --------------------------------------------------------------------------
def local_pp_energy(nelectrons, natoms, ndim, rn_local, local_coefficient, local_exponent):
    """pseudopotential.py:86-117 -> (..., N, A) table; r^(n-2) (quirk Q15)."""
    rn = torch.as_tensor(rn_local) - 2

    def pp_local_part_energy(data: AINetData):
        ae = data.positions.reshape(*data.positions.shape[:-1], -1, 1, ndim) - data.atoms
        r_ae = torch.linalg.norm(ae, dim=-1)
        part1 = -1 * torch.as_tensor(data.charges).to(r_ae.dtype) / r_ae
        r = r_ae[..., None]
        dt = r.dtype
        part2 = (torch.as_tensor(local_coefficient).to(dt) * r ** rn.to(dt) *
                 torch.exp(-torch.as_tensor(local_exponent).to(dt) * r ** 2)).sum(dim=-1)
        return part1 + part2
    return pp_local_part_energy


def get_non_v_l(ndim, nelectrons, natoms, rn_non_local, non_local_coefficient, non_local_exponent):
    """pseudopotential.py:134-165 -> v_l(r_ia) (..., N, A, L); r^n here (quirk Q15)."""
    def get_non_local_coe(data: AINetData):
        ae = data.positions.reshape(*data.positions.shape[:-1], -1, 1, ndim) - data.atoms
        r = torch.linalg.norm(ae, dim=-1)[..., None, None]
        dt = r.dtype
        out = (torch.as_tensor(non_local_coefficient).to(dt) * r ** torch.as_tensor(rn_non_local).to(dt) *
               torch.exp(-torch.as_tensor(non_local_exponent).to(dt) * r ** 2))
        return out.sum(dim=-1)
    return get_non_local_coe


def generate_quadrature_grids():
    """pseudopotential.py:181-225: 6+12+8+24 octahedral points (8-digit literals as written there)."""
    a, b = 0.70710678, 0.57735027
    OA = np.array([[-1, 0, 0], [0, -1, 0], [0, 0, -1], [0, 0, 1], [0, 1, 0], [1, 0, 0]], dtype=np.float64)
    OB = np.array([[-a, -a, 0.], [-a, 0., -a], [-a, 0., a], [-a, a, 0.], [0., -a, -a], [0., -a, a],
                   [0., a, -a], [0., a, a], [a, -a, 0.], [a, 0., -a], [a, 0., a], [a, a, 0.]])
    OC = np.array([[-b, -b, -b], [-b, -b, b], [-b, b, -b], [-b, b, b], [b, -b, -b], [b, -b, b],
                   [b, b, -b], [b, b, b]])
    d1 = OC * math.sqrt(3 / 11)
    OD1 = np.stack([d1[:, 0], d1[:, 1], d1[:, 2] * 3], axis=1)
    OD2 = np.stack([d1[:, 0], d1[:, 1] * 3, d1[:, 2]], axis=1)
    OD3 = np.stack([d1[:, 0] * 3, d1[:, 1], d1[:, 2]], axis=1)
    OD = np.concatenate([OD1, OD2, OD3], axis=0)
    weights = np.array([[4 / 315], [64 / 2835], [27 / 1280], [14641 / 725760]])
    return OA, OB, OC, OD, weights


def quadrature_table():
    """All 50 points and per-point weights in the reference's OA,OB,OC,OD order."""
    OA, OB, OC, OD, w = generate_quadrature_grids()
    pts = np.concatenate([OA, OB, OC, OD], axis=0)
    wts = np.concatenate([np.full(len(g), w[k, 0]) for k, g in enumerate((OA, OB, OC, OD))])
    return pts, wts


def random_rotations(rng: np.random.Generator, n: int):
    """Stand-in for jax.random.orthogonal (pseudopotential.py:234): Haar via QR."""
    q, r = np.linalg.qr(rng.standard_normal((n, 3, 3)))
    return q * np.sign(np.diagonal(r, axis1=-2, axis2=-1))[:, None, :]


def get_rot(rot):
    """pseudopotential.py:233-241 with the rotation supplied: Points = O @ rot, rot (...,3,3)."""
    OA, OB, OC, OD, weights = generate_quadrature_grids()
    dt = rot.dtype
    pts = tuple(torch.einsum('ik,...kl->...il', torch.as_tensor(o).to(dt), rot) for o in (OA, OB, OC, OD))
    return pts + (torch.as_tensor(weights).to(dt),)


def P_l(x, list_l):
    """pseudopotential.py:250-269 (== DMC/Tmoves.py:10-29): (2l+1)/(4pi) P_l(x)."""
    out = [1 / (4 * PI) * torch.ones_like(x), 3 / (4 * PI) * x, 5 / (4 * PI) * 0.5 * (3 * x * x - 1),
           7 / (4 * PI) * 0.5 * (5 * x * x * x - 3 * x)]
    return out[:list_l + 1]


def get_P_l(nelectrons, natoms, ndim, log_network_inner):
    """pseudopotential.py:272-318; Points (..., npts, 3) per walker."""
    def generate_points_information(data: AINetData, params, Points, weights):
        pos = data.positions
        lead = pos.shape[:-1]
        x2 = pos.reshape(*lead, nelectrons, ndim)
        ae = x2[..., :, None, :] - data.atoms
        r_ae = torch.linalg.norm(ae, dim=-1)[..., None]                          # (...,N,A,1)
        denominator = log_network_inner(params, pos, data.spins, data.atoms, data.charges)
        roted = r_ae[..., None] * Points[..., None, None, :, :]                    # (...,N,A,P,3)
        # quirk Q14: Frobenius norm over all points in the denominator
        cos_theta = (ae[..., None, :] * roted).sum(-1) / (
            torch.linalg.norm(ae, dim=-1)[..., None] * torch.linalg.norm(roted, dim=(-1, -2))[..., None])
        npts = Points.shape[-2]
        # quirk Q13: electron i is placed AT r_ia * n_hat (atom position not added)
        conf = x2[..., None, None, None, :, :].expand(*lead, nelectrons, natoms, npts, nelectrons, ndim).clone()
        for i in range(nelectrons):
            conf[..., i, :, :, i, :] = roted[..., i, :, :, :]
        conf = conf.reshape(*lead, nelectrons, natoms, npts, nelectrons * ndim)
        val = log_network_inner(params, conf, data.spins, data.atoms, data.charges)
        d = denominator
        for _ in range(3):
            d = d[..., None]
        ratios = val / d * weights                                                # quirk Q12
        return cos_theta, ratios, conf, weights, roted
    return generate_points_information


def total_energy_pseudopotential(get_local_pp_energy, get_nonlocal_pp_coes, get_P_l_fn, list_l):
    """pp_energy_test.py:45-105; `key` is the per-walker rotation matrix (...,3,3)."""
    def get_total_pp_energy(params, key, data: AINetData):
        local = get_local_pp_energy(data).sum(dim=(-1, -2))
        v_l = get_nonlocal_pp_coes(data)                                            # (...,N,A,L)
        *points, weights = get_rot(key)
        nonlocal_energy = 0.0
        for g, pts in enumerate(points):
            cos_theta, ratios, _, _, _ = get_P_l_fn(data, params, pts, weights[g])
            pl = torch.stack(P_l(cos_theta, list_l), dim=0)                         # (L,...,N,A,P)
            out = (pl * ratios).sum(dim=-1)                                         # (L,...,N,A)
            nonlocal_energy = nonlocal_energy + (out * v_l.movedim(-1, 0)).sum(dim=(0, -1, -2))
        return local + nonlocal_energy
    return get_total_pp_energy


def local_energy_ecp(f, lognetwork, charges, nspins, rn_local, local_coes, local_exps, rn_non_local,
                     non_local_coes, non_local_exps, natoms, nelectrons, ndim, list_l,
                     use_scan=False, complex_output=False):
    """pphamiltonian.py:130-190 -> _e_l(params, key, data) -> (complex E_L, None); no e-n Coulomb term."""
    ke = local_kinetic_energy(f, complex_output=complex_output)
    loc = local_pp_energy(nelectrons, natoms, ndim, rn_local, local_coes, local_exps)
    nl = get_non_v_l(ndim, nelectrons, natoms, rn_non_local, non_local_coes, non_local_exps)
    pts = get_P_l(nelectrons, natoms, ndim, lognetwork)
    pp = total_energy_pseudopotential(loc, nl, pts, list_l)
    charges = torch.as_tensor(charges)

    def _e_l(params, key, data: AINetData):
        _, _, _, r_ee = construct_input_features(data.positions, data.atoms)
        potential = potential_electron_electron(r_ee) + potential_nuclear_nuclear(
            charges.to(r_ee.dtype), data.atoms)
        kinetic = ke(params, data)
        return pp(params, key, data) + kinetic + potential, None
    return _e_l


def total_energy(local_energy_fn):
    """DMC/total_energy.py:9-32 / Loss/pploss.py:157-167 (forward part, single device)."""
    def _total(params, key, data):
        e_l, _ = local_energy_fn(params, key, data)
        loss = e_l.mean()
        diff = e_l - loss
        variance = (diff * diff.conj()).mean()
        return e_l, loss, variance
    return _total


def clip_local_values(local_values, mean_local_values, clip_scale, clip_from_median, center_at_clipped_value,
                      complex_output=False):
    """Loss/pploss.py:73-135 on one device (pmean / all_gather are identities)."""
    def clip_at_total_variation(values, center, scale):
        tv = torch.mean(torch.abs(values - center))
        return torch.clamp(values, min=center - scale * tv, max=center + scale * tv)
    if clip_from_median:
        clip_center = torch.as_tensor(np.median(local_values.real.detach().numpy()))      # jnp.median
    else:
        clip_center = mean_local_values
    if complex_output:
        ci = clip_center.imag if torch.is_complex(clip_center) else torch.zeros(())
        clipped = torch.complex(clip_at_total_variation(local_values.real, clip_center.real, clip_scale),
                                clip_at_total_variation(local_values.imag, ci, clip_scale))
    else:
        clipped = clip_at_total_variation(local_values, clip_center, clip_scale)
    diff_center = torch.mean(clipped) if center_at_clipped_value else mean_local_values
    return diff_center, clipped - diff_center


def make_loss(network_apply, local_energy_fn, clip_local_energy=0.0, clip_from_median=True,
              center_at_clipped_energy=True, complex_output=True):
    """Loss/pploss.py:137-223.  total_energy(params, key, data) -> (loss, aux dict); value_and_grad restates what
    jax.value_and_grad sees through total_energy_jvp: the tangent functional (:204-222) is linear in psi_tangent =
    J t, so its gradient is autograd of the same expression with psi(params) in place of psi_tangent and the local
    energies held constant."""
    def total_energy(params, key, data):
        e_l, _ = local_energy_fn(params, key, data)
        loss = e_l.mean()
        diff = e_l - loss
        variance = (diff * diff.conj()).mean().real
        return loss, dict(variance=variance, local_energy=e_l, clipped_energy=e_l)

    def value_and_grad(params, key, data):
        loss, aux = total_energy(params, key, data)             # (the kinetic term differentiates inside: no no_grad)
        loss = loss.detach()
        aux = {k: v.detach() for k, v in aux.items()}
        if clip_local_energy > 0.0:
            aux['clipped_energy'], diff = clip_local_values(aux['local_energy'], loss, clip_local_energy,
                                                            clip_from_median, center_at_clipped_energy, complex_output)
        else:
            diff = aux['local_energy'] - loss
        leaves = []
        def req(t):
            if isinstance(t, dict):
                return {k: req(v) for k, v in t.items()}
            if isinstance(t, (list, tuple)):
                return [req(v) for v in t]
            v = t.clone().requires_grad_(True)
            leaves.append(v)
            return v
        p = req(params)
        phase, logabs = network_apply(p, data.positions, data.spins[0], data.atoms[0], None)
        B = data.positions.shape[0]
        if complex_output:
            psi = torch.complex(logabs, phase)                      # log_network = mag + 1j * phase
            clipped_el = diff + aux['clipped_energy']
            term1 = torch.sum(clipped_el * psi.conj()) + torch.sum(clipped_el.conj() * psi)
            term2 = torch.sum(aux['clipped_energy'] * psi.real)
            surrogate = (term1 - 2 * term2).real / B
            out_loss = loss.real
        else:
            surrogate = torch.dot(logabs, diff.real) / B
            out_loss = loss.real if torch.is_complex(loss) else loss
        grads = torch.autograd.grad(surrogate, leaves, allow_unused=True)
        it = iter(grads)
        def rebuild(t):
            if isinstance(t, dict):
                return {k: rebuild(v) for k, v in t.items()}
            if isinstance(t, (list, tuple)):
                return [rebuild(v) for v in t]
            g = next(it)
            return g if g is not None else torch.zeros_like(t)
        return (out_loss, aux), rebuild(params)

    total_energy.value_and_grad = value_and_grad
    return total_energy


# --------------------------------------------------------------------------
# all-electron Metropolis-Hastings (AIQMCrelease2/MonteCarloSample/mcstep.py:12-124 = ferminet/mcmc.py:67-150)
# --------------------------------------------------------------------------
def _harmonic_mean(x, atoms):
    """mcstep.py:12-16; x (B,N,1,3), atoms (A,3) -> (B,N,1,1)."""
    ae = x - atoms[None, ...]
    r_ae = torch.linalg.norm(ae, dim=-1, keepdim=True)
    return 1.0 / torch.mean(1.0 / r_ae, dim=-2, keepdim=True)


def _log_prob_gaussian(x, mu, sigma):
    """mcstep.py:19-23."""
    numer = torch.sum(-0.5 * ((x - mu) ** 2) / (sigma ** 2), dim=[1, 2, 3])
    denom = x.shape[-1] * torch.sum(torch.log(sigma), dim=[1, 2, 3])
    return numer - denom


def mh_update(params, f, data: AINetData, rand, lp_1, num_accepts, stddev=0.02, atoms=None, ndim=3):
    """mcstep.py:37-68 with the two draws explicit: rand = dict(noise (B,N,1,3), u (B,))."""
    x1 = data.positions
    n = x1.shape[0]
    x1 = x1.reshape(n, -1, 1, ndim)
    hmean1 = _harmonic_mean(x1, atoms)
    x2 = x1 + stddev * hmean1 * rand['noise'].reshape(x1.shape)
    lp_2 = 2.0 * f(params, x2.reshape(n, -1), data.spins, data.atoms, data.charges)
    hmean2 = _harmonic_mean(x2, atoms)
    lq_1 = _log_prob_gaussian(x1, x2, stddev * hmean1)
    lq_2 = _log_prob_gaussian(x2, x1, stddev * hmean2)
    ratio = lp_2 + lq_2 - lp_1 - lq_1
    cond = ratio > torch.log(rand['u'])                                    # mh_accept :26-34
    x_new = torch.where(cond[..., None], x2.reshape(n, -1), x1.reshape(n, -1))
    lp_new = torch.where(cond, lp_2, lp_1)
    return replace(data, positions=x_new), lp_new, num_accepts + cond.sum(), cond


def make_mcmc_step(batch_network, batch_per_device, steps=10, atoms=None, ndim=3, blocks=1):
    """mcstep.py:71-104; rand = dict(noise (steps,B,3N), u (steps,B))."""
    def mcmc_step(params, data: AINetData, rand, width):
        lp = 2.0 * batch_network(params, data.positions, data.spins, data.atoms, data.charges)
        acc = torch.zeros((), dtype=torch.float64)
        masks = []
        for s in range(steps * blocks):
            data, lp, acc, cond = mh_update(params, batch_network, data, dict(noise=rand['noise'][s], u=rand['u'][s]), lp,
                                            acc, stddev=width, atoms=atoms, ndim=ndim)
            masks.append(cond)
        return data, acc / (steps * blocks * batch_per_device), torch.stack(masks), lp
    return mcmc_step


def update_mcmc_width(t, width, adapt_frequency, pmove, pmoves, pmove_max=0.55, pmove_min=0.5):
    """mcstep.py:107-124."""
    t_since = t % adapt_frequency
    pmoves[t_since] = float(pmove)
    if t > 0 and t_since == 0:
        if np.mean(pmoves) > pmove_max:
            width *= 1.1
        elif np.mean(pmoves) < pmove_min:
            width /= 1.1
    return width, pmoves


# --------------------------------------------------------------------------
# correlated sampling (correlatedsamples/corrsamples.py:23-47, jacobianWeights.py:22-51), one walker at a time
# --------------------------------------------------------------------------
def correlated_samples(atoms, new_atoms, pos):
    deltaR = new_atoms - atoms
    ae, ee, r_ae, r_ee = construct_input_features(pos, atoms, ndim=3)
    k_r_R = 1 / (r_ae ** 4)                                                       # (N,A,1)
    denominator = torch.sum(torch.sum(k_r_R, dim=-1), dim=-1, keepdim=True)     # (N,1)
    output = k_r_R / denominator[:, None, :]                                      # vmap(devided)
    move = torch.sum(output * deltaR[None, :, :], dim=1)                          # vmap(multiply), sum over atoms
    return pos + move.reshape(-1)


def weights_jacobian(pos, atoms, new_atoms):
    ae, ee, r_ae, r_ee = construct_input_features(pos, atoms, ndim=3)
    deltaR = new_atoms - atoms

    def jacobian_element(ae_inner, atoms_inner):
        temp1 = torch.sum(-4 * torch.abs(ae_inner) ** (-5) * (1 - atoms_inner), dim=-1, keepdim=True)
        temp2 = deltaR[:, 0] * (-4 * torch.abs(ae_inner) ** (-5) * (1 - atoms_inner))
        return torch.sum(temp2 / temp1, dim=-1, keepdim=True) + 1

    x = jacobian_element(ae[:, :, 0], atoms[:, 0])
    y = jacobian_element(ae[:, :, 1], atoms[:, 1])
    z = jacobian_element(ae[:, :, 2], atoms[:, 2])
    return torch.prod(x * y * z)


# --------------------------------------------------------------------------
# DMC  (DMC/drift_diffusion.py, S_matrix.py, branch.py, dmc.py, main_dmc.py)
# --------------------------------------------------------------------------
def propose_drift_diffusion(logabs_f, tstep, ndim, nelectrons, batch_size):
    """drift_diffusion.py:25-107 -> (new_data, tdamp, grad_eff_old, grad_new_eff_s)."""
    def drift_diffusion(params, rand, data: AINetData):
        new_data, aux = walkers_update(logabs_f, params, data, rand, tstep, ndim, nelectrons, batch_size,
                                       signed=True, return_aux=True)
        x_new = new_data.positions.reshape(batch_size, nelectrons, ndim)
        tdamp = torch.sum(x_new) / torch.sum(aux['changed'])          # quirk Q19
        atoms, spins = data.atoms[0], data.spins[0]
        f = lambda x: logabs_f(params, x, spins, atoms, None)
        _, g_s, _ = value_and_grad(f, new_data.positions)
        return new_data, tdamp, aux['grad_eff'], limdrift(g_s, tstep, 0.25), aux
    return drift_diffusion


def comput_S(e_trial, e_est, branchcut, v2, tau, eloc, nelec):
    """S_matrix.py:4-25; the e_cut clamp is a GLOBAL min (quirk Q20)."""
    v2 = torch.sum(v2, dim=-1)
    eloc = torch.real(eloc) if torch.is_complex(eloc) else eloc
    e_est = torch.as_tensor(e_est)
    e_trial = torch.as_tensor(e_trial)
    e_cut = e_est - eloc
    m = torch.min(torch.stack([torch.abs(e_cut).expand_as(eloc), torch.as_tensor(branchcut).expand_as(eloc)]))
    e_cut = m * torch.sign(e_cut)
    denominator = 1 + (v2 * tau / nelec) ** 2
    return e_trial - e_est + e_cut / denominator


def branch(weights: torch.Tensor, u: float):
    """branch.py:10-34 with the uniform `u` supplied -> (new weight, newinds)."""
    n = weights.shape[0]
    probability = torch.cumsum(weights, dim=0)
    wtot = probability[-1]
    base = u * wtot
    comb = torch.remainder(base + torch.arange(n, dtype=weights.dtype) * (wtot / n), wtot)
    newinds = torch.searchsorted(probability, comb)          # side='left', as jnp.searchsorted
    return wtot / n, newinds


def reconfigure(positions: torch.Tensor, newinds: torch.Tensor, noise: torch.Tensor):
    """main_dmc.py:218-231: keep unique survivors (sorted), pad with last + U(0,1) noise rows."""
    uniq = torch.unique(newinds)
    temp = positions[uniq]
    nmiss = positions.shape[0] - uniq.shape[0]
    if nmiss > 0:
        temp = torch.cat([temp, temp[-1][None, :] + noise[:nmiss]], dim=0)
    return temp, uniq.shape[0]


# --------------------------------------------------------------------------
# T-moves  (DMC/Tmoves.py:32-225)
# --------------------------------------------------------------------------
def _lex_gt(a: torch.Tensor, b: torch.Tensor):
    """jnp `>` on complex numbers is lexicographic: (real, then imag)   (quirk Q25)."""
    return (a.real > b.real) | ((a.real == b.real) & (a.imag > b.imag))


def _lex_le(a: torch.Tensor, b: torch.Tensor):
    return (a.real < b.real) | ((a.real == b.real) & (a.imag <= b.imag))


def searchsorted_scan(arr: torch.Tensor, query: complex):
    """jnp.searchsorted(arr, query) with side='left', method='scan' (jax/_src/numpy/lax_numpy.py
    `_searchsorted_via_scan`): a fixed-trip-count bisection that is well defined on UNSORTED input,
    which is what Tmoves.py:141-149 feeds it (complex cdf, lexicographic compare)."""
    n = arr.shape[0]
    low, high = 0, n
    q = torch.as_tensor(query, dtype=arr.dtype)
    for _ in range(int(np.ceil(np.log2(n + 1)))):
        mid = (low + high) // 2
        if bool(_lex_le(q, arr[mid])):
            high = mid
        else:
            low = mid
    return high


def compute_tmoves(list_l, tstep, nelectrons, natoms, ndim, lognetwork, Rn_non_local, Non_local_coes,
                   Non_local_exps):
    """Tmoves.py:32-225 for ONE walker: calculate_ratio_weight_tmoves(data, params, key) with
    key = dict(rot (3,3), u scalar in [0,1), rnd (N,1)) -- the three draws the reference makes from its key
    (:69, :146, :216-217).  Returns (final_configuration (3N,), acceptance (N,1))."""
    get_P = get_P_l(nelectrons, natoms, ndim, lognetwork)
    get_v = get_non_v_l(ndim, nelectrons, natoms, Rn_non_local, Non_local_coes, Non_local_exps)
    N = nelectrons

    def calculate_ratio_weight_tmoves(data: AINetData, params, key):
        *points, weights = get_rot(key['rot'])
        v_l = get_v(data)                                            # (N,A,L)
        fwd, ratios_g, coords_g = [], [], []
        for g, pts in enumerate(points):
            cos_theta, ratios, _, _, roted = get_P(data, params, pts, weights[g])
            pl = torch.stack(P_l(cos_theta, list_l), dim=0)          # (L,N,A,P)
            wts = ((torch.exp(-tstep * v_l.movedim(-1, 0)) - 1)[..., None] * pl).sum(dim=0)     # (N,A,P)
            t_amp = ratios * wts
            fwd.append(torch.where(_lex_gt(t_amp, torch.zeros_like(t_amp)), t_amp, torch.zeros_like(t_amp)))
            ratios_g.append(ratios)
            coords_g.append(roted)
        norm = 1 + sum((weights[g] * fwd[g]).sum() for g in range(4))
        total = torch.cat(fwd, dim=-1).reshape(N, -1)                # (N, A*50): atom-major, points OA..OD
        total = torch.cat([torch.ones(N, 1, dtype=total.dtype), total], dim=-1)
        cdf = torch.cumsum(total / norm, dim=-1)
        r = complex(float(key['u']) + 1.0, 0.0)
        nmov = total.shape[1]
        sel = [searchsorted_scan(cdf[i], r) for i in range(N)]
        sel = [s if s < nmov else 0 for s in sel]
        coords = torch.cat(coords_g, dim=2).reshape(N, -1, ndim)
        x1 = data.positions.reshape(N, ndim)
        coords = torch.cat([x1[:, None, :], coords], dim=1)          # (N, 1+50A, 3)
        ratio_total = torch.cat(ratios_g, dim=-1).reshape(N, -1)
        ratio_total = torch.cat([torch.ones(N, 1, dtype=ratio_total.dtype), ratio_total], dim=1)
        new_conf, back = [], []
        for i in range(N):
            mv = sel[i]
            new_conf.append(coords[i, mv])
            # quirk Q18: t_amp[move] indexes the ELECTRON axis with a move index (clamped read)
            back.append(total[min(mv, N - 1)] * (1 / ratio_total[i, mv]))
        new_conf, back = torch.stack(new_conf), torch.stack(back)    # (N,3), (N,1+50A)
        wf = torch.cat([torch.zeros(1, 1, dtype=weights.dtype), weights])           # (5,1)
        sl = [(1, 19), (19, 55), (55, 79), (79, 151)]                # hard-coded for A=3 (quirk Q18)
        back_norm = 1.0 + wf[0] * back[:, 0]
        for g, (lo, hi) in enumerate(sl):
            back_norm = back_norm + (wf[g + 1] * back[:, lo:hi]).sum(dim=-1)
        acceptance = (norm / back_norm).real.reshape(-1, 1)
        cond = acceptance > key['rnd'].reshape(-1, 1)
        final = torch.where(cond, new_conf, x1).reshape(-1)
        aux = dict(selected=torch.tensor(sel), norm=norm, back_norm=back_norm, cdf=cdf, total=total, accept=cond)
        return final, acceptance, aux
    return calculate_ratio_weight_tmoves


def dmc_propagate(signed_network, tstep, nelectrons, natoms, ndim, batch_size, charges, rn_local, local_coes,
                  local_exps, rn_non_local, non_local_coes, non_local_exps, list_l=2):
    """DMC/dmc.py:72-93 composed from the pieces above (single device).  key = dict(tmove=dict(rot (B,3,3), u (B,),
    rnd (B,N)), sweep=dict(gauss1, gauss2, rnd), rot (B,3,3))."""
    lognet = make_log_network(signed_network)
    tm = compute_tmoves(list_l, tstep, nelectrons, natoms, ndim, lognet, rn_non_local, non_local_coes, non_local_exps)
    dd = propose_drift_diffusion(select_output(signed_network, 1), tstep, ndim, nelectrons, batch_size)
    le = local_energy_ecp(signed_network, lognet, charges, None, rn_local, local_coes, local_exps, rn_non_local,
                          non_local_coes, non_local_exps, natoms, nelectrons, ndim, list_l)

    def dmc_propagate_run(params, key, data: AINetData, weights, branchcut_start, e_trial, e_est):
        B = data.positions.shape[0]
        atoms, spins, ch = data.atoms[0], data.spins[0], data.charges[0]
        pos = []
        for b in range(B):                                   # tmoves_pmap: vmap over walkers
            d1 = AINetData(positions=data.positions[b], spins=spins, atoms=atoms, charges=ch)
            k = key['tmove']
            final, _, _ = tm(d1, params, dict(rot=k['rot'][b], u=float(k['u'][b]), rnd=k['rnd'][b]))
            pos.append(final)
        t_move_data = AINetData(positions=torch.stack(pos), spins=data.spins, atoms=data.atoms, charges=data.charges)
        new_data, tdamp, grad_eff_old, grad_new_eff, _ = dd(params, key['sweep'], t_move_data)
        flat = lambda d: AINetData(positions=d.positions, spins=spins, atoms=atoms, charges=ch)
        eloc_old, _ = le(params, key['rot'], flat(data))
        eloc_new, _ = le(params, key['rot'], flat(new_data))
        s_old = comput_S(e_trial, e_est, branchcut_start, grad_eff_old ** 2, tstep, eloc_old, nelectrons)
        s_new = comput_S(e_trial, e_est, branchcut_start, grad_new_eff ** 2, tstep, eloc_new, nelectrons)
        return eloc_new, torch.exp(tstep * tdamp * (0.5 * s_new + 0.5 * s_old)) * weights, new_data
    return dmc_propagate_run
