"""Host-side static description of a system and the flat parameter pack the kernels read.

Mirrors make_ai_net's arguments (AIQMCrelease3/wavefunction_Ynlm/nn.py:511-526), the index
tables of AIQMCrelease3/spin_indices.py:5-46 and the parameter pytree of nn.py:203-278,370-407.
ctypes structures match include/aiqmc_b200.h field for field.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any, Mapping, Sequence

import numpy as np

MAX_ELEC, MAX_ATOMS, ECP_MAX_L, ECP_MAX_K, NQUAD = 32, 16, 4, 4, 50


class AiqmcSystem(C.Structure):
    _fields_ = [("n_elec", C.c_int32), ("n_atoms", C.c_int32), ("n_up", C.c_int32), ("n_dn", C.c_int32),
                ("n_up_rows", C.c_int32), ("sigma", C.c_int32 * MAX_ELEC)]


class AiqmcEcp(C.Structure):
    _fields_ = [("k_loc", C.c_int32), ("n_l", C.c_int32), ("k_nl", C.c_int32), ("pad_", C.c_int32),
                ("rn_local", C.c_double * ECP_MAX_K * MAX_ATOMS),
                ("local_coes", C.c_double * ECP_MAX_K * MAX_ATOMS),
                ("local_exps", C.c_double * ECP_MAX_K * MAX_ATOMS),
                ("rn_non_local", C.c_double * ECP_MAX_K * ECP_MAX_L * MAX_ATOMS),
                ("non_local_coes", C.c_double * ECP_MAX_K * ECP_MAX_L * MAX_ATOMS),
                ("non_local_exps", C.c_double * ECP_MAX_K * ECP_MAX_L * MAX_ATOMS),
                ("quad_pts", C.c_double * 3 * NQUAD), ("quad_wts", C.c_double * NQUAD)]


class AiqmcLayout(C.Structure):
    _fields_ = [("conv_w", C.c_int32 * 3), ("conv_b", C.c_int32 * 3), ("sing_w", C.c_int32 * 3),
                ("sing_b", C.c_int32 * 3), ("dbl_w", C.c_int32 * 2), ("dbl_b", C.c_int32 * 2),
                ("yn_w", C.c_int32 * 3), ("yn_b", C.c_int32 * 3), ("orb_w", C.c_int32 * 2),
                ("orb_b", C.c_int32 * 2), ("y_w", C.c_int32), ("jas_alpha", C.c_int32), ("jas_cusp", C.c_int32),
                ("jas_beta", C.c_int32), ("jas_c34", C.c_int32), ("jas_c14", C.c_int32), ("env_pi", C.c_int32),
                ("env_sx", C.c_int32), ("env_alpha", C.c_int32), ("env_beta", C.c_int32), ("atoms", C.c_int32),
                ("charges", C.c_int32), ("total", C.c_int32)]


def jastrow_indices_ee(spins, nelectrons: int):
    """spin_indices.py:5-19: (2, n) index lists of parallel / antiparallel pairs i<j, row-major."""
    s = np.asarray(spins, dtype=np.float64).reshape(nelectrons)
    prod = np.triu(s[None, :] * s[:, None], k=1)
    par = np.array(np.nonzero(prod > 0))
    anti = np.array(np.nonzero(prod < 0))
    return par, anti, par.shape[1], anti.shape[1]


def spin_indices_h(spins):
    """spin_indices.py:38-46."""
    s = np.asarray(spins, dtype=np.float64)
    return np.nonzero(s > 0)[0], np.nonzero(s < 0)[0]


@dataclass
class SystemSpec:
    """Everything static that signed_network closes over in the reference."""
    nelectrons: int
    natoms: int
    nspins: tuple
    atoms: np.ndarray            # (A,3)
    charges: np.ndarray          # (A,)
    spin_up_indices: np.ndarray
    spin_down_indices: np.ndarray
    parallel_indices: np.ndarray
    antiparallel_indices: np.ndarray

    @staticmethod
    def from_spins(atoms, charges, spins, nspins=None) -> "SystemSpec":
        spins = np.asarray(spins, dtype=np.float64)
        n = spins.shape[0]
        up, dn = spin_indices_h(spins)
        par, anti, _, _ = jastrow_indices_ee(spins, n)
        nspins = tuple(nspins) if nspins is not None else (len(up), len(dn))
        atoms = np.asarray(atoms, dtype=np.float64).reshape(-1, 3)
        return SystemSpec(n, atoms.shape[0], nspins, atoms, np.asarray(charges, dtype=np.float64).reshape(-1),
                          up, dn, par, anti)

    def c_struct(self) -> AiqmcSystem:
        n = self.nelectrons
        if not (2 <= n <= MAX_ELEC and 1 <= self.natoms <= MAX_ATOMS):
            raise ValueError(f"unsupported system size N={n}, A={self.natoms}")
        if self.nspins[0] <= 0 or self.nspins[1] <= 0 or sum(self.nspins) != n:
            raise ValueError("both spin blocks must be non-empty and sum to nelectrons")
        if len(self.spin_up_indices) + len(self.spin_down_indices) != n:
            raise ValueError("spin index lists must cover every electron")
        s = AiqmcSystem()
        s.n_elec, s.n_atoms, s.n_up, s.n_dn = n, self.natoms, int(self.nspins[0]), int(self.nspins[1])
        s.n_up_rows = len(self.spin_up_indices)
        for k, e in enumerate(list(self.spin_up_indices) + list(self.spin_down_indices)):
            s.sigma[k] = int(e)
        return s


def _np(x) -> np.ndarray:
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x, dtype=np.float64)


def pack_params(layout: AiqmcLayout, params: Mapping[str, Any], spec: SystemSpec) -> np.ndarray:
    """Flatten the reference parameter pytree into the kernel's packed float64 buffer.

    Pure re-layout except for three host-side precomputations that depend on parameters only:
    the row normalisation of params['y'][0]['w'] (nn.py:449-451), sigma*xi of the envelope
    (envelope.py:30) and the (N,N) pair tables for the e-e Pade Jastrow (Jastrow.py:23-41).
    """
    n, a = spec.nelectrons, spec.natoms
    buf = np.zeros(layout.total, dtype=np.float64)

    def put(off, arr, size):
        arr = _np(arr).reshape(-1)
        if arr.size != size:
            raise ValueError(f"parameter leaf has {arr.size} elements, expected {size}")
        buf[off:off + size] = arr

    d = [12 * a + 8, 20, 20]
    ky = [4 * a + 2, 6, 6]
    st, sty = params['layers']['streams'], params['layers']['streams_y']
    for l in range(3):
        put(layout.conv_w[l], st[l]['convolutional']['w'], n * d[l])
        put(layout.conv_b[l], st[l]['convolutional']['b'], n * d[l] // 4)
        put(layout.sing_w[l], st[l]['single']['w'], d[l])
        put(layout.sing_b[l], st[l]['single']['b'], 4)
        if l < 2:
            put(layout.dbl_w[l], st[l]['double']['w'], 16)
            put(layout.dbl_b[l], st[l]['double']['b'], 4)
        put(layout.yn_w[l], sty[l]['single_Ynlm']['w'], ky[l] * 6)
        put(layout.yn_b[l], sty[l]['single_Ynlm']['b'], 6)
    for s in range(2):
        put(layout.orb_w[s], params['orbitals'][s]['w'], 8 * n)
        put(layout.orb_b[s], params['orbitals'][s]['b'], 2 * n)
    wy = _np(params['y'][0]['w'])
    put(layout.y_w, wy / np.linalg.norm(wy, axis=-1, keepdims=True), 6 * n)
    alpha, cusp = np.ones((n, n)), np.zeros((n, n))
    for idx, c, leaf in ((spec.parallel_indices, 0.25, 'ee_par'), (spec.antiparallel_indices, 0.5, 'ee_anti')):
        vals = _np(params['jastrow_ee'][leaf]).reshape(-1)
        idx = np.asarray(idx).reshape(2, -1)
        alpha[idx[0], idx[1]] = vals
        cusp[idx[0], idx[1]] = c
    put(layout.jas_alpha, alpha, n * n)
    put(layout.jas_cusp, cusp, n * n)
    put(layout.jas_beta, params['jastrow_ae']['ae'], n * a)
    put(layout.jas_c34, (2.0 * spec.charges) ** 0.75, a)
    put(layout.jas_c14, (2.0 * spec.charges) ** 0.25, a)
    env = params['envelope']
    put(layout.env_pi, np.stack([_np(e['pi']) for e in env]), n * a * 3)
    put(layout.env_sx, np.stack([_np(e['sigma']) * _np(e['xi']) for e in env]), n * a * 3)
    put(layout.env_alpha, np.stack([_np(e['alpha']) for e in env]), n)
    put(layout.env_beta, np.stack([_np(e['beta']) for e in env]), n * a)
    put(layout.atoms, spec.atoms, 3 * a)
    put(layout.charges, spec.charges, a)
    return buf


def unpack_param_grad(layout: AiqmcLayout, gpacked, params: Mapping[str, Any], spec: SystemSpec) -> dict:
    """Transpose of pack_params: a gradient w.r.t. the packed buffer (aiqmc_psi_param_grad) -> a gradient pytree with
    the structure of the reference's params (nn.py:203-278,370-407).  The three host-side precomputations of
    pack_params get their chain rule here: row normalisation of params['y'][0]['w'], sigma * xi of the envelope, the
    scatter of ee_par / ee_anti into the (N,N) pair table.  Leaves the wavefunction does not use come back as zeros."""
    n, a = spec.nelectrons, spec.natoms
    g = _np(gpacked).reshape(-1)
    if g.size != layout.total:
        raise ValueError(f"packed gradient has {g.size} elements, expected {layout.total}")
    take = lambda off, shape: g[off:off + int(np.prod(shape))].reshape(shape).copy()
    d = [12 * a + 8, 20, 20]
    ky = [4 * a + 2, 6, 6]
    streams, streams_y = [], []
    for l in range(3):
        lp = {'convolutional': {'w': take(layout.conv_w[l], (n, d[l])), 'b': take(layout.conv_b[l], (n, d[l] // 4))},
              'single': {'w': take(layout.sing_w[l], (d[l] // 4, 4)), 'b': take(layout.sing_b[l], (4,))}}
        if l < 2:
            lp['double'] = {'w': take(layout.dbl_w[l], (4, 4)), 'b': take(layout.dbl_b[l], (4,))}
        streams.append(lp)
        streams_y.append({'single_Ynlm': {'w': take(layout.yn_w[l], (ky[l], 6)), 'b': take(layout.yn_b[l], (6,))}})
    out = {'layers': {'input': {}, 'streams': streams, 'streams_y': streams_y}}
    out['orbitals'] = [{'w': take(layout.orb_w[s], (4, 2 * n)), 'b': take(layout.orb_b[s], (2 * n,))} for s in range(2)]
    wy = _np(params['y'][0]['w'])
    nrm = np.linalg.norm(wy, axis=-1, keepdims=True)
    what, gy = wy / nrm, take(layout.y_w, (6, n))
    out['y'] = [{'w': (gy - what * np.sum(what * gy, axis=-1, keepdims=True)) / nrm}]
    ga = take(layout.jas_alpha, (n, n))
    ee = {}
    for idx, leaf in ((spec.parallel_indices, 'ee_par'), (spec.antiparallel_indices, 'ee_anti')):
        idx = np.asarray(idx).reshape(2, -1)
        ee[leaf] = ga[idx[0], idx[1]].reshape(_np(params['jastrow_ee'][leaf]).shape)
    out['jastrow_ee'] = ee
    out['jastrow_ae'] = {'ae': take(layout.jas_beta, (n, a))}
    gpi, gsx = take(layout.env_pi, (n, a, 3)), take(layout.env_sx, (n, a, 3))
    gal, gbe = take(layout.env_alpha, (n,)), take(layout.env_beta, (n, a))
    env = []
    for e, pe in enumerate(params['envelope']):
        sigma, xi = _np(pe['sigma']), _np(pe['xi'])
        leaf = {'pi': gpi[e], 'sigma': gsx[e] * xi, 'xi': np.sum(gsx[e] * sigma).reshape(xi.shape),
                'alpha': gal[e].reshape(_np(pe['alpha']).shape), 'beta': gbe[e]}
        for k in pe:
            if k not in leaf:
                leaf[k] = np.zeros_like(_np(pe[k]))
        env.append(leaf)
    out['envelope'] = env
    return out


def quadrature_table():
    """pseudopotential.py:181-225: the 50 unrotated points (8-digit literals as written in the
    reference) and their per-point weights, order OA, OB, OC, OD."""
    a, b = 0.70710678, 0.57735027
    OA = np.array([[-1, 0, 0], [0, -1, 0], [0, 0, -1], [0, 0, 1], [0, 1, 0], [1, 0, 0]], dtype=np.float64)
    OB = np.array([[-a, -a, 0.], [-a, 0., -a], [-a, 0., a], [-a, a, 0.], [0., -a, -a], [0., -a, a],
                   [0., a, -a], [0., a, a], [a, -a, 0.], [a, 0., -a], [a, 0., a], [a, a, 0.]])
    OC = np.array([[-b, -b, -b], [-b, -b, b], [-b, b, -b], [-b, b, b], [b, -b, -b], [b, -b, b],
                   [b, b, -b], [b, b, b]])
    d1 = OC * np.sqrt(3 / 11)
    OD = np.concatenate([np.stack([d1[:, 0], d1[:, 1], d1[:, 2] * 3], axis=1),
                         np.stack([d1[:, 0], d1[:, 1] * 3, d1[:, 2]], axis=1),
                         np.stack([d1[:, 0] * 3, d1[:, 1], d1[:, 2]], axis=1)], axis=0)
    w = [4 / 315, 64 / 2835, 27 / 1280, 14641 / 725760]
    pts = np.concatenate([OA, OB, OC, OD], axis=0)
    wts = np.concatenate([np.full(len(g), w[k]) for k, g in enumerate((OA, OB, OC, OD))])
    return pts, wts


def make_ecp(natoms: int, rn_local, local_coes, local_exps, rn_non_local, non_local_coes, non_local_exps,
             list_l: int) -> AiqmcEcp:
    """Pack the arrays of example/single_atom_C/single_atom_C.py:13-23 into the padded C struct."""
    e = AiqmcEcp()
    rl, lc, le = (np.asarray(x, dtype=np.float64).reshape(natoms, -1) for x in (rn_local, local_coes, local_exps))
    rn, nc, ne = (np.asarray(x, dtype=np.float64) for x in (rn_non_local, non_local_coes, non_local_exps))
    rn, nc, ne = (x.reshape(natoms, x.shape[1], -1) for x in (rn, nc, ne))
    e.k_loc, e.n_l, e.k_nl = rl.shape[1], list_l + 1, rn.shape[2]
    if e.k_loc > ECP_MAX_K or e.k_nl > ECP_MAX_K or e.n_l > ECP_MAX_L or rn.shape[1] != e.n_l:
        raise ValueError("ECP table exceeds compiled limits or list_l+1 != number of channels")
    for a in range(natoms):
        for k in range(e.k_loc):
            e.rn_local[a][k], e.local_coes[a][k], e.local_exps[a][k] = rl[a, k], lc[a, k], le[a, k]
        for l in range(e.n_l):
            for k in range(e.k_nl):
                e.rn_non_local[a][l][k], e.non_local_coes[a][l][k], e.non_local_exps[a][l][k] = \
                    rn[a, l, k], nc[a, l, k], ne[a, l, k]
    pts, wts = quadrature_table()
    for p in range(NQUAD):
        e.quad_wts[p] = wts[p]
        for c in range(3):
            e.quad_pts[p][c] = pts[p, c]
    return e
