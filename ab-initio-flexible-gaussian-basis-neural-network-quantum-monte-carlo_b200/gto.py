"""Host side of the contracted-Gaussian basis kernel: basis-file parsing in the format of
AIQMC/C.cc-pVDZ.nwchem, the shell table of include/aiqmc_b200.h, and `GaussianBasis.eval` which mirrors
ferminet/utils/gto.py:338-389 `Mol.eval_gto(coords) -> [G, nAO]` (plus gradient and Laplacian)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import lib as _lib

MAX_PRIM, MAX_SHELLS, MAX_CENTRES = 16, 48, 16


class AiqmcGtoShell(C.Structure):
    _fields_ = [("l", C.c_int32), ("n_prim", C.c_int32), ("centre", C.c_int32), ("ao_offset", C.c_int32),
                ("alpha", C.c_double * MAX_PRIM), ("coef", C.c_double * MAX_PRIM)]


def parse_nwchem_basis(text: str):
    """[(element, l, exponents, coefficients)] from "El shell" headers followed by "exponent coefficient" lines."""
    shells, cur = [], None
    for line in text.splitlines():
        t = line.split()
        if not t or t[0].startswith("#"):
            continue
        if len(t) == 2 and t[1].lower() in ("s", "p", "d", "f"):
            try:
                float(t[0])
            except ValueError:
                cur = (t[0], "spdf".index(t[1].lower()), [], [])
                shells.append(cur)
                continue
        if cur is None:
            raise ValueError(f"basis line before any shell header: {line!r}")
        cur[2].append(float(t[0]))
        cur[3].append(float(t[1]))
    return [(e, l, np.asarray(a), np.asarray(c)) for e, l, a, c in shells]


class GaussianBasis:
    """shells: [(centre index, l, exponents, coefficients)] in AO order; centres (n_centres, 3)."""

    def __init__(self, shells: Sequence[Tuple[int, int, np.ndarray, np.ndarray]], centres, device=None):
        if not torch.cuda.is_available():
            raise _lib.AiqmcError("GaussianBasis needs a CUDA device (there is no CPU fallback)")
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.centres = np.ascontiguousarray(np.asarray(centres, dtype=np.float64).reshape(-1, 3))
        if len(shells) > MAX_SHELLS or self.centres.shape[0] > MAX_CENTRES:
            raise ValueError("basis exceeds the compiled table size")
        self.table = (AiqmcGtoShell * len(shells))()
        off = 0
        for k, (c, l, al, co) in enumerate(shells):
            al, co = np.asarray(al, dtype=np.float64).ravel(), np.asarray(co, dtype=np.float64).ravel()
            if not (0 <= l <= 3) or al.size != co.size or not (1 <= al.size <= MAX_PRIM):
                raise ValueError(f"shell {k}: l must be 0..3 and 1..{MAX_PRIM} primitives")
            s = self.table[k]
            s.l, s.n_prim, s.centre, s.ao_offset = int(l), int(al.size), int(c), off
            for p in range(al.size):
                s.alpha[p], s.coef[p] = float(al[p]), float(co[p])
            off += 2 * l + 1
        self.nao = off

    @classmethod
    def from_nwchem(cls, text: str, atoms, elements: Optional[Sequence[str]] = None, device=None):
        """One copy of every shell of `text` whose element matches, on every atom (file order inside an atom)."""
        parsed = parse_nwchem_basis(text)
        atoms = np.asarray(atoms, dtype=np.float64).reshape(-1, 3)
        shells = []
        for a in range(atoms.shape[0]):
            for el, l, al, co in parsed:
                if elements is None or el.lower() == elements[a].lower():
                    shells.append((a, l, al, co))
        return cls(shells, atoms, device=device)

    def eval(self, points, want_grad: bool = True, want_lap: bool = True):
        """points (..., 3) -> val (..., nAO) [, grad (..., nAO, 3)] [, lap (..., nAO)]  (float64, on the device)."""
        p = torch.as_tensor(points).to(device=self.device, dtype=torch.float64)
        lead = p.shape[:-1]
        p = p.reshape(-1, 3).contiguous()
        n = p.shape[0]
        val = torch.empty((n, self.nao), dtype=torch.float64, device=self.device)
        grad = torch.empty((n, self.nao, 3), dtype=torch.float64, device=self.device) if want_grad else None
        lap = torch.empty((n, self.nao), dtype=torch.float64, device=self.device) if want_lap else None
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
        with torch.cuda.device(self.device):
            rc = self.lib.aiqmc_gto_eval(C.cast(self.table, C.c_void_p), len(self.table),
                                         C.c_void_p(self.centres.ctypes.data), self.centres.shape[0], ptr(p), n,
                                         self.nao, ptr(val), ptr(grad), ptr(lap),
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "aiqmc_gto_eval")
        out = [val.reshape(*lead, self.nao)]
        if want_grad:
            out.append(grad.reshape(*lead, self.nao, 3))
        if want_lap:
            out.append(lap.reshape(*lead, self.nao))
        return tuple(out) if len(out) > 1 else out[0]
