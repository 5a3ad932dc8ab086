"""CPU restatement of row A0: contracted Gaussian primitives times real solid harmonics.

TEST INFRASTRUCTURE ONLY (imported by tests/, __graft_entry__.smoke() and bench.py's cpu leg; never by the product).

Follows /root/reference:
  AIQMC/Gaussian_orbitals.py:11-13   primitive = c * r^l * exp(-alpha r^2) * Y_lm   (complex scipy sph_harm there)
  AIQMC/C.cc-pVDZ.nwchem:1-27        the basis text format ("El shell" header lines, then "exponent coefficient")
  ferminet/utils/gto.py:100-135      cart2sph + solid_harmonic: REAL solid harmonics r^l Y_lm built from the
                                     orthonormalised associated Legendre functions (jss.lpmn_values(..., True)),
                                     m < 0 -> sqrt2 (-1)^m P_l^|m| sin(|m| phi), m > 0 -> sqrt2 (-1)^m P_l^m cos(m phi)
  ferminet/utils/gto.py:338-389      eval_gto: radial = sum_p w_p exp(-alpha_p r^2) per contracted shell, AO =
                                     radial * angular, AO order = shells in file order, m = -l..l inside a shell
Derivatives (gradient, Laplacian) by torch.autograd on this restatement.

Pinned by tests/test_gto.py: closed-form real harmonics to l = 3, orthonormality on the 50-point rule's exactness
range, |Y_lm| against scipy.special.sph_harm (the function Gaussian_orbitals.py:13 calls), Laplacian of r^l Y_lm == 0.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.special
import torch


def parse_nwchem_basis(text: str):
    """[(element, l, exponents (P,), coefficients (P,))] from the text of C.cc-pVDZ.nwchem."""
    shells, cur = [], None
    for line in text.splitlines():
        t = line.split()
        if not t:
            continue
        if len(t) == 2 and t[1].lower() in "spdfgh" and not _is_number(t[0]):
            cur = (t[0], "spdfgh".index(t[1].lower()), [], [])
            shells.append(cur)
        elif cur is not None and len(t) >= 2:
            cur[2].append(float(t[0]))
            cur[3].append(float(t[1]))
    return [(e, l, np.array(a), np.array(c)) for e, l, a, c in shells]


def _is_number(s):
    try:
        float(s)
        return True
    except ValueError:
        return False


def _norm_legendre(l: int, m: int, x: torch.Tensor) -> torch.Tensor:
    """Orthonormalised associated Legendre function (Condon-Shortley phase included), what
    jax.scipy.special.lpmn_values(l_max, l_max, x, is_normalized=True)[m, l] returns (gto.py:121)."""
    # P_m^m
    pmm = torch.ones_like(x)
    if m > 0:
        somx2 = torch.sqrt((1 - x) * (1 + x))
        fact = 1.0
        for _ in range(m):
            pmm = -pmm * fact * somx2
            fact += 2.0
    if l == m:
        p = pmm
    else:
        pmmp1 = x * (2 * m + 1) * pmm
        if l == m + 1:
            p = pmmp1
        else:
            p = pmmp1
            for ll in range(m + 2, l + 1):
                p = (x * (2 * ll - 1) * pmmp1 - (ll + m - 1) * pmm) / (ll - m)
                pmm, pmmp1 = pmmp1, p
    norm = math.sqrt((2 * l + 1) / (4 * math.pi) * math.factorial(l - m) / math.factorial(l + m))
    return norm * p


def solid_harmonic(r: torch.Tensor, l: int, m: int) -> torch.Tensor:
    """r^l Y_lm (real), gto.py:117-135 for a single (l, m); r (...,3)."""
    rho = torch.linalg.norm(r, dim=-1)
    phi = torch.atan2(r[..., 1], r[..., 0])
    cos_theta = r[..., 2] / rho
    am = abs(m)
    leg = _norm_legendre(l, am, cos_theta)
    if m == 0:
        h = leg
    elif m > 0:
        h = math.sqrt(2.0) * (-1) ** am * leg * torch.cos(am * phi)
    else:
        h = math.sqrt(2.0) * (-1) ** am * leg * torch.sin(am * phi)
    return h * rho ** l


def eval_gto(points: torch.Tensor, shells, centres: torch.Tensor) -> torch.Tensor:
    """[G, nAO] contracted GTO values; shells = [(centre index, l, alphas, coefs)] in AO order (gto.py:338-389)."""
    out = []
    for c, l, al, co in shells:
        d = points - centres[c]
        r2 = (d * d).sum(-1)
        radial = (torch.as_tensor(co, dtype=points.dtype) * torch.exp(-torch.as_tensor(al, dtype=points.dtype) * r2[..., None])).sum(-1)
        for m in range(-l, l + 1):
            out.append(radial * solid_harmonic(d, l, m))
    return torch.stack(out, dim=-1)


def eval_gto_with_derivatives(points: torch.Tensor, shells, centres: torch.Tensor):
    """(val [G,nAO], grad [G,nAO,3], lap [G,nAO]) by autograd (double backward for the Laplacian)."""
    x = points.detach().clone().requires_grad_(True)
    val = eval_gto(x, shells, centres)
    nao = val.shape[-1]
    grads, laps = [], []
    for k in range(nao):
        g, = torch.autograd.grad(val[:, k].sum(), x, create_graph=True)
        lap = 0.0
        for c in range(3):
            h, = torch.autograd.grad(g[:, c].sum(), x, retain_graph=True)
            lap = lap + h[:, c]
        grads.append(g.detach())
        laps.append(lap.detach())
    return val.detach(), torch.stack(grads, dim=1), torch.stack(laps, dim=1)


def complex_sph_harm(l: int, m: int, r: np.ndarray) -> np.ndarray:
    """Y_l^m as AIQMC/Gaussian_orbitals.py:13 evaluates it (scipy convention: azimuth first)."""
    rho = np.linalg.norm(r, axis=-1)
    az, pol = np.arctan2(r[..., 1], r[..., 0]), np.arccos(r[..., 2] / rho)
    if hasattr(scipy.special, "sph_harm"):                      # scipy < 1.17: sph_harm(m, l, azimuth, polar)
        return scipy.special.sph_harm(m, l, az, pol)
    return scipy.special.sph_harm_y(l, m, pol, az)              # its replacement: sph_harm_y(l, m, polar, azimuth)
