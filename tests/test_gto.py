"""Row A0: contracted Gaussians x real solid harmonics.  CPU: pins of the oracle restatement (oracle/gto_oracle.py).
GPU: the kernel (aiqmc_gto_eval) against the oracle's values and autograd derivatives on the carbon cc-pVDZ basis."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import gto_oracle as G

torch.set_default_dtype(torch.float64)

# AIQMC/C.cc-pVDZ.nwchem:1-27 verbatim (the reference's own basis file; 2s + 2p + 1d = 13 AOs from 21 primitives)
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


def test_parse_reference_basis_file():
    sh = G.parse_nwchem_basis(C_CC_PVDZ)
    assert [(e, l, len(a)) for e, l, a, c in sh] == [("C", 0, 9), ("C", 0, 1), ("C", 1, 9), ("C", 1, 1), ("C", 2, 1)]
    assert sum(len(a) for _, _, a, _ in sh) == 21 and sum(2 * l + 1 for _, l, _, _ in sh) == 13
    assert sh[0][2][0] == 13.073594 and sh[0][3][2] == -0.1978471


def test_real_solid_harmonics_closed_forms_and_scipy():
    rng = np.random.default_rng(0)
    r = torch.tensor(rng.normal(size=(50, 3)))
    x, y, z = r[:, 0], r[:, 1], r[:, 2]
    r2 = (r * r).sum(-1)
    pi = math.pi
    closed = {(0, 0): 0.5 / math.sqrt(pi) * torch.ones_like(x), (1, -1): math.sqrt(3 / (4 * pi)) * y,
              (1, 0): math.sqrt(3 / (4 * pi)) * z, (1, 1): math.sqrt(3 / (4 * pi)) * x,
              (2, -2): 0.5 * math.sqrt(15 / pi) * x * y, (2, -1): 0.5 * math.sqrt(15 / pi) * y * z,
              (2, 0): 0.25 * math.sqrt(5 / pi) * (3 * z * z - r2), (2, 1): 0.5 * math.sqrt(15 / pi) * x * z,
              (2, 2): 0.25 * math.sqrt(15 / pi) * (x * x - y * y),
              (3, -3): 0.25 * math.sqrt(35 / (2 * pi)) * y * (3 * x * x - y * y), (3, -2): 0.5 * math.sqrt(105 / pi) * x * y * z,
              (3, -1): 0.25 * math.sqrt(21 / (2 * pi)) * y * (5 * z * z - r2), (3, 0): 0.25 * math.sqrt(7 / pi) * z * (5 * z * z - 3 * r2),
              (3, 1): 0.25 * math.sqrt(21 / (2 * pi)) * x * (5 * z * z - r2), (3, 2): 0.25 * math.sqrt(105 / pi) * z * (x * x - y * y),
              (3, 3): 0.25 * math.sqrt(35 / (2 * pi)) * x * (x * x - 3 * y * y)}
    for (l, m), want in closed.items():
        np.testing.assert_allclose(G.solid_harmonic(r, l, m).numpy(), want.numpy(), rtol=1e-12, atol=1e-13)
    # against the complex Y_l^m that Gaussian_orbitals.py:13 calls: real m>0 = sqrt2 (-1)^m Re Y, m<0 = sqrt2 (-1)^m Im Y_l^|m|
    rho = torch.linalg.norm(r, dim=-1)
    for l in range(4):
        for m in range(-l, l + 1):
            yc = G.complex_sph_harm(l, abs(m), r.numpy())
            want = yc.real if m == 0 else math.sqrt(2) * (-1) ** m * (yc.real if m > 0 else yc.imag)
            np.testing.assert_allclose(G.solid_harmonic(r, l, m).numpy(), want * rho.numpy() ** l, rtol=1e-10, atol=1e-12)


def test_solid_harmonics_are_harmonic_and_orthonormal():
    rng = np.random.default_rng(1)
    pts = torch.tensor(rng.normal(size=(6, 3)))
    basis = [(0, 3, np.array([1e-30]), np.array([1.0]))]                 # f == 1: AO = S_3m
    _, _, lap = G.eval_gto_with_derivatives(pts, basis, torch.zeros(1, 3))
    assert np.abs(lap.numpy()).max() < 1e-9                                # lap(r^l Y_lm) = 0
    # orthonormality on the sphere with a 200k-point Monte-Carlo-free rule: Gauss-Legendre x uniform phi
    xs, ws = np.polynomial.legendre.leggauss(16)
    phi = (np.arange(32) + 0.5) * 2 * math.pi / 32
    ct, ph = np.meshgrid(xs, phi, indexing="ij")
    st = np.sqrt(1 - ct ** 2)
    r = torch.tensor(np.stack([st * np.cos(ph), st * np.sin(ph), ct], -1).reshape(-1, 3))
    w = np.repeat(ws, 32) * 2 * math.pi / 32
    Y = np.stack([G.solid_harmonic(r, l, m).numpy() for l in range(4) for m in range(-l, l + 1)], 1)
    np.testing.assert_allclose((Y * w[:, None]).T @ Y, np.eye(16), atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("natoms", [1, 2])
def test_kernel_matches_oracle_on_carbon_cc_pvdz(natoms):
    import aiqmc_b200
    rng = np.random.default_rng(5)
    atoms = np.array([[0.0, 0.0, 0.0], [0.3, -0.2, 2.1]])[:natoms]
    basis = aiqmc_b200.GaussianBasis.from_nwchem(C_CC_PVDZ, atoms)
    assert basis.nao == 13 * natoms
    pts = rng.normal(size=(257, 3)) * 1.5
    val, grad, lap = (t.cpu().numpy() for t in basis.eval(pts))
    shells = [(a, l, al, co) for a in range(natoms) for _, l, al, co in G.parse_nwchem_basis(C_CC_PVDZ)]
    v0, g0, l0 = G.eval_gto_with_derivatives(torch.tensor(pts), shells, torch.tensor(atoms))
    np.testing.assert_allclose(val, v0.numpy(), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(grad, g0.numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(lap, l0.numpy(), rtol=1e-9, atol=1e-11)
    only = basis.eval(pts, want_grad=False, want_lap=False).cpu().numpy()
    assert np.array_equal(only, val)
    assert basis.eval(pts[:0])[0].shape == (0, 13 * natoms)                 # empty input


@pytest.mark.gpu
def test_streaming_kernel_many_tiles_per_warp_matches_oracle():
    """More 32-point tiles than resident warps (every persistent warp loops, prefetches and re-uses its shared-memory tile)
    plus a ragged tail; also the rolled-loop instantiation used beyond 2^20 points agrees with the unrolled one."""
    import aiqmc_b200
    rng = np.random.default_rng(7)
    atoms = np.zeros((1, 3))
    basis = aiqmc_b200.GaussianBasis.from_nwchem(C_CC_PVDZ, atoms)
    n = 32 * 4 * 148 * 4 + 17
    pts = rng.normal(size=(n, 3)) * 2.0
    val, grad, lap = (t.cpu().numpy() for t in basis.eval(pts))
    shells = [(0, l, al, co) for _, l, al, co in G.parse_nwchem_basis(C_CC_PVDZ)]
    v0, g0, l0 = G.eval_gto_with_derivatives(torch.tensor(pts), shells, torch.tensor(atoms))
    np.testing.assert_allclose(val, v0.numpy(), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(grad, g0.numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(lap, l0.numpy(), rtol=1e-9, atol=1e-11)
    big = np.tile(pts[:65536], (17, 1))                                    # 1,114,112 points > 2^20
    vb = basis.eval(big, want_grad=False, want_lap=False).cpu().numpy()
    np.testing.assert_allclose(vb[-65536:], val[:65536], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(vb[:65536], val[:65536], rtol=1e-13, atol=1e-15)


@pytest.mark.gpu
def test_kernel_f_shell_and_bad_arguments():
    import aiqmc_b200
    rng = np.random.default_rng(6)
    shells = [(0, 3, np.array([0.7, 0.2]), np.array([0.4, 0.6])), (0, 1, np.array([1.1]), np.array([1.0]))]
    basis = aiqmc_b200.GaussianBasis(shells, np.array([[0.1, 0.2, -0.3]]))
    pts = rng.normal(size=(64, 3))
    val, grad, lap = (t.cpu().numpy() for t in basis.eval(pts))
    v0, g0, l0 = G.eval_gto_with_derivatives(torch.tensor(pts), shells, torch.tensor([[0.1, 0.2, -0.3]]))
    np.testing.assert_allclose(val, v0.numpy(), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(grad, g0.numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(lap, l0.numpy(), rtol=1e-9, atol=1e-11)
    with pytest.raises(ValueError):
        aiqmc_b200.GaussianBasis([(0, 4, np.array([1.0]), np.array([1.0]))], np.zeros((1, 3)))
