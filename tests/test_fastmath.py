"""fastmath.cuh (host build of the same bodies the kernels inline) against libm."""
import ctypes as C

import numpy as np

import hostcore_util as H


def _run(which, x):
    lib = H.load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    lib.hc_fastmath(C.c_int(which), H.dptr(x), C.c_long(x.size), H.dptr(out))
    return out


def test_fexp_relative_error():
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-50, 50, 200000), rng.uniform(-1, 1, 100000), [0.0, -700.0, 700.0, 1e-300]])
    rel = np.abs(_run(0, x) / np.exp(x) - 1.0)
    assert rel.max() < 1e-15, rel.max()
    assert np.isfinite(_run(0, np.array([-1e4, 1e4]))).all()      # clamped, not inf/nan


def test_ftanh_absolute_error_and_sign():
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.uniform(-25, 25, 300000), rng.normal(0, 1, 200000), rng.normal(0, 1e-6, 1000),
                        [0.0, -0.0, 1e-300, 19.0, -19.0, 40.0, -40.0, 1e3, -1e3, 1e7, -1e7]])
    got = _run(1, x)
    err = np.abs(got - np.tanh(x))
    assert err.max() < 2e-13, err.max()
    big = np.abs(x) > 1e-12                                          # tanh = 1 - 2/(1 + exp(2x)): absolute accuracy only,
    assert (np.signbit(got) == np.signbit(x))[big].all()             # so the sign of a result below 1e-13 is not defined
    assert (np.abs(got) <= 1.0).all()


def test_ftanh_batched_variants():
    rng = np.random.default_rng(4)
    x = np.concatenate([rng.uniform(-25, 25, 200000), rng.normal(0, 1, 200000), rng.uniform(-400, 400, 3996),
                        [1e3, -1e3, 1e5, -1e5]])                       # saturation far beyond the exponent clamp
    assert x.size % 4 == 0
    assert np.array_equal(_run(4, x), _run(1, x))                     # ACC=0 batched == scalar, bit for bit
    got = _run(5, x)                                                  # ACC=1: 512-entry table, quadratic, one Newton step
    err = np.abs(got - np.tanh(x))
    assert err.max() < 5e-11, err.max()
    assert (np.abs(got) <= 1.0).all() and np.isfinite(got).all()


def test_frcp_frsqrt():
    rng = np.random.default_rng(3)
    x = np.exp(rng.uniform(-60, 60, 200000))
    assert np.abs(_run(2, x) * x - 1.0).max() < 1e-15
    assert np.abs(_run(2, -x) * -x - 1.0).max() < 1e-15
    assert np.abs(_run(3, x) ** 2 * x - 1.0).max() < 2e-15
