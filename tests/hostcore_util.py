"""Test helper: loads the TEST-ONLY host build of csrc/psi_core.cuh (tests/hostcore)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostcore", "hostcore.cpp")
LIB = os.path.join(HERE, "hostcore", "libhostcore.so")
CSRC = os.path.join(os.path.dirname(HERE), "ab-initio-flexible-gaussian-basis-neural-network-quantum-monte-carlo_b200", "csrc")
CORE = os.path.join(CSRC, "psi_core.cuh")


def build(force=False):
    newest = max(os.path.getmtime(p) for p in [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))])
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        subprocess.check_call(["g++", "-O2", "-fopenmp", "-std=c++17", "-shared", "-fPIC", "-o", LIB, SRC])
    return LIB


def load():
    lib = C.CDLL(build())
    return lib


def dptr(a):
    return a.ctypes.data_as(C.c_void_p)


def host_psi(lib, sys_struct, packed, pos, mode):
    n = sys_struct.n_elec
    pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3 * n)
    ncfg = pos.shape[0]
    phase, logabs = np.zeros(ncfg), np.zeros(ncfg)
    grad, lap = np.zeros((ncfg, 3 * n)), np.zeros(ncfg)
    rc = lib.hc_psi(C.byref(sys_struct), dptr(packed), dptr(pos), C.c_long(ncfg), C.c_int(mode), dptr(phase),
                    dptr(logabs), dptr(grad), dptr(lap))
    assert rc == 0, "no host instantiation for this (N, A)"
    return phase, logabs, grad, lap


def host_param_grad(lib, sys_struct, packed, pos, alpha, beta, cached=False):
    """Per-walker d(alpha log|psi| + beta phase)/d(packed params), shape (ncfg, len(packed))."""
    n = sys_struct.n_elec
    pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3 * n)
    ncfg = pos.shape[0]
    out = np.zeros((ncfg, packed.size))
    rc = lib.hc_param_grad(C.byref(sys_struct), dptr(packed), dptr(pos), C.c_long(ncfg),
                           dptr(np.ascontiguousarray(alpha, dtype=np.float64)),
                           dptr(np.ascontiguousarray(beta, dtype=np.float64)), dptr(out), C.c_int(1 if cached else 0))
    assert rc == 0, "no host instantiation for this (N, A)"
    return out
