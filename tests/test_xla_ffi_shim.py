"""csrc/xla_ffi.cc (the jax.ffi / XLA custom-call boundary) against the genuine XLA FFI header when jax is
installed (jax.ffi.include_dir()), else against the minimal mock under tests/ffi_mock: a syntax + type check that every
handler's parameter list matches its Ffi::Bind() description and every C-ABI call matches include/aiqmc_b200.h.  The
shim can only be EXERCISED where jax exists; this keeps it from rotting where it does not."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "ab-initio-flexible-gaussian-basis-neural-network-quantum-monte-carlo_b200", "csrc", "xla_ffi.cc")


def _ffi_include():
    try:
        import jax.ffi
        return jax.ffi.include_dir(), "genuine"
    except Exception:
        return os.path.join(ROOT, "tests", "ffi_mock"), "mock"


def test_xla_ffi_shim_type_checks():
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    inc, kind = _ffi_include()
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not available")
    r = subprocess.run([gxx, "-std=c++17", "-fsyntax-only", "-I", inc, "-I", cuda_inc, SRC], capture_output=True, text=True)
    assert r.returncode == 0, f"xla_ffi.cc does not compile against the {kind} FFI header:\n{r.stderr[:4000]}"


def test_every_handler_symbol_is_documented():
    src = open(SRC).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    import re
    names = re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((\w+),", src)
    assert len(names) >= 8
    for n in names:
        assert n in doc, f"{n} is not described in INTEGRATION.md"
