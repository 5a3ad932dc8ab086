"""The C-ABI library loads on a CPU-only box and exports every symbol include/aiqmc_b200.h declares."""
import ctypes as C
import os
import re

import pytest

import aiqmc_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "aiqmc_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(aiqmc_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = aiqmc_b200.lib.load()
    names = declared_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(aiqmc_b200.lib.EXPORTS) == names


def test_layout_and_support_queries_without_gpu():
    lib = aiqmc_b200.lib.load()
    assert lib.aiqmc_supported(4, 1) == 1 and lib.aiqmc_supported(30, 12) == 1 and lib.aiqmc_supported(7, 3) == 0
    lay = aiqmc_b200.lib.param_layout(4, 1)
    assert lay.conv_w[0] == 0 and lay.total > 0
    assert lib.aiqmc_param_layout(0, 1, C.byref(lay)) == -2     # AIQMC_E_BADARG
    assert lib.aiqmc_psi_fwd(None, None, None, 0, None, None, None) == -2


def test_struct_sizes_match_header():
    # sizes implied by the header's constants (int32 fields first, then doubles)
    assert C.sizeof(aiqmc_b200.system.AiqmcSystem) == 4 * (5 + 32)
    assert C.sizeof(aiqmc_b200.system.AiqmcEcp) == 16 + 8 * (3 * 16 * 4 + 3 * 16 * 4 * 4 + 50 * 3 + 50)
    assert C.sizeof(aiqmc_b200.system.AiqmcLayout) == 4 * 39


def test_product_path_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from common import CASES, Case
    case = Case(**CASES["C_ecp"])
    with pytest.raises(aiqmc_b200.lib.AiqmcError):
        aiqmc_b200.WalkerEngine(case.spec(), case.params)
