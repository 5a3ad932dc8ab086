"""Per-system kernel plugins: the core library is small and system-agnostic, every (n_elec, n_atoms)
system is its own shared object bound on first use, and a system outside the prebuilt set (csrc/dispatch.h) is compiled
on demand -- no source list to edit.  CPU-only: nvcc cross-compiles sm_100a, binding a plugin needs no GPU."""
import os

import pytest

import aiqmc_b200
from aiqmc_b200 import build as B


def test_core_library_is_small_and_every_prebuilt_system_has_a_plugin():
    lib = aiqmc_b200.lib.load()
    assert os.path.getsize(aiqmc_b200.lib.LIB_PATH) < 30 * 2 ** 20
    for n, a in B.systems():
        p = B.plugin_path(n, a)
        assert os.path.exists(p), p
        assert os.path.getsize(p) < 30 * 2 ** 20, (p, os.path.getsize(p))
        assert lib.aiqmc_supported(n, a) == 1
    assert lib.aiqmc_supported(33, 1) == 0 and lib.aiqmc_supported(4, 17) == 0


@pytest.mark.timeout(900)
def test_a_system_outside_the_prebuilt_set_is_built_on_demand():
    lib = aiqmc_b200.lib.load()
    n, a = 3, 2
    assert (n, a) not in B.systems()
    path = B.ensure_system(n, a)
    assert os.path.exists(path)
    lib.aiqmc_rescan_systems()
    assert lib.aiqmc_supported(n, a) == 1
    assert B.ensure_system(n, a) == path                     # up to date: no rebuild
