"""ctypes binding of libaiqmc_b200.so (the C ABI of include/aiqmc_b200.h).

There is deliberately NO fallback: if the CUDA library is missing or cannot be loaded every
product entry point raises.  The CPU oracle under oracle/ is test infrastructure only.
"""
from __future__ import annotations

import ctypes as C
import os

from .system import AiqmcEcp, AiqmcLayout, AiqmcSystem

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AIQMC_LIB", os.path.join(HERE, "libaiqmc_b200.so"))

ERRORS = {-1: "AIQMC_E_UNSUPPORTED: the kernel plugin of this (n_elec, n_atoms) is not built; aiqmc_b200.build.ensure_system(n, a)",
          -2: "AIQMC_E_BADARG", -3: "AIQMC_E_CUDA", -4: "AIQMC_E_WORKSPACE",
          -5: "AIQMC_E_NCCL: libnccl.so.2 could not be bound (set AIQMC_NCCL_LIB) or an NCCL call failed"}

EXPORTS = ["aiqmc_param_layout", "aiqmc_supported", "aiqmc_rescan_systems", "aiqmc_last_cuda_error", "aiqmc_launch_count", "aiqmc_version", "aiqmc_psi_fwd",
           "aiqmc_psi_workspace_bytes", "aiqmc_psi_grad", "aiqmc_psi_fwdlap", "aiqmc_vmc_workspace_bytes", "aiqmc_vmc_sweep",
           "aiqmc_energy_workspace_bytes", "aiqmc_local_energy_ae", "aiqmc_local_energy_ecp", "aiqmc_local_energy_ecp_stages", "aiqmc_energy_stats",
           "aiqmc_dmc_tmove_workspace_bytes", "aiqmc_dmc_tmove", "aiqmc_dmc_ecut_min", "aiqmc_dmc_s", "aiqmc_dmc_weights", "aiqmc_branch_workspace_bytes",
           "aiqmc_branch_comb", "aiqmc_gather_walkers", "aiqmc_bench_dfma", "aiqmc_gto_eval", "aiqmc_param_grad_workspace_bytes",
           "aiqmc_psi_param_grad", "aiqmc_mh_workspace_bytes", "aiqmc_mh_step", "aiqmc_correlated_samples",
           "aiqmc_weights_jacobian", "aiqmc_vmc_sweep_compact", "aiqmc_rng_sweep", "aiqmc_rng_rotations", "aiqmc_rng_uniform",
           "aiqmc_nccl_available", "aiqmc_nccl_unique_id", "aiqmc_nccl_comm_init", "aiqmc_nccl_comm_destroy", "aiqmc_energy_allreduce",
           "aiqmc_ecut_allreduce_min", "aiqmc_rebalance_workspace_bytes", "aiqmc_rebalance_nccl",
           "aiqmc_energy_stats_workspace_bytes", "aiqmc_energy_stats_ws"]

_lib = None


class AiqmcError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Loads the CUDA library built by build.py; raises if it is absent (no CPU fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AiqmcError(f"{LIB_PATH} is missing: run `python __graft_entry__.py` (or aiqmc_b200.build.build()) "
                         "to compile the sm_100a kernels; there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    sysp, ecpp = C.POINTER(AiqmcSystem), C.POINTER(AiqmcEcp)
    sig = {
        "aiqmc_param_layout": (C.c_int, [i32, i32, C.POINTER(AiqmcLayout)]),
        "aiqmc_supported": (C.c_int, [i32, i32]),
        "aiqmc_rescan_systems": (None, []),
        "aiqmc_last_cuda_error": (C.c_int, []),
        "aiqmc_launch_count": (i64, []),
        "aiqmc_version": (C.c_char_p, []),
        "aiqmc_psi_fwd": (C.c_int, [sysp, vp, vp, i64, vp, vp, vp]),
        "aiqmc_psi_workspace_bytes": (i64, [sysp, i64, i32]),
        "aiqmc_psi_grad": (C.c_int, [sysp, vp, vp, i64, vp, vp, vp, vp, i64, vp]),
        "aiqmc_psi_fwdlap": (C.c_int, [sysp, vp, vp, i64, vp, vp, vp, vp, vp, i64, vp]),
        "aiqmc_vmc_workspace_bytes": (i64, [sysp, i64]),
        "aiqmc_vmc_sweep": (C.c_int, [sysp, vp, vp, vp, vp, vp, i64, f64, f64, i32, vp, vp, vp, vp, i64, vp]),
        "aiqmc_vmc_sweep_compact": (C.c_int, [sysp, vp, vp, vp, vp, vp, i64, f64, f64, i32, vp, vp, vp, vp, i64, vp]),
        "aiqmc_rng_sweep": (C.c_int, [C.c_uint64, C.c_uint32, i64, i64, i32, f64, vp, vp, vp, vp]),
        "aiqmc_rng_rotations": (C.c_int, [C.c_uint64, C.c_uint32, i64, i64, vp, vp]),
        "aiqmc_rng_uniform": (C.c_int, [C.c_uint64, C.c_uint32, i64, i64, i32, C.c_uint32, vp, vp]),
        "aiqmc_nccl_available": (C.c_int, []),
        "aiqmc_nccl_unique_id": (C.c_int, [vp]),
        "aiqmc_nccl_comm_init": (C.c_int, [i32, i32, vp, C.POINTER(vp)]),
        "aiqmc_nccl_comm_destroy": (C.c_int, [vp]),
        "aiqmc_energy_allreduce": (C.c_int, [vp, vp, vp]),
        "aiqmc_ecut_allreduce_min": (C.c_int, [vp, vp, vp]),
        "aiqmc_rebalance_workspace_bytes": (i64, [i64, i32, i32]),
        "aiqmc_rebalance_nccl": (C.c_int, [vp, vp, i64, i32, f64, i32, i32, vp, i32, vp, vp, vp, C.POINTER(i64), vp, i64, vp]),
        "aiqmc_energy_workspace_bytes": (i64, [sysp, i64, i32]),
        "aiqmc_local_energy_ae": (C.c_int, [sysp, vp, vp, i64, vp, vp, i64, vp]),
        "aiqmc_local_energy_ecp": (C.c_int, [sysp, ecpp, vp, vp, vp, i64, vp, vp, i64, vp]),
        "aiqmc_local_energy_ecp_stages": (C.c_int, [sysp, ecpp, vp, vp, vp, i64, vp, vp, i64, i32, vp]),
        "aiqmc_dmc_tmove_workspace_bytes": (i64, [sysp, i64]),
        "aiqmc_dmc_tmove": (C.c_int, [sysp, ecpp, vp, vp, vp, vp, vp, i64, f64, vp, vp, vp, vp, i64, vp]),
        "aiqmc_param_grad_workspace_bytes": (i64, [sysp, i64]),
        "aiqmc_psi_param_grad": (C.c_int, [sysp, vp, vp, i64, vp, vp, vp, vp, vp, vp, i64, vp]),
        "aiqmc_mh_workspace_bytes": (i64, [sysp, i64]),
        "aiqmc_mh_step": (C.c_int, [sysp, vp, vp, vp, vp, vp, i64, f64, vp, vp, vp, i64, vp]),
        "aiqmc_correlated_samples": (C.c_int, [vp, vp, i32, vp, i64, i32, vp, vp]),
        "aiqmc_weights_jacobian": (C.c_int, [vp, vp, i32, vp, i64, i32, vp, vp]),
        "aiqmc_gto_eval": (C.c_int, [vp, i32, vp, i32, vp, i64, i32, vp, vp, vp, vp]),
        "aiqmc_bench_dfma": (C.c_int, [i64, vp, C.POINTER(C.c_double), vp]),
        "aiqmc_energy_stats": (C.c_int, [vp, i32, i64, vp, vp]),
        "aiqmc_energy_stats_workspace_bytes": (i64, [i64]),
        "aiqmc_energy_stats_ws": (C.c_int, [vp, i32, i64, vp, vp, i64, vp]),
        "aiqmc_dmc_ecut_min": (C.c_int, [vp, i32, i64, f64, vp, vp, vp]),
        "aiqmc_dmc_s": (C.c_int, [vp, i32, vp, i64, i32, f64, f64, vp, f64, vp, vp]),
        "aiqmc_dmc_weights": (C.c_int, [vp, vp, vp, i64, f64, f64, vp]),
        "aiqmc_branch_workspace_bytes": (i64, [i64]),
        "aiqmc_branch_comb": (C.c_int, [vp, i64, f64, vp, vp, vp, i64, vp]),
        "aiqmc_gather_walkers": (C.c_int, [vp, vp, i64, i32, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError here == header/library mismatch
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        extra = f" (cudaError {load().aiqmc_last_cuda_error()})" if rc == -3 else ""
        raise AiqmcError(f"{what} failed: {ERRORS.get(rc, rc)}{extra}")


def param_layout(n_elec: int, n_atoms: int) -> AiqmcLayout:
    lay = AiqmcLayout()
    check(load().aiqmc_param_layout(n_elec, n_atoms, C.byref(lay)), "aiqmc_param_layout")
    return lay
