"""aiqmc_b200 -- B200-native (sm_100a) walker engine for the VMC/DMC inner loop of AIQMCrelease3.

Importable as `aiqmc_b200` (see the alias package at the repository root)."""
from . import api, build, checkpoint, engine, gto, lib, parallel, system, workloads  # noqa: F401
from .gto import GaussianBasis  # noqa: F401
from .api import (AINetData, Network, PackedParams, branch, comput_S, compute_tmoves, dmc_propagate, estimate_energy, reconfigure, trial_energy, local_energy, main_monte_carlo,  # noqa: F401
                  make_ai_net, propose_drift_diffusion, random_rotations, total_energy, branch_global, make_loss,
                  clip_local_values, AuxiliaryLossData, make_mcmc_step, update_mcmc_width,
                  correlated_samples, weights_jacobian, read_ecp_nwchem)
from .engine import HostStepPipeline, NcclComm, WalkerEngine  # noqa: F401
from .system import SystemSpec, jastrow_indices_ee, make_ecp, pack_params, spin_indices_h, unpack_param_grad  # noqa: F401

__all__ = ["AINetData", "Network", "PackedParams", "WalkerEngine", "SystemSpec", "make_ai_net", "main_monte_carlo",
           "local_energy", "propose_drift_diffusion", "comput_S", "compute_tmoves", "dmc_propagate", "estimate_energy", "reconfigure", "trial_energy", "branch", "make_ecp", "pack_params",
           "jastrow_indices_ee", "spin_indices_h", "random_rotations", "total_energy", "branch_global", "parallel", "GaussianBasis", "gto", "make_loss", "clip_local_values",
           "AuxiliaryLossData", "unpack_param_grad", "make_mcmc_step", "update_mcmc_width", "checkpoint", "correlated_samples",
           "weights_jacobian", "read_ecp_nwchem", "HostStepPipeline"]
