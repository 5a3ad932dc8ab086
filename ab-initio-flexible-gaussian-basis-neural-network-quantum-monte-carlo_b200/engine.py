"""Device-resident walker engine: thin, explicit wrapper over the C ABI.

torch is used only for device memory, streams and (in throughput mode) the device RNG; every
numerical operation of the hot path runs in the hand-written kernels of csrc/.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Mapping, Optional

import numpy as np
import torch

from . import lib as _lib
from .system import AiqmcEcp, SystemSpec, pack_params


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_UNSET = object()


def _nbytes(rc: int, what: str) -> int:
    """A *_workspace_bytes query returns the size, or a negative AIQMC_E_* code."""
    if rc < 0:
        _lib.check(int(rc), what)
    return int(rc)


class WalkerEngine:
    """Owns the packed parameters + workspaces of one system on one GPU."""

    def __init__(self, spec: SystemSpec, params: Optional[Mapping[str, Any]] = None, ecp: Optional[AiqmcEcp] = None,
                 device: Optional[torch.device] = None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.AiqmcError("aiqmc_b200 needs a CUDA device; there is no CPU fallback")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.spec = spec
        self.sys = spec.c_struct()
        self.n, self.a = spec.nelectrons, spec.natoms
        if not self.lib.aiqmc_supported(self.n, self.a):
            # not in the prebuilt set (csrc/dispatch.h): compile this system's kernel plugin now, if nvcc is here
            from . import build as _build
            import os
            try:
                if os.environ.get("AIQMC_NO_AUTOBUILD"):
                    raise RuntimeError("AIQMC_NO_AUTOBUILD is set")
                _build.ensure_system(self.n, self.a)
            except (OSError, RuntimeError, ValueError) as exc:
                raise _lib.AiqmcError(f"system N={self.n}, A={self.a}: its kernel plugin is not built and could not be "
                                      f"compiled here ({exc})") from exc
            self.lib.aiqmc_rescan_systems()
            if not self.lib.aiqmc_supported(self.n, self.a):
                _lib.check(-1, f"system N={self.n}, A={self.a}")
        self.layout = _lib.param_layout(self.n, self.a)
        self.ecp = ecp
        self.params_dev: Optional[torch.Tensor] = None
        self._ws = {}
        if params is not None:
            self.set_params(params)

    # ---- parameters -------------------------------------------------------------------
    def set_params(self, params) -> None:
        packed = params if isinstance(params, np.ndarray) else pack_params(self.layout, params, self.spec)
        self.params_dev = torch.from_numpy(np.ascontiguousarray(packed)).to(self.device)

    def _workspace(self, key: str, nbytes: int) -> torch.Tensor:
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws

    def _pos(self, pos) -> torch.Tensor:
        t = torch.as_tensor(pos)
        return t.to(device=self.device, dtype=torch.float64).contiguous()

    def _arg(self, t, shape, name: str, dtype=torch.float64) -> torch.Tensor:
        """Every tensor whose pointer crosses the ABI: on this engine's device, `dtype`, contiguous, exact shape
        (converted when it is not; a wrong element count raises instead of reading out of bounds in the kernel)."""
        t = torch.as_tensor(t)
        if t.numel() != int(np.prod(shape)):
            raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
        if t.device != self.device or t.dtype != dtype or not t.is_contiguous() or tuple(t.shape) != tuple(shape):
            t = t.to(device=self.device, dtype=dtype).reshape(shape).contiguous()
        return t

    # ---- wavefunction -----------------------------------------------------------------
    def psi(self, pos, mode: int = 0):
        """mode 0: (phase, logabs); 1: + grad; 2: + grad, lap.  pos (..., 3N)."""
        p = self._pos(pos)
        lead = p.shape[:-1]
        p2 = p.reshape(-1, 3 * self.n)
        ncfg = p2.shape[0]
        phase = torch.empty(ncfg, dtype=torch.float64, device=self.device)
        logabs = torch.empty_like(phase)
        grad = torch.empty((ncfg, 3 * self.n), dtype=torch.float64, device=self.device) if mode >= 1 else None
        lap = torch.empty_like(phase) if mode == 2 else None
        ws = None
        if mode >= 1:
            ws = self._workspace("psi", _nbytes(self.lib.aiqmc_psi_workspace_bytes(C.byref(self.sys), ncfg, 1 if mode == 2 else 0),
                                                "aiqmc_psi_workspace_bytes"))
        with torch.cuda.device(self.device):
            if mode == 0:
                rc = self.lib.aiqmc_psi_fwd(C.byref(self.sys), _ptr(self.params_dev), _ptr(p2), ncfg, _ptr(phase),
                                            _ptr(logabs), _stream())
            elif mode == 1:
                rc = self.lib.aiqmc_psi_grad(C.byref(self.sys), _ptr(self.params_dev), _ptr(p2), ncfg, _ptr(phase),
                                             _ptr(logabs), _ptr(grad), _ptr(ws), ws.numel(), _stream())
            else:
                rc = self.lib.aiqmc_psi_fwdlap(C.byref(self.sys), _ptr(self.params_dev), _ptr(p2), ncfg, _ptr(phase),
                                               _ptr(logabs), _ptr(grad), _ptr(lap), _ptr(ws), ws.numel(), _stream())
        _lib.check(rc, "aiqmc_psi")
        out = [phase.reshape(lead), logabs.reshape(lead)]
        if mode >= 1:
            out.append(grad.reshape(*lead, 3 * self.n))
        if mode == 2:
            out.append(lap.reshape(lead))
        return tuple(out)

    # ---- VMC sweep --------------------------------------------------------------------
    def vmc_sweep(self, pos: torch.Tensor, gauss1: torch.Tensor, gauss2: torch.Tensor, rnd: torch.Tensor,
                  tstep: float, acyrus: float = 0.25, signed_ratio: bool = False, want_accept: bool = True,
                  want_drift: bool = False, want_aux: bool = False):
        """In-place sweep on pos (B,3N) float64 cuda.  Returns dict(accept, grad_eff_old, aux)."""
        B = pos.shape[0]
        if not (pos.is_cuda and pos.device == self.device and pos.dtype == torch.float64 and pos.is_contiguous()
                and tuple(pos.shape) == (B, 3 * self.n)):
            raise ValueError("vmc_sweep updates pos in place: it must be a contiguous float64 (B,3N) tensor on the engine's device")
        gauss1 = self._arg(gauss1, (B, 3 * self.n), "gauss1")
        compact = torch.as_tensor(gauss2).numel() == B * self.n * 3 and self.n > 1     # (B,N,3): diagonal blocks only
        gauss2 = self._arg(gauss2, (B, self.n, 3) if compact else (B, self.n, 3 * self.n), "gauss2")
        rnd = self._arg(rnd, (B, self.n), "rnd")
        nbytes = _nbytes(self.lib.aiqmc_vmc_workspace_bytes(C.byref(self.sys), B), "aiqmc_vmc_workspace_bytes")
        ws = self._workspace("vmc", nbytes)
        accept = torch.empty((B, self.n), dtype=torch.uint8, device=self.device) if want_accept else None
        drift = torch.empty((B, 3 * self.n), dtype=torch.float64, device=self.device) if want_drift else None
        aux = torch.empty(4, dtype=torch.float64, device=self.device) if want_aux else None
        with torch.cuda.device(self.device):
            fn = self.lib.aiqmc_vmc_sweep_compact if compact else self.lib.aiqmc_vmc_sweep
            rc = fn(C.byref(self.sys), _ptr(self.params_dev), _ptr(pos), _ptr(gauss1),
                                          _ptr(gauss2), _ptr(rnd), B, float(tstep), float(acyrus),
                                          1 if signed_ratio else 0, _ptr(accept), _ptr(drift), _ptr(aux), _ptr(ws),
                                          ws.numel(), _stream())
        _lib.check(rc, "aiqmc_vmc_sweep")
        return dict(accept=accept, grad_eff_old=drift, aux=aux)

    # ---- device-side random inputs (throughput mode; csrc/rng.cu) ------------------------
    def rng_sweep(self, seed: int, step: int, walker0: int, B: int, tstep: float, out=None):
        """(gauss1 (B,3N), gauss2c (B,N,3), rnd (B,N)) from Philox keyed by (seed, global walker id, step)."""
        if out is None:
            out = (torch.empty((B, 3 * self.n), dtype=torch.float64, device=self.device),
                   torch.empty((B, self.n, 3), dtype=torch.float64, device=self.device),
                   torch.empty((B, self.n), dtype=torch.float64, device=self.device))
        g1, g2c, u = out
        with torch.cuda.device(self.device):
            _lib.check(self.lib.aiqmc_rng_sweep(int(seed), int(step), int(walker0), B, self.n, float(tstep), _ptr(g1),
                                                _ptr(g2c), _ptr(u), _stream()), "aiqmc_rng_sweep")
        return out

    def rng_rotations(self, seed: int, step: int, walker0: int, B: int, out=None) -> torch.Tensor:
        rot = out if out is not None else torch.empty((B, 3, 3), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.aiqmc_rng_rotations(int(seed), int(step), int(walker0), B, _ptr(rot), _stream()),
                       "aiqmc_rng_rotations")
        return rot

    def rng_uniform(self, seed: int, step: int, walker0: int, B: int, cols: int, tag: int) -> torch.Tensor:
        out = torch.empty((B, cols), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.aiqmc_rng_uniform(int(seed), int(step), int(walker0), B, int(cols), int(tag), _ptr(out),
                                                  _stream()), "aiqmc_rng_uniform")
        return out

    # ---- local energy -----------------------------------------------------------------
    def local_energy(self, pos: torch.Tensor, rot: Optional[torch.Tensor] = None, stages: int = 7,
                     out: Optional[torch.Tensor] = None, ecp=_UNSET) -> torch.Tensor:
        """All-electron (no ECP table): real (B,).  ccECP: complex128 (B,) (quirk Q25).  `ecp` overrides the engine's
        table for this call only (None = all-electron); closures pass their own table instead of mutating the engine."""
        p = self._pos(pos).reshape(-1, 3 * self.n)
        B = p.shape[0]
        ecp_tab = self.ecp if ecp is _UNSET else ecp
        with_ecp = ecp_tab is not None
        nbytes = _nbytes(self.lib.aiqmc_energy_workspace_bytes(C.byref(self.sys), B, 1 if with_ecp else 0),
                         "aiqmc_energy_workspace_bytes")
        ws = self._workspace("energy", nbytes)
        with torch.cuda.device(self.device):
            if with_ecp:
                if rot is None:
                    raise ValueError("ccECP local energy needs the per-walker rotation matrices (B,3,3)")
                r = self._arg(rot, (B, 9), "rot")
                e = out if out is not None else torch.empty((B, 2), dtype=torch.float64, device=self.device)
                if out is not None:
                    e = self._arg(out, (B, 2), "out")
                    if e.data_ptr() != out.data_ptr():
                        raise ValueError("out must be a contiguous float64 (B,2) tensor on the engine's device")
                rc = self.lib.aiqmc_local_energy_ecp_stages(C.byref(self.sys), C.byref(ecp_tab),
                                                            _ptr(self.params_dev), _ptr(p), _ptr(r), B, _ptr(e),
                                                            _ptr(ws), ws.numel(), int(stages), _stream())
                _lib.check(rc, "aiqmc_local_energy_ecp")
                return torch.view_as_complex(e)
            e = torch.empty(B, dtype=torch.float64, device=self.device)
            rc = self.lib.aiqmc_local_energy_ae(C.byref(self.sys), _ptr(self.params_dev), _ptr(p), B, _ptr(e),
                                                _ptr(ws), ws.numel(), _stream())
            _lib.check(rc, "aiqmc_local_energy_ae")
            return e

    def energy_stats(self, e_l: torch.Tensor) -> torch.Tensor:
        """[sum Re E, sum Im E, sum |E|^2, count] on device (the partials of pploss.py:165-167)."""
        if e_l.is_complex():
            raw, stride = torch.view_as_real(e_l.contiguous()), 2
        else:
            raw, stride = e_l.contiguous(), 1
        out = torch.empty(4, dtype=torch.float64, device=self.device)
        nbytes = _nbytes(self.lib.aiqmc_energy_stats_workspace_bytes(e_l.shape[0]), "aiqmc_energy_stats_workspace_bytes")
        with torch.cuda.device(self.device):
            if nbytes > 0:                 # beyond 2^18 walkers: chunk partials in a workspace, fixed-order sum
                ws = self._workspace("stats", nbytes)
                _lib.check(self.lib.aiqmc_energy_stats_ws(_ptr(raw), stride, e_l.shape[0], _ptr(out), _ptr(ws), ws.numel(),
                                                          _stream()), "aiqmc_energy_stats_ws")
            else:
                _lib.check(self.lib.aiqmc_energy_stats(_ptr(raw), stride, e_l.shape[0], _ptr(out), _stream()),
                           "aiqmc_energy_stats")
        return out

    # ---- DMC pieces -------------------------------------------------------------------
    def dmc_ecut_min(self, e_l: torch.Tensor, e_est: float, branchcut: torch.Tensor) -> torch.Tensor:
        raw, stride = (torch.view_as_real(e_l.contiguous()), 2) if e_l.is_complex() else (e_l.contiguous(), 1)
        out = torch.empty(1, dtype=torch.float64, device=self.device)
        branchcut = self._arg(branchcut, (e_l.shape[0],), "branchcut")
        with torch.cuda.device(self.device):
            _lib.check(self.lib.aiqmc_dmc_ecut_min(_ptr(raw), stride, e_l.shape[0], float(e_est), _ptr(branchcut),
                                                   _ptr(out), _stream()), "aiqmc_dmc_ecut_min")
        return out

    def dmc_s(self, e_l: torch.Tensor, drift: torch.Tensor, e_trial: float, e_est: float, ecut_min: torch.Tensor,
              tau: float) -> torch.Tensor:
        raw, stride = (torch.view_as_real(e_l.contiguous()), 2) if e_l.is_complex() else (e_l.contiguous(), 1)
        B = e_l.shape[0]
        s = torch.empty(B, dtype=torch.float64, device=self.device)
        drift = self._arg(drift, (B, 3 * self.n), "drift")
        ecut_min = self._arg(ecut_min, (1,), "ecut_min")
        with torch.cuda.device(self.device):
            _lib.check(self.lib.aiqmc_dmc_s(_ptr(raw), stride, _ptr(drift), B, self.n, float(e_trial),
                                            float(e_est), _ptr(ecut_min), float(tau), _ptr(s), _stream()),
                       "aiqmc_dmc_s")
        return s

    def dmc_weights(self, weights: torch.Tensor, s_old: torch.Tensor, s_new: torch.Tensor, tau: float,
                    tdamp: float) -> None:
        B = weights.shape[0]
        if not (weights.device == self.device and weights.dtype == torch.float64 and weights.is_contiguous()):
            raise ValueError("dmc_weights updates weights in place: contiguous float64 on the engine's device")
        s_old, s_new = self._arg(s_old, (B,), "s_old"), self._arg(s_new, (B,), "s_new")
        with torch.cuda.device(self.device):
            _lib.check(self.lib.aiqmc_dmc_weights(_ptr(weights), _ptr(s_old), _ptr(s_new), weights.shape[0],
                                                  float(tau), float(tdamp), _stream()), "aiqmc_dmc_weights")

    def branch_comb(self, weights: torch.Tensor, u: float):
        B = weights.shape[0]
        weights = self._arg(weights, (B,), "weights")
        ws = self._workspace("branch", _nbytes(self.lib.aiqmc_branch_workspace_bytes(B), "aiqmc_branch_workspace_bytes"))
        inds = torch.empty(B, dtype=torch.int32, device=self.device)
        neww = torch.empty(1, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.aiqmc_branch_comb(_ptr(weights.contiguous()), B, float(u), _ptr(inds), _ptr(neww),
                                                  _ptr(ws), ws.numel(), _stream()), "aiqmc_branch_comb")
        return neww, inds

    def rebalance(self, weights: torch.Tensor, pos: torch.Tensor, u: float, comm: Optional["NcclComm"] = None,
                  mode: str = "balanced"):
        """Cross-GPU systematic comb + migration through the C ABI (aiqmc_rebalance_nccl): only the block totals of the
        blocked weight scan and the walkers that change rank are exchanged.  mode "ordered": slot k of rank r gets the
        walker of tooth r*B + k (identical to the single-GPU comb, moves almost every walker); "balanced": the same
        multiset of walkers, every rank keeps its own and only the population imbalance moves.  Returns (new weight
        (device scalar), new positions (B,row), source rank of every new walker (B,) int32, bytes this rank sent)."""
        if mode not in ("ordered", "balanced"):
            raise ValueError("mode must be 'ordered' or 'balanced'")
        B, row = pos.shape[0], pos.shape[1]
        world, rank = (comm.world, comm.rank) if comm is not None else (1, 0)
        weights = self._arg(weights, (B,), "weights")
        pos = self._arg(pos, (B, row), "pos")
        ws = self._workspace("rebalance", _nbytes(self.lib.aiqmc_rebalance_workspace_bytes(B, row, world),
                                                  "aiqmc_rebalance_workspace_bytes"))
        out = torch.empty_like(pos)
        neww = torch.empty(1, dtype=torch.float64, device=self.device)
        src = torch.empty(B, dtype=torch.int32, device=self.device)
        moved = C.c_int64(0)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.aiqmc_rebalance_nccl(_ptr(weights), _ptr(pos), B, row, float(u), world, rank,
                                                     comm.handle if comm is not None else None,
                                                     1 if mode == "balanced" else 0, _ptr(out), _ptr(neww),
                                                     _ptr(src), C.byref(moved), _ptr(ws), ws.numel(), _stream()),
                       "aiqmc_rebalance_nccl")
        return neww, out, src, int(moved.value)

    def gather_walkers(self, pos: torch.Tensor, inds: torch.Tensor) -> torch.Tensor:
        """out[k] = pos[inds[k]]; pos may hold more rows than inds selects (cross-GPU population control)."""
        out = torch.empty((inds.shape[0], pos.shape[1]), dtype=pos.dtype, device=pos.device)
        inds = inds.to(torch.int32).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.aiqmc_gather_walkers(_ptr(pos.contiguous()), _ptr(inds), inds.shape[0], pos.shape[1],
                                                     _ptr(out), _stream()), "aiqmc_gather_walkers")
        return out

    def param_grad(self, pos, alpha, beta):
        """sum_w alpha_w d log|psi_w|/d params + beta_w d phase_w/d params in the PACKED layout (device, (P,)), plus
        (phase, log|psi|) of every walker.  One reverse sweep per walker (csrc/param_grad.cuh); replaces the
        jax.jvp(batch_network) of Loss/pploss.py:204 as seen through jax.grad."""
        p = self._pos(pos).reshape(-1, 3 * self.n)
        B = p.shape[0]
        cv = lambda a: torch.as_tensor(a).to(device=self.device, dtype=torch.float64).reshape(B).contiguous()
        al, be = cv(alpha), cv(beta)
        nbytes = _nbytes(self.lib.aiqmc_param_grad_workspace_bytes(C.byref(self.sys), B), "aiqmc_param_grad_workspace_bytes")
        ws = self._workspace("pgrad", nbytes)
        g = torch.empty(self.layout.total, dtype=torch.float64, device=self.device)
        ph = torch.empty(B, dtype=torch.float64, device=self.device)
        la = torch.empty(B, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.aiqmc_psi_param_grad(C.byref(self.sys), _ptr(self.params_dev), _ptr(p), B, _ptr(al),
                                                     _ptr(be), _ptr(g), _ptr(ph), _ptr(la), _ptr(ws), ws.numel(),
                                                     _stream()), "aiqmc_psi_param_grad")
        return g, ph, la

    def mh_step(self, pos: torch.Tensor, lp: torch.Tensor, noise, u, stddev: float, num_accepts: torch.Tensor,
                want_accept: bool = False):
        """One all-electron Metropolis-Hastings move (MonteCarloSample/mcstep.py:37-68), in place on pos (B,3N) and
        lp (B) = 2 log|psi|; num_accepts: device int64 scalar, incremented.  Returns the accept mask if asked."""
        B = pos.shape[0]
        cv = lambda a, shape: torch.as_tensor(a).to(device=self.device, dtype=torch.float64).reshape(shape).contiguous()
        nz, uu = cv(noise, (B, 3 * self.n)), cv(u, (B,))
        for name, t, shape, dt in (("pos", pos, (B, 3 * self.n), torch.float64), ("lp", lp, (B,), torch.float64),
                                   ("num_accepts", num_accepts, (), torch.int64)):
            if not (t.device == self.device and t.dtype == dt and t.is_contiguous() and tuple(t.shape) == shape):
                raise ValueError(f"mh_step updates {name} in place: contiguous {dt} {shape} on the engine's device")
        ws = self._workspace("mh", _nbytes(self.lib.aiqmc_mh_workspace_bytes(C.byref(self.sys), B), "aiqmc_mh_workspace_bytes"))
        acc = torch.empty(B, dtype=torch.uint8, device=self.device) if want_accept else None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.aiqmc_mh_step(C.byref(self.sys), _ptr(self.params_dev), _ptr(pos), _ptr(lp), _ptr(nz),
                                              _ptr(uu), B, float(stddev), _ptr(acc) if acc is not None else None,
                                              _ptr(num_accepts), _ptr(ws), ws.numel(), _stream()), "aiqmc_mh_step")
        return acc

    def dmc_tmove(self, pos: torch.Tensor, rot: torch.Tensor, u: torch.Tensor, rnd: torch.Tensor, tstep: float,
                  ecp=None):
        """DMC/Tmoves.py:32-225 for the whole batch: (new positions (B,3N), acceptance (B,N), selected move (B,N)).
        `ecp` overrides the engine's table for this call only."""
        ecp_tab = ecp if ecp is not None else self.ecp
        if ecp_tab is None:
            raise ValueError("T-moves need the ccECP tables (engine.ecp or the ecp argument)")
        p = self._pos(pos).reshape(-1, 3 * self.n)
        B = p.shape[0]
        cv = lambda a, shape: torch.as_tensor(a).to(device=self.device, dtype=torch.float64).reshape(shape).contiguous()
        r, uu, rr = cv(rot, (B, 9)), cv(u, (B,)), cv(rnd, (B, self.n))
        ws = self._workspace("tmove", _nbytes(self.lib.aiqmc_dmc_tmove_workspace_bytes(C.byref(self.sys), B),
                                              "aiqmc_dmc_tmove_workspace_bytes"))
        out = torch.empty_like(p)
        acc = torch.empty((B, self.n), dtype=torch.float64, device=self.device)
        sel = torch.empty((B, self.n), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.aiqmc_dmc_tmove(C.byref(self.sys), C.byref(ecp_tab), _ptr(self.params_dev), _ptr(p),
                                                _ptr(r), _ptr(uu), _ptr(rr), B, float(tstep), _ptr(out), _ptr(acc),
                                                _ptr(sel), _ptr(ws), ws.numel(), _stream()), "aiqmc_dmc_tmove")
        return out, acc, sel


class NcclComm:
    """An ncclComm_t owned by the C library (csrc/population.cu), created from a torch.distributed process group: rank 0
    makes the 128-byte unique id, the group broadcasts it, every rank joins.  Used by the C-ABI collectives
    (aiqmc_energy_allreduce, aiqmc_ecut_allreduce_min, aiqmc_rebalance_nccl)."""
    _cache = {}

    def __init__(self, process_group=None, device=None):
        import torch.distributed as dist
        self.lib = _lib.load()
        self.group = process_group
        self.rank, self.world = dist.get_rank(process_group), dist.get_world_size(process_group)
        if not self.lib.aiqmc_nccl_available():
            _lib.check(-5, "NCCL")
        dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        ident = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_ubyte * 128)()
            _lib.check(self.lib.aiqmc_nccl_unique_id(buf), "aiqmc_nccl_unique_id")
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        bdev = dev if dist.get_backend(process_group) == "nccl" else torch.device("cpu")
        ident = ident.to(bdev)
        dist.broadcast(ident, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0, group=process_group)
        raw = bytes(ident.cpu().tolist())
        self.handle = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(self.lib.aiqmc_nccl_comm_init(self.world, self.rank, raw, C.byref(self.handle)), "aiqmc_nccl_comm_init")

    @classmethod
    def for_group(cls, process_group=None, device=None) -> "NcclComm":
        key = (id(process_group), str(device))
        if key not in cls._cache:
            cls._cache[key] = cls(process_group, device)
        return cls._cache[key]

    def close(self):
        if self.handle:
            self.lib.aiqmc_nccl_comm_destroy(self.handle)
            self.handle = C.c_void_p()


class HostStepPipeline:
    """VMC walker steps driven from HOST buffers (the end-to-end path bench.py times as `e2e`): every step copies its
    inputs host -> device and its results device -> host, with the copies arranged around the kernels:
      * the step's random arrays (gauss1, gauss2, rnd, rot) do not depend on the previous step: the set of step k+1
        travels on a copy stream while step k computes, and the host issues that copy after launching step k's kernels
        so that its own work hides behind them;
      * the walker positions make a host round trip per step (they serialise the steps); the sweep is their last
        writer, so they leave for the host while the local-energy kernels run;
      * one synchronisation per step, when the host reads positions and energy statistics.
    All host tensors must be pinned.  `reduce_stats`, if given, is applied to the 4-double statistics on the device
    (e.g. a torch.distributed all-reduce) before they are copied out."""

    def __init__(self, engine: WalkerEngine, tstep: float, reduce_stats=None):
        self.eng, self.tstep, self.reduce_stats = engine, float(tstep), reduce_stats
        self.copy_stream = torch.cuda.Stream(device=engine.device)
        self.swept = torch.cuda.Event()
        self.e_l = None
        self._rng_bufs = None

    def run_seeded(self, pos_host: torch.Tensor, seed: int, step0: int, nsteps: int, stats_host: torch.Tensor,
                   walker0: int = 0) -> None:
        """Throughput mode: the host supplies positions and a seed only; gauss1 / gauss2 / uniforms / rotations are
        generated on the device by the counter-based Philox kernels (csrc/rng.cu), as the reference draws them inside
        its jitted graph.  Per step: positions host -> device, sweep, positions device -> host (overlapping the local
        energy), statistics device -> host."""
        eng, dev = self.eng, self.eng.device
        B = pos_host.shape[0]
        if self.e_l is None or self.e_l.shape[0] != B:
            self.e_l = torch.empty((B, 2), dtype=torch.float64, device=dev)
        if self._rng_bufs is None or self._rng_bufs[0].shape[0] != B:
            self._rng_bufs = (torch.empty((B, 3 * eng.n), dtype=torch.float64, device=dev),
                              torch.empty((B, eng.n, 3), dtype=torch.float64, device=dev),
                              torch.empty((B, eng.n), dtype=torch.float64, device=dev),
                              torch.empty((B, 3, 3), dtype=torch.float64, device=dev))
        g1, g2c, u, rot = self._rng_bufs
        cur = torch.cuda.current_stream(dev)
        for k in range(nsteps):
            p = pos_host.to(dev, non_blocking=True)
            eng.rng_sweep(seed, step0 + k, walker0, B, self.tstep, out=(g1, g2c, u))
            eng.vmc_sweep(p, g1, g2c, u, self.tstep, want_accept=False)
            self.swept.record(cur)
            self.copy_stream.wait_event(self.swept)
            with torch.cuda.stream(self.copy_stream):
                pos_host.copy_(p, non_blocking=True)
            p.record_stream(self.copy_stream)
            if eng.ecp is not None:
                eng.rng_rotations(seed, step0 + k, walker0, B, out=rot)
                eng.local_energy(p, rot, out=self.e_l)
                e = torch.view_as_complex(self.e_l)
            else:
                e = eng.local_energy(p)
            stats = eng.energy_stats(e)
            if self.reduce_stats is not None:
                self.reduce_stats(stats)
            stats_host.copy_(stats, non_blocking=True)
            cur.synchronize()
            self.copy_stream.synchronize()

    def _prefetch(self, host_set):
        with torch.cuda.stream(self.copy_stream):
            dev_set = {k: v.to(self.eng.device, non_blocking=True) for k, v in host_set.items()}
            done = torch.cuda.Event()
            done.record(self.copy_stream)
        return dev_set, done

    def run(self, pos_host: torch.Tensor, host_sets, stats_host: torch.Tensor) -> None:
        """Runs len(host_sets) steps (sweep + local energy + statistics); pos_host (B,3N) is updated in place after
        every step, stats_host (4,) holds [sum Re E, sum Im E, sum |E|^2, count] of the last step."""
        eng, dev = self.eng, self.eng.device
        B = pos_host.shape[0]
        if self.e_l is None or self.e_l.shape[0] != B:
            self.e_l = torch.empty((B, 2), dtype=torch.float64, device=dev)
        if not host_sets:
            return
        cur = torch.cuda.current_stream(dev)
        nxt = self._prefetch(host_sets[0])
        for k in range(len(host_sets)):
            s, done = nxt
            cur.wait_event(done)
            p = pos_host.to(dev, non_blocking=True)
            eng.vmc_sweep(p, s["gauss1"], s["gauss2"], s["rnd"], self.tstep, want_accept=False)
            self.swept.record(cur)
            self.copy_stream.wait_event(self.swept)
            with torch.cuda.stream(self.copy_stream):
                pos_host.copy_(p, non_blocking=True)
            p.record_stream(self.copy_stream)
            if eng.ecp is not None:
                eng.local_energy(p, s["rot"], out=self.e_l)
                e = torch.view_as_complex(self.e_l)
            else:
                e = eng.local_energy(p)
            stats = eng.energy_stats(e)
            if self.reduce_stats is not None:
                self.reduce_stats(stats)
            stats_host.copy_(stats, non_blocking=True)
            for v in s.values():
                v.record_stream(cur)
            if k + 1 < len(host_sets):
                nxt = self._prefetch(host_sets[k + 1])
            cur.synchronize()
            self.copy_stream.synchronize()
