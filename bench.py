#!/usr/bin/env python
"""bench.py -- VMC walker-steps/s incl. local energy (BASELINE.json metric) on the B200-native walker engine.

One "step" = one walkers_update sweep (AIQMCrelease3/VMC/VMCmcstep.py:28-111) + one local_energy
(Energy/pphamiltonian.py:130-190, Energy/hamiltonian.py:236-260) + the energy mean/variance partials (one 4-double
NCCL all-reduce, Loss/pploss.py:165-167) over one batch of synthetic walkers; for the DMC workload one
dmc_propagate_run (DMC/dmc.py:72-93) + the cross-GPU systematic comb / walker migration (DMC/branch.py:10-34,
DMC/main_dmc.py:208-242 made global).  Walkers are the independent units: sharded over ranks, weak scaling.

  python bench.py --gpus N --steps K --warmup W                 # headline: BASELINE configs[1] (C atom ccECP, 65,536 walkers/GPU)
  python bench.py --workload {c_ecp,c_ae,n2,c6h6,dmc} ...       # any BASELINE configuration as the headline
  python bench.py --impl reference ...                          # the reference's CPU path timed on host cores

Unless --no-side is given, the line also carries "workloads": the OTHER BASELINE configurations (n2 at 65,536 walkers
per GPU, dmc with the cross-GPU population control, c6h6, c_ae), each measured in this same run at this same N with the
same timing rules (CUDA events, max over ranks), so a 1/2/4/8 sweep of this command covers every configuration.
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import ctypes as C
import csv
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "VMC walker-steps/s incl. local energy"
UNIT = "walker-steps/s"
SEED = 20260101
TSTEP = 0.05
SIDE_DEFAULTS = {"n2": dict(walkers=65536, steps=2, warmup=1), "dmc": dict(walkers=65536, steps=3, warmup=1),
                 "c6h6": dict(walkers=2368, steps=1, warmup=1), "c_ae": dict(walkers=4096, steps=5, warmup=2)}


# ------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: an in-process NVML thread polling every 20 ms;
    nvidia-smi is the fallback when pynvml is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v for v in vis.split(",") if v.strip().isdigit()]
        self.gpu = int(ids[gpu_index]) if gpu_index < len(ids) else gpu_index
        self.rows = []          # (sm_mhz, max_mhz, set(reasons))
        self.proc = None
        self._stop = threading.Event()
        self._thread = None

    def _nvml_loop(self, nv, h):
        masks = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.rows.append((sm, mx, {k for k, m in masks.items() if bits & m}))
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            self._thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self._thread.start()
            return
        except Exception:
            self._thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                self.rows.append((float(r[1]), float(r[2]),
                                  {name for k, name in enumerate(names) if r[5 + k].lower().startswith("active")}))
            except Exception:
                pass

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=1.0)
        if self.proc is not None:
            self.proc.terminate()
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = set()
        for r in self.rows:
            reasons |= r[2]
        return {"sm_mhz": float(np.median([r[0] for r in self.rows])), "sm_max_mhz": float(max(r[1] for r in self.rows)),
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
class Dist:
    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device(f"cuda:{self.local_rank}")
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, seconds: float) -> float:
        t = torch.tensor([seconds], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t)

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def _events(n):
    return [torch.cuda.Event(enable_timing=True) for _ in range(n)]


def measure_vmc(D: Dist, name: str, B: int, steps: int, warmup: int, flush, detail: bool):
    """walker-steps/s of sweep + local energy + statistics (+ all-reduce) with every input resident in HBM.
    Returns (result dict, context for the e2e / roofline legs)."""
    import aiqmc_b200
    from aiqmc_b200 import workloads as W
    wl = W.build(name, B, seed=SEED, rank=D.rank)
    eng = wl.engine(D.dev)
    n, a = wl.n, wl.a
    with_ecp = wl.ecp is not None
    walker0 = D.rank * B
    nsets = steps + warmup
    # every step's random inputs, generated ahead of the timed region by the library's Philox kernels (counter =
    # (global walker id, step): independent of the sharding) and left resident in HBM
    sets = []
    for k in range(nsets):
        g1, g2c, u = eng.rng_sweep(SEED, k, walker0, B, TSTEP)
        rot = eng.rng_rotations(SEED, k, walker0, B) if with_ecp else None
        sets.append((g1, g2c, u, rot))
    pos = torch.from_numpy(wl.pos.copy()).to(D.dev)
    e_l = torch.empty((B, 2), dtype=torch.float64, device=D.dev)

    def step(s, ev=None):
        g1, g2c, u, rot = s
        if ev is not None:
            ev[0].record()
        eng.vmc_sweep(pos, g1, g2c, u, TSTEP, want_accept=False)
        if ev is not None:
            ev[1].record()
        if with_ecp:
            eng.local_energy(pos, rot, stages=1, out=e_l)
            if ev is not None:
                ev[2].record()
            eng.local_energy(pos, rot, stages=2, out=e_l)
            if ev is not None:
                ev[3].record()
            eng.local_energy(pos, rot, stages=4, out=e_l)
            stats = eng.energy_stats(torch.view_as_complex(e_l))
        else:
            e = eng.local_energy(pos)
            if ev is not None:
                ev[2].record()
                ev[3].record()
            stats = eng.energy_stats(e)
        if D.world > 1:
            D.dist.all_reduce(stats)                         # the path's only collective (pploss.py:165-167)
        if ev is not None:
            ev[4].record()
        return stats

    for w in range(warmup):
        step(sets[w])
    D.barrier()
    evs = [_events(5) for _ in range(steps)]
    launches0 = eng.lib.aiqmc_launch_count()
    t0 = time.perf_counter()
    last = None
    for k in range(steps):
        flush.zero_()                                        # L2 flush between timed iterations (untimed)
        last = step(sets[warmup + k], evs[k])
    launches = int(eng.lib.aiqmc_launch_count() - launches0)
    D.barrier()
    wall = time.perf_counter() - t0
    ms = np.array([[e[i].elapsed_time(e[i + 1]) for i in range(4)] for e in evs])        # sweep, kinetic, quad, rest
    ms_step = np.array([e[0].elapsed_time(e[4]) for e in evs])
    total = D.max_over_ranks(float(ms_step.sum()) / 1e3)
    value = D.world * B * steps / total
    stats = last.cpu().numpy()
    res = {"label": wl.label, "value": value, "unit": UNIT, "walkers_per_gpu": B, "global_walkers": B * D.world,
           "steps": steps, "warmup": warmup, "ms_per_step": total / steps * 1e3, "n_elec": n, "n_atoms": a,
           "algorithmic_flops_per_walker_step": W.flops_walker_step(name),
           "energy_mean_last_step": float(stats[0] / stats[3])}
    if detail:
        res["gpu_launches"] = launches
        res["wall_s_timed_region"] = wall
    ctx = dict(wl=wl, eng=eng, sets=sets, ms_stage=ms.mean(axis=0), ms_step=float(ms_step.mean()), pos0=wl.pos, B=B)
    return res, ctx


def stage_table(ctx, peak_tflops):
    from aiqmc_b200 import workloads as W
    wl, B = ctx["wl"], ctx["B"]
    fl = W.stage_flops(wl.n, wl.a, wl.ecp is not None)
    names = ["sweep", "kinetic", "quadrature"] if wl.ecp is not None else ["sweep", "kinetic"]
    kernels = {"sweep": "grad x1 + grad of the N single-electron moves + accept (A11-A13)",
               "kinetic": "forward Laplacian + Coulomb + local ECP channel (A15-A18)",
               "quadrature": "ccECP non-local 50-point quadrature (A19-A23)"}
    out = {}
    for i, s in enumerate(names):
        ms = float(ctx["ms_stage"][i])
        tf = fl[s] * B / (ms * 1e-3) / 1e12 if ms > 0 else None
        out[s] = {"ms": ms, "algorithmic_tflops": tf, "frac": (tf / peak_tflops) if (tf and peak_tflops) else None,
                  "share_of_step": ms / ctx["ms_step"], "what": kernels[s]}
    return out


def measure_dmc(D: Dist, B: int, steps: int, warmup: int, flush, args_mode: str = "balanced"):
    """DMC walker-steps/s: dmc_propagate_run (T-move, drift-diffusion sweep, two ccECP local energies, S, weights)
    + the cross-GPU population control (global systematic comb + migration of the selected walkers over NCCL)."""
    import aiqmc_b200
    from aiqmc_b200 import workloads as W
    wl = W.build("dmc", B, seed=SEED, rank=D.rank)
    n, a = wl.n, wl.a
    net = aiqmc_b200.make_ai_net(wl.spec.nspins, wl.spec.charges, wl.spec.parallel_indices, wl.spec.antiparallel_indices,
                                 wl.spec.spin_up_indices, wl.spec.spin_down_indices, wl.spec.parallel_indices.shape[1],
                                 wl.spec.antiparallel_indices.shape[1], 3, a, n, device=D.dev)
    group = D.dist.group.WORLD if D.world > 1 else None
    run = aiqmc_b200.dmc_propagate(net.apply, net.apply, TSTEP, n, a, 3, B, wl.spec.charges, process_group=group, **wl.tables)
    packed = net.pack(wl.params, torch.tensor(wl.spec.atoms))
    eng = packed.engine
    walker0 = D.rank * B
    nsets = steps + warmup
    keys = []
    for k in range(nsets):
        g1, g2c, u = eng.rng_sweep(SEED, 1000 + k, walker0, B, TSTEP)
        keys.append(dict(tmove=dict(rot=eng.rng_rotations(SEED, 2000 + k, walker0, B),
                                    u=eng.rng_uniform(SEED, 1000 + k, walker0, B, 1, 1).reshape(B),
                                    rnd=eng.rng_uniform(SEED, 1000 + k, walker0, B, n, 2)),
                         sweep=dict(gauss1=g1, gauss2=g2c, rnd=u), rot=eng.rng_rotations(SEED, 1000 + k, walker0, B)))
    state = {"data": aiqmc_b200.AINetData(positions=torch.from_numpy(wl.pos.copy()).to(D.dev), spins=torch.tensor(wl.spins),
                                          atoms=torch.tensor(wl.spec.atoms), charges=torch.tensor(wl.spec.charges)),
             "w": torch.ones(B, dtype=torch.float64, device=D.dev)}
    branchcut = torch.full((B,), 3.0, dtype=torch.float64, device=D.dev)
    moved = []

    def step(k):
        e_new, w, new_data = run(packed, keys[k], state["data"], state["w"], branchcut, -5.39, -5.41)
        neww, new_pos, src, imported = aiqmc_b200.branch_global(eng, w, new_data.positions, 0.37 + 0.01 * k, group, mode=args_mode)
        state["data"] = aiqmc_b200.AINetData(positions=new_pos, spins=state["data"].spins, atoms=state["data"].atoms,
                                             charges=state["data"].charges)
        state["w"] = torch.ones(B, dtype=torch.float64, device=D.dev) * neww
        moved.append(int(imported))

    for k in range(warmup):
        step(k)
    D.barrier()
    evs = [_events(2) for _ in range(steps)]
    launches0 = eng.lib.aiqmc_launch_count()
    for k in range(steps):
        flush.zero_()
        evs[k][0].record()
        step(warmup + k)
        evs[k][1].record()
    launches = int(eng.lib.aiqmc_launch_count() - launches0)
    D.barrier()
    total = D.max_over_ranks(sum(e[0].elapsed_time(e[1]) for e in evs) / 1e3)
    value = D.world * B * steps / total
    fl = W.flops_walker_step("dmc")
    return {"label": wl.label, "value": value, "unit": "DMC walker-steps/s", "walkers_per_gpu": B, "global_walkers": B * D.world,
            "steps": steps, "warmup": warmup, "ms_per_step": total / steps * 1e3, "n_elec": n, "n_atoms": a,
            "algorithmic_flops_per_walker_step": fl, "algorithmic_tflops_per_gpu": value / D.world * fl / 1e12,
            "walkers_imported_from_other_ranks_last_step": moved[-1] if moved else 0, "gpu_launches": launches,
            "population_control": f"global systematic comb over all ranks + NCCL migration every step, mode '{args_mode}' "
                                  "(balanced: the comb's survivors and multiplicities, every rank keeps its own walkers, only "
                                  "the population imbalance crosses NVLink; ordered: the single-GPU slot order, moves ~all)"}


def measure_e2e(D: Dist, ctx, steps: int, warmup: int):
    """The same metric through the package's host-buffer entry point (aiqmc_b200.HostStepPipeline): every step copies
    its inputs from pinned host memory and its results back inside the timed region.
      seeded : the drop-in call mc_step(params, data, key) -- the host supplies positions + a seed, the random
               arrays are drawn on the device (Philox), as the reference draws them inside its jitted graph;
      parity : explicit gauss1 / gauss2 (compact) / uniforms / rotations travel from the host too."""
    import aiqmc_b200
    eng, B, wl = ctx["eng"], ctx["B"], ctx["wl"]
    n = wl.n
    reduce = (lambda st: D.dist.all_reduce(st)) if D.world > 1 else None
    pipe = aiqmc_b200.HostStepPipeline(eng, TSTEP, reduce_stats=reduce)
    out = {}
    pos_host = torch.from_numpy(ctx["pos0"].copy()).pin_memory()
    stats_host = torch.empty(4, dtype=torch.float64).pin_memory()
    # seeded
    pipe.run_seeded(pos_host, SEED, 0, min(warmup, 3), stats_host, walker0=D.rank * B)
    D.barrier()
    e0, e1 = _events(2)
    e0.record()
    pipe.run_seeded(pos_host, SEED, warmup, steps, stats_host, walker0=D.rank * B)
    e1.record()
    D.barrier()
    t = D.max_over_ranks(e0.elapsed_time(e1) / 1e3)
    out["seeded"] = {"value": D.world * B * steps / t, "unit": UNIT, "h2d_bytes_per_step": int(pos_host.numel() * 8 + 8),
                     "d2h_bytes_per_step": int(pos_host.numel() * 8 + 32)}
    # parity inputs from the host
    host_sets = []
    for (g1, g2c, u, rot) in ctx["sets"][:steps + warmup]:
        hs = {"gauss1": g1.cpu().pin_memory(), "gauss2": g2c.cpu().pin_memory(), "rnd": u.cpu().pin_memory()}
        if rot is not None:
            hs["rot"] = rot.cpu().pin_memory()
        host_sets.append(hs)
    pos_host.copy_(torch.from_numpy(ctx["pos0"]))
    pipe.run(pos_host, host_sets[:min(warmup, 3)], stats_host)
    D.barrier()
    e0, e1 = _events(2)
    e0.record()
    pipe.run(pos_host, host_sets[warmup:warmup + steps], stats_host)
    e1.record()
    D.barrier()
    t = D.max_over_ranks(e0.elapsed_time(e1) / 1e3)
    h2d = sum(v.numel() * v.element_size() for v in host_sets[0].values()) + pos_host.numel() * 8
    out["parity"] = {"value": D.world * B * steps / t, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                     "d2h_bytes_per_step": int(pos_host.numel() * 8 + 32)}
    return out


def measure_fp64_peak(eng, dev):
    """DFMA throughput of this GPU, measured live (MEASURED_PEAKS.json has no FP64 figure)."""
    sink = torch.zeros(8, dtype=torch.float64, device=dev)
    fl = C.c_double(0.0)
    st_ptr = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    best = 0.0
    for it in range(6):
        a0, a1 = _events(2)
        a0.record()
        eng.lib.aiqmc_bench_dfma(200000, C.c_void_p(sink.data_ptr()), C.byref(fl), st_ptr)
        a1.record()
        torch.cuda.synchronize()
        if it >= 1:
            best = max(best, fl.value / (a0.elapsed_time(a1) * 1e-3) / 1e12)
    return best


def ncu_metric(path, metric, kernel=None):
    """One metric value (mean over the launches of `kernel`) out of a committed `ncu --page raw --csv` file (the bench
    line cites it with its source)."""
    full = os.path.join(ROOT, path)
    if not os.path.exists(full):
        return None
    try:
        rows = list(csv.reader(open(full)))
        hdr = next(r for r in rows if metric in r)
        col = hdr.index(metric)
        vals = []
        kcol = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
        for r in rows[rows.index(hdr) + 1:]:
            if kernel is not None and kcol is not None and kernel not in r[kcol]:
                continue
            try:
                vals.append(float(r[col].replace(",", "")))
            except (ValueError, IndexError):
                pass
        return float(np.mean(vals)) if vals else None
    except Exception:
        return None


def time_hbm_kernels(eng, B, n, dev, flush):
    """The HBM-bound kernels north_star wants evidenced by achieved GB/s: gather (A29), energy statistics (A24)."""
    out = {}
    pos = torch.randn((B, 3 * n), dtype=torch.float64, device=dev)
    inds = torch.randint(0, B, (B,), dtype=torch.int32, device=dev)
    e = torch.randn((B, 2), dtype=torch.float64, device=dev)
    w = torch.rand(B, dtype=torch.float64, device=dev) + 0.5

    def timed(fn, reps=5):
        fn()
        ts = []
        for _ in range(reps):
            flush.zero_()
            flush.sum()                 # then READ it: L2 is left full of clean lines, so the write-back of the flush itself
                                        # (up to 126 MB of dirty lines) is not charged to a 12-200 MB kernel
            a0, a1 = _events(2)
            a0.record(); fn(); a1.record(); torch.cuda.synchronize()
            ts.append(a0.elapsed_time(a1))
        return float(np.min(ts))
    ms = timed(lambda: eng.gather_walkers(pos, inds))
    out["gather_walkers"] = {"ms": ms, "GB_per_s": (B * (48 * n + 4)) / (ms * 1e-3) / 1e9, "bytes": B * (48 * n + 4)}
    ms = timed(lambda: eng.energy_stats(torch.view_as_complex(e)))
    out["energy_stats"] = {"ms": ms, "GB_per_s": (B * 16) / (ms * 1e-3) / 1e9, "bytes": B * 16}
    ms = timed(lambda: eng.branch_comb(w, 0.37))
    out["branch_comb"] = {"ms": ms, "GB_per_s": (B * 28) / (ms * 1e-3) / 1e9, "bytes": B * 28}
    # A0 micro-benchmark: the reference's C cc-pVDZ basis (13 AOs) with gradient and Laplacian at
    # walkers x 6 electron positions: 24 B in, 65 doubles out per point
    import aiqmc_b200
    from aiqmc_b200 import workloads as W
    basis = aiqmc_b200.GaussianBasis.from_nwchem(W.C_CC_PVDZ, np.zeros((1, 3)), device=dev)
    pts = torch.randn((B * 6, 3), dtype=torch.float64, device=dev)
    ms = timed(lambda: basis.eval(pts))
    nbytes = B * 6 * (24 + 65 * 8)
    out["gto_eval_A0"] = {"ms": ms, "GB_per_s": nbytes / (ms * 1e-3) / 1e9, "bytes": nbytes, "points": B * 6,
                          "points_per_s": B * 6 / (ms * 1e-3), "what": "C cc-pVDZ, 13 AOs, value + gradient + Laplacian"}
    # the same kernels on a 16x batch: at the BASELINE size gather / statistics / comb move 1-13 MB and are bounded by
    # launch latency (~10 us), not by HBM; these rows show what the kernels sustain once the batch is HBM-sized
    Bl = 16 * B
    posl = torch.randn((Bl, 3 * n), dtype=torch.float64, device=dev)
    indl = torch.randint(0, Bl, (Bl,), dtype=torch.int32, device=dev)
    ms = timed(lambda: eng.gather_walkers(posl, indl))
    out["gather_walkers_x16"] = {"ms": ms, "GB_per_s": (Bl * (48 * n + 4)) / (ms * 1e-3) / 1e9, "bytes": Bl * (48 * n + 4)}
    del posl, indl
    el = torch.randn((Bl, 2), dtype=torch.float64, device=dev)
    ms = timed(lambda: eng.energy_stats(torch.view_as_complex(el)))
    out["energy_stats_x16"] = {"ms": ms, "GB_per_s": (Bl * 16) / (ms * 1e-3) / 1e9, "bytes": Bl * 16,
                               "what": "aiqmc_energy_stats_ws: chunk partials + fixed-order sum (beyond 2^18 walkers)"}
    del el
    ptl = torch.randn((Bl * 6, 3), dtype=torch.float64, device=dev)
    ms = timed(lambda: basis.eval(ptl))
    nbl = Bl * 6 * (24 + 65 * 8)
    out["gto_eval_A0_x16"] = {"ms": ms, "GB_per_s": nbl / (ms * 1e-3) / 1e9, "bytes": nbl, "points": Bl * 6}
    del ptl
    torch.cuda.empty_cache()
    return out


def make_config(label, head, B, world):
    """The `config` object of the JSON line -- identical for the b200 and the reference arm of the same command."""
    return {"workload": label, "workload_key": head, "walkers_per_gpu": B, "global_walkers": B * world,
            "tstep": TSTEP, "nsteps_per_step": 1, "params": "random-init (reference init scales)",
            "parallelism": f"walker-sharded x{world}",
            "l2": "256 MiB flush write between timed iterations (untimed)",
            "rng": "per-step gauss/uniform/rotation arrays generated ahead of the timed region by the library's "
                   "Philox kernels (counter = global walker id, step) and resident in HBM",
            "tables": "carbon ccECP verbatim from the reference example; N/H atoms re-use the carbon table and "
                      "the N2 / C6H6 geometries are builder-defined (the reference ships neither)"}


def run_gpu(args):
    D = Dist()
    from aiqmc_b200 import workloads as W
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=D.dev)      # 256 MiB > 126 MB L2
    head = args.workload
    B = args.walkers if args.walkers else W.SYSTEMS[head]["walkers"]
    sampler = ClockSampler(D.local_rank)
    sampler.start()
    ctx = None
    if head == "dmc":
        res = measure_dmc(D, B, args.steps, args.warmup, flush)
    else:
        res, ctx = measure_vmc(D, head, B, args.steps, args.warmup, flush, detail=True)
    clocks = sampler.stop()
    e2e = measure_e2e(D, ctx, args.steps, args.warmup) if ctx is not None else None

    side = {}
    if not args.no_side:
        for name, cfg in SIDE_DEFAULTS.items():
            if name == head:
                continue
            try:
                if name == "dmc":
                    side[name] = measure_dmc(D, cfg["walkers"], cfg["steps"], cfg["warmup"], flush)
                else:
                    r, c = measure_vmc(D, name, cfg["walkers"], cfg["steps"], cfg["warmup"], flush, detail=False)
                    r["algorithmic_tflops_per_gpu"] = r["value"] / D.world * r["algorithmic_flops_per_walker_step"] / 1e12
                    r["stage_ms"] = {k: float(v) for k, v in zip(("sweep", "kinetic", "quadrature"), c["ms_stage"][:3])}
                    side[name] = r
                    del c
                torch.cuda.empty_cache()
            except Exception as exc:                        # never lose the headline line to a side measurement
                side[name] = {"error": repr(exc)}
        if "c6h6" in side and "error" not in side["c6h6"]:
            side["c6h6"]["note"] = ("bounded sample of the 32,768-walker per-GPU shard of the 262,144-walker configuration "
                                    "(a full shard step takes ~14 s; run --workload c6h6 for it)")

    if D.rank != 0:
        D.close()
        return

    # ---- roofline of the dominant kernel: FP64 FMA peak measured live; per-stage table --------------------
    roofline, cpu, hbm = None, None, None
    if ctx is not None:
        eng = ctx["eng"]
        peak = measure_fp64_peak(eng, D.dev)
        stages = stage_table(ctx, peak)
        dom = max(stages, key=lambda s: stages[s]["ms"])
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if head == "c_ecp" and os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("k_ecp_pt_dram_bytes_per_step")
            except Exception:
                traffic = None
        ncu_src = "profiles/r2_v6_ecp_pt_raw.csv"
        roofline = {"bound": "fp64", "kernel": f"{dom} stage of {head}", "achieved": stages[dom]["algorithmic_tflops"],
                    "peak": peak, "unit": "TFLOP/s", "frac": stages[dom]["frac"], "traffic": traffic,
                    "stages": stages,
                    "whole_step_frac": (res["value"] / D.world) * res["algorithmic_flops_per_walker_step"] / 1e12 / peak if peak else None,
                    "fp64_pipe_busy_ncu": {"value": ncu_metric(ncu_src, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "k_ecp_pt") if head == "c_ecp" else None,
                                           "source": ncu_src, "note": "parsed from the committed ncu page at run time; not measured live"},
                    "note": "compute-bound on the FP64 pipe (SURVEY 8d), not HBM/tensor: achieved = SURVEY's FIXED algorithmic "
                            "flops per walker (a tanh priced at ~1 flop) x walkers / CUDA-event time of the stage; peak = DFMA "
                            "microbenchmark of this run (nominal 37.2 TFLOP/s; MEASURED_PEAKS.json has no FP64 entry). "
                            "profiles/r2_fp64_pipes.txt: FP64 DMMA (mma.sync f64) runs on the same units as DFMA on B200 "
                            "(37.0 vs 33.9 TFLOP/s alone, 35 mixed), so there is no second FP64 pipe to move work to."}
        try:
            hbm = time_hbm_kernels(eng, ctx["B"], ctx["wl"].n, D.dev, flush)
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
            for v in hbm.values():
                v["frac_of_hbm_peak"] = v["GB_per_s"] / peaks.get("hbm_gbs", 6537.6)
        except Exception as exc:
            hbm = {"error": repr(exc)}
    if ctx is None:                      # DMC headline: whole-step roofline against the live DFMA peak
        import aiqmc_b200
        from aiqmc_b200 import workloads as W
        peak = measure_fp64_peak(W.build("dmc", 2).engine(D.dev), D.dev)
        roofline = {"bound": "fp64", "kernel": "whole DMC step (T-move + drift-diffusion sweep + 2 ccECP local energies + S + comb)",
                    "achieved": res["algorithmic_tflops_per_gpu"], "peak": peak, "unit": "TFLOP/s",
                    "frac": res["algorithmic_tflops_per_gpu"] / peak if peak else None, "traffic": None,
                    "note": "algorithmic flops (150NA + 9N + 11) F(N,A) per walker (SURVEY 8d) / CUDA-event step time / DFMA microbenchmark"}
    if D.world == 1 and not args.no_cpu_baseline:
        cw, cs = cpu_sample(head, args.cpu_walkers)
        cpu = cpu_baseline(head if head != "dmc" else "c_ecp", cw, cs, 1)

    line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": D.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": make_config(res["label"], head, B, D.world),
            "clocks": clocks, "gpu_launches": res.get("gpu_launches"),
            "e2e": ({**e2e["seeded"], "mode": "seeded: positions + seed from the host, randoms drawn on the device",
                     "parity_inputs": e2e["parity"]} if e2e else None),
            "roofline": roofline, "cpu_baseline": cpu, "workloads": side, "hbm_kernels": hbm,
            "headline_detail": {k: v for k, v in res.items() if k not in ("value", "unit", "label")}}
    print(json.dumps(line), flush=True)
    D.close()


# ------------------------------------------------------------------------------------------
# CPU arms: the genuine reference under JAX when it is importable, else the oracle port
# ------------------------------------------------------------------------------------------
def _oracle_case(name, nwalkers):
    """Oracle-side network + parameters + walkers of a product workload (CPU legs only)."""
    from oracle import aiqmc_oracle as O
    from aiqmc_b200 import workloads as W
    wl = W.build(name, nwalkers, seed=SEED)
    sp = wl.spec
    kw = dict(nspins=sp.nspins, charges=sp.charges, parallel_indices=sp.parallel_indices,
              antiparallel_indices=sp.antiparallel_indices, spin_up_indices=sp.spin_up_indices,
              spin_down_indices=sp.spin_down_indices, n_parallel=sp.parallel_indices.shape[1],
              n_antiparallel=sp.antiparallel_indices.shape[1], ndim=3, natoms=wl.a, nelectrons=wl.n)
    spins = np.zeros(wl.n)
    spins[sp.spin_up_indices] = 1.0
    spins[sp.spin_down_indices] = -1.0
    return O, wl, kw, spins


def cpu_step_fn(name, nwalkers, dtype=torch.float32):
    """fn() runs ONE walker-step batch (sweep + local energy) on the CPU oracle in the reference's own dtype
    (float32 / complex64, quirk Q1)."""
    O, wl, kw, spins = _oracle_case(name, nwalkers)
    net = O.make_ai_net(**kw, dtype=dtype)
    params = O.tree_map(lambda t: torch.as_tensor(np.asarray(t, dtype=np.float64)).to(dtype), wl.params)
    rng = np.random.default_rng(SEED + 7)
    logabs = O.select_output(net.apply, 1)
    B, n, a = nwalkers, wl.n, wl.a
    if wl.tables is not None:
        t = wl.tables
        le = O.local_energy_ecp(net.apply, O.make_log_network(net.apply), wl.spec.charges, None, t['rn_local'], t['local_coes'],
                                t['local_exps'], t['rn_non_local'], t['non_local_coes'], t['non_local_exps'], a, n, 3, 2)
    else:
        le = O.local_energy_ae(net.apply, wl.spec.charges)
    state = {"pos": torch.tensor(wl.pos, dtype=dtype)}
    atoms, spins_t = torch.tensor(wl.spec.atoms, dtype=dtype), torch.tensor(spins, dtype=dtype)
    charges = torch.tensor(wl.spec.charges, dtype=dtype)

    def fn():
        s = math.sqrt(TSTEP)
        rand = dict(gauss1=torch.tensor(rng.standard_normal((B, 3 * n)) * s, dtype=dtype),
                    gauss2=torch.tensor(rng.standard_normal((B, n, 3 * n)) * s, dtype=dtype),
                    rnd=torch.tensor(rng.uniform(size=(B, n)), dtype=dtype))
        rot = torch.tensor(O.random_rotations(rng, B), dtype=dtype) if wl.tables is not None else None
        data = O.AINetData(positions=state["pos"], spins=spins_t.expand(B, n), atoms=atoms.expand(B, a, 3),
                           charges=charges.expand(B, a))
        new = O.walkers_update(logabs, params, data, rand, TSTEP, 3, n, B)
        state["pos"] = new.positions
        e, _ = le(params, rot, O.AINetData(positions=new.positions, spins=spins_t, atoms=atoms, charges=charges))
        return float(e.real.mean())
    return fn


def reference_step_fn(name, nwalkers):
    """The GENUINE reference (unmodified AIQMCrelease3 under JAX, its own float32 default), if importable."""
    from oracle import reference_jax as RJ
    ok, why, _ = RJ.probe()
    if not ok:
        return None, why
    try:
        O, wl, kw, spins = _oracle_case(name, nwalkers)
        if wl.tables is None:
            return None, "the genuine-reference timing is wired for the ccECP workloads"
        h = RJ.ReferenceHarness(kw, wl.params, wl.spec.atoms, wl.spec.charges, spins, x64=False)
        step = h.make_timed_step(wl.tables, nwalkers, TSTEP)
        state = {"pos": wl.pos.astype(np.float32), "k": 0}

        def fn():
            state["pos"], e = step(state["pos"], state["k"])
            state["k"] += 1
            return float(np.real(e).mean())
        fn()                                                   # compile outside the timed region
        return fn, "ok"
    except Exception as exc:
        return None, f"the reference failed to build or run: {exc!r}"


# bounded CPU samples (about 10-30 s of host work per workload: the oracle's ccECP energy is 50 N A psi evaluations per
# walker -- 2,048 benzene walkers would be 37 M evaluations of a 30-electron network and ran the host out of memory)
CPU_SAMPLE = {"c_ecp": (2048, 5), "c_ae": (2048, 5), "dmc": (2048, 5), "n2": (96, 3), "c6h6": (2, 2)}


def cpu_sample(head, cpu_walkers):
    """(walkers, steps) of the CPU leg: --cpu-walkers if given, else the workload's bounded default."""
    w, st = CPU_SAMPLE[head]
    return (cpu_walkers if cpu_walkers > 0 else w), st


def cpu_baseline(name, nwalkers, steps, warmup):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fn, why = reference_step_fn(name, nwalkers)
    kind = "reference"
    if fn is None:
        fn, kind = cpu_step_fn(name, nwalkers), "port"
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = time.perf_counter() - t0
    what = ("unmodified AIQMCrelease3 under JAX (float32)" if kind == "reference"
            else f"oracle port, torch CPU float32 (genuine reference unavailable: {why})")
    return {"value": nwalkers * steps / dt, "unit": UNIT, "cores": cores if kind == "reference" else torch.get_num_threads(),
            "kind": kind, "sample": f"{nwalkers} walkers x {steps} steps of the same workload ({what}, {dt / steps:.2f} s/step)"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores -- the genuine
    AIQMCrelease3 modules when `import jax` works (kind "reference"), else the oracle port (kind "port")."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from aiqmc_b200 import workloads as W
    head = args.workload if args.workload != "dmc" else "c_ecp"
    B = args.walkers if args.walkers else W.SYSTEMS[args.workload]["walkers"]
    nwalk, _ = cpu_sample(args.workload, args.cpu_walkers)
    cpu = cpu_baseline(head, nwalk, args.steps, args.warmup)
    value = cpu["value"]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": nwalk / value * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": make_config(W.SYSTEMS[args.workload]["label"], args.workload, B, args.gpus),
            "reference_arm_note": "CPU arm on a bounded sample of the same workload: " + cpu["sample"],
            "cpu_baseline": cpu, "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c_ecp", choices=["c_ecp", "c_ae", "n2", "c6h6", "dmc"])
    ap.add_argument("--walkers", type=int, default=0, help="walkers per GPU (default: the workload's BASELINE size)")
    ap.add_argument("--cpu-walkers", type=int, default=0,
                    help="walkers in the bounded CPU sample (0: per-workload default, e.g. 2,048 for carbon, 2 for benzene)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side", "--no-other-systems", dest="no_side", action="store_true",
                    help="skip the other BASELINE configurations measured next to the headline")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
