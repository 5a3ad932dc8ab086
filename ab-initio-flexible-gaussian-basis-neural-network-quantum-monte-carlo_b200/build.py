"""Builds the CUDA libraries (sm_100a) in-tree with nvcc: libaiqmc_b200.so -- the C ABI of include/aiqmc_b200.h, the
system-independent kernels and the plugin loader -- plus one libaiqmc_sys_<N>_<A>.so per (n_elec, n_atoms) instantiation
of the per-system kernels (csrc/engine_impl.cuh), compiled in parallel.  csrc/dispatch.h names the prebuilt set;
ensure_system(n, a) adds any other system at run time."""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(CSRC, os.environ.get("AIQMC_BUILD_DIR", "build"))
LIB = os.path.join(HERE, os.environ.get("AIQMC_LIB_NAME", "libaiqmc_b200.so"))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def systems():
    if os.environ.get("AIQMC_SYSTEMS"):          # debugging aid: "10,2;4,1" builds only those instantiations
        return [tuple(int(v) for v in s.split(",")) for s in os.environ["AIQMC_SYSTEMS"].split(";")]
    txt = open(os.path.join(CSRC, "dispatch.h")).read()
    body = txt[txt.index("#define AIQMC_FOR_EACH_SYSTEM"):]
    return [(int(a), int(b)) for a, b in re.findall(r"X\((\d+),\s*(\d+)\)", body)]


def _digest(files):
    h = hashlib.sha256()
    for p in files:
        h.update(os.path.basename(p).encode())
        h.update(open(p, "rb").read())
    h.update(" ".join(FLAGS + ARCH + _extra_flags()).encode())
    return h.hexdigest()


def _headers():
    out = []
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        out += [os.path.join(root, n) for n in sorted(os.listdir(root)) if n.endswith((".cuh", ".h"))]
    return out


def _sources_digest():
    """What a per-system plugin depends on: every header (the .cu files of the core library do not enter)."""
    return _digest(_headers())


def _extra_flags():
    extra = os.environ.get("AIQMC_EXTRA_FLAGS", "").split()
    if os.environ.get("AIQMC_SYSTEMS"):
        os.makedirs(BUILD, exist_ok=True)            # nvcc splits -D values at commas: use a pre-include
        hdr = os.path.join(BUILD, "systems_override.h")
        with open(hdr, "w") as f:
            f.write("#define AIQMC_SYSTEM_LIST " + " ".join(f"X({n},{a})" for n, a in systems()) + "\n")
        extra += ["-include", hdr]
    return extra


def plugin_path(n: int, a: int) -> str:
    """$AIQMC_PLUGIN_DIR (the directory the loader searches first) also redirects where plugins are built: A/B variants."""
    return os.path.join(os.environ.get("AIQMC_PLUGIN_DIR") or HERE, f"libaiqmc_sys_{n}_{a}.so")


def _inst_source(n: int, a: int) -> str:
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(BUILD, f"inst_{n}_{a}.cu")
    with open(src, "w") as f:
        f.write('#include "' + os.path.join(CSRC, 'engine_impl.cuh') + '"\n'
                f'extern "C" const aiqmc::OpsTable* aiqmc_ops_{n}_{a}() {{ return aiqmc::Launch<{n}, {a}>::table(); }}\n')
    return src


def _link_plugin(n: int, a: int, obj: str) -> str:
    """One shared object per system, depending on the core library (g_launch_count / g_last_cuda_error live there)."""
    out = plugin_path(n, a)
    cmd = [NVCC] + ARCH + ["-shared", "-o", out, obj, "-L", HERE, "-l:" + os.path.basename(LIB), "-Xlinker", "-rpath",
                           "-Xlinker", "$ORIGIN", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link of {out} failed:\n{r.stdout}\n{r.stderr}")
    return out


def ensure_system(n: int, a: int, verbose: bool = False) -> str:
    """Builds the kernel plugin of one (n_elec, n_atoms) system if it is missing or stale (any n <= 32, a <= 16; the
    prebuilt set is csrc/dispatch.h) and makes the loaded library look again.  Needs nvcc at run time."""
    if not (2 <= n <= 32 and 1 <= a <= 16):
        raise ValueError(f"system size out of range: N={n}, A={a}")
    build()                                            # the core library (and the prebuilt set) first
    out = plugin_path(n, a)
    stamp = os.path.join(BUILD, f"inst_{n}_{a}.stamp")
    digest = _sources_digest()
    if os.path.exists(out) and os.path.exists(stamp) and open(stamp).read() == digest:
        return out
    obj = _compile(_inst_source(n, a), os.path.join(BUILD, f"inst_{n}_{a}.o"), os.path.join(BUILD, f"inst_{n}_{a}.log"))
    _link_plugin(n, a, obj)
    with open(stamp, "w") as f:
        f.write(digest)
    if verbose:
        print("built", out, file=sys.stderr)
    try:
        import ctypes
        ctypes.CDLL(LIB).aiqmc_rescan_systems()
    except OSError:
        pass
    return out


def _compile(src, obj, log):
    cmd = [NVCC] + ARCH + FLAGS + _extra_flags() + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def _fresh(stamp_file, digest, *outputs):
    return all(os.path.exists(o) for o in outputs) and os.path.exists(stamp_file) and open(stamp_file).read() == digest


def build(force: bool = False, verbose: bool = False) -> str:
    """Compiles what is stale: a per-system plugin depends on the headers only, a core object on the headers and its own
    .cu file -- editing abi.cu / rng.cu / population.cu does not recompile the ~8 MB per-system translation units."""
    os.makedirs(BUILD, exist_ok=True)
    hdr = _sources_digest()
    jobs = []                                     # (kind, key, src, obj, log, stamp, digest)
    for n, a in systems():
        stamp = os.path.join(BUILD, f"inst_{n}_{a}.stamp")
        obj = os.path.join(BUILD, f"inst_{n}_{a}.o")
        if force or not _fresh(stamp, hdr, obj, plugin_path(n, a)):
            jobs.append(("sys", (n, a), _inst_source(n, a), obj, os.path.join(BUILD, f"inst_{n}_{a}.log"), stamp, hdr))
    core_objs, core_stale = [], False
    for name in ("abi", "wsizes", "gto", "rng", "population"):
        src, obj = os.path.join(CSRC, name + ".cu"), os.path.join(BUILD, name + ".o")
        stamp, dg = os.path.join(BUILD, name + ".stamp"), _digest(_headers() + [src])
        core_objs.append(obj)
        if force or not _fresh(stamp, dg, obj):
            jobs.append(("core", name, src, obj, os.path.join(BUILD, name + ".log"), stamp, dg))
            core_stale = True
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(lambda j: _compile(j[2], j[3], j[4]), jobs))
    if core_stale or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + core_objs + ["-lcudart", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    for kind, key, src, obj, log, stamp, dg in jobs:
        if kind == "sys":
            _link_plugin(key[0], key[1], obj)         # one plugin per system, bound by the core on first use
        with open(stamp, "w") as f:
            f.write(dg)
    if verbose:
        print("built", LIB, f"({len(jobs)} translation units recompiled)", file=sys.stderr)
    build_xla_ffi(verbose)
    return LIB


def build_xla_ffi(verbose: bool = False):
    """csrc/xla_ffi.cc -> libaiqmc_b200_xla.so, only where the XLA FFI headers exist (jax.ffi.include_dir())."""
    try:
        import jax.ffi
        inc = jax.ffi.include_dir()
    except Exception:
        return None                                   # no jax in this image: the shim is only type-checked (tests/)
    out = os.path.join(HERE, "libaiqmc_b200_xla.so")
    cmd = [NVCC, "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-I", inc, os.path.join(CSRC, "xla_ffi.cc"),
           "-o", out, "-L", HERE, "-l:" + os.path.basename(LIB), "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"xla_ffi.cc failed to build against {inc}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print("built", out, file=sys.stderr)
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
