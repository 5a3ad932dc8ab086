"""Builds libaiqmc_b200.so (sm_100a) in-tree: one translation unit per (n_elec, n_atoms)
instantiation, compiled in parallel with nvcc, linked into a plain CUDA-runtime shared
library that exports the C ABI of include/aiqmc_b200.h."""
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


def _sources_digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for name in sorted(os.listdir(root)):
            p = os.path.join(root, name)
            if os.path.isfile(p) and name.endswith((".cu", ".cuh", ".h")):
                h.update(name.encode())
                h.update(open(p, "rb").read())
    h.update(" ".join(FLAGS + ARCH + _extra_flags()).encode())
    return h.hexdigest()


def _extra_flags():
    extra = os.environ.get("AIQMC_EXTRA_FLAGS", "").split()
    if os.environ.get("AIQMC_SYSTEMS"):
        os.makedirs(BUILD, exist_ok=True)            # nvcc splits -D values at commas: use a pre-include
        hdr = os.path.join(BUILD, "systems_override.h")
        with open(hdr, "w") as f:
            f.write("#define AIQMC_SYSTEM_LIST " + " ".join(f"X({n},{a})" for n, a in systems()) + "\n")
        extra += ["-include", hdr]
    return extra


def _compile(src, obj, log):
    cmd = [NVCC] + ARCH + FLAGS + _extra_flags() + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    digest = _sources_digest()
    stamp = os.path.join(BUILD, "stamp.txt")
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    jobs = []
    for n, a in systems():
        src = os.path.join(BUILD, f"inst_{n}_{a}.cu")
        with open(src, "w") as f:
            f.write('#include "' + os.path.join(CSRC, 'engine_impl.cuh') + '"\n'
                    f'extern "C" const aiqmc::OpsTable* aiqmc_ops_{n}_{a}() {{ return aiqmc::Launch<{n}, {a}>::table(); }}\n')
        jobs.append((src, os.path.join(BUILD, f"inst_{n}_{a}.o"), os.path.join(BUILD, f"inst_{n}_{a}.log")))
    for name in ("abi", "wsizes", "gto", "rng", "population"):
        jobs.append((os.path.join(CSRC, name + ".cu"), os.path.join(BUILD, name + ".o"),
                     os.path.join(BUILD, name + ".log")))
    with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(lambda j: _compile(*j), jobs))
    cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    if verbose:
        print("built", LIB, file=sys.stderr)
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
