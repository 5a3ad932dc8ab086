"""A0 micro-benchmark: C cc-pVDZ (13 AOs) value + gradient + Laplacian at walkers x 6 points.
    python tools/gto_bench.py [--points 393216,1572864]
Prints one line per size: ms (min of 7, 256 MiB L2 flush between), GB/s (24 B in + 65 doubles out per point).
GTO_BENCH_CLEAN_L2=1 reads the flush buffer back after writing it, so the kernel is not charged the write-back of up to
126 MB of dirty flush lines; GTO_BENCH_MEMSET=1 adds the time of a plain memset of the same number of bytes."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aiqmc_b200                                     # noqa: E402
from aiqmc_b200 import workloads as W                 # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", default="393216,1572864")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    basis = aiqmc_b200.GaussianBasis.from_nwchem(W.C_CC_PVDZ, np.zeros((1, 3)), device=dev)
    for n in [int(v) for v in args.points.split(",")]:
        pts = torch.randn((n, 3), dtype=torch.float64, device=dev)
        ref = basis.eval(pts)
        ts = []
        for _ in range(7):
            flush.zero_()
            if os.environ.get("GTO_BENCH_CLEAN_L2"):
                flush.sum()                           # leaves L2 full of CLEAN lines (no write-back charged to the kernel)
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(); out = basis.eval(pts); a1.record(); torch.cuda.synchronize()
            ts.append(a0.elapsed_time(a1))
        ms = min(ts)
        nbytes = n * (24 + 65 * 8)
        chk = float(sum(float(o.double().abs().sum()) for o in (out if isinstance(out, (tuple, list)) else [out])))
        if os.environ.get("GTO_BENCH_MEMSET"):
            buf = torch.empty(nbytes // 8, dtype=torch.float64, device=dev)
            t2 = []
            for _ in range(7):
                flush.zero_()
                if os.environ.get("GTO_BENCH_CLEAN_L2"):
                    flush.sum()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record(); buf.zero_(); a1.record(); torch.cuda.synchronize()
                t2.append(a0.elapsed_time(a1))
            print(f"  memset of the same bytes: ms={min(t2):.4f} GB/s={nbytes / min(t2) / 1e6:.1f}")
        print(f"points={n} ms={ms:.4f} GB/s={nbytes / ms / 1e6:.1f} checksum={chk:.10e}")


if __name__ == "__main__":
    main()
