#!/usr/bin/env python
"""Opcode histogram per kernel of a built library (cuobjdump -sass): python tools/sass_histogram.py lib.so [kernel-regex] [top]
Committed under profiles/ as the SASS evidence for each instantiation (UBLKCP = TMA bulk copy, SYNCS = mbarrier,
DFMA/DMUL/DADD = FP64 pipe, MUFU.RCP64H / RSQ64H = FP64 seeds, REDUX = warp reduce, UCGABAR* = cluster barrier)."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 14
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name, hist = None, collections.Counter()
out = []


def flush():
    if name and hist and (pat is None or pat.search(name)):
        tot = sum(hist.values())
        fp64 = sum(v for k, v in hist.items() if k.split(".")[0] in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
        notable = {k: v for k, v in hist.items() if k.split(".")[0] in ("UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "REDUX", "UCGABAR_ARV", "UCGABAR_WAIT", "LDL", "STL", "DMMA", "HMMA", "MUFU", "RED", "ATOMG", "SHFL", "MAPA")}
        out.append((name, tot, fp64, hist.most_common(top), notable))


for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        flush()
        name, hist = m.group(1), collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m:
        op = m.group(1)
        base = op.split(".")[0]
        key = op if base in ("MUFU", "UBLKCP", "SYNCS", "REDUX", "LDS", "STS", "LDL", "STL", "LDG", "STG", "LDC", "RED", "ATOMG", "DMMA", "UCGABAR_ARV", "UCGABAR_WAIT") else base
        hist[key] += 1
flush()
demangle = subprocess.run(["c++filt"] + [o[0] for o in out], capture_output=True, text=True).stdout.splitlines() if out else []
for (nm, tot, fp64, common, notable), dn in zip(out, demangle):
    short = re.sub(r"\(.*", "", dn)
    print(f"{short}: {tot} SASS instructions, {fp64} FP64-pipe ({100.0 * fp64 / tot:.1f} %)")
    print("    " + "  ".join(f"{k} {v}" for k, v in common))
    if notable:
        print("    notable: " + "  ".join(f"{k} {v}" for k, v in sorted(notable.items())))
