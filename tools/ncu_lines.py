#!/usr/bin/env python
"""Stall samples / executed instructions per CUDA source line from `ncu --page source --csv --print-source cuda,sass`
(development aid): python tools/ncu_lines.py file.ncu-rep [top]"""
import collections, csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
sec = None; agg = collections.OrderedDict(); hdr = None
for r in csv.reader(io.StringIO(raw)):
    if not r: continue
    if r[0] == "File Path": sec = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; si = hdr.index("# Samples"); ei = hdr.index("Instructions Executed"); continue
    if hdr is None: continue
    try: ln = int(r[0]); s = int(float(r[si] or 0)); e = int(float(r[ei] or 0))
    except ValueError: continue
    a = agg.setdefault((sec, ln), [0, 0, r[1][:100]]); a[0] += s; a[1] += e
tot = sum(a[0] for a in agg.values()) or 1; tote = sum(a[1] for a in agg.values()) or 1
pf = collections.Counter(); pe = collections.Counter()
for (f, l), a in agg.items(): pf[f] += a[0]; pe[f] += a[1]
for f in pf: print(f"{f:24s} {100*pf[f]/tot:5.1f}% samples {100*pe[f]/tote:5.1f}% exec")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{f:18s} {l:4d}  smp {100*a[0]/tot:5.1f}%  exec {100*a[1]/tote:5.1f}%  {a[2].strip()}")
