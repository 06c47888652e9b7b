#!/usr/bin/env python
"""Opcode histogram (warp-level executed instructions) of a kernel from an .ncu-rep, SASS view.
usage: tools/ncu_sass_mix.py prof.ncu-rep [top-N]"""
import csv, io, subprocess, sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
agg = defaultdict(lambda: [0, 0])
for r in rows:
    if not r: continue
    if r[0] in ("Address", "#"):
        hdr = r; continue
    if hdr is None or len(r) != len(hdr): 
        if r and r[0].startswith("Address"): hdr = r
        continue
    d = dict(zip(hdr, r))
    src = d.get("Source", "")
    try:
        inst = int(float(d.get("Instructions Executed", "0") or 0)); tinst = int(float(d.get("Thread Instructions Executed", "0") or 0))
    except ValueError:
        continue
    toks = src.replace("@!", "@").split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.split(".")[0]
    agg[op][0] += inst; agg[op][1] += tinst
tot = sum(v[0] for v in agg.values()) or 1
print(f"total warp inst {tot:.3e}")
for op, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{op:12s} {100*v[0]/tot:6.2f}%  avg active {v[1]/max(v[0],1):5.1f}")
