#!/usr/bin/env python
"""Per-SASS-instruction listing of a kernel from an .ncu-rep with executed counts, for segmenting a kernel into
regions by hand.  usage: tools/ncu_sass_regions.py prof.ncu-rep [kernel-substring] > listing.txt
Columns: index, warp instructions executed, average active threads, stall samples, SASS text."""
import csv, io, subprocess, sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
i = 0
tot = 0
lines = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    try:
        inst = int(float(d["Instructions Executed"] or 0)); tinst = int(float(d["Thread Instructions Executed"] or 0))
        samp = int(float(d["# Samples"] or 0))
    except (ValueError, KeyError):
        continue
    lines.append((i, inst, tinst, samp, d["Source"].strip()))
    tot += inst
    i += 1
cum = 0
for i, inst, tinst, samp, src in lines:
    cum += inst
    print(f"{i:5d} {inst:12d} {tinst / max(inst, 1):5.1f} {samp:6d} {100 * cum / max(tot, 1):6.2f}%  {src}")
