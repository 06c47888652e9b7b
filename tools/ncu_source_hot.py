#!/usr/bin/env python
"""Aggregate the ncu source page by CUDA source line: warp instructions, thread instructions, average
active threads.  usage: tools/ncu_source_hot.py prof.ncu-rep [top-N]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    # the CSV has blocks: per file a header ("File Path"), then rows
    rows = list(csv.reader(io.StringIO(out)))
    agg = defaultdict(lambda: [0, 0, 0, ""])   # (file,line) -> inst, thread inst, samples, text
    cur_file = None
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        try:
            inst = int(float(d.get("Instructions Executed", "0") or 0))
            tinst = int(float(d.get("Thread Instructions Executed", "0") or 0))
            samp = int(float(d.get("# Samples", "0") or 0))
        except ValueError:
            continue
        key = (cur_file, d["Line No"])
        a = agg[key]
        a[0] += inst
        a[1] += tinst
        a[2] += samp
        a[3] = r[1][:110]
    tot_i = sum(a[0] for a in agg.values())
    tot_t = sum(a[1] for a in agg.values())
    print(f"total warp inst {tot_i:.3e}, thread inst {tot_t:.3e}, avg active {tot_t / max(tot_i, 1):.2f}")
    print(f"{'file:line':34s} {'%inst':>6s} {'act':>5s} {'%samp':>6s}  source")
    tot_s = sum(a[2] for a in agg.values()) or 1
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{f + ':' + ln:34s} {100 * a[0] / tot_i:6.2f} {a[1] / max(a[0], 1):5.1f} {100 * a[2] / tot_s:6.2f}  {a[3]}")


if __name__ == "__main__":
    main()
