#!/usr/bin/env python
"""How far does the image move when a WGSL toolchain resolves an open convention differently from the oracle?

The reference's shader cannot run in this image, so the conventions DESIGN.md §3 fixes (no FMA contraction, pow by
multiplication, normalize by division, IEEE minNum/maxNum, short-circuit `||`) cannot be checked against naga + a driver.
This tool rebuilds the oracle with each plausible alternative (macros in oracle/bvr_oracle.cpp, builds under
oracle/_variants/, git-ignored) and compares against the strict oracle on C1 (1280x720, 1 spp, 4 bounces) and on the
C2 frame at 8 spp: pixels whose radiance changed, primary-hit id mismatches, per-channel RMSE and PSNR.  CPU only.
usage: python tools/oracle_sensitivity.py > profiles/r02_oracle_sensitivity.txt"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import oracle  # noqa: E402

SRC = os.path.join(ROOT, "oracle", "bvr_oracle.cpp")
OUT = os.path.join(ROOT, "oracle", "_variants")
BASE = ["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-fno-fast-math", "-fno-unsafe-math-optimizations"]
VARIANTS = [
    ("strict (the oracle)", ["-ffp-contract=off"]),
    ("FMA contraction on", ["-ffp-contract=fast", "-mfma"]),
    ("pow(x,5) = exp2(5 log2 x)", ["-ffp-contract=off", "-DBVRO_VAR_POW_EXP2LOG2"]),
    ("normalize = v * (1/sqrt)", ["-ffp-contract=off", "-DBVRO_VAR_RSQRT"]),
    ("NaN-propagating min/max", ["-ffp-contract=off", "-DBVRO_VAR_NAN_MINMAX"]),
    ("`||` evaluates both sides", ["-ffp-contract=off", "-DBVRO_VAR_EAGER_OR"]),
    ("all of the above", ["-ffp-contract=fast", "-mfma", "-DBVRO_VAR_POW_EXP2LOG2", "-DBVRO_VAR_RSQRT", "-DBVRO_VAR_NAN_MINMAX",
                          "-DBVRO_VAR_EAGER_OR"]),
]
CASES = [("C1 1280x720 1spp 4 bounces", "c1", None), ("C2 1920x1080 8spp 10 bounces", "c2", 8)]
if len(sys.argv) > 1:      # e.g. `oracle_sensitivity.py 100`: the C2 frame at its full 100 spp only (minutes of CPU time)
    CASES = [(f"C2 1920x1080 {int(a)}spp 10 bounces", "c2", int(a)) for a in sys.argv[1:]]


def build(i, flags):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, f"variant{i}.so")
    subprocess.check_call(BASE + flags + ["-o", path, SRC])
    lib = C.CDLL(path)
    lib.bvro_render.restype = oracle.lib.bvro_render.restype
    lib.bvro_render.argtypes = oracle.lib.bvro_render.argtypes
    return lib


def render(lib, key, spp):
    wl = bench.WORKLOADS[key]
    models, materials, nodes, cam = bench.fixture_scene(oracle, key, wl)
    if spp:
        cam.sample_count = spp
    saved, oracle.lib = oracle.lib, lib
    try:
        planes, cnt = oracle.render(models, materials, nodes, cam, oracle.make_level(3), oracle.make_window(bench.BASE_SEED, wl["height"]),
                                    wl["width"])
    finally:
        oracle.lib = saved
    return planes, cnt


def main():
    libs = [build(i, flags) for i, (_, flags) in enumerate(VARIANTS)]
    print("Sensitivity of the image to conventions WGSL leaves open (tools/oracle_sensitivity.py; strict = the oracle).")
    print("pixels = texels whose fp32 radiance differs in any bit; id = primary-hit id mismatches; RMSE / PSNR per channel over RGB")
    print("(peak 1.0) against the strict oracle.  The north-star bar is RMSE <= 1e-3 and PSNR >= 50 dB at equal spp.\n")
    for title, key, spp in CASES:
        ref, cnt0 = render(libs[0], key, spp)
        n = ref["rgba"].shape[0] * ref["rgba"].shape[1]
        print(f"{title}: {n} pixels, {cnt0['rays']} rays")
        print(f"  {'variant':32s} {'pixels changed':>16s} {'id mismatches':>14s} {'rays':>12s} {'RMSE':>10s} {'PSNR dB':>9s}")
        for (name, _), lib in zip(VARIANTS[1:], libs[1:]):
            got, cnt = render(lib, key, spp)
            diff = (got["rgba"].view(np.uint32) != ref["rgba"].view(np.uint32)).any(axis=2)
            ids = int((got["primary_id"] != ref["primary_id"]).sum())
            err = (got["rgba"][..., :3].astype(np.float64) - ref["rgba"][..., :3].astype(np.float64))
            rmse = float(np.sqrt((err ** 2).mean()))
            psnr = float("inf") if rmse == 0 else 20 * np.log10(1.0 / rmse)
            print(f"  {name:32s} {int(diff.sum()):9d} ({100 * diff.mean():5.2f}%) {ids:14d} {cnt['rays'] - cnt0['rays']:+12d} {rmse:10.2e} {psnr:9.1f}")
        print()


if __name__ == "__main__":
    main()
