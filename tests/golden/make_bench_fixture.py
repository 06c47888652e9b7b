#!/usr/bin/env python
"""Generates tests/golden/bench_rtiow.npz: the scene and camera bytes of the benchmark workloads C1-C3 (BASELINE.json
configs[0..2]) in the reference's buffer layout, so that `bench.py --impl reference` (the CPU arm) needs nothing from the
product library.  tests/test_golden_oracle.py checks that the host layer still produces exactly these bytes.
Run from the repo root:  python tests/golden/make_bench_fixture.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import bevyray_b200 as bvr  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bench_rtiow.npz")


def main():
    sc = bvr.Scene.rtiow(bench.SCENE_SEED)
    cams = {}
    for key in ("c1", "c2", "c3"):
        cams["camera_" + key] = np.frombuffer(bytes(bench.make_cam(bvr, bench.WORKLOADS[key])), dtype=np.uint8).copy()
    np.savez_compressed(OUT, models=sc.models.view(np.uint8).reshape(-1, 32), materials=sc.materials.view(np.uint8).reshape(-1, 32),
                        nodes=sc.nodes.view(np.uint8).reshape(-1, 48), scene_seed=np.uint32(bench.SCENE_SEED), **cams)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(sc.models), "spheres")


if __name__ == "__main__":
    main()
