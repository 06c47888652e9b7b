#!/usr/bin/env python
"""Generates tests/golden/*.npz: small scenes (reference buffer layout), uniforms, and the CPU oracle's output
planes.  The reference itself cannot run in this image (no Rust toolchain / Vulkan), so these are the
ORACLE's outputs — they pin the oracle against accidental change and give the CUDA path a fixture that
does not depend on the host scene generator or libm.  Run from the repo root:  python tests/golden/make_golden.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bevyray_b200 as bvr  # noqa: E402
from oracle import oracle  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def struct_bytes(s):
    return np.frombuffer(bytes(s), dtype=np.uint8).copy()


def case(name, scene, cam, level, seed, W, H, raster=False):
    win = bvr.make_window(seed, H)
    rgba = depth = None
    if raster:
        rs = np.random.RandomState(11)
        rgba = rs.rand(H, W, 4).astype(np.float32)
        depth = (rs.rand(H, W) * 0.05).astype(np.float32)
    planes, cnt = oracle.render(scene.models, scene.materials, scene.nodes, cam, bvr.make_level(level), win, W, rgba, depth)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        models=scene.models.view(np.uint8).reshape(-1, 32), materials=scene.materials.view(np.uint8).reshape(-1, 32),
        nodes=scene.nodes.view(np.uint8).reshape(-1, 48), camera=struct_bytes(cam), level=np.uint32(level),
        seed=np.float32(seed), width=np.uint32(W), height=np.uint32(H),
        raster_rgba=rgba if raster else np.zeros(0, np.float32), raster_depth=depth if raster else np.zeros(0, np.float32),
        rays=np.uint64(cnt["rays"]), **{"out_" + k: v for k, v in planes.items()})
    print(name, "rays", cnt["rays"], "trunc", cnt["stack_truncations"])


def main():
    rt = bvr.Scene.rtiow(1)
    case("rtiow_repo_cam", rt, bvr.make_camera(sample_count=2, bounces=4, aspect=64 / 36), 3, 0.37, 64, 36)
    case("rtiow_book_cam", rt, bvr.make_camera(position=(13, 2, 3), target=(0, 0, 0), fov=float(np.deg2rad(20)),
                                               aspect=64 / 36, sample_count=3, bounces=10), 3, 0.5, 64, 36)
    rnd = bvr.Scene.random(9, 200, 12.0, 0.1, 0.6)
    case("random200_composite", rnd, bvr.make_camera(position=(0, 0, 14), target=(0, 0, 0), aspect=48 / 32,
                                                     sample_count=2, bounces=3), 2, 0.81, 48, 32, raster=True)


if __name__ == "__main__":
    main()
