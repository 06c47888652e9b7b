#!/usr/bin/env python
"""Randomised parity sweep for the big-scene production layout (4-wide records on the 16-bit grid, 4-byte stack
entries, GPU-side validation above 32 k nodes) against the oracle: sizes 2 k - 120 k spheres, four decades of scale,
clustered and uniform placements, host PLOC and GPU LBVH trees.  usage (GPU box): python tools/big_fuzz.py [n]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bevyray_b200 as bvr  # noqa: E402
from oracle import oracle  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
W, H = 160, 90
ctx = bvr.Context(0)
bad_total = 0
for it in range(N):
    rs = np.random.RandomState(500 + it)
    n = int(rs.choice([2000, 6000, 20000, 50000, 120000]))
    scale = float(10.0 ** rs.uniform(-2, 2))
    side = float((n / 0.125) ** (1 / 3.0))                     # C4's density
    models = np.zeros(n, bvr.MODEL_DTYPE)
    if rs.rand() < 0.5:
        pos = rs.uniform(-0.5, 0.5, (n, 3)) * side
    else:                                                      # clustered: a few dense blobs and empty space between
        centres = rs.uniform(-0.5, 0.5, (8, 3)) * side
        pos = centres[rs.randint(0, 8, n)] + rs.normal(size=(n, 3)) * side * 0.04
    models["position"] = (pos * scale).astype(np.float32)
    models["radius"] = (rs.uniform(0.05, 0.25, n) * scale).astype(np.float32)
    models["material_id"] = rs.randint(0, 4, n)
    mats = np.zeros(4, bvr.MATERIAL_DTYPE)
    mats["base_color"] = rs.uniform(0.2, 0.95, (4, 3)).astype(np.float32)
    mats["metallic"] = [0.0, 1.0, 0.0, 0.4]
    mats["roughness"] = [0.5, 0.1, 0.0, 0.6]
    mats["ior"] = 1.5
    mats["specular_transmission"] = [0.0, 0.0, 1.0, 0.3]
    gpu_bvh = bool(rs.rand() < 0.4)
    if gpu_bvh:
        nodes = ctx.upload_scene_gpu_bvh(models, mats, want_nodes=True)
    else:
        nodes = bvr.build_ploc(models)
        ctx.upload_scene(models, mats, nodes)
    d = rs.normal(size=3); d /= np.linalg.norm(d)
    dist = side * scale * float(rs.uniform(0.2, 1.5))
    cam = bvr.make_camera(position=tuple(d * dist), target=(0, 0, 0), fov=float(rs.uniform(0.3, 1.0)), aspect=W / H,
                          near=0.1 * scale, far=1e5 * scale, sample_count=2, bounces=8)
    win = bvr.make_window(float(rs.rand()), H)
    got = ctx.render(cam, 3, win, bvr.make_options(W, kernel=1))
    rays = ctx.stats()["rays"]
    want, cnt = oracle.render(models, mats, nodes, cam, bvr.make_level(3), win, W)
    bad = sum(int((np.ascontiguousarray(got[k]).view(np.uint32) != np.ascontiguousarray(want[k]).view(np.uint32)).sum()) for k in want)
    ok = bad == 0 and rays == cnt["rays"] and cnt["stack_truncations"] == 0
    bad_total += 0 if ok else 1
    print(f"{it:3d} n={n:6d} scale={scale:8.3g} {'gpu-lbvh ' if gpu_bvh else 'host-ploc'} hit={float((want['primary_id'] != 0xFFFFFFFF).mean()):.2f} "
          f"rays={rays} trunc={cnt['stack_truncations']} {'OK' if ok else 'MISMATCH words=%d oracle_rays=%d' % (bad, cnt['rays'])}", flush=True)
print("TOTAL mismatching scenes:", bad_total)
