#!/usr/bin/env python
"""Randomised parity sweep for the small-scene production layout (4-wide records, tight boxes, tie rule) against the
oracle: scenes at very different scales, radius ratios and camera distances.  usage (GPU box): python tools/tight_fuzz.py [n]
FUZZ_SIZES=1025,1500,2500,4000 sweeps the sizes just above the 4-wide shared-memory layout (child-pair records in shared
memory, then fp32 records in HBM/L2)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bevyray_b200 as bvr  # noqa: E402
from oracle import oracle  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 48
W, H = 176, 99
ctx = bvr.Context(0)
bad_total = 0
for it in range(N):
    rs = np.random.RandomState(1000 + it)
    sizes = [int(x) for x in os.environ.get("FUZZ_SIZES", "2,7,40,200,600,1024").split(",")]
    n = int(rs.choice(sizes))
    scale = float(10.0 ** rs.uniform(-2, 2))                  # world units per scene unit
    spread = float(rs.choice([2.0, 8.0, 30.0]))
    models = np.zeros(n, bvr.MODEL_DTYPE)
    models["position"] = (rs.uniform(-1, 1, (n, 3)) * spread * scale).astype(np.float32)
    rad = 10.0 ** rs.uniform(-1.5, 0.3, n)
    if rs.rand() < 0.5:                                       # a ground-like giant sphere, others resting near it
        rad[0] = 10.0 ** rs.uniform(2, 3.3)
        models["position"][0] = (0, -rad[0] * scale, 0)
        models["position"][1:, 1] = (rad[1:] * scale * rs.choice([1.0, 1.5], n - 1)).astype(np.float32)
    models["radius"] = (rad * scale).astype(np.float32)
    models["material_id"] = rs.randint(0, 4, n)
    mats = np.zeros(4, bvr.MATERIAL_DTYPE)
    mats["base_color"] = rs.uniform(0.2, 0.95, (4, 3)).astype(np.float32)
    mats["metallic"] = [0.0, 1.0, 0.0, 0.4]
    mats["roughness"] = [0.5, 0.1, 0.0, 0.6]
    mats["ior"] = 1.5
    mats["specular_transmission"] = [0.0, 0.0, 1.0, 0.3]
    nodes = bvr.build_ploc(models)
    dist = float(spread * scale * 10.0 ** rs.uniform(-0.3, 1.6))
    direction = rs.normal(size=3); direction[1] = abs(direction[1]) * 0.5 + 0.05; direction /= np.linalg.norm(direction)
    cam = bvr.make_camera(position=tuple(direction * dist), target=(0, 0, 0), fov=float(rs.uniform(0.2, 1.2)), aspect=W / H,
                          near=0.1 * scale, far=1000.0 * scale * 50, sample_count=4, bounces=8)
    win = bvr.make_window(float(rs.rand()), H)
    ctx.upload_scene(models, mats, nodes)
    got = ctx.render(cam, 3, win, bvr.make_options(W, kernel=int(os.environ.get("FUZZ_KERNEL", "1")), traversal=int(os.environ.get("FUZZ_TRAVERSAL", "0"))))
    rays = ctx.stats()["rays"]
    want, cnt = oracle.render(models, mats, nodes, cam, bvr.make_level(3), win, W)
    bad = sum(int((np.ascontiguousarray(got[k]).view(np.uint32) != np.ascontiguousarray(want[k]).view(np.uint32)).sum()) for k in want)
    hits = float((want["primary_id"] != 0xFFFFFFFF).mean())
    bad_total += bad + (rays != cnt["rays"])
    print(f"{it:3d} n={n:5d} scale={scale:9.3g} spread={spread:5.1f} cam_dist={dist:10.3g} hit_frac={hits:.2f} rays={rays} "
          f"{'OK' if bad == 0 and rays == cnt['rays'] else 'MISMATCH words=%d oracle_rays=%d' % (bad, cnt['rays'])}")
print("TOTAL mismatching scenes/words:", bad_total)
