#!/usr/bin/env python
"""Small render through every kernel variant, meant to run under compute-sanitizer
(memcheck / racecheck / synccheck):  compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bevyray_b200 as bvr  # noqa: E402

scene = bvr.Scene.rtiow(1)
W, H = 64, 36
cam = bvr.make_camera(position=(13, 2, 3), target=(0, 0, 0), fov=float(np.deg2rad(20)), aspect=W / H, sample_count=2, bounces=4)
win = bvr.make_window(0.37, H)
ctx = bvr.Context(0)
ctx.upload_scene(scene.models, scene.materials, scene.nodes)
ref = None
for kernel, traversal, name in [(1, 0, "megakernel"), (1, 1, "reference-order"), (2, 0, "wavefront"), (3, 0, "cta-wavefront")]:
    out = ctx.render(cam, 3, win, bvr.make_options(W, kernel=kernel, traversal=traversal))
    if ref is None:
        ref = out
    same = all(np.array_equal(out[k].view(np.uint32), ref[k].view(np.uint32)) for k in ref)
    print(name, "rays", ctx.stats()["rays"], "identical" if same else "DIFFERENT")
nodes = ctx.upload_scene_gpu_bvh(scene.models, scene.materials, want_nodes=True)
out = ctx.render(cam, 3, win, bvr.make_options(W))
print("gpu-bvh", bvr.validate_bvh(nodes, scene.models), all(np.array_equal(out[k].view(np.uint32), ref[k].view(np.uint32)) for k in ref))
# megakernel variants chosen through the environment (tests/test_gpu_layouts.py)
ctx.upload_scene(scene.models, scene.materials, scene.nodes)
# v5's rings of entry ids are synchronised by flags and counters between warps (megakernel_v5.cu): racecheck, which
# only knows barriers, reports every such hand-over as a hazard -> SANITIZE_SKIP_V5=1 for the racecheck run
variants = [{"BVR_NO_TIGHT": "1"}, {"BVR_NO_BVH4": "1"}, {"BVR_MK_VARIANT": "4"}, {"BVR_MK_VARIANT": "5"}]
if os.environ.get("SANITIZE_SKIP_V5"):
    variants = [v for v in variants if v.get("BVR_MK_VARIANT") != "5"]
for env in variants:
    os.environ.update(env)
    out = ctx.render(cam, 3, win, bvr.make_options(W, kernel=1))
    for k in env:
        del os.environ[k]
    print(env, all(np.array_equal(out[k].view(np.uint32), ref[k].view(np.uint32)) for k in ref))
# scenes walked in HBM/L2: 4-wide 16-bit records, 2-wide 16-bit records, fp32 records
big = bvr.Scene.random(7, 6000, 36.0, 0.05, 0.25)
bref = None
for env in ({}, {"BVR_NO_BVH4": "1"}, {"BVR_NO_Q16": "1"}):
    os.environ.update(env)
    ctx.upload_scene(big.models, big.materials, big.nodes)
    out = ctx.render(cam, 3, win, bvr.make_options(W))
    for k in env:
        del os.environ[k]
    bref = bref or out
    print("big", env, all(np.array_equal(out[k].view(np.uint32), bref[k].view(np.uint32)) for k in bref))
# structural validation on the GPU (scene_validate.cu), forced on the small scene
os.environ["BVR_GPU_VALIDATE"] = "1"
ctx.upload_scene(scene.models, scene.materials, scene.nodes)
out = ctx.render(cam, 3, win, bvr.make_options(W, kernel=1))
del os.environ["BVR_GPU_VALIDATE"]
print("gpu-validate", all(np.array_equal(out[k].view(np.uint32), ref[k].view(np.uint32)) for k in ref))
print("done")
