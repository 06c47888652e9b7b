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
variants = [{"BVR_NO_TIGHT": "1"}, {"BVR_NO_BVH4": "1"}, {"BVR_MK_V1": "1"}, {"BVR_MK_THREADS": "512"}]
for env in variants:
    os.environ.update(env)
    ctx.reload_tuning()
    ctx.upload_scene(scene.models, scene.materials, scene.nodes)
    out = ctx.render(cam, 3, win, bvr.make_options(W, kernel=1))
    for k in env:
        del os.environ[k]
    ctx.reload_tuning()
    print(env, all(np.array_equal(out[k].view(np.uint32), ref[k].view(np.uint32)) for k in ref))
# scenes walked in HBM/L2: 4-wide 16-bit records, 2-wide 16-bit records, fp32 records
big = bvr.Scene.random(7, 6000, 36.0, 0.05, 0.25)
bref = None
for env in ({}, {"BVR_NO_BVH4": "1"}, {"BVR_NO_Q16": "1"}):
    os.environ.update(env)
    ctx.reload_tuning()
    ctx.upload_scene(big.models, big.materials, big.nodes)
    out = ctx.render(cam, 3, win, bvr.make_options(W))
    for k in env:
        del os.environ[k]
    ctx.reload_tuning()
    bref = bref or out
    print("big", env, all(np.array_equal(out[k].view(np.uint32), bref[k].view(np.uint32)) for k in bref))
# structural validation on the GPU (scene_validate.cu), forced on the small scene
os.environ["BVR_GPU_VALIDATE"] = "1"
ctx.reload_tuning()
ctx.upload_scene(scene.models, scene.materials, scene.nodes)
out = ctx.render(cam, 3, win, bvr.make_options(W, kernel=1))
del os.environ["BVR_GPU_VALIDATE"]
print("gpu-validate", all(np.array_equal(out[k].view(np.uint32), ref[k].view(np.uint32)) for k in ref))
print("done")
