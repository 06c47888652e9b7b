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
for env in ({}, {"BVR_NO_BVH4": "1"}, {"BVR_NO_Q16": "1"}, {"BVR_W4_LEAN": "0"}, {"BVR_W4_LEAN": "0", "BVR_TOP_RECORDS": "300"},
            {"BVR_NO_TOP": "1"}, {"BVR_HOT_RECORDS": "40"}):
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
# pixel-queue order fed back from the previous frame (tile_order.cu), uneven sample counts, weighted stores, slot sum
os.environ["BVR_TILE_ORDER"] = "2"
ctx.reload_tuning()
ctx.upload_scene(scene.models, scene.materials, scene.nodes)
for frame in range(3):
    out = ctx.render(cam, 3, win, bvr.make_options(W, kernel=1))
    print("tile order, frame", frame, all(np.array_equal(out[k].view(np.uint32), ref[k].view(np.uint32)) for k in ref))
del os.environ["BVR_TILE_ORDER"]
ctx.reload_tuning()
o = bvr.make_options(W, output_weight=0.125)
o.flags |= bvr.capi.render_extra_sample_bits(3, 1, 2)
out = ctx.render(cam, 3, win, o)
print("extra sample + weight", ctx.stats()["paths"])
import torch  # noqa: E402
slots = torch.rand((3, 4 * 256), device="cuda")
dst = torch.empty(4 * 256, device="cuda")
ctx.sum_slots_device(slots.data_ptr(), 4 * 256, 3, 0b101, dst.data_ptr(), 4 * 256)
ctx.sync()
print("sum slots", bool(torch.equal(dst, slots[0] + slots[2])))
print("done")
