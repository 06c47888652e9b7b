"""Whole 1080p frames of three more 2^20-sphere scenes (other seeds and cameras, host PLOC and GPU LBVH trees) against the
oracle at 1 spp.  usage (GPU box): python tools/c4_full_frames.py"""
import sys, numpy as np
sys.path.insert(0, '.')
import bevyray_b200 as bvr
from oracle import oracle
W, H = 1920, 1080
ctx = bvr.Context(0)
for seed, campos, fov in ((8, (0, 0, 130), np.pi / 4), (9, (60, 40, -90), 0.5), (10, (5, 3, 2), 1.0)):
    scene = bvr.Scene.random(seed, 1 << 20, 200.0, 0.05, 0.25)
    cam = bvr.make_camera(position=campos, target=(0, 0, 0), fov=float(fov), aspect=W / H, sample_count=1, bounces=10)
    win = bvr.make_window(0.13 * seed % 1.0, H)
    for gpu_bvh in (False, True):
        if gpu_bvh:
            nodes = ctx.upload_scene_gpu_bvh(scene.models, scene.materials, want_nodes=True)
        else:
            nodes = scene.nodes
            ctx.upload_scene(scene.models, scene.materials, nodes)
        got = ctx.render(cam, 3, win, bvr.make_options(W)); st = ctx.stats()
        want, cnt = oracle.render(scene.models, scene.materials, nodes, cam, bvr.make_level(3), win, W)
        bad = sum(int((np.ascontiguousarray(got[k]).view(np.uint32) != np.ascontiguousarray(want[k]).view(np.uint32)).sum()) for k in want)
        print('seed', seed, 'cam', campos, 'gpu_bvh', gpu_bvh, 'bad words', bad, 'rays', st['rays'], cnt['rays'], 'trunc', cnt['stack_truncations'], 'ms', round(st['last_render_ms'], 2), flush=True)
