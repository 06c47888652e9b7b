"""Odd image sizes (1x1 ... 1000x1), sample/bounce counts and tile-shard geometries against the oracle (RTIOW scene).
usage (GPU box): python tools/size_shard_sweep.py"""
import sys, numpy as np
sys.path.insert(0, '.')
import bevyray_b200 as bvr
from oracle import oracle
from bevyray_b200.distributed import shard_global_rows
scene = bvr.Scene.rtiow(1)
ctx = bvr.Context(0); ctx.upload_scene(scene.models, scene.materials, scene.nodes)
bad = 0
for (W, H) in [(1, 1), (3, 2), (8, 4), (9, 5), (31, 33), (257, 3), (5, 259), (1000, 1), (1, 1000), (640, 361)]:
    for spp, bounces in ((1, 0), (3, 4), (2, 11)):
        cam = bvr.make_camera(position=(13, 2, 3), target=(0, 0, 0), fov=0.35, aspect=W / H, sample_count=spp, bounces=bounces)
        win = bvr.make_window(0.77, H)
        got = ctx.render(cam, 3, win, bvr.make_options(W))
        want, cnt = oracle.render(scene.models, scene.materials, scene.nodes, cam, bvr.make_level(3), win, W)
        b = sum(int((np.ascontiguousarray(got[k]).view(np.uint32) != np.ascontiguousarray(want[k]).view(np.uint32)).sum()) for k in want) + (ctx.stats()['rays'] != cnt['rays'])
        # tile shards with odd strip heights reassemble to the same frame
        for count, strip in ((3, 1), (5, 7)):
            full = np.zeros_like(want['rgba'])
            for r in range(count):
                part = ctx.render(cam, 3, win, bvr.make_options(W, shard_index=r, shard_count=count, strip_rows=strip), want=("rgba",))
                rows = shard_global_rows(H, r, count, strip); valid = rows < H
                full[rows[valid]] = part['rgba'][valid]
            b += int((full.view(np.uint32) != want['rgba'].view(np.uint32)).sum())
        bad += b
        if b: print('MISMATCH', W, H, spp, bounces, b)
print('size/shard sweep mismatches:', bad)
