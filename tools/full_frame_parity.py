"""Whole C2 benchmark frame (1080p x N spp) against the oracle for several megakernel variants; lists the pixels that differ.
usage (on a GPU box): python tools/full_frame_parity.py [spp]"""
import sys, os, time, numpy as np
sys.path.insert(0, '.')
import bevyray_b200 as bvr, bench
from oracle import oracle
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 100
wl = bench.WORKLOADS['c2']; W, H = wl['width'], wl['height']
scene = bvr.Scene.rtiow(1)
cam = bench.make_cam(bvr, wl, spp); win = bvr.make_window(0.37, H)
t = time.time(); want, cnt = oracle.render(scene.models, scene.materials, scene.nodes, cam, bvr.make_level(3), win, W); print('oracle s', time.time() - t)
ctx = bvr.Context(0)
for env in ({}, {"BVR_NO_TIGHT": "1"}, {"BVR_NO_BVH4": "1"}, {"BVR_TIGHT_PAD": "200"}, {"BVR_TIGHT_PAD": "400"}):
    for k in ("BVR_NO_TIGHT", "BVR_NO_BVH4", "BVR_TIGHT_PAD"): os.environ.pop(k, None)
    os.environ.update(env)
    ctx.reload_tuning()
    ctx.upload_scene(scene.models, scene.materials, scene.nodes)
    got = ctx.render(cam, 3, win, bvr.make_options(W)); st = ctx.stats()
    bad = np.zeros((H, W), bool)
    for k in want:
        d = np.ascontiguousarray(got[k]).view(np.uint32) != np.ascontiguousarray(want[k]).view(np.uint32)
        bad |= d.any(axis=2) if d.ndim == 3 else d
    ys, xs = np.nonzero(bad)
    print(env, 'pixels differing', len(ys), list(zip(xs.tolist(), ys.tolist()))[:8], 'rays', st['rays'], cnt['rays'], 'ms', round(st['last_render_ms'], 3))
    for x, y in list(zip(xs.tolist(), ys.tolist()))[:3]:
        print('   ', (x, y), 'got', got['rgba'][y, x], got['rt_depth'][y, x], 'want', want['rgba'][y, x], want['rt_depth'][y, x])
