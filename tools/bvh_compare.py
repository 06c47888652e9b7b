#!/usr/bin/env python
"""Compares the trees of the three builders on the benchmark scenes: surface-area-heuristic cost, depth, and whether the
GPU PLOC tree is the host PLOC tree (same set of node boxes).  usage (GPU box): python tools/bvh_compare.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bevyray_b200 as bvr  # noqa: E402


def sah(nodes):
    d = nodes["bounds_max"].astype(np.float64) - nodes["bounds_min"].astype(np.float64)
    a = d[:, 0] * d[:, 1] + d[:, 1] * d[:, 2] + d[:, 2] * d[:, 0]
    return float(a[nodes["model_count"] == 0].sum() / a[0])


def boxes(nodes):
    b = np.concatenate([nodes["bounds_min"], nodes["bounds_max"]], axis=1)
    return b[np.lexsort(b.T[::-1])]


ctx = bvr.Context(0)
for name, scene in (("rtiow", bvr.Scene.rtiow(1)), ("random 10k", bvr.Scene.random(11, 10000, 43.0, 0.05, 0.25)),
                    ("random 2^20", bvr.Scene.random(7, 1 << 20, 200.0, 0.05, 0.25))):
    host = scene.nodes
    os.environ.pop("BVR_GPU_LBVH", None)
    ctx.reload_tuning()
    ploc = ctx.upload_scene_gpu_bvh(scene.models, scene.materials, want_nodes=True)
    t_ploc = ctx.stats()["last_upload_ms"]
    os.environ["BVR_GPU_LBVH"] = "1"
    ctx.reload_tuning()
    lbvh = ctx.upload_scene_gpu_bvh(scene.models, scene.materials, want_nodes=True)
    t_lbvh = ctx.stats()["last_upload_ms"]
    os.environ.pop("BVR_GPU_LBVH", None)
    ctx.reload_tuning()
    d = [bvr.traversal_ranks(t, len(scene.models))[1] for t in (host, ploc, lbvh)]
    same = np.array_equal(boxes(host), boxes(ploc))
    print(f"{name:12s} SAH host-PLOC {sah(host):9.3f} GPU-PLOC {sah(ploc):9.3f} GPU-LBVH {sah(lbvh):9.3f} | depth {d} | "
          f"GPU PLOC boxes == host PLOC boxes: {same} | upload+build ms PLOC {t_ploc:.3f} LBVH {t_lbvh:.3f}")
