import ctypes as C
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    """Render cases (make_golden.py); bench_*.npz / nversion_*.npz are fixtures of another kind."""
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in names if not n.startswith(("bench_", "nversion_"))]


def load_golden(bvr, name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["models"] = np.ascontiguousarray(g["models"]).view(bvr.MODEL_DTYPE).reshape(-1)
    g["materials"] = np.ascontiguousarray(g["materials"]).view(bvr.MATERIAL_DTYPE).reshape(-1)
    g["nodes"] = np.ascontiguousarray(g["nodes"]).view(bvr.BVH_NODE_DTYPE).reshape(-1)
    cam = bvr.capi.BvrCamera()
    C.memmove(C.addressof(cam), g["camera"].tobytes(), 80)
    g["camera"] = cam
    g["level"], g["width"], g["height"] = int(g["level"]), int(g["width"]), int(g["height"])
    g["seed"] = float(g["seed"])
    if g["raster_rgba"].size == 0:
        g["raster_rgba"] = g["raster_depth"] = None
    return g
