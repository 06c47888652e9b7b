"""N-version check of the oracle: tests/nversion/wgsl_numpy.py is a second restatement of the reference's `fragment`
(data-parallel numpy, written from the WGSL text alone).  Both must produce the same bits on every golden scene — all
four planes, every pixel.  The reference's own shader cannot run in this image (no Rust toolchain, no Vulkan), so this
does not pin the oracle against the reference; it does remove the oracle as a single point of misreading."""
import numpy as np
import pytest

from golden_util import golden_names, load_golden
from nversion import wgsl_numpy

PLANES = ("rgba", "rt_depth", "primary_id", "primary_depth")


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("name", golden_names())
def test_numpy_restatement_matches_oracle_on_golden_scenes(bvr, oracle, name):
    g = load_golden(bvr, name)
    got = wgsl_numpy.render(g["models"], g["materials"], g["nodes"], g["camera"], g["level"], g["seed"], g["width"], g["height"],
                            g["raster_rgba"], g["raster_depth"])
    live, _ = oracle.render(g["models"], g["materials"], g["nodes"], g["camera"], bvr.make_level(g["level"]),
                            bvr.make_window(g["seed"], g["height"]), g["width"], g["raster_rgba"], g["raster_depth"])
    for k in PLANES:
        assert np.array_equal(bits(got[k]), bits(g["out_" + k])), (name, k, "vs frozen oracle output")
        assert np.array_equal(bits(got[k]), bits(live[k])), (name, k, "vs live oracle")


@pytest.mark.parametrize("level", [0, 1, 2, 3])
def test_numpy_restatement_matches_oracle_on_every_level(bvr, oracle, level):
    """Mixed materials (metal, glass with ior < 1 and > 1, fractional metallic / transmission), several spheres per leaf
    reached through a hand-built tree, 6 bounces, all four raytrace levels with a raster colour / depth."""
    rs = np.random.RandomState(21)
    n = 48
    models = np.zeros(n, bvr.MODEL_DTYPE)
    models["position"] = rs.uniform(-3, 3, (n, 3)).astype(np.float32)
    models["position"][:, 2] -= 8
    models["radius"] = rs.uniform(0.2, 0.9, n).astype(np.float32)
    models["material_id"] = rs.permutation(n)
    mats = np.zeros(n, bvr.MATERIAL_DTYPE)
    mats["base_color"] = rs.uniform(0.05, 0.95, (n, 3)).astype(np.float32)
    mats["metallic"] = rs.choice([0.0, 0.3, 1.0], n).astype(np.float32)
    mats["roughness"] = rs.uniform(0, 0.8, n).astype(np.float32)
    mats["ior"] = rs.choice([0.7, 1.3, 1.5, 2.4], n).astype(np.float32)
    mats["specular_transmission"] = rs.choice([0.0, 0.5, 1.0], n).astype(np.float32)
    # a tree with 3 models per leaf: 16 leaves under a balanced binary tree (node 0 = root, children adjacent)
    order = np.argsort(models["position"][:, 0])
    models = models[order]
    pad = models["radius"] + np.float32(0.1)
    lo, hi = models["position"] - pad[:, None], models["position"] + pad[:, None]
    nodes = np.zeros(31, bvr.BVH_NODE_DTYPE)

    def build(node, first, count, free):
        nodes["bounds_min"][node], nodes["bounds_max"][node] = lo[first:first + count].min(axis=0), hi[first:first + count].max(axis=0)
        if count <= 3:
            nodes["index"][node], nodes["model_count"][node] = first, count
            return free
        left = (count // 3 // 2) * 3
        nodes["index"][node], nodes["model_count"][node] = free, 0
        c0, c1 = free, free + 1
        free = build(c0, first, left, free + 2)
        return build(c1, first + left, count - left, free)

    assert build(0, 0, n, 1) == 31
    assert bvr.validate_bvh(nodes, models) is None
    W, H = 40, 30
    cam = bvr.make_camera(position=(0.3, 0.2, 1.0), target=(0, 0, -8), aspect=W / H, sample_count=3, bounces=6)
    raster = rs.rand(H, W, 4).astype(np.float32)
    depth = (rs.rand(H, W) * 0.03).astype(np.float32)
    want, cnt = oracle.render(models, mats, nodes, cam, bvr.make_level(level), bvr.make_window(0.62, H), W, raster, depth)
    got = wgsl_numpy.render(models, mats, nodes, cam, level, 0.62, W, H, raster, depth)
    for k in PLANES:
        assert np.array_equal(bits(got[k]), bits(want[k])), (level, k)
    if level:
        assert cnt["hits_shaded"] > 500
