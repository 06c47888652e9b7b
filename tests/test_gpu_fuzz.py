"""Randomised parity of the small-scene production layout (4-wide records, tight boxes, tie rule) against the oracle:
scenes over four decades of scale, radius ratios up to 10^5 (a ground-like giant with spheres resting on it), cameras
from inside the cluster to 40 cluster sizes away.  tools/tight_fuzz.py is the longer version of the same sweep."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, H = 144, 81


def make_case(bvr, it):
    rs = np.random.RandomState(7000 + it)
    n = int(rs.choice([2, 7, 40, 200, 600, 1024]))
    scale = float(10.0 ** rs.uniform(-2, 2))
    spread = float(rs.choice([2.0, 8.0, 30.0]))
    models = np.zeros(n, bvr.MODEL_DTYPE)
    models["position"] = (rs.uniform(-1, 1, (n, 3)) * spread * scale).astype(np.float32)
    rad = 10.0 ** rs.uniform(-1.5, 0.3, n)
    if rs.rand() < 0.5:
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
    dist = float(spread * scale * 10.0 ** rs.uniform(-0.3, 1.6))
    d = rs.normal(size=3)
    d[1] = abs(d[1]) * 0.5 + 0.05
    d /= np.linalg.norm(d)
    cam = bvr.make_camera(position=tuple(d * dist), target=(0, 0, 0), fov=float(rs.uniform(0.2, 1.2)), aspect=W / H,
                          near=0.1 * scale, far=5e4 * scale, sample_count=3, bounces=8)
    return models, mats, cam, float(rs.rand())


@pytest.mark.parametrize("it", range(24))
def test_random_scene_scales(bvr, oracle, ctx, it):
    models, mats, cam, seed = make_case(bvr, it)
    nodes = bvr.build_ploc(models)
    win = bvr.make_window(seed, H)
    ctx.upload_scene(models, mats, nodes)
    got = ctx.render(cam, 3, win, bvr.make_options(W, kernel=1))
    rays = ctx.stats()["rays"]
    want, cnt = oracle.render(models, mats, nodes, cam, bvr.make_level(3), win, W)
    for k in ("primary_id", "primary_depth", "rt_depth", "rgba"):
        assert np.array_equal(np.ascontiguousarray(got[k]).view(np.uint32), np.ascontiguousarray(want[k]).view(np.uint32)), k
    assert rays == cnt["rays"]


def make_big_case(bvr, it):
    """tools/big_fuzz.py's generator: 2 k - 120 k spheres at C4's density, uniform or clustered, four decades of scale."""
    rs = np.random.RandomState(500 + it)
    n = int(rs.choice([2000, 6000, 20000, 50000, 120000]))
    scale = float(10.0 ** rs.uniform(-2, 2))
    side = float((n / 0.125) ** (1 / 3.0))
    models = np.zeros(n, bvr.MODEL_DTYPE)
    if rs.rand() < 0.5:
        pos = rs.uniform(-0.5, 0.5, (n, 3)) * side
    else:
        centres = rs.uniform(-0.5, 0.5, (8, 3)) * side
        pos = centres[rs.randint(0, 8, n)] + rs.normal(size=(n, 3)) * side * 0.04
    models["position"] = (pos * scale).astype(np.float32)
    models["radius"] = (rs.uniform(0.05, 0.25, n) * scale).astype(np.float32)
    models["material_id"] = rs.randint(0, 4, n)
    mats = np.zeros(4, bvr.MATERIAL_DTYPE)
    mats["base_color"] = rs.uniform(0.2, 0.95, (4, 3)).astype(np.float32)
    mats["metallic"] = [0.0, 1.0, 0.0, 0.4]
    mats["roughness"] = [0.5, 0.1, 0.0, 0.6]
    mats["ior"] = 1.5
    mats["specular_transmission"] = [0.0, 0.0, 1.0, 0.3]
    gpu_bvh = bool(rs.rand() < 0.4)
    d = rs.normal(size=3)
    d /= np.linalg.norm(d)
    dist = side * scale * float(rs.uniform(0.2, 1.5))
    cam = bvr.make_camera(position=tuple(d * dist), target=(0, 0, 0), fov=float(rs.uniform(0.3, 1.0)), aspect=160 / 90,
                          near=0.1 * scale, far=1e5 * scale, sample_count=2, bounces=8)
    return models, mats, cam, float(rs.rand()), gpu_bvh


@pytest.mark.parametrize("it", [0, 2, 3, 5, 6, 11, 13, 16])
def test_random_big_scenes(bvr, oracle, ctx, it):
    """Scenes walked in HBM/L2.  Case 11 (extent 3400 units, spheres of radius 2-10, the reference's pad of 0.1 far
    below the f32 noise of hit_sphere there) is the one that made the 16-bit grid conditional on its step size."""
    models, mats, cam, seed, gpu_bvh = make_big_case(bvr, it)
    if gpu_bvh:
        nodes = ctx.upload_scene_gpu_bvh(models, mats, want_nodes=True)
    else:
        nodes = bvr.build_ploc(models)
        ctx.upload_scene(models, mats, nodes)
    win = bvr.make_window(seed, 90)
    got = ctx.render(cam, 3, win, bvr.make_options(160, kernel=1))
    rays = ctx.stats()["rays"]
    want, cnt = oracle.render(models, mats, nodes, cam, bvr.make_level(3), win, 160)
    assert cnt["stack_truncations"] == 0
    for k in ("primary_id", "primary_depth", "rt_depth", "rgba"):
        assert np.array_equal(np.ascontiguousarray(got[k]).view(np.uint32), np.ascontiguousarray(want[k]).view(np.uint32)), k
    assert rays == cnt["rays"]


@pytest.mark.parametrize("shift", [3.0e3, 2.0e4, 1.0e5, 1.0e6])
def test_scenes_far_from_the_origin(bvr, oracle, ctx, shift):
    """The culling-only slab test computes c/d - o/d, whose rounding error grows with the distance from the origin: the
    library drops the tight boxes beyond 4096 units and the fast box arithmetic beyond 32768 (bvr_api.cu).  The same
    scene translated far away must still match the oracle bit for bit."""
    rs = np.random.RandomState(5)
    n = 300
    models = np.zeros(n, bvr.MODEL_DTYPE)
    models["position"] = (rs.uniform(-6, 6, (n, 3)) + np.array([shift, -shift * 0.5, shift * 0.25])).astype(np.float32)
    models["radius"] = rs.uniform(0.2, 0.6, n).astype(np.float32)
    models["material_id"] = np.arange(n)
    mats = np.zeros(n, bvr.MATERIAL_DTYPE)
    mats["base_color"] = rs.uniform(0.1, 0.9, (n, 3)).astype(np.float32)
    mats["metallic"] = (rs.rand(n) < 0.2).astype(np.float32)
    mats["roughness"] = 0.5
    mats["ior"] = 1.5
    mats["specular_transmission"] = (rs.rand(n) < 0.1).astype(np.float32)
    scene = bvr.Scene.from_arrays(models, mats)
    W, H = 160, 90
    centre = (shift, -shift * 0.5, shift * 0.25)
    cam = bvr.make_camera(position=(centre[0], centre[1], centre[2] + 22.0), target=centre, aspect=W / H, sample_count=3, bounces=6)
    win = bvr.make_window(0.23, H)
    want, cnt = oracle.render(scene.models, scene.materials, scene.nodes, cam, bvr.make_level(3), win, W)
    assert cnt["hits_shaded"] > 1000
    for gpu_bvh in (False, True):
        if gpu_bvh:
            nodes = ctx.upload_scene_gpu_bvh(scene.models, scene.materials, want_nodes=True)
            want2, cnt2 = oracle.render(scene.models, scene.materials, nodes, cam, bvr.make_level(3), win, W)
        else:
            ctx.upload_scene(scene.models, scene.materials, scene.nodes)
            want2, cnt2 = want, cnt
        got = ctx.render(cam, 3, win, bvr.make_options(W))
        for k in ("primary_id", "primary_depth", "rt_depth", "rgba"):
            assert np.array_equal(np.ascontiguousarray(got[k]).view(np.uint32), np.ascontiguousarray(want2[k]).view(np.uint32)), (shift, gpu_bvh, k)
        assert ctx.stats()["rays"] == cnt2["rays"]
