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
