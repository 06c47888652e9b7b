"""GPU BVH builder: bvr_upload_scene_gpu_bvh builds the BVHNode array on the device — PLOC over the
Morton order (default; the reference's own algorithm, extract.rs:316-321) or the plain LBVH (BVR_GPU_LBVH=1).
Checked against the reference contract with the HOST validator, rendered through the oracle with the
downloaded nodes (bit-exact), and compared with the image obtained from the host PLOC tree (the closest hit
does not depend on topology)."""
import numpy as np
import pytest

import os

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["ploc", "lbvh"])
def builder(request, ctx):
    """Both GPU builders behind bvr_upload_scene_gpu_bvh (the knob is read at bvr_create / bvr_reload_tuning)."""
    saved = os.environ.pop("BVR_GPU_LBVH", None)
    if request.param == "lbvh":
        os.environ["BVR_GPU_LBVH"] = "1"
    ctx.reload_tuning()
    yield request.param
    os.environ.pop("BVR_GPU_LBVH", None)
    if saved is not None:
        os.environ["BVR_GPU_LBVH"] = saved
    ctx.reload_tuning()


def sah_cost(nodes):
    """Surface-area-heuristic cost of a tree in the reference layout: sum of inner-node areas / root area."""
    d = nodes["bounds_max"].astype(np.float64) - nodes["bounds_min"].astype(np.float64)
    area = d[:, 0] * d[:, 1] + d[:, 1] * d[:, 2] + d[:, 2] * d[:, 0]
    inner = nodes["model_count"] == 0
    return float(area[inner].sum() / area[0])


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def random_models(bvr, n, seed, spread=10.0, coincident=False):
    rs = np.random.RandomState(seed)
    m = np.zeros(n, bvr.MODEL_DTYPE)
    if not coincident:
        m["position"] = rs.uniform(-spread, spread, (n, 3)).astype(np.float32)
    m["position"][:, 2] -= 30
    m["radius"] = rs.uniform(0.2, 0.8, n).astype(np.float32)
    m["material_id"] = np.arange(n)
    mats = np.zeros(n, bvr.MATERIAL_DTYPE)
    mats["base_color"] = rs.uniform(0.1, 0.9, (n, 3)).astype(np.float32)
    mats["roughness"] = 0.5
    mats["ior"] = 1.5
    mats["metallic"] = (rs.rand(n) < 0.2).astype(np.float32)
    return m, mats


@pytest.mark.parametrize("n", [1, 2, 3, 5, 64, 1000, 20000])
def test_contract_and_parity(bvr, oracle, ctx, builder, n):
    models, mats = random_models(bvr, n, seed=n)
    nodes = ctx.upload_scene_gpu_bvh(models, mats, want_nodes=True)
    assert len(nodes) == 2 * n - 1
    assert bvr.validate_bvh(nodes, models) is None
    leaves = nodes[nodes["model_count"] > 0]
    assert np.all(leaves["model_count"] == 1) and sorted(leaves["index"].tolist()) == list(range(n))
    W, H = 96, 64
    cam = bvr.make_camera(position=(0, 0, 0), target=(0, 0, -1), aspect=W / H, sample_count=2, bounces=4)
    win = bvr.make_window(0.7, H)
    got = ctx.render(cam, 3, win, bvr.make_options(W))
    want, cnt = oracle.render(models, mats, nodes, cam, bvr.make_level(3), win, W)
    assert cnt["stack_truncations"] == 0
    for k in ("primary_id", "primary_depth", "rt_depth", "rgba"):
        assert np.array_equal(bits(got[k]), bits(want[k])), k
    # same image as with the host PLOC tree
    ctx.upload_scene(models, mats, bvr.build_ploc(models))
    ref = ctx.render(cam, 3, win, bvr.make_options(W))
    for k in ref:
        assert np.array_equal(bits(got[k]), bits(ref[k])), k


def test_coincident_spheres_and_demo_scene(bvr, oracle, ctx, builder, rtiow):
    models, mats = random_models(bvr, 40, seed=1, coincident=True)   # identical Morton keys
    nodes = ctx.upload_scene_gpu_bvh(models, mats, want_nodes=True)
    assert bvr.validate_bvh(nodes, models) is None
    nodes = ctx.upload_scene_gpu_bvh(rtiow.models, rtiow.materials, want_nodes=True)
    assert bvr.validate_bvh(nodes, rtiow.models) is None
    W, H = 160, 90
    cam = bvr.make_camera(position=(13, 2, 3), target=(0, 0, 0), fov=float(np.deg2rad(20)), aspect=W / H, sample_count=2, bounces=6)
    win = bvr.make_window(0.37, H)
    got = ctx.render(cam, 3, win, bvr.make_options(W))
    want, _ = oracle.render(rtiow.models, rtiow.materials, nodes, cam, bvr.make_level(3), win, W)
    for k in ("primary_id", "primary_depth", "rt_depth", "rgba"):
        assert np.array_equal(bits(got[k]), bits(want[k])), k


def test_dirty_models_rebuild(bvr, ctx, builder):
    models, mats = random_models(bvr, 500, seed=3)
    ctx.upload_scene_gpu_bvh(models, mats)
    h0 = ctx.stats()["h2d_bytes"]
    models["position"][17] += np.float32(0.5)
    nodes = ctx.upload_scene_gpu_bvh(models, mats, ranges=[(bvr.capi.ARRAY_MODELS, 17, 1)], want_nodes=True)
    assert ctx.stats()["h2d_bytes"] - h0 == 32
    assert bvr.validate_bvh(nodes, models) is None
    with pytest.raises(bvr.BvrError):
        ctx.upload_scene_gpu_bvh(models, mats, ranges=[(bvr.capi.ARRAY_BVH_NODES, 0, 1)])


def test_ploc_tree_is_deterministic_and_as_good_as_the_host_ploc(bvr, ctx):
    """Two PLOC builds of the same models give the same bytes (node ids come from a scan, not from atomics), and the tree
    has the surface-area-heuristic cost of the host PLOC builder's tree (the restated obvhs call, extract.rs:316-321):
    same algorithm, same search radius, same order.  (On uniformly random spheres a plain LBVH is a few per cent cheaper
    by that measure; on the structured RTIOW scene PLOC renders faster — profiles/r02_tuning_sweeps.txt.)"""
    os.environ.pop("BVR_GPU_LBVH", None)
    ctx.reload_tuning()
    for n, seed in ((506, 1), (5000, 2), (60000, 3)):      # one CTA / one CTA / cooperative grid
        models, mats = random_models(bvr, n, seed=seed, spread=30.0)
        a = ctx.upload_scene_gpu_bvh(models, mats, want_nodes=True)
        b = ctx.upload_scene_gpu_bvh(models, mats, want_nodes=True)
        assert a.tobytes() == b.tobytes()
        c_gpu, c_host = sah_cost(a), sah_cost(bvr.build_ploc(models))
        assert abs(c_gpu - c_host) < 0.02 * c_host, (n, c_gpu, c_host)


def test_refit_keeps_topology_and_matches_the_oracle(bvr, oracle, ctx, builder):
    """bvr_refit_scene_gpu_bvh: spheres move a little, the tree keeps its topology (same index / model_count columns),
    its boxes follow the spheres, and the image equals the oracle's on the refitted nodes."""
    models, mats = random_models(bvr, 3000, seed=9, spread=14.0)
    nodes0 = ctx.upload_scene_gpu_bvh(models, mats, want_nodes=True)
    rs = np.random.RandomState(4)
    moved = np.sort(rs.choice(len(models), 400, replace=False))
    models["position"][moved] += rs.uniform(-0.3, 0.3, (400, 3)).astype(np.float32)
    models["radius"][moved[:50]] *= np.float32(1.3)
    h0 = ctx.stats()["h2d_bytes"]
    ranges = [(bvr.capi.ARRAY_MODELS, int(i), 1) for i in moved]
    nodes1 = ctx.upload_scene_gpu_bvh(models, mats, ranges=ranges, want_nodes=True, refit=True)
    assert ctx.stats()["h2d_bytes"] - h0 == 400 * 32
    assert np.array_equal(nodes1["index"], nodes0["index"]) and np.array_equal(nodes1["model_count"], nodes0["model_count"])
    assert not np.array_equal(nodes1["bounds_min"], nodes0["bounds_min"])
    assert bvr.validate_bvh(nodes1, models) is None
    W, H = 128, 80
    cam = bvr.make_camera(position=(0, 0, 0), target=(0, 0, -1), aspect=W / H, sample_count=2, bounces=5)
    win = bvr.make_window(0.55, H)
    got = ctx.render(cam, 3, win, bvr.make_options(W))
    want, cnt = oracle.render(models, mats, nodes1, cam, bvr.make_level(3), win, W)
    for k in ("primary_id", "primary_depth", "rt_depth", "rgba"):
        assert np.array_equal(bits(got[k]), bits(want[k])), k
    assert ctx.stats()["rays"] == cnt["rays"]
    # a refit needs a tree of the library's own for the same counts
    with pytest.raises(bvr.BvrError):
        ctx.upload_scene_gpu_bvh(models[:100], mats[:100], refit=True)
    ctx.upload_scene(models, mats, bvr.build_ploc(models))
    with pytest.raises(bvr.BvrError):
        ctx.upload_scene_gpu_bvh(models, mats, refit=True)


def test_app_mirror_with_gpu_bvh(bvr, oracle):
    """RayTracingNode::gpu_bvh: prepare_buffers skips the host PLOC build and the library builds the tree;
    the frame equals the one rendered with the host-built tree."""
    import ctypes as C
    lib = bvr.capi.lib
    frames = []
    for gpu in (0, 1):
        app = C.c_void_p(lib.bvrh_app_create())
        cam_e = lib.bvrh_app_setup_demo(app, 1)
        assert lib.bvrh_app_add_raytrace_plugin(app, 0) == 0
        lib.bvrh_app_set_gpu_bvh(app, gpu)
        lib.bvrh_app_set_window_size(app, 160, 90)
        lib.bvrh_app_set_seed(app, 0.37)
        assert lib.bvrh_app_update(app) == 1, lib.bvrh_app_last_error(app)
        w, h = C.c_uint32(), C.c_uint32()
        ptr = lib.bvrh_app_frame(app, cam_e, C.byref(w), C.byref(h))
        frames.append(np.frombuffer((C.c_char * (160 * 90 * 16)).from_address(ptr), np.float32).copy())
        nn = C.c_size_t()
        lib.bvrh_app_buffers(app, None, None, None, C.byref(nn))
        assert (nn.value == 0) == bool(gpu)          # no host tree was built in GPU mode
        lib.bvrh_app_destroy(app)
    assert np.array_equal(bits(frames[0]), bits(frames[1]))
