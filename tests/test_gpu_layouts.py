"""Every node layout / kernel variant of the megakernel against the oracle (bit-exact).  The variants are picked by
the library from the scene; the environment knobs below force the others for A/B runs:
  small scenes (staged in shared memory): 4-wide fp32 records (default), child-pair records (BVR_NO_BVH4),
      one thread per pixel (BVR_MK_V1);
  big scenes (walked in HBM/L2): 4-wide 16-bit records (default), 2-wide 16-bit records (BVR_NO_BVH4),
      fp32 child-pair records (BVR_NO_Q16, chosen at upload); the 16-bit records numbered in upload order and nothing staged
      (BVR_NO_TOP), the first N records staged in shared memory (BVR_TOP_RECORDS=N), the first N records kept in L1 and the
      others loaded without allocating (BVR_HOT_RECORDS=N)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KNOBS = ("BVR_NO_BVH4", "BVR_NO_Q16", "BVR_NO_TIGHT", "BVR_MK_V1", "BVR_MK_THREADS", "BVR_GPU_VALIDATE", "BVR_NO_BOTH", "BVR_SELFCHECK", "BVR_NO_TOP", "BVR_TOP_RECORDS", "BVR_HOT_RECORDS",
         "BVR_TILE_ORDER", "BVR_W4_LEAN")


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.fixture
def knobs(ctx):
    """The library reads its experiment knobs once per context (bvr_create); bvr_reload_tuning re-reads them."""
    saved = {k: os.environ.pop(k, None) for k in KNOBS}

    def set_(**kw):
        for k in KNOBS:
            os.environ.pop(k, None)
        os.environ.update({k: str(v) for k, v in kw.items()})
        ctx.reload_tuning()
    yield set_
    for k in KNOBS:
        os.environ.pop(k, None)
        if saved[k] is not None:
            os.environ[k] = saved[k]
    ctx.reload_tuning()


def check(got, want, cnt, stats, tag):
    for k in ("primary_id", "primary_depth", "rt_depth", "rgba"):
        assert np.array_equal(bits(got[k]), bits(want[k])), (tag, k)
    assert stats["rays"] == cnt["rays"], tag


SMALL = [dict(), dict(BVR_NO_BVH4=1), dict(BVR_NO_TIGHT=1), dict(BVR_NO_BOTH=1), dict(BVR_MK_V1=1),
         dict(BVR_MK_THREADS=512), dict(BVR_NO_BVH4=1, BVR_MK_THREADS=768)]


@pytest.mark.parametrize("env", SMALL, ids=lambda e: "-".join(f"{k[4:]}={v}" for k, v in e.items()) or "default")
def test_small_scene_layouts(bvr, oracle, ctx, rtiow, knobs, env):
    """RTIOW final scene (506 spheres, in shared memory), book camera, 10 bounces, ragged image size."""
    W, H = 333, 187
    cam = bvr.make_camera(position=(13, 2, 3), target=(0, 0, 0), fov=float(np.deg2rad(20)), aspect=W / H,
                          sample_count=5, bounces=10)
    win = bvr.make_window(0.37, H)
    want, cnt = oracle.render(rtiow.models, rtiow.materials, rtiow.nodes, cam, bvr.make_level(3), win, W)
    knobs(**env)
    ctx.upload_scene(rtiow.models, rtiow.materials, rtiow.nodes)
    got = ctx.render(cam, 3, win, bvr.make_options(W, kernel=1))
    check(got, want, cnt, ctx.stats(), env)


@pytest.mark.parametrize("gpu_bvh", [False, True], ids=["host-ploc", "gpu-lbvh"])
@pytest.mark.parametrize("env", [dict(), dict(BVR_NO_BVH4=1), dict(BVR_NO_Q16=1), dict(BVR_MK_THREADS=768), dict(BVR_NO_TOP=1),
                                 dict(BVR_TOP_RECORDS=1), dict(BVR_TOP_RECORDS=37), dict(BVR_TOP_RECORDS=100000), dict(BVR_HOT_RECORDS=300),
                                 dict(BVR_TOP_RECORDS=21, BVR_HOT_RECORDS=85), dict(BVR_W4_LEAN=0), dict(BVR_W4_LEAN=0, BVR_MK_THREADS=768), dict(BVR_W4_LEAN=0, BVR_HOT_RECORDS=0), dict(BVR_NO_BVH4=1, BVR_NO_TOP=1)],
                         ids=lambda e: "-".join(f"{k[4:]}={v}" for k, v in e.items()) or "default")
def test_big_scene_layouts(bvr, oracle, ctx, knobs, env, gpu_bvh):
    """30k random spheres (C4's density): the records live in HBM/L2.  Also through the GPU-built tree, whose
    nodes are read back and handed to the oracle."""
    scene = bvr.Scene.random(3, 30000, 62.0, 0.05, 0.25)
    W, H = 200, 120
    cam = bvr.make_camera(position=(0, 0, 42), target=(0, 0, 0), aspect=W / H, sample_count=2, bounces=8)
    win = bvr.make_window(0.81, H)
    knobs(**env)
    if gpu_bvh:
        nodes = ctx.upload_scene_gpu_bvh(scene.models, scene.materials, want_nodes=True)
        assert bvr.validate_bvh(nodes, scene.models) is None
    else:
        nodes = scene.nodes
        ctx.upload_scene(scene.models, scene.materials, nodes)
    got = ctx.render(cam, 3, win, bvr.make_options(W, kernel=1))
    want, cnt = oracle.render(scene.models, scene.materials, nodes, cam, bvr.make_level(3), win, W)
    assert cnt["stack_truncations"] == 0
    check(got, want, cnt, ctx.stats(), (env, gpu_bvh))


def test_quantised_records_refuse_boxes_outside_the_root(bvr, oracle, ctx, knobs):
    """A child box that pokes out of the root box cannot be put on the root's 16-bit grid without shrinking it:
    the library must notice and fall back to the fp32 records (same image as the oracle either way)."""
    scene = bvr.Scene.random(5, 6000, 36.0, 0.05, 0.25)
    nodes = scene.nodes.copy()
    inner = np.nonzero(nodes["model_count"] == 0)[0]
    victim = int(nodes["index"][inner[len(inner) // 2]])
    nodes["bounds_max"][victim] += np.float32(500.0)          # still conservative, but outside the root box
    W, H = 160, 100
    cam = bvr.make_camera(position=(0, 0, 26), target=(0, 0, 0), aspect=W / H, sample_count=2, bounces=6)
    win = bvr.make_window(0.5, H)
    knobs()
    ctx.upload_scene(scene.models, scene.materials, nodes)
    got = ctx.render(cam, 3, win, bvr.make_options(W, kernel=1))
    want, cnt = oracle.render(scene.models, scene.materials, nodes, cam, bvr.make_level(3), win, W)
    check(got, want, cnt, ctx.stats(), "box outside the root")


@pytest.mark.parametrize("gpu_bvh", [False, True], ids=["host-ploc", "gpu-lbvh"])
@pytest.mark.parametrize("env", [dict(), dict(BVR_NO_TIGHT=1), dict(BVR_NO_BVH4=1),
                                 dict(BVR_MK_V1=1), dict(BVR_GPU_VALIDATE=1)],
                         ids=lambda e: "-".join(f"{k[4:]}={v}" for k, v in e.items()) or "default")
def test_exact_ties_go_to_the_sphere_the_reference_reaches_first(bvr, oracle, ctx, knobs, env, gpu_bvh):
    """Two spheres with bit-identical t: the reference keeps the one it visits first (strict <, raytrace.wgsl:354),
    and its visiting order is fixed by the tree.  Every sphere of this scene exists twice, with different materials,
    so every hit is a tie; any visiting order must resolve it like the reference does (trace.cuh: model_rank)."""
    rs = np.random.RandomState(11)
    n = 120
    base = np.zeros(n, bvr.MODEL_DTYPE)
    base["position"] = rs.uniform(-4, 4, (n, 3)).astype(np.float32)
    base["position"][:, 2] -= 9
    base["radius"] = rs.uniform(0.2, 0.7, n).astype(np.float32)
    models = np.zeros(2 * n, bvr.MODEL_DTYPE)      # (np.concatenate would pack the 32-byte records to 20 bytes)
    models[:n], models[n:] = base, base
    models = models[rs.permutation(2 * n)]
    models["material_id"] = np.arange(2 * n)
    mats = np.zeros(2 * n, bvr.MATERIAL_DTYPE)
    mats["base_color"] = rs.uniform(0.1, 0.95, (2 * n, 3)).astype(np.float32)
    mats["metallic"] = (rs.rand(2 * n) < 0.3).astype(np.float32)
    mats["roughness"] = rs.uniform(0, 0.6, 2 * n).astype(np.float32)
    mats["ior"] = 1.5
    mats["specular_transmission"] = (rs.rand(2 * n) < 0.2).astype(np.float32)
    W, H = 224, 126
    cam = bvr.make_camera(position=(0, 0, 0), target=(0, 0, -1), aspect=W / H, sample_count=3, bounces=6)
    win = bvr.make_window(0.29, H)
    knobs(**env)
    if gpu_bvh:
        nodes = ctx.upload_scene_gpu_bvh(models, mats, want_nodes=True)
    else:
        nodes = bvr.build_ploc(models)
        ctx.upload_scene(models, mats, nodes)
    got = ctx.render(cam, 3, win, bvr.make_options(W, kernel=1))
    want, cnt = oracle.render(models, mats, nodes, cam, bvr.make_level(3), win, W)
    check(got, want, cnt, ctx.stats(), (env, gpu_bvh))
    # the test means something only if ties are frequent and NOT simply resolved by the lower model index (which is
    # what a visiting order that ignores the tree would do)
    ids = want["primary_id"][want["primary_id"] != 0xFFFFFFFF]
    twin = {}
    for i in range(2 * n):
        twin.setdefault(models["position"][i].tobytes() + models["radius"][i].tobytes(), []).append(i)
    first_of_pair = {min(v) for v in twin.values()}
    hit_first = np.isin(ids, list(first_of_pair)).mean()
    assert len(ids) > 1000 and hit_first < 0.9


def test_gpu_side_validation_rejects_what_the_host_walk_rejects(bvr, rtiow, knobs):
    """Node arrays of 32 k entries and more are validated on the GPU (scene_validate.cu); forced here on the small
    scene so that the same malformed trees go through both implementations."""
    cp = bvr.capi
    for gv in (0, 1):
        knobs(BVR_GPU_VALIDATE=gv)
        c = bvr.Context(0)
        cases = []
        bad = rtiow.nodes.copy(); bad["index"][0] = 0; cases.append(bad)                     # root points at itself
        bad = rtiow.nodes.copy()
        leaf = int(np.nonzero(bad["model_count"] > 0)[0][0])
        bad["index"][leaf] = len(rtiow.models); cases.append(bad)                              # leaf beyond the models
        bad = rtiow.nodes.copy()
        inner = np.nonzero(bad["model_count"] == 0)[0]
        bad["index"][inner[3]] = len(bad) - 1; cases.append(bad)                               # second child out of range
        bad = rtiow.nodes.copy(); bad["index"][inner[5]] = bad["index"][inner[7]]; cases.append(bad)   # two parents
        bad = rtiow.nodes.copy(); bad["model_count"][leaf] = 200; cases.append(bad)           # leaf too large
        # a node the root does not reach, with a child index far outside the array (the derive kernels visit it too)
        bad = np.zeros(len(rtiow.nodes) + 1, bvr.BVH_NODE_DTYPE); bad[:-1] = rtiow.nodes
        bad["index"][-1], bad["model_count"][-1] = 0x7ffffff0, 0; cases.append(bad)
        bad = bad.copy(); bad["index"][-1], bad["model_count"][-1] = len(rtiow.models) + 9, 1; cases.append(bad)
        for k, nodes in enumerate(cases):
            with pytest.raises(bvr.BvrError) as e:
                c.upload_scene(rtiow.models, rtiow.materials, nodes)
            assert e.value.status == cp.BVR_ERR_BAD_SCENE, (gv, k)
        # the context stays usable after errors
        c.upload_scene(rtiow.models, rtiow.materials, rtiow.nodes)
        W, H = 64, 36
        cam = bvr.make_camera(sample_count=1, bounces=3, aspect=W / H)
        out = c.render(cam, 3, bvr.make_window(0.2, H), bvr.make_options(W))
        st = c.stats()
        c.close()
        if gv == 0:
            ref, ref_rays = out, st["rays"]
        else:
            assert st["rays"] == ref_rays and all(np.array_equal(bits(out[k]), bits(ref[k])) for k in ref)


def test_dirty_range_upload_of_a_big_scene(bvr, oracle, knobs):
    """40 k spheres (80 k nodes: validated on the GPU), a few hundred of them moved, the tree rebuilt on the host and
    only the dirty ranges uploaded: same image as the oracle's on the new scene, far fewer bytes than a full upload."""
    knobs()
    scene = bvr.Scene.random(9, 40000, 68.0, 0.05, 0.25)
    models, mats = scene.models.copy(), scene.materials.copy()
    W, H = 160, 90
    cam = bvr.make_camera(position=(0, 0, 46), target=(0, 0, 0), aspect=W / H, sample_count=2, bounces=6)
    win = bvr.make_window(0.44, H)
    c = bvr.Context(0)
    c.upload_scene(models, mats, scene.nodes)
    c.render(cam, 3, win, bvr.make_options(W))
    h2d0 = c.stats()["h2d_bytes"]
    rs = np.random.RandomState(1)
    moved = np.sort(rs.choice(len(models), 300, replace=False))
    models["position"][moved] += rs.uniform(-0.2, 0.2, (300, 3)).astype(np.float32)
    nodes = bvr.build_ploc(models)
    changed = np.nonzero((nodes.view(np.uint8).reshape(-1, 48) != scene.nodes.view(np.uint8).reshape(-1, 48)).any(axis=1))[0]
    ranges = [(bvr.capi.ARRAY_MODELS, int(i), 1) for i in moved]
    ranges.append((bvr.capi.ARRAY_BVH_NODES, int(changed.min()), int(changed.max() - changed.min() + 1)))
    c.upload_scene(models, mats, nodes, ranges)
    got = c.render(cam, 3, win, bvr.make_options(W))
    sent = c.stats()["h2d_bytes"] - h2d0
    assert sent == 300 * 32 + int(changed.max() - changed.min() + 1) * 48     # ranks come from the GPU-side validation
    want, cnt = oracle.render(models, mats, nodes, cam, bvr.make_level(3), win, W)
    check(got, want, cnt, c.stats(), "dirty big scene")
    c.close()


def test_selfcheck_mode_retraces_sampled_rays_in_reference_order(bvr, oracle, ctx, rtiow, knobs):
    """BVR_SELFCHECK=1: about one finished ray in 1024 is traced again, in the kernel, with the reference's own traversal
    order and boxes; the culling shortcuts (tight boxes, 16-bit grid, FMA slab test) must never change a closest hit."""
    W, H = 480, 270
    cam = bvr.make_camera(position=(13, 2, 3), target=(0, 0, 0), fov=float(np.deg2rad(20)), aspect=W / H, sample_count=8, bounces=10)
    win = bvr.make_window(0.37, H)
    big = bvr.Scene.random(3, 30000, 62.0, 0.05, 0.25)
    bcam = bvr.make_camera(position=(0, 0, 42), target=(0, 0, 0), aspect=W / H, sample_count=4, bounces=8)
    for scene, c, env in ((rtiow, cam, dict()), (rtiow, cam, dict(BVR_NO_BOTH=1)), (rtiow, cam, dict(BVR_NO_TIGHT=1)),
                          (big, bcam, dict()), (big, bcam, dict(BVR_NO_BVH4=1))):
        knobs(**env)
        ctx.upload_scene(scene.models, scene.materials, scene.nodes)
        off = ctx.render(c, 3, win, bvr.make_options(W, kernel=1))
        st = ctx.stats()
        assert st["selfcheck_rays"] == 0 and st["selfcheck_mismatches"] == 0
        knobs(BVR_SELFCHECK=1, **env)
        on = ctx.render(c, 3, win, bvr.make_options(W, kernel=1))
        st1 = ctx.stats()
        assert st1["rays"] == st["rays"]
        assert st["rays"] / 4096 < st1["selfcheck_rays"] < st["rays"] / 256, (env, st1)
        assert st1["selfcheck_mismatches"] == 0, (env, st1)
        for k in off:
            assert np.array_equal(bits(on[k]), bits(off[k])), (env, k)


@pytest.mark.parametrize("order", [0, 1, 2, 3], ids=["row-major", "reversed", "heaviest-first", "lightest-first"])
def test_tile_order_cannot_change_a_frame(bvr, oracle, ctx, rtiow, knobs, order):
    """The pixel queue hands out tiles heaviest first, judged by the previous frame's per-tile ray counts (tile_order.cu):
    the first frame of a size runs row-major, the following ones in the fed-back order, a new size starts over.  Every one
    of them has to be the oracle's frame, bit for bit, with the oracle's ray count."""
    knobs(BVR_TILE_ORDER=order)
    ctx.upload_scene(rtiow.models, rtiow.materials, rtiow.nodes)
    for (W, H) in ((333, 187), (120, 67)):
        cam = bvr.make_camera(position=(13, 2, 3), target=(0, 0, 0), fov=float(np.deg2rad(20)), aspect=W / H,
                              sample_count=4, bounces=10)
        win = bvr.make_window(0.37, H)
        want, cnt = oracle.render(rtiow.models, rtiow.materials, rtiow.nodes, cam, bvr.make_level(3), win, W)
        for frame in range(3):
            got = ctx.render(cam, 3, win, bvr.make_options(W, kernel=1))
            check(got, want, cnt, ctx.stats(), (order, W, frame))
