"""GPU edge cases, golden fixtures, sharding, dirty uploads, error behaviour and the App mirror — all
through the C ABI, checked against the CPU oracle (bit-exact)."""
import ctypes as C

import numpy as np
import pytest

from golden_util import golden_names, load_golden

pytestmark = pytest.mark.gpu

KERNELS = [(1, 0), (1, 1), (2, 0), (3, 0)]
KERNEL_IDS = ["megakernel-near-first", "megakernel-reference-order", "wavefront", "cta-wavefront"]


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("kernel,traversal", KERNELS, ids=KERNEL_IDS)
def test_output_weight_is_one_rounding_per_texel(bvr, ctx, rtiow, kernel, traversal):
    """BvrRenderOptions.output_weight (a rank's share in a sample-sharded frame): rgba and rt_depth leave the library
    multiplied by it — inside the store of the production kernel, as a pass of its own after the others — and the
    result is the unweighted plane times the weight, one IEEE multiplication per word; the other planes are untouched."""
    W, H = 200, 113
    cam = bvr.make_camera(position=(13, 2, 3), target=(0, 0, 0), fov=float(np.deg2rad(20)), aspect=W / H, sample_count=3, bounces=6)
    win = bvr.make_window(0.37, H)
    ctx.upload_scene(rtiow.models, rtiow.materials, rtiow.nodes)
    plain = ctx.render(cam, 3, win, bvr.make_options(W, kernel, traversal))
    wgt = np.float32(3.0 / 13.0)
    got = ctx.render(cam, 3, win, bvr.make_options(W, kernel, traversal, output_weight=float(wgt)))
    assert np.array_equal(bits(got["rgba"]), bits(plain["rgba"] * wgt))
    assert np.array_equal(bits(got["rt_depth"]), bits(plain["rt_depth"] * wgt))
    assert np.array_equal(got["primary_id"], plain["primary_id"])
    assert np.array_equal(bits(got["primary_depth"]), bits(plain["primary_depth"]))
    one = ctx.render(cam, 3, win, bvr.make_options(W, kernel, traversal, output_weight=1.0))
    assert np.array_equal(bits(one["rgba"]), bits(plain["rgba"]))


def test_extra_sample_tiles(bvr, ctx, rtiow):
    """BVR_RENDER_EXTRA_SAMPLE: the pixels of the tiles of class (tx + ty + phase) % modulus < count take one sample more.
    A pixel is its own RNG stream, so such a frame is the n-sample frame on those tiles and the (n-1)-sample frame on the
    others, each weighted by its own sample count x output_weight; all four planes, and the ray count adds up."""
    W, H, n, S = 203, 117, 3, 31
    cam = lambda k: bvr.make_camera(position=(13, 2, 3), target=(0, 0, 0), fov=float(np.deg2rad(20)), aspect=W / H, sample_count=k, bounces=6)
    win = bvr.make_window(0.37, H)
    ctx.upload_scene(rtiow.models, rtiow.materials, rtiow.nodes)
    lo = ctx.render(cam(n), 3, win, bvr.make_options(W))
    rays_lo = ctx.stats()["rays"]
    hi = ctx.render(cam(n + 1), 3, win, bvr.make_options(W))
    ty, tx = np.mgrid[0:H, 0:W]
    ty, tx = ty // 4, tx // 8
    for modulus, phase, count in ((2, 0, 1), (2, 1, 1), (8, 3, 4), (5, 2, 0), (3, 1, 3)):
        o = bvr.make_options(W, output_weight=1.0 / S)
        o.flags |= bvr.capi.render_extra_sample_bits(modulus, phase, count)
        got = ctx.render(cam(n), 3, win, o)
        st = ctx.stats()
        extra = ((tx + ty + phase) % modulus) < count
        w_lo, w_hi = np.float32(n) * np.float32(1.0 / S), np.float32(n + 1) * np.float32(1.0 / S)
        want_rgba = np.where(extra[..., None], hi["rgba"] * w_hi, lo["rgba"] * w_lo)
        want_depth = np.where(extra, hi["rt_depth"] * w_hi, lo["rt_depth"] * w_lo)
        assert np.array_equal(bits(got["rgba"]), bits(want_rgba)), (modulus, phase, count)
        assert np.array_equal(bits(got["rt_depth"]), bits(want_depth))
        assert np.array_equal(got["primary_id"], lo["primary_id"])
        assert st["paths"] == W * H * n + int(extra.sum())
        assert rays_lo <= st["rays"]
    for kernel, traversal in ((2, 0), (3, 0), (1, 1)):      # only the megakernel's near-first walk takes uneven samples
        o = bvr.make_options(W, kernel, traversal)
        o.flags |= bvr.capi.render_extra_sample_bits(2, 0, 1)
        with pytest.raises(RuntimeError):
            ctx.render(cam(n), 3, win, o)
    o = bvr.make_options(W)
    o.flags |= bvr.capi.render_extra_sample_bits(2, 0, 1)
    with pytest.raises(RuntimeError):
        ctx.render(cam(0), 3, win, o)


def test_sum_slots_adds_in_slot_order(bvr, ctx):
    """bvr_sum_slots_device: dst = ((s0 + s1) + s2) + ... over the slots in the mask, first term copied."""
    import torch
    n, world = 4 * 1000, 5
    g = torch.Generator(device="cuda").manual_seed(3)
    slots = (torch.rand((world, n + 8), device="cuda", generator=g) * 3.0 - 1.0).contiguous()
    dst = torch.full((n,), 7.0, device="cuda")
    for mask in (0b11111, 0b10110, 0b00001, 0):
        ctx.sum_slots_device(slots.data_ptr(), n + 8, world, mask, dst.data_ptr(), n)
        ctx.sync()
        want = torch.zeros(n, device="cuda")
        first = True
        for s in range(world):
            if (mask >> s) & 1:
                want = slots[s, :n].clone() if first else want + slots[s, :n]
                first = False
        assert torch.equal(dst, want), mask
    with pytest.raises(RuntimeError):
        ctx.sum_slots_device(slots.data_ptr(), n + 8, 65, 1, dst.data_ptr(), n)


def check(got, want, keys=("primary_id", "primary_depth", "rt_depth", "rgba")):
    for k in keys:
        nbad = int((bits(got[k]) != bits(want[k])).sum())
        assert nbad == 0, f"{k}: {nbad} words differ"


@pytest.mark.parametrize("kernel,traversal", KERNELS, ids=KERNEL_IDS)
@pytest.mark.parametrize("name", golden_names())
def test_golden_fixtures(bvr, ctx, name, kernel, traversal):
    g = load_golden(bvr, name)
    ctx.upload_scene(g["models"], g["materials"], g["nodes"])
    got = ctx.render(g["camera"], g["level"], bvr.make_window(g["seed"], g["height"]),
                     bvr.make_options(g["width"], kernel=kernel, traversal=traversal), g["raster_rgba"], g["raster_depth"])
    check(got, {k[4:]: v for k, v in g.items() if k.startswith("out_")})
    assert ctx.stats()["rays"] == int(g["rays"])


@pytest.mark.parametrize("kernel,traversal", KERNELS, ids=KERNEL_IDS)
def test_empty_scene_every_ray_misses(bvr, oracle, ctx, kernel, traversal):
    """n_models == 0 is legal: the sky gradient everywhere (the reference would skip the frame, pipeline.rs:141-151)."""
    e = (np.zeros(0, bvr.MODEL_DTYPE), np.zeros(0, bvr.MATERIAL_DTYPE), np.zeros(0, bvr.BVH_NODE_DTYPE))
    W, H = 50, 30
    cam = bvr.make_camera(sample_count=2, bounces=3, aspect=W / H)
    win = bvr.make_window(0.4, H)
    ctx.upload_scene(*e)
    got = ctx.render(cam, 3, win, bvr.make_options(W, kernel=kernel, traversal=traversal))
    want, cnt = oracle.render(*e, cam, bvr.make_level(3), win, W)
    check(got, want)
    assert np.all(got["primary_id"] == 0xFFFFFFFF) and ctx.stats()["rays"] == cnt["rays"] == W * H * 2


@pytest.mark.parametrize("kernel,traversal", KERNELS, ids=KERNEL_IDS)
def test_single_sphere_leaf_root_and_ragged_size(bvr, oracle, ctx, kernel, traversal):
    """N == 1: the root is a leaf.  Image size is not a multiple of the 8x4 warp tile."""
    models = np.zeros(1, bvr.MODEL_DTYPE)
    models["position"][0] = (0.3, -0.2, -4)
    models["radius"][0] = 1.5
    mats = np.zeros(1, bvr.MATERIAL_DTYPE)
    mats["base_color"][0] = (0.8, 0.3, 0.2)
    mats["roughness"][0] = 0.5
    mats["ior"][0] = 1.5
    nodes = bvr.build_ploc(models)
    W, H = 37, 23
    cam = bvr.make_camera(position=(0, 0, 0), target=(0, 0, -1), aspect=W / H, sample_count=3, bounces=5)
    win = bvr.make_window(0.9, H)
    ctx.upload_scene(models, mats, nodes)
    got = ctx.render(cam, 3, win, bvr.make_options(W, kernel=kernel, traversal=traversal))
    want, cnt = oracle.render(models, mats, nodes, cam, bvr.make_level(3), win, W)
    check(got, want)
    assert (got["primary_id"] == 0).any() and (got["primary_id"] == 0xFFFFFFFF).any()


@pytest.mark.parametrize("kernel,traversal", KERNELS, ids=KERNEL_IDS)
def test_mixed_materials_and_multi_model_leaves(bvr, oracle, ctx, kernel, traversal):
    """Fractional metallic / transmission (every scatter branch reachable from one material), material ids
    that are not the identity, and hand-built leaves holding several models (the contract allows it)."""
    rs = np.random.RandomState(4)
    n = 24
    models = np.zeros(n, bvr.MODEL_DTYPE)
    models["position"] = rs.uniform(-3, 3, (n, 3)).astype(np.float32)
    models["position"][:, 2] -= 8
    models["radius"] = rs.uniform(0.3, 0.9, n).astype(np.float32)
    models["material_id"] = rs.randint(0, 5, n)
    mats = np.zeros(5, bvr.MATERIAL_DTYPE)
    mats["base_color"] = rs.uniform(0.2, 0.9, (5, 3)).astype(np.float32)
    mats["metallic"] = [0.0, 0.5, 1.0, 0.3, 0.0]
    mats["roughness"] = [0.5, 0.2, 0.0, 0.7, 0.1]
    mats["ior"] = [1.5, 1.3, 1.5, 0.8, 2.4]      # ior < 1: the cannot_refract branch
    mats["specular_transmission"] = [0.0, 0.5, 0.0, 0.6, 1.0]
    # a two-level tree by hand: root -> 2 leaves of 12 models each
    nodes = np.zeros(3, bvr.BVH_NODE_DTYPE)
    for k, (a, b) in enumerate([(0, 12), (12, 24)]):
        pad = (models["radius"][a:b] + np.float32(0.1))[:, None]
        nodes["bounds_min"][1 + k] = (models["position"][a:b] - pad).min(axis=0)
        nodes["bounds_max"][1 + k] = (models["position"][a:b] + pad).max(axis=0)
        nodes["index"][1 + k], nodes["model_count"][1 + k] = a, b - a
    nodes["bounds_min"][0] = nodes["bounds_min"][1:].min(axis=0)
    nodes["bounds_max"][0] = nodes["bounds_max"][1:].max(axis=0)
    nodes["index"][0], nodes["model_count"][0] = 1, 0
    assert bvr.validate_bvh(nodes, models) is None
    W, H = 120, 80
    cam = bvr.make_camera(position=(0, 0, 0), target=(0, 0, -1), aspect=W / H, sample_count=8, bounces=12)
    win = bvr.make_window(0.23, H)
    ctx.upload_scene(models, mats, nodes)
    got = ctx.render(cam, 3, win, bvr.make_options(W, kernel=kernel, traversal=traversal))
    want, cnt = oracle.render(models, mats, nodes, cam, bvr.make_level(3), win, W)
    check(got, want)
    assert ctx.stats()["rays"] == cnt["rays"]


@pytest.mark.parametrize("kernel,traversal", KERNELS, ids=KERNEL_IDS)
@pytest.mark.parametrize("level", [0, 1, 2, 3])
def test_composite_levels(bvr, oracle, ctx, rtiow, level, kernel, traversal):
    """fragment's level switch and depth test (raytrace.wgsl:97-122), fused into the render kernels."""
    W, H = 160, 90
    cam = bvr.make_camera(sample_count=2, bounces=3, aspect=W / H)
    win = bvr.make_window(0.66, H)
    rs = np.random.RandomState(level)
    raster = rs.rand(H, W, 4).astype(np.float32)
    depth = (rs.rand(H, W) * 0.08).astype(np.float32)
    ctx.upload_scene(rtiow.models, rtiow.materials, rtiow.nodes)
    got = ctx.render(cam, level, win, bvr.make_options(W, kernel=kernel, traversal=traversal), raster, depth)
    want, _ = oracle.render(rtiow.models, rtiow.materials, rtiow.nodes, cam, bvr.make_level(level), win, W, raster, depth)
    check(got, want, keys=("rgba",) if level == 0 else ("primary_id", "primary_depth", "rt_depth", "rgba"))
    if level in (1, 2):
        assert (got["rgba"] == raster).all(axis=-1).any() and (got["rgba"] != raster).any(axis=-1).any()


@pytest.mark.parametrize("kernel,traversal", KERNELS, ids=KERNEL_IDS)
def test_large_random_scene_global_memory_path(bvr, oracle, ctx, kernel, traversal):
    """C4-style scene (BASELINE configs[3]) at reduced size: 20k spheres do not fit in shared memory, so the
    kernels walk the child-pair records in HBM/L2.  Same density as the 2^20-sphere benchmark scene."""
    scene = bvr.Scene.random(7, 20000, 54.0, 0.05, 0.25)
    assert bvr.validate_bvh(scene.nodes, scene.models) is None
    W, H = 192, 108
    cam = bvr.make_camera(position=(0, 0, 36), target=(0, 0, 0), aspect=W / H, sample_count=2, bounces=6)
    win = bvr.make_window(0.42, H)
    ctx.upload_scene(scene.models, scene.materials, scene.nodes)
    got = ctx.render(cam, 3, win, bvr.make_options(W, kernel=kernel, traversal=traversal))
    want, cnt = oracle.render(scene.models, scene.materials, scene.nodes, cam, bvr.make_level(3), win, W)
    check(got, want)
    assert ctx.stats()["rays"] == cnt["rays"] and cnt["stack_truncations"] == 0


def test_srgb8_store(bvr, oracle, ctx, rtiow):
    """Rgba8UnormSrgb store conversion of the colour target (pipeline.rs:311-315), +-1 LSB (powf on device)."""
    W, H = 128, 72
    cam = bvr.make_camera(sample_count=2, bounces=4, aspect=W / H)
    ctx.upload_scene(rtiow.models, rtiow.materials, rtiow.nodes)
    got = ctx.render(cam, 3, bvr.make_window(0.5, H), bvr.make_options(W), want=("rgba", "srgb8"))
    want = oracle.store_srgb8(got["rgba"])
    assert np.abs(got["srgb8"].astype(int) - want.astype(int)).max() <= 1
    assert (got["srgb8"] == want).mean() > 0.99


@pytest.mark.parametrize("kernel", [1, 2, 3], ids=["megakernel", "wavefront", "cta-wavefront"])
@pytest.mark.parametrize("world,strip", [(2, 4), (3, 8), (8, 1)])
def test_tile_shards_reassemble_bit_exact(bvr, oracle, ctx, rtiow, world, strip, kernel):
    """Tile sharding: the union of all shards equals the unsharded image bit for bit."""
    from bevyray_b200.distributed import shard_global_rows
    W, H = 128, 70
    cam = bvr.make_camera(sample_count=2, bounces=4, aspect=W / H)
    win = bvr.make_window(0.37, H)
    ctx.upload_scene(rtiow.models, rtiow.materials, rtiow.nodes)
    full = ctx.render(cam, 3, win, bvr.make_options(W, kernel=kernel))
    rays_full = ctx.stats()["rays"]
    out = {k: np.zeros_like(v) for k, v in full.items()}
    rays = 0
    for r in range(world):
        opts = bvr.make_options(W, kernel=kernel, shard_index=r, shard_count=world, strip_rows=strip)
        part = ctx.render(cam, 3, win, opts)
        rays += ctx.stats()["rays"]
        rows = shard_global_rows(H, r, world, strip)
        assert part["rgba"].shape[0] == len(rows)
        valid = rows < H
        for k in out:
            out[k][rows[valid]] = part[k][valid]
    check(out, full)
    assert rays == rays_full


def test_unshard_and_axpby_device_kernels(bvr, ctx):
    import torch
    world, strip, W, H, ch = 3, 4, 40, 30, 4
    from bevyray_b200.distributed import shard_global_rows
    rows = [shard_global_rows(H, r, world, strip) for r in range(world)]
    full = torch.arange(H * W * ch, dtype=torch.float32, device="cuda").reshape(H, W, ch)
    gathered = torch.zeros((world, len(rows[0]), W, ch), dtype=torch.float32, device="cuda")
    for r in range(world):
        valid = rows[r] < H
        gathered[r, torch.from_numpy(np.nonzero(valid)[0]).cuda()] = full[torch.from_numpy(rows[r][valid]).cuda()]
    out = torch.zeros_like(full)
    ctx.unshard_device(gathered.data_ptr(), len(rows[0]) * W * ch, out.data_ptr(), W, H, ch, world, strip)
    ctx.sync()
    assert torch.equal(out, full)
    a = torch.rand(1000, device="cuda")
    b = torch.rand(1000, device="cuda")
    want = a * 0.25 + b * 0.75
    ctx.axpby_device(a.data_ptr(), 0.25, b.data_ptr(), 0.75, 1000)
    ctx.sync()
    assert torch.allclose(a, want, atol=1e-7)


def test_dirty_range_upload_equals_full_upload(bvr, oracle, rtiow):
    """Only the dirty element ranges travel (pinned staging -> HBM); the result equals a full re-upload."""
    W, H = 96, 54
    cam = bvr.make_camera(sample_count=1, bounces=4, aspect=W / H)
    win = bvr.make_window(0.37, H)
    models, mats = rtiow.models.copy(), rtiow.materials.copy()
    c = bvr.Context(0)
    c.upload_scene(models, mats, rtiow.nodes)
    base = c.render(cam, 3, win, bvr.make_options(W))
    h2d0 = c.stats()["h2d_bytes"]
    # move two spheres and recolour one material; rebuild the BVH like prepare_buffers does every frame
    models["position"][5] += np.float32(0.05)
    models["position"][-1][1] += np.float32(0.5)
    mats["base_color"][0] = (0.9, 0.1, 0.1)
    nodes = bvr.build_ploc(models)
    changed = np.nonzero((nodes.view(np.uint8).reshape(-1, 48) != rtiow.nodes.view(np.uint8).reshape(-1, 48)).any(axis=1))[0]
    ranges = [(bvr.capi.ARRAY_MODELS, 5, 1), (bvr.capi.ARRAY_MODELS, len(models) - 1, 1), (bvr.capi.ARRAY_MATERIALS, 0, 1),
              (bvr.capi.ARRAY_BVH_NODES, int(changed.min()), int(changed.max() - changed.min() + 1))]
    c.upload_scene(models, mats, nodes, ranges)
    got = c.render(cam, 3, win, bvr.make_options(W))
    sent = c.stats()["h2d_bytes"] - h2d0
    # dirty elements + 4 bytes per model of traversal ranks, which are derived from the new tree on the host
    assert sent == 3 * 32 + (int(changed.max() - changed.min() + 1)) * 48 + 4 * len(models)
    c2 = bvr.Context(0)
    c2.upload_scene(models, mats, nodes)
    want = c2.render(cam, 3, win, bvr.make_options(W))
    check(got, want)
    assert (bits(got["rgba"]) != bits(base["rgba"])).any()
    ora, _ = oracle.render(models, mats, nodes, cam, bvr.make_level(3), win, W)
    check(got, ora)
    c.close()
    c2.close()


def test_error_codes(bvr, rtiow):
    cp = bvr.capi
    c = bvr.Context(0)
    cam = bvr.make_camera()
    win = bvr.make_window(0.1, 16)
    with pytest.raises(bvr.BvrError) as e:
        c.render(cam, 3, win, bvr.make_options(16))
    assert e.value.status == cp.BVR_ERR_NO_SCENE
    c.upload_scene(rtiow.models, rtiow.materials, rtiow.nodes)
    ortho = bvr.make_camera()
    ortho.projection = 1
    with pytest.raises(bvr.BvrError) as e:
        c.render(ortho, 3, win, bvr.make_options(16))
    assert e.value.status == cp.BVR_ERR_UNSUPPORTED_PROJECTION     # extract.rs:148
    with pytest.raises(bvr.BvrError) as e:
        c.render(cam, 2, win, bvr.make_options(16))                # level 2 without raster inputs
    assert e.value.status == cp.BVR_ERR_INVALID_ARGUMENT
    with pytest.raises(bvr.BvrError) as e:
        c.render(cam, 7, win, bvr.make_options(16))
    assert e.value.status == cp.BVR_ERR_INVALID_ARGUMENT
    bad = rtiow.nodes.copy()
    bad["index"][0] = 0                                            # root points at itself: a cycle
    with pytest.raises(bvr.BvrError) as e:
        c.upload_scene(rtiow.models, rtiow.materials, bad)
    assert e.value.status == cp.BVR_ERR_BAD_SCENE
    bad = rtiow.nodes.copy()
    leaf = int(np.nonzero(bad["model_count"] > 0)[0][0])
    bad["index"][leaf] = len(rtiow.models)                         # leaf beyond the model buffer
    with pytest.raises(bvr.BvrError) as e:
        c.upload_scene(rtiow.models, rtiow.materials, bad)
    assert e.value.status == cp.BVR_ERR_BAD_SCENE
    with pytest.raises(bvr.BvrError) as e:
        c.upload_scene(rtiow.models[:10], rtiow.materials, rtiow.nodes, [(0, 0, 1)])   # ranges with changed counts
    assert e.value.status == cp.BVR_ERR_INVALID_ARGUMENT
    # the context stays usable after errors
    c.upload_scene(rtiow.models, rtiow.materials, rtiow.nodes)
    assert c.render(cam, 3, win, bvr.make_options(16))["rgba"].shape == (16, 16, 4)
    c.close()


def test_deep_tree_falls_back_to_reference_order(bvr, oracle, ctx):
    """A degenerate 70-level chain is deeper than the near-first stack: the library switches to the
    reference-order kernel, whose 32-entry stack semantics (raytrace.wgsl:320) are the oracle's."""
    n = 70
    models = np.zeros(n, bvr.MODEL_DTYPE)
    models["position"][:, 0] = np.arange(n) * 0.5 - 17
    models["position"][:, 2] = -30
    models["radius"] = 0.3
    models["material_id"] = 0
    mats = np.zeros(1, bvr.MATERIAL_DTYPE)
    mats["base_color"][0] = (0.7, 0.7, 0.7)
    mats["roughness"][0] = 0.5
    mats["ior"][0] = 1.5
    nodes = np.zeros(2 * n - 1, bvr.BVH_NODE_DTYPE)
    pad = models["radius"] + np.float32(0.1)
    lo, hi = models["position"] - pad[:, None], models["position"] + pad[:, None]
    # node 2k = inner (children 2k+1 leaf k, 2k+2 rest); last node = leaf n-1
    for k in range(n - 1):
        nodes["index"][2 * k], nodes["model_count"][2 * k] = 2 * k + 1, 0
        nodes["bounds_min"][2 * k], nodes["bounds_max"][2 * k] = lo[k:].min(axis=0), hi[k:].max(axis=0)
        nodes["index"][2 * k + 1], nodes["model_count"][2 * k + 1] = k, 1
        nodes["bounds_min"][2 * k + 1], nodes["bounds_max"][2 * k + 1] = lo[k], hi[k]
    nodes["index"][-1], nodes["model_count"][-1] = n - 1, 1
    nodes["bounds_min"][-1], nodes["bounds_max"][-1] = lo[-1], hi[-1]
    assert bvr.validate_bvh(nodes, models) is None
    W, H = 96, 32
    cam = bvr.make_camera(position=(0, 0, 0), target=(0, 0, -1), aspect=W / H, sample_count=2, bounces=3)
    win = bvr.make_window(0.3, H)
    ctx.upload_scene(models, mats, nodes)
    got = ctx.render(cam, 3, win, bvr.make_options(W))
    want, cnt = oracle.render(models, mats, nodes, cam, bvr.make_level(3), win, W)
    check(got, want)


def test_uploaded_trees_of_32_to_63_levels_truncate_like_the_reference(bvr, oracle, ctx):
    """raytrace.wgsl:320 abandons a traversal once its stack index reaches 32, which a tree of 32+ levels can trigger.
    For an uploaded tree that deep the library runs the verbatim reference-order kernel: here a 40-level
    tree of overlapping spheres whose layout leaves one stack entry behind per level, so the oracle does truncate."""
    n = 40
    rs = np.random.RandomState(4)
    models = np.zeros(n, bvr.MODEL_DTYPE)
    # all spheres on one line of sight, overlapping boxes: a ray down the line enters every box
    models["position"][:, 0] = rs.uniform(-0.3, 0.3, n)
    models["position"][:, 1] = rs.uniform(-0.3, 0.3, n)
    models["position"][:, 2] = -8 - 0.8 * np.arange(n)
    models["radius"] = 0.5
    models["material_id"] = np.arange(n) % 2
    mats = np.zeros(2, bvr.MATERIAL_DTYPE)
    mats["base_color"] = [(0.8, 0.6, 0.2), (0.3, 0.6, 0.9)]
    mats["roughness"] = 0.5
    mats["ior"] = 1.5
    pad = models["radius"] + np.float32(0.1)
    lo, hi = models["position"] - pad[:, None], models["position"] + pad[:, None]
    # node layout: inner k at index 2k (k = 0..n-2), children at 2k+1 (LEAF k) and 2k+2 (the rest: inner k+1 or the
    # last leaf).  The reference pushes 2k+1 first and 2k+2 second, pops the REST first: the leaves pile up on its stack,
    # one per level -> the stack index reaches 32 at level 31 and the traversal is abandoned.
    nodes = np.zeros(2 * n - 1, bvr.BVH_NODE_DTYPE)
    for k in range(n - 1):
        nodes["index"][2 * k], nodes["model_count"][2 * k] = 2 * k + 1, 0
        nodes["bounds_min"][2 * k], nodes["bounds_max"][2 * k] = lo[k:].min(axis=0), hi[k:].max(axis=0)
        nodes["index"][2 * k + 1], nodes["model_count"][2 * k + 1] = k, 1
        nodes["bounds_min"][2 * k + 1], nodes["bounds_max"][2 * k + 1] = lo[k], hi[k]
    nodes["index"][-1], nodes["model_count"][-1] = n - 1, 1
    nodes["bounds_min"][-1], nodes["bounds_max"][-1] = lo[-1], hi[-1]
    assert bvr.validate_bvh(nodes, models) is None
    ranks, depth = bvr.traversal_ranks(nodes, n)
    assert 32 <= depth == n < 64
    W, H = 64, 48
    cam = bvr.make_camera(position=(0, 0, 0), target=(0, 0, -1), fov=0.3, aspect=W / H, sample_count=2, bounces=3)
    win = bvr.make_window(0.3, H)
    ctx.upload_scene(models, mats, nodes)
    got = ctx.render(cam, 3, win, bvr.make_options(W))
    want, cnt = oracle.render(models, mats, nodes, cam, bvr.make_level(3), win, W)
    assert cnt["stack_truncations"] > 0                      # the case is real: the reference does give up here
    check(got, want)
    assert ctx.stats()["rays"] == cnt["rays"]
    # the same spheres with a tree built by the library (bvr_upload_scene_gpu_bvh) are walked completely: the closest
    # hit of every camera ray equals brute force over all spheres
    ctx.upload_scene_gpu_bvh(models, mats)
    got2 = ctx.render(cam, 3, win, bvr.make_options(W))
    brute, _ = oracle.render(models, mats, nodes, cam, bvr.make_level(3), win, W, brute_force=True)
    assert np.array_equal(got2["primary_id"], brute["primary_id"]) and np.array_equal(bits(got2["primary_depth"]), bits(brute["primary_depth"]))
    assert not np.array_equal(want["primary_id"], brute["primary_id"])   # truncation changed what the reference sees


def test_app_mirror_renders_demo_scene(bvr, oracle):
    """RaytracePlugin + one frame of the schedule (extract -> prepare_buffers -> RayTracingNode::run),
    src/raytracing/{mod,extract,pipeline}.rs, at the demo defaults (FallbackRaytraced, 4 spp, 4 bounces)."""
    lib = bvr.capi.lib
    app = C.c_void_p(lib.bvrh_app_create())
    try:
        cam_e = lib.bvrh_app_setup_demo(app, 1)
        assert lib.bvrh_app_add_raytrace_plugin(app, 0) == 0
        W, H = 320, 180
        lib.bvrh_app_set_window_size(app, W, H)
        lib.bvrh_app_set_seed(app, 0.37)
        assert lib.bvrh_app_update(app) == 1, lib.bvrh_app_last_error(app)
        w, h = C.c_uint32(), C.c_uint32()
        ptr = lib.bvrh_app_frame(app, cam_e, C.byref(w), C.byref(h))
        assert (w.value, h.value) == (W, H)
        frame = np.frombuffer((C.c_char * (W * H * 16)).from_address(ptr), np.float32).reshape(H, W, 4).copy()
        scene = bvr.Scene.rtiow(1)
        cam = bvr.make_camera(sample_count=4, bounces=4, aspect=W / H)
        raster = np.ones((H, W, 4), np.float32)          # clear colour WHITE (main.rs:60), nothing rasterised
        depth = np.zeros((H, W), np.float32)
        want, _ = oracle.render(scene.models, scene.materials, scene.nodes, cam, bvr.make_level(2), bvr.make_window(0.37, H), W, raster, depth)
        assert np.array_equal(bits(frame), bits(want["rgba"]))
        # second frame: nothing changed -> no scene bytes travel (dirty-range upload), same seed -> same image
        st0 = bvr.capi.BvrStats()
        lib.bvrh_app_get_stats(app, C.byref(st0))
        assert lib.bvrh_app_update(app) == 1
        st1 = bvr.capi.BvrStats()
        lib.bvrh_app_get_stats(app, C.byref(st1))
        assert st1.h2d_bytes - st0.h2d_bytes == W * H * 20       # only the raster colour + depth inputs
        # move one sphere: only its model, and the BVH nodes that changed, are uploaded
        lib.bvrh_app_set_translation(app, cam_e + 5, 0.5, 0.2, 1.5)
        assert lib.bvrh_app_update(app) == 1
        st2 = bvr.capi.BvrStats()
        lib.bvrh_app_get_stats(app, C.byref(st2))
        scene_bytes = scene.models.nbytes + scene.materials.nbytes + scene.nodes.nbytes
        assert 32 <= st2.h2d_bytes - st1.h2d_bytes - W * H * 20 < scene_bytes
        # orthographic cameras are never extracted (extract.rs:148): no view, nothing rendered
        app2 = C.c_void_p(lib.bvrh_app_create())
        lib.bvrh_app_spawn_window(app2, 64, 64)
        f3 = C.c_float * 3
        lib.bvrh_app_spawn_camera(app2, f3(0, 0, 5), f3(0, 0, 0), f3(0, 1, 0), 0.785, 1.0, 0.1, 1000.0, 3, 1, 1, 1)
        assert lib.bvrh_app_add_raytrace_plugin(app2, 0) == 0
        assert lib.bvrh_app_update(app2) == 0
        lib.bvrh_app_destroy(app2)
    finally:
        lib.bvrh_app_destroy(app)


def test_render_async_with_two_contexts(bvr, rtiow):
    """bvr_render_async enqueues input copies, kernels and output copies and returns; bvr_sync completes the frame.
    Two contexts on one device pipeline an animated scene (frame f renders while frame f+1 is uploaded): every frame must
    equal the synchronous render of the same scene, and pageable buffers are refused."""
    import torch
    W, H = 192, 108
    cam = bvr.make_camera(sample_count=2, bounces=4, aspect=W / H)
    rs = np.random.RandomState(3)
    raster = torch.from_numpy(rs.rand(H, W, 4).astype(np.float32)).pin_memory().numpy()
    depth = torch.from_numpy((rs.rand(H, W) * 0.02).astype(np.float32)).pin_memory().numpy()
    scene = bvr.Scene.random(4, 3000, 30.0, 0.1, 0.4)
    cam = bvr.make_camera(position=(0, 0, 28), target=(0, 0, 0), aspect=W / H, sample_count=2, bounces=4)
    ctxs = [bvr.Context(0), bvr.Context(0)]
    outs = [{"rgba": torch.empty((H, W, 4), dtype=torch.float32).pin_memory().numpy(),
             "primary_id": torch.empty((H, W), dtype=torch.int32).pin_memory().numpy().view(np.uint32)} for _ in ctxs]
    ref_ctx = bvr.Context(0)
    opts = bvr.make_options(W)
    frames = []
    pending = [None, None]
    for f in range(6):
        k = f % 2
        ctxs[k].sync()
        if pending[k] is not None:
            frames.append((pending[k], outs[k]["rgba"].copy(), outs[k]["primary_id"].copy()))
        scene.animate(f + 1, rebuild_bvh=False)
        ctxs[k].upload_scene_gpu_bvh(scene.models, scene.materials)
        ctxs[k].render(cam, 2, bvr.make_window(0.1 * (f + 1), H), opts, raster, depth, want=("rgba", "primary_id"), out=outs[k],
                       asynchronous=True)
        pending[k] = (f, scene.models.copy())
    for k in range(2):
        ctxs[k].sync()
        frames.append((pending[k], outs[k]["rgba"].copy(), outs[k]["primary_id"].copy()))
    assert len(frames) == 6
    for (f, models), rgba, pid in frames:
        ref_ctx.upload_scene_gpu_bvh(models, scene.materials)
        want = ref_ctx.render(cam, 2, bvr.make_window(0.1 * (f + 1), H), opts, raster, depth, want=("rgba", "primary_id"))
        assert np.array_equal(bits(rgba), bits(want["rgba"])) and np.array_equal(pid, want["primary_id"]), f
    with pytest.raises(bvr.BvrError) as e:
        ctxs[0].render(cam, 2, bvr.make_window(0.5, H), opts, raster, depth, want=("rgba",), out={"rgba": np.empty((H, W, 4), np.float32)},
                       asynchronous=True)
    assert e.value.status == bvr.capi.BVR_ERR_INVALID_ARGUMENT
    for c in ctxs + [ref_ctx]:
        c.close()
