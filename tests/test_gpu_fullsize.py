"""Checks at BASELINE.json's full sizes.  The oracle cannot render a whole 1080p x 100 spp frame in seconds,
so the full-size frame is checked (a) against the oracle on sampled row bands (rows are independent, so a band
of the full frame is a full-fidelity sample), and (b) through size-independent properties: every kernel variant
produces the same bits, re-rendering is deterministic, tile shards reassemble to the full frame, and the ray
counter adds up."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def book_camera(bvr, W, H, spp, bounces):
    return bvr.make_camera(position=(13.0, 2.0, 3.0), target=(0.0, 0.0, 0.0), fov=float(np.deg2rad(20.0)),
                           aspect=W / H, sample_count=spp, bounces=bounces)


def test_c2_full_frame_against_oracle_bands(bvr, oracle, ctx, rtiow):
    """BASELINE configs[1]: RTIOW final scene, 1920x1080, 100 spp, 10 bounces (the benchmark workload)."""
    W, H = 1920, 1080
    cam = book_camera(bvr, W, H, 100, 10)
    win = bvr.make_window(0.37, H)
    ctx.upload_scene(rtiow.models, rtiow.materials, rtiow.nodes)
    got = ctx.render(cam, 3, win, bvr.make_options(W))
    rays = ctx.stats()["rays"]
    assert 2 * W * H * 100 < rays < 11 * W * H * 100
    for y0 in (0, 377, 700, 1076):      # sky, horizon, spheres, bottom edge
        want, _ = oracle.render(rtiow.models, rtiow.materials, rtiow.nodes, cam, bvr.make_level(3), win, W, rows=(y0, y0 + 4))
        for k in ("primary_id", "primary_depth", "rt_depth", "rgba"):
            assert np.array_equal(bits(got[k][y0:y0 + 4]), bits(want[k][y0:y0 + 4])), (k, y0)
    # deterministic: a second render of the same frame is identical, and so is the ray count
    again = ctx.render(cam, 3, win, bvr.make_options(W), want=("rgba",))
    assert np.array_equal(bits(again["rgba"]), bits(got["rgba"])) and ctx.stats()["rays"] == rays
    # radiance is finite and in range (gamma-encoded averages of values in [0,1])
    rgb = got["rgba"][..., :3]
    assert np.isfinite(rgb).all() and rgb.min() >= 0.0 and rgb.max() <= 1.0 + 1e-6 and np.all(got["rgba"][..., 3] == 1.0)


def test_c2_whole_frame_against_oracle(bvr, oracle, ctx, rtiow):
    """The benchmark frame itself — 1920x1080, 100 spp, 10 bounces, 516 M rays — every plane, every pixel, against
    the oracle (about half a minute of CPU).  This is the frame on which a small sphere resting on the ground
    produces a bit-exact tie in t at pixel (741, 412): see trace.cuh test_leaf."""
    W, H = 1920, 1080
    cam = book_camera(bvr, W, H, 100, 10)
    win = bvr.make_window(0.37, H)
    want, cnt = oracle.render(rtiow.models, rtiow.materials, rtiow.nodes, cam, bvr.make_level(3), win, W)
    ctx.upload_scene(rtiow.models, rtiow.materials, rtiow.nodes)
    got = ctx.render(cam, 3, win, bvr.make_options(W))
    for k in ("primary_id", "primary_depth", "rt_depth", "rgba"):
        assert np.array_equal(bits(got[k]), bits(want[k])), k
    assert ctx.stats()["rays"] == cnt["rays"] == 515999699


def test_c2_size_kernel_variants_and_shards_agree(bvr, ctx, rtiow):
    """Full 1080p frame at 8 spp: megakernel == reference-order == wavefront == CTA wavefront, and 8 tile
    shards reassemble to the same frame with the same total ray count."""
    from bevyray_b200.distributed import shard_global_rows
    W, H = 1920, 1080
    cam = book_camera(bvr, W, H, 8, 10)
    win = bvr.make_window(0.37, H)
    ctx.upload_scene(rtiow.models, rtiow.materials, rtiow.nodes)
    ref = ctx.render(cam, 3, win, bvr.make_options(W, kernel=1, traversal=0))
    rays = ctx.stats()["rays"]
    for kernel, traversal in [(1, 1), (2, 0), (3, 0)]:
        out = ctx.render(cam, 3, win, bvr.make_options(W, kernel=kernel, traversal=traversal))
        for k in ref:
            assert np.array_equal(bits(out[k]), bits(ref[k])), (kernel, traversal, k)
        assert ctx.stats()["rays"] == rays
    full = np.zeros_like(ref["rgba"])
    total = 0
    for r in range(8):
        part = ctx.render(cam, 3, win, bvr.make_options(W, shard_index=r, shard_count=8, strip_rows=4), want=("rgba",))
        total += ctx.stats()["rays"]
        rows = shard_global_rows(H, r, 8, 4)
        valid = rows < H
        full[rows[valid]] = part["rgba"][valid]
    assert np.array_equal(bits(full), bits(ref["rgba"])) and total == rays


def test_c3_size_4k_band(bvr, oracle, ctx, rtiow):
    """BASELINE configs[2] geometry (3840x2160): one shard of an 8-way tile sharding at 16 spp against the oracle."""
    from bevyray_b200.distributed import shard_global_rows
    W, H = 3840, 2160
    cam = book_camera(bvr, W, H, 16, 10)
    win = bvr.make_window(0.37, H)
    ctx.upload_scene(rtiow.models, rtiow.materials, rtiow.nodes)
    part = ctx.render(cam, 3, win, bvr.make_options(W, shard_index=5, shard_count=8, strip_rows=4), want=("rgba", "primary_id"))
    rows = shard_global_rows(H, 5, 8, 4)
    ly = 200                              # a strip in the middle of the shard
    gy = int(rows[ly])
    want, _ = oracle.render(rtiow.models, rtiow.materials, rtiow.nodes, cam, bvr.make_level(3), win, W, rows=(gy, gy + 4))
    assert np.array_equal(bits(part["rgba"][ly:ly + 4]), bits(want["rgba"][gy:gy + 4]))
    assert np.array_equal(part["primary_id"][ly:ly + 4], want["primary_id"][gy:gy + 4])


def test_c4_size_full_frame_against_oracle(bvr, oracle, ctx):
    """BASELINE configs[3] scene and resolution (2^20 random spheres, 1920x1080, 10 bounces) at 1 spp: the whole
    frame against the oracle, for the production path of scenes walked in HBM/L2 (4-wide 16-bit records, 4-byte
    stack entries), the 2-wide quantised records, the fp32 records and the reference-order kernel."""
    import os
    W, H = 1920, 1080
    scene = bvr.Scene.random(7, 1 << 20, 200.0, 0.05, 0.25)
    cam = bvr.make_camera(position=(0.0, 0.0, 130.0), target=(0.0, 0.0, 0.0), aspect=W / H, sample_count=1, bounces=10)
    win = bvr.make_window(0.37, H)
    want, cnt = oracle.render(scene.models, scene.materials, scene.nodes, cam, bvr.make_level(3), win, W)
    assert cnt["stack_truncations"] == 0
    ctx.upload_scene(scene.models, scene.materials, scene.nodes)
    variants = [({}, 0), ({"BVR_NO_BVH4": "1"}, 0), ({"BVR_NO_Q16": "1"}, 0), ({}, 1)]
    try:
        for env, traversal in variants:
            for k in ("BVR_NO_BVH4", "BVR_NO_Q16"):
                os.environ.pop(k, None)
            os.environ.update(env)
            ctx.reload_tuning()
            if "BVR_NO_Q16" in env:
                ctx.upload_scene(scene.models, scene.materials, scene.nodes)   # the records are chosen at upload
            got = ctx.render(cam, 3, win, bvr.make_options(W, traversal=traversal))
            for k in ("primary_id", "primary_depth", "rt_depth", "rgba"):
                assert np.array_equal(bits(got[k]), bits(want[k])), (env, traversal, k)
            assert ctx.stats()["rays"] == cnt["rays"]
    finally:
        for k in ("BVR_NO_BVH4", "BVR_NO_Q16"):
            os.environ.pop(k, None)
        ctx.reload_tuning()
