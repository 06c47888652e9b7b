"""Known-answer tests that pin the CPU oracle.

The reference has no tests, golden vectors or fixtures, so every vector here is DERIVED
from the WGSL text: assets/shaders/random.wgsl:3-30, assets/shaders/raytrace.wgsl:95, 371-398.  The
RNG / seed vectors were computed independently of the oracle with plain Python integer / numpy-f32
arithmetic (recomputed below as well)."""
import ctypes as C

import numpy as np


def py_rng_next(state):
    """random.wgsl:8-15 in Python integers (independent restatement)."""
    old = (state + 747796405 + 2891336453) & 0xFFFFFFFF
    word = (((old >> ((old >> 28) + 4)) ^ old) * 277803737) & 0xFFFFFFFF
    return ((word >> 22) ^ word) & 0xFFFFFFFF


def test_additive_constant():
    assert (747796405 + 2891336453) % 2**32 == 3639132858


def test_rng_state_sequences(oracle):
    kat = {0: [0xA8BEEA3C, 0x0A2A1484, 0x1E93BE90, 0x75134D09],
           1: [0xB94DD992, 0x7D3246CC, 0xCB994A9C, 0x4DD1F399],
           12345: [0x21EBFEE8, 0x06C77023, 0x3D3393C9, 0xE142C31A],
           0xDEADBEEF: [0xCC79F6D5, 0xD2E80976, 0xE12301FF, 0xAE983DC4]}
    for seed, want in kat.items():
        assert oracle.rng_sequence(seed, 4) == want
        s, got = seed, []
        for _ in range(4):
            s = py_rng_next(s)
            got.append(s)
        assert got == want


def test_rng_matches_python_restatement_on_random_states(oracle):
    rs = np.random.RandomState(7)
    for s in rs.randint(0, 2**32, size=2000, dtype=np.uint64):
        assert oracle.lib.bvro_rng_next_int(int(s)) == py_rng_next(int(s))


def test_rng_float_conversion(oracle):
    # f32(0xffffffffu) == 4294967296.0, so rngNextFloat = f32(state) * 2^-32 and can return exactly 1.0
    assert np.float32(0xFFFFFFFF) == np.float32(4294967296.0)
    f = oracle.lib.bvro_rng_float_of_state
    np.testing.assert_array_equal(
        np.array([f(s) for s in (0xA8BEEA3C, 0x0A2A1484, 0x1E93BE90, 0x75134D09)], np.float32),
        np.array([0.6591631, 0.03970459, 0.1194419, 0.45732576], np.float32))
    assert f(0xFFFFFF80) == 1.0 and f(0xFFFFFF7F) < 1.0 and f(0) == 0.0
    for s in (1, 12345, 0x80000000, 0xFFFFFFFF):
        assert np.float32(f(s)) == np.float32(np.float32(s) / np.float32(4294967296.0))


def np_pixel_seed(seed, x, y, w, h):
    """raytrace.wgsl:95 in numpy f32 (independent restatement), left-associative product."""
    f = np.float32
    u = (f(x) + f(0.5)) / f(w)
    v = (f(y) + f(0.5)) / f(h)
    return int(np.uint32(((f(seed) * f(10000.0)) * (u * f(402.0))) * (v * f(31.5))))


def test_pixel_seed_kat(oracle):
    ps = oracle.lib.bvro_pixel_seed
    assert ps(0.5, 0, 0, 1280, 720) == 17
    assert ps(0.5, 639, 359, 1280, 720) == 15794417
    assert ps(0.5, 1279, 719, 1280, 720) == 63246312
    rs = np.random.RandomState(3)
    for _ in range(500):
        x, y = int(rs.randint(0, 1920)), int(rs.randint(0, 1080))
        seed = float(np.float32(rs.rand()))
        assert ps(seed, x, y, 1920, 1080) == np_pixel_seed(seed, x, y, 1920, 1080)
    # the largest possible product fits u32: no saturation
    assert 10000 * 402 * 31.5 < 2**32
    # saturating conversion for out-of-contract seeds
    assert ps(-1.0, 5, 5, 64, 64) == 0


def test_seed_collisions_are_reproduced_not_fixed(oracle):
    """Many pixels share an RNG stream: 1,280,711 distinct seeds for 1920x1080 at seed 0.5."""
    f = np.float32
    xs = (np.arange(1920, dtype=np.float32) + f(0.5)) / f(1920)
    ys = (np.arange(1080, dtype=np.float32) + f(0.5)) / f(1080)
    a = (f(0.5) * f(10000.0)) * (xs * f(402.0))
    seeds = (a[None, :] * (ys * f(31.5))[:, None]).astype(np.uint32)
    assert len(np.unique(seeds)) == 1280711
    for (x, y) in [(0, 0), (77, 901), (1919, 1079)]:
        assert oracle.lib.bvro_pixel_seed(0.5, x, y, 1920, 1080) == int(seeds[y, x])


def _f3(*v):
    return (C.c_float * 3)(*v)


def test_hit_sphere_analytic(bvr, oracle):
    m = bvr.capi.BvrModel()
    m.position[:] = [0.0, 0.0, -5.0]
    m.radius = 1.0
    hs = oracle.lib.bvro_hit_sphere
    # head-on: near root at t = 4
    assert hs(C.byref(m), _f3(0, 0, 0), _f3(0, 0, -1)) == 4.0
    # un-normalised direction: t is parametric (raytrace.wgsl:373)
    assert hs(C.byref(m), _f3(0, 0, 0), _f3(0, 0, -2)) == 2.0
    # clean miss returns exactly -1.0
    assert hs(C.byref(m), _f3(0, 0, 0), _f3(0, 1, 0)) == -1.0
    # from inside: only the NEAR root is used (raytrace.wgsl:382) -> negative t, rejected by t > 0.001
    assert hs(C.byref(m), _f3(0, 0, -5), _f3(0, 0, -1)) == -1.0 * 1.0
    # behind the origin
    assert hs(C.byref(m), _f3(0, 0, -10), _f3(0, 0, -1)) == -6.0


def test_ray_bounding_dst(oracle):
    rb = oracle.lib.bvro_ray_bounding_dst
    INF = np.float32(3.40282347e+38)
    mn, mx = _f3(-1, -1, -1), _f3(1, 1, 1)
    assert rb(_f3(0, 0, 5), _f3(0, 0, -1), mn, mx) == 4.0          # entry distance
    assert rb(_f3(0, 0, 0), _f3(0, 0, -1), mn, mx) == 0.0          # origin inside -> 0
    assert np.float32(rb(_f3(0, 0, 5), _f3(0, 0, 1), mn, mx)) == INF   # box behind
    assert np.float32(rb(_f3(3, 0, 5), _f3(0, 0, -1), mn, mx)) == INF  # passes beside (dir.x == 0 -> inf slabs)
    # direction component 0 with the origin exactly on a slab plane: (1-1)*inf = NaN; min/max ignore the NaN
    # operand (IEEE minNum/maxNum), so both t1.x and t2.x become -inf and the box is missed
    assert np.float32(rb(_f3(1, 0, 5), _f3(0, 0, -1), mn, mx)) == INF
    # just inside the slab it is an ordinary hit
    assert rb(_f3(0.999, 0, 5), _f3(0, 0, -1), mn, mx) == 4.0


def test_tan_half_fov(oracle):
    assert np.float32(oracle.lib.bvro_tan_half_fov(np.float32(np.pi / 2))) == np.float32(np.tan(np.float64(np.float32(np.pi / 2) * np.float32(0.5))))
    assert abs(oracle.lib.bvro_tan_half_fov(np.float32(np.pi / 4)) - 0.41421357) < 1e-7


def test_bvh_equals_brute_force(bvr, oracle, rtiow):
    """Closest hit does not depend on BVH topology: id, depth and radiance planes of the
    BVH traversal are bit-identical to testing every sphere."""
    W, H = 256, 144
    cam = bvr.make_camera(sample_count=2, bounces=6, aspect=W / H)
    win = bvr.make_window(0.61, H)
    a, ca = oracle.render(rtiow.models, rtiow.materials, rtiow.nodes, cam, bvr.make_level(3), win, W)
    b, cb = oracle.render(rtiow.models, rtiow.materials, rtiow.nodes, cam, bvr.make_level(3), win, W, brute_force=True)
    for k in a:
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k
    assert ca["rays"] == cb["rays"] and ca["rng_draws"] == cb["rng_draws"]
    assert ca["stack_truncations"] == 0 and ca["max_stack"] < 32
    assert cb["sphere_tests"] == cb["rays"] * len(rtiow.models)


def test_single_sphere_depth_and_id(bvr, oracle):
    """One diffuse sphere straight ahead: centre pixel hits it at distance 4, corners see the sky."""
    models = np.zeros(1, bvr.MODEL_DTYPE)
    models["position"][0] = (0, 0, -5)
    models["radius"][0] = 1.0
    mats = np.zeros(1, bvr.MATERIAL_DTYPE)
    mats["base_color"][0] = (0.5, 0.5, 0.5)
    mats["roughness"][0] = 0.5
    mats["ior"][0] = 1.5
    nodes = bvr.build_ploc(models)
    assert len(nodes) == 1 and nodes["model_count"][0] == 1
    W = H = 65
    cam = bvr.make_camera(position=(0, 0, 0), target=(0, 0, -1), aspect=1.0, sample_count=1, bounces=0)
    planes, cnt = oracle.render(models, mats, nodes, cam, bvr.make_level(3), bvr.make_window(0.5, H), W)
    assert planes["primary_id"][H // 2, W // 2] == 0
    assert abs(planes["primary_depth"][H // 2, W // 2] - 4.0) < 1e-3
    assert planes["primary_id"][0, 0] == 0xFFFFFFFF and planes["primary_depth"][0, 0] == bvr.INF
    # bounces = 0: a hit scatters once and the loop is exhausted -> black (raytrace.wgsl:214-216)
    assert np.all(planes["rgba"][H // 2, W // 2, :3] == 0.0) and planes["rgba"][H // 2, W // 2, 3] == 1.0
    # a miss returns the sqrt-encoded sky gradient
    assert np.all(planes["rgba"][0, 0, :3] > 0.5)
    # level 3 miss depth falls back to far - 1 (raytrace.wgsl:177-182, 219-221)
    assert planes["rt_depth"][0, 0] == np.float32(999.0)


def test_composite_levels(bvr, oracle, rtiow):
    """fragment's level switch (raytrace.wgsl:97-122): 0 = raster, 1/2 = reverse-Z depth test, 3 = raytraced."""
    W, H = 96, 54
    cam = bvr.make_camera(sample_count=1, bounces=2, aspect=W / H)
    win = bvr.make_window(0.2, H)
    rs = np.random.RandomState(0)
    raster = rs.rand(H, W, 4).astype(np.float32)
    depth = np.zeros((H, W), np.float32)
    depth[:, W // 2:] = 1.0          # right half: raster geometry at the near plane -> raster wins
    args = (rtiow.models, rtiow.materials, rtiow.nodes, cam)
    p3, _ = oracle.render(*args, bvr.make_level(3), win, W)
    p0, _ = oracle.render(*args, bvr.make_level(0), win, W, raster, depth)
    assert np.array_equal(p0["rgba"], raster)
    f = np.float32
    for lvl in (1, 2):
        p, _ = oracle.render(*args, bvr.make_level(lvl), win, W, raster, depth)
        # independent numpy restatement of raytrace.wgsl:104-120 on the oracle's own averaged depth
        rt_depth = p["rt_depth"]
        with np.errstate(divide="ignore"):
            rd = np.where(rt_depth > f(cam.far_plane), f(-1.0), f(cam.near_plane) / rt_depth).astype(np.float32)
        raster_wins = depth > rd
        want = np.where(raster_wins[..., None], raster, p3["rgba"])
        assert np.array_equal(p["rgba"], want)
        assert raster_wins.any() and (~raster_wins).any()
        miss = p["primary_id"] == 0xFFFFFFFF
        # 1 spp: a miss's depth is the fallback: far+10 at level 1 (-> -1: raster always wins, even at
        # depth 0), far-1 at level 2 (-> near/(far-1) > 0: the raytraced sky wins over an empty raster)
        assert np.all(rt_depth[miss] == (f(cam.far_plane) + f(10.0) if lvl == 1 else f(cam.far_plane) - f(1.0)))
        left_miss = miss.copy()
        left_miss[:, W // 2:] = False
        assert left_miss.any()
        if lvl == 1:
            assert np.array_equal(p["rgba"][left_miss], raster[left_miss])
        else:
            assert np.array_equal(p["rgba"][left_miss], p3["rgba"][left_miss])


def test_orthographic_is_rejected(bvr, oracle, rtiow):
    cam = bvr.make_camera()
    cam.projection = 1
    try:
        oracle.render(rtiow.models, rtiow.materials, rtiow.nodes, cam, bvr.make_level(3), bvr.make_window(0.1, 8), 8)
    except RuntimeError as e:
        assert "3" in str(e)
    else:
        raise AssertionError("orthographic camera must be rejected")


def test_store_srgb8(oracle):
    x = np.array([[0.0, 0.0031308, 0.5, 1.0], [2.0, -1.0, 0.2, 0.5]], np.float32)
    out = oracle.store_srgb8(x)
    assert out.tolist() == [[0, 10, 188, 255], [255, 0, 124, 128]]


def test_exact_ties_go_to_the_leaf_popped_first(bvr, oracle):
    """raytrace.wgsl:354 accepts a hit only if it is strictly nearer, and raytrace.wgsl:329-341 pops child `index + 1`
    before child `index`: of two identical spheres in sibling leaves the reference reports the one in the SECOND leaf.
    (The kernels reproduce this through the model ranks, bvr_scene_traversal_ranks.)"""
    models = np.zeros(2, bvr.MODEL_DTYPE)
    models["position"][:] = (0.0, 0.0, -5.0)
    models["radius"][:] = 1.0
    models["material_id"] = [0, 1]
    mats = np.zeros(2, bvr.MATERIAL_DTYPE)
    mats["base_color"] = [(0.9, 0.1, 0.1), (0.1, 0.9, 0.1)]
    mats["roughness"] = 0.5
    mats["ior"] = 1.5
    nodes = np.zeros(3, bvr.BVH_NODE_DTYPE)
    nodes["bounds_min"][:] = (-1.1, -1.1, -6.1)
    nodes["bounds_max"][:] = (1.1, 1.1, -3.9)
    nodes["index"][0], nodes["model_count"][0] = 1, 0
    nodes["index"][1], nodes["model_count"][1] = 0, 1
    nodes["index"][2], nodes["model_count"][2] = 1, 1
    W, H = 16, 16
    cam = bvr.make_camera(position=(0, 0, 0), target=(0, 0, -1), aspect=1.0, sample_count=1, bounces=1)
    win = bvr.make_window(0.5, H)
    tree, _ = oracle.render(models, mats, nodes, cam, bvr.make_level(3), win, W)
    brute, _ = oracle.render(models, mats, nodes, cam, bvr.make_level(3), win, W, brute_force=True)
    hit = tree["primary_id"] != 0xFFFFFFFF
    assert hit.any() and np.array_equal(tree["primary_depth"], brute["primary_depth"])
    assert (tree["primary_id"][hit] == 1).all()        # second leaf: popped first
    assert (brute["primary_id"][hit] == 0).all()       # buffer order: first model first
    ranks, _ = bvr.traversal_ranks(nodes, 2)
    assert list(ranks) == [1, 0]
