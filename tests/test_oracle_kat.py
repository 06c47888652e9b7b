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


# ---------------------------------------------------------------------------------------------------------------------
# Per-function known answers for the path logic (raytrace.wgsl:139-299, 364-369, 400-416), derived here statement by
# statement in scalar numpy float32 — one IEEE operation per Python operator, in the WGSL's order.
# ---------------------------------------------------------------------------------------------------------------------
f = np.float32


def _dot(a, b):
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]


def _vec(*xs):
    return np.array(xs, np.float32)


def _normalize(v):
    return v / np.sqrt(_dot(v, v))


def _next_float(state):
    state = py_rng_next(state)
    return f(state) / f(4294967296.0), state


def _ball(state):
    """random.wgsl:17-26 on Python integers / f32 scalars; returns (p, state, iterations)."""
    it = 0
    while True:
        it += 1
        x, state = _next_float(state)
        y, state = _next_float(state)
        z, state = _next_float(state)
        p = f(2.0) * _vec(x, y, z) - f(1.0)
        if _dot(p, p) <= f(1.0):
            return p, state, it


def _material(bvr, **kw):
    m = np.zeros(1, bvr.MATERIAL_DTYPE)
    m["base_color"] = (0.8, 0.6, 0.2)
    m["roughness"], m["ior"] = 0.5, 1.5
    for k, v in kw.items():
        m[k] = v
    return m


def test_reflect_refract_reflectance_background(oracle):
    v = _normalize(_vec(0.3, -0.8, 0.52))
    n = _normalize(_vec(0.1, 1.0, -0.2))
    # reflect, raytrace.wgsl:400-402:  v - 2 * dot(v, n) * n
    want = v - (f(2.0) * _dot(v, n)) * n
    assert np.array_equal(oracle.reflect(v, n).view(np.uint32), want.view(np.uint32))
    # refract, raytrace.wgsl:404-409
    for ratio in (f(1.0) / f(1.5), f(1.5), f(0.7)):
        cos_theta = min(_dot(-v, n), f(1.0))
        perp = ratio * (v + cos_theta * n)
        parallel = -np.sqrt(np.abs(f(1.0) - _dot(perp, perp))) * n
        want = perp + parallel
        assert np.array_equal(oracle.refract(v, n, ratio).view(np.uint32), want.view(np.uint32))
    # reflectance, raytrace.wgsl:411-416 with pow(x, 5) = ((x x)(x x)) x
    for cosine, ri in ((f(0.0), f(1.5)), (f(0.37), f(1.0) / f(1.5)), (f(1.0), f(2.4)), (f(0.9), f(0.7))):
        r0 = (f(1.0) - ri) / (f(1.0) + ri)
        r0 = r0 * r0
        x = f(1.0) - cosine
        want = r0 + (f(1.0) - r0) * (((x * x) * (x * x)) * x)
        assert oracle.reflectance(cosine, ri) == want
    assert oracle.reflectance(1.0, 1.5) == (f(-0.5) / f(2.5)) * (f(-0.5) / f(2.5))   # head-on glass: r0 = ((1 - 1.5) / 2.5)^2 = 0.04
    # background_gradient, raytrace.wgsl:364-369
    for d in (_vec(0, 1, 0), _vec(0, -3, 0), _vec(1.0, 0.25, -2.0)):
        unit = _normalize(d)
        a = f(0.5) * (unit[1] + f(1.0))
        want = (f(1.0) - a) * _vec(1, 1, 1) + a * _vec(0.5, 0.7, 1.0)
        assert np.array_equal(oracle.background_gradient(d).view(np.uint32), want.view(np.uint32))
    assert np.array_equal(oracle.background_gradient(_vec(0, 1, 0)), _vec(0.5, 0.7, 1.0))     # straight up: sky blue
    assert np.array_equal(oracle.background_gradient(_vec(0, -1, 0)), _vec(1, 1, 1))          # straight down: white


def test_random_ray_from_uv(bvr, oracle):
    """raytrace.wgsl:139-156: two draws (x then y), jitter of one pixel, ray through the jittered ndc point."""
    W, H = 1280, 720
    cam = bvr.make_camera(position=(13, 2, 3), target=(0, 0, 0), fov=float(np.deg2rad(20)), aspect=W / H)
    win = bvr.make_window(0.5, H)
    direction, up, position = _vec(*cam.direction), _vec(*cam.up), _vec(*cam.position)
    for (x, y, state) in ((0, 0, 17), (639, 359, 15794417), (1279, 719, 63246312), (100, 650, 0xDEADBEEF)):
        u, v = (f(x) + f(0.5)) / f(W), (f(y) + f(0.5)) / f(H)
        r1, s1 = _next_float(state)
        r2, s2 = _next_float(s1)
        rand_square = (r1 - f(0.5), r2 - f(0.5))
        height = f(H)
        width = f(H) * f(cam.aspect)
        delta_u = (f(1.0) / width) * rand_square[0]
        delta_v = (f(1.0) / height) * rand_square[1]
        ndc_x = (u * f(2.0) - f(1.0)) + delta_u
        ndc_y = (f(1.0) - v * f(2.0)) + delta_v
        right = _vec(direction[1] * up[2] - direction[2] * up[1], direction[2] * up[0] - direction[0] * up[2],
                     direction[0] * up[1] - direction[1] * up[0])
        scale = f(np.tan(np.float64(f(cam.fov) * f(0.5))))
        want = _normalize((direction + (ndc_x * f(cam.aspect) * scale) * right) + (ndc_y * scale) * up)
        o, d, s_out = oracle.random_ray_from_uv(cam, win, u, v, state)
        assert s_out == s2
        assert np.array_equal(o, position) and np.array_equal(d.view(np.uint32), want.view(np.uint32))
        assert abs(float(_dot(d, d)) - 1.0) < 1e-6


def test_scatter_metal_branch(bvr, oracle):
    """raytrace.wgsl:234-246: u1 < metallic; direction = normalize(reflect(d, n)) + roughness * ball; the ball is drawn
    even when roughness is 0; absorbed iff the fuzzed direction points below the surface."""
    n = _vec(0, 1, 0)
    hit = _vec(1.5, 0.25, -2.0)
    d_in = _vec(0.6, -0.3, 0.2)                                   # not normalised: the reference never normalises bounces
    seen = set()
    for state in (1, 2, 3, 12345, 99, 1000, 31337, 5):
        for rough in (0.0, 0.5, 1.0):
            m = _material(bvr, metallic=1.0, roughness=rough)
            u1, s = _next_float(state)
            assert u1 < f(1.0) or u1 == f(1.0)
            if not (u1 < f(1.0)):
                continue
            b, s, _ = _ball(s)
            want = _normalize(d_in - (f(2.0) * _dot(d_in, n)) * n) + f(rough) * b
            absorbed, d, att, s_out = oracle.scatter(m, d_in, hit, n, True, state)
            assert s_out == s and np.array_equal(d.view(np.uint32), want.view(np.uint32))
            assert np.array_equal(att, m["base_color"][0])
            assert absorbed == bool(_dot(want, n) < f(0.0))
            seen.add(absorbed)
    # grazing incidence + full fuzz: some rays end up below the surface
    d_graze = _vec(1.0, -0.02, 0.0)
    for state in range(40):
        m = _material(bvr, metallic=1.0, roughness=1.0)
        absorbed, d, _, _ = oracle.scatter(m, d_graze, hit, n, True, state)
        assert absorbed == bool(_dot(d, n) < f(0.0))
        seen.add(absorbed)
    assert seen == {True, False}


def test_scatter_diffuse_branch(bvr, oracle):
    """raytrace.wgsl:283-298: u1 >= metallic, u2 >= transmission; direction = normal + ball1 + roughness * ball2 (drawn
    in that order), unnormalised; absorbed iff it points below the surface."""
    n = _normalize(_vec(0.2, 0.9, -0.3))
    hit = _vec(-4.0, 0.2, 1.0)
    d_in = _vec(0.1, -1.0, 0.3)
    seen = set()
    for state in range(1, 300):
        rough = (0.5, 1.0)[state % 2]
        m = _material(bvr, metallic=0.0, specular_transmission=0.0, roughness=rough)
        u1, s = _next_float(state)
        u2, s = _next_float(s)
        b1, s, _ = _ball(s)
        b2, s, _ = _ball(s)
        want = (n + b1) + f(rough) * b2
        absorbed, d, att, s_out = oracle.scatter(m, d_in, hit, n, True, state)
        assert s_out == s
        assert np.array_equal(d.view(np.uint32), want.view(np.uint32)) and np.array_equal(att, m["base_color"][0])
        assert absorbed == bool(_dot(want, n) < f(0.0))
        seen.add(absorbed)
    assert seen == {True, False}       # |ball1| + roughness |ball2| can exceed 1: a few per cent of diffuse bounces are absorbed
    # fractional metallic: the branch is picked by the first draw alone
    m = _material(bvr, metallic=0.3)
    kinds = set()
    for state in range(1, 40):
        u1, _ = _next_float(state)
        _, d, att, s_out = oracle.scatter(m, d_in, hit, n, True, state)
        s = py_rng_next(state)
        if u1 < f(0.3):
            b, s, _ = _ball(s)
            kinds.add("metal")
        else:
            s = py_rng_next(s)
            _, s, _ = _ball(s)
            _, s, _ = _ball(s)
            kinds.add("diffuse")
        assert s_out == s
    assert kinds == {"metal", "diffuse"}


def test_scatter_glass_branch(bvr, oracle):
    """raytrace.wgsl:248-282: ri = 1/ior on a front face, ior otherwise; total internal reflection takes NO draw
    (`||` short-circuits), otherwise one draw against Schlick; attenuation 1, never absorbed."""
    n = _vec(0, 0, 1)
    hit = _vec(0.5, 0.5, 3.0)
    taken = set()
    for ior, front, d_in in ((1.5, True, _vec(0.2, 0.1, -1.0)), (1.5, True, _vec(3.0, 0.0, -0.2)),
                             (1.5, False, _vec(0.9, 0.0, -0.3)),          # back face, ri = 1.5: total internal reflection
                             (0.7, True, _vec(1.0, 0.2, -0.25)),          # ior < 1 on a front face: ri = 1/0.7 > 1
                             (2.4, True, _vec(0.0, 0.0, -2.0))):
        for state in (7, 8, 9, 10, 11, 12, 4242, 777):
            m = _material(bvr, metallic=0.0, specular_transmission=1.0, ior=ior)
            u1, s = _next_float(state)
            u2, s = _next_float(s)
            if not (u2 < f(1.0)):
                continue
            ri = f(1.0) / f(ior) if front else f(ior)
            unit = _normalize(d_in)
            cos_theta = min(_dot(-unit, n), f(1.0))
            sin_theta = np.sqrt(f(1.0) - cos_theta * cos_theta)
            cannot = bool(ri * sin_theta > f(1.0))
            if cannot:
                reflects = True
                taken.add("tir")
            else:
                r0 = (f(1.0) - ri) / (f(1.0) + ri)
                r0 = r0 * r0
                x = f(1.0) - cos_theta
                schlick = r0 + (f(1.0) - r0) * (((x * x) * (x * x)) * x)
                u3, s = _next_float(s)
                reflects = bool(schlick > u3)
                taken.add("reflect" if reflects else "refract")
            if reflects:
                want = unit - (f(2.0) * _dot(unit, n)) * n
            else:
                ct = min(_dot(-unit, n), f(1.0))
                perp = ri * (unit + ct * n)
                want = perp + (-np.sqrt(np.abs(f(1.0) - _dot(perp, perp)))) * n
            absorbed, d, att, s_out = oracle.scatter(m, d_in, hit, n, front, state)
            assert not absorbed and np.array_equal(att, _vec(1, 1, 1))
            assert s_out == s, (ior, front, state)
            assert np.array_equal(d.view(np.uint32), want.view(np.uint32))
    assert taken == {"tir", "reflect", "refract"}


def test_raytrace_exits(bvr, oracle):
    """The three ways out of raytrace's loop (raytrace.wgsl:188-224), on one-sphere scenes through the whole fragment:
    miss -> gamma-encoded sky, depth = fallback; absorbed -> black; loop exhausted (bounce_count + 1 hits) -> black."""
    def scene(radius, **mat):
        models = np.zeros(1, bvr.MODEL_DTYPE)
        models["position"], models["radius"], models["material_id"] = (0, 0, -5), radius, 0
        mats = _material(bvr, **mat)
        nodes = np.zeros(1, bvr.BVH_NODE_DTYPE)
        nodes["bounds_min"], nodes["bounds_max"] = models["position"][0] - (radius + 0.1), models["position"][0] + (radius + 0.1)
        nodes["index"], nodes["model_count"] = 0, 1
        return models, mats, nodes
    W = H = 9
    far = 1000.0
    # miss: a pixel in the corner never meets the small sphere; level 1 falls back to far + 10, the others to far - 1
    models, mats, nodes = scene(0.5)
    for level, fallback in ((1, f(far) + f(10.0)), (2, f(far) - f(1.0)), (3, f(far) - f(1.0))):
        cam = bvr.make_camera(position=(0, 0, 0), target=(0, 0, -1), aspect=1.0, sample_count=1, bounces=3, far=far)
        raster, depth = np.zeros((H, W, 4), np.float32), np.zeros((H, W), np.float32)
        planes, _ = oracle.render(models, mats, nodes, cam, bvr.make_level(level), bvr.make_window(0.4, H), W, raster, depth)
        assert planes["primary_id"][0, 0] == 0xFFFFFFFF and planes["rt_depth"][0, 0] == fallback
        if level == 3:
            # the colour of a miss is sqrt(1 * sky(direction)) per channel: recompute it from the camera ray
            u, v = (f(0) + f(0.5)) / f(W), (f(0) + f(0.5)) / f(H)
            seed = np_pixel_seed(0.4, 0, 0, W, H)
            _, d, _ = oracle.random_ray_from_uv(cam, bvr.make_window(0.4, H), u, v, seed)
            sky = oracle.background_gradient(d)
            assert np.array_equal(planes["rgba"][0, 0, :3].view(np.uint32), np.sqrt(sky).view(np.uint32))
            assert planes["rgba"][0, 0, 3] == 1.0
    # loop exhausted: camera INSIDE a big mirror-less glass... simpler: bounces = 0 and a hit -> one scatter, then black
    models, mats, nodes = scene(2.0)
    cam = bvr.make_camera(position=(0, 0, 0), target=(0, 0, -1), aspect=1.0, sample_count=3, bounces=0)
    planes, cnt = oracle.render(models, mats, nodes, cam, bvr.make_level(3), bvr.make_window(0.4, H), W)
    c = H // 2
    assert planes["primary_id"][c, c] == 0 and np.array_equal(planes["rgba"][c, c], _vec(0, 0, 0, 1))
    assert np.isclose(planes["rt_depth"][c, c], 3.0, atol=0.01) and 3.0 <= planes["primary_depth"][c, c] < 3.01   # jittered centre pixel
    # absorbed: a fully fuzzed metal seen at grazing incidence absorbs some samples -> their contribution is 0, the
    # others see sky through the tinted reflection; with one sample per pixel a pixel is either exactly black or not
    models, mats, nodes = scene(2.0, metallic=1.0, roughness=1.0)
    cam = bvr.make_camera(position=(0, 0, 0), target=(0, 0, -1), aspect=1.0, sample_count=1, bounces=50)
    planes, cnt = oracle.render(models, mats, nodes, cam, bvr.make_level(3), bvr.make_window(0.9, 64), 64)
    on_sphere = planes["primary_id"] == 0
    black = (planes["rgba"][..., :3] == 0).all(axis=2)
    assert (black & on_sphere).sum() > 0 and (~black & on_sphere).sum() > 0 and not (black & ~on_sphere).any()
