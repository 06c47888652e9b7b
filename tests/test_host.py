"""Host layer (C++ mirror of src/raytracing/{mod,extract}.rs and src/main.rs): BVH producer contract,
scene recipe, extract semantics, plugin surface.  CPU only."""
import ctypes as C

import numpy as np
import pytest


def test_ploc_contract_on_demo_scene(bvr, rtiow):
    m, nodes = rtiow.models, rtiow.nodes
    assert len(nodes) == 2 * len(m) - 1
    assert bvr.validate_bvh(nodes, m) is None
    # PLOC leaves hold exactly one primitive whose index is the model index (extract.rs:323-332 relies on it)
    leaves = nodes[nodes["model_count"] > 0]
    assert np.all(leaves["model_count"] == 1)
    assert sorted(leaves["index"].tolist()) == list(range(len(m)))
    # leaf bounds = centre +- (radius + 0.1), extract.rs:220-227
    pad = (m["radius"] + np.float32(0.1))[:, None]
    want_min, want_max = m["position"] - pad, m["position"] + pad
    order = leaves["index"]
    assert np.array_equal(leaves["bounds_min"], want_min[order]) and np.array_equal(leaves["bounds_max"], want_max[order])
    # root is node 0 and encloses everything; children of inner nodes are adjacent
    assert nodes["model_count"][0] == 0
    assert np.all(nodes["bounds_min"][0] <= nodes["bounds_min"].min(axis=0))
    assert np.all(nodes["bounds_max"][0] >= nodes["bounds_max"].max(axis=0))


@pytest.mark.parametrize("n", [0, 1, 2, 3, 7, 100, 1000])
def test_ploc_sizes_and_validity(bvr, n):
    rs = np.random.RandomState(n)
    models = np.zeros(n, bvr.MODEL_DTYPE)
    models["position"] = rs.uniform(-10, 10, (n, 3)).astype(np.float32)
    models["radius"] = rs.uniform(0.05, 0.5, n).astype(np.float32)
    models["material_id"] = np.arange(n)
    nodes = bvr.build_ploc(models)
    assert len(nodes) == max(2 * n - 1, 0)
    assert bvr.validate_bvh(nodes, models) is None
    if n == 1:
        assert nodes["model_count"][0] == 1 and nodes["index"][0] == 0


def test_ploc_coincident_spheres(bvr):
    """Degenerate input: all centres identical (zero Morton extent)."""
    models = np.zeros(33, bvr.MODEL_DTYPE)
    models["radius"] = 0.25
    nodes = bvr.build_ploc(models)
    assert bvr.validate_bvh(nodes, models) is None


def test_validate_bvh_rejects_broken_trees(bvr, rtiow):
    nodes = rtiow.nodes.copy()
    nodes["index"][0] = len(nodes)            # child out of range
    assert bvr.validate_bvh(nodes, rtiow.models) is not None
    nodes = rtiow.nodes.copy()
    leaf = int(np.nonzero(nodes["model_count"] > 0)[0][0])
    nodes["bounds_max"][leaf] -= 1.0          # no longer encloses its sphere
    assert bvr.validate_bvh(nodes, rtiow.models) is not None


def test_demo_scene_recipe(bvr):
    """src/main.rs:87-239: ground r=1000, 22x23 grid candidates minus the exclusion zone around (4,0.2,0),
    three r=1 spheres; 80/15/5 % diffuse/metal/glass; one material per sphere, material_id == index."""
    s = bvr.Scene.rtiow(1)
    m, mats = s.models, s.materials
    n = len(m)
    assert 23 * 22 + 4 - 12 <= n <= 23 * 22 + 4
    assert len(mats) == n and np.array_equal(m["material_id"], np.arange(n))
    assert m["radius"][0] == 1000.0 and tuple(m["position"][0]) == (0.0, -1000.0, 0.0)
    assert np.all(m["radius"][1:-3] == np.float32(0.2)) and np.all(m["position"][1:-3, 1] == np.float32(0.2))
    assert np.all(m["radius"][-3:] == 1.0)
    assert [tuple(p) for p in m["position"][-3:]] == [(0, 1, 0), (-4, 1, 0), (4, 1, 0)]
    small = m["position"][1:-3]
    assert np.all(np.linalg.norm(small - np.array([4, 0.2, 0], np.float32), axis=1) > 0.9)
    # ground: Color::srgb(0.5,0.5,0.5).to_linear(), StandardMaterial defaults for the rest
    lin = bvr.capi.lib.bvrh_srgb_to_linear(0.5)
    assert abs(lin - 0.21404114) < 1e-6
    assert np.allclose(mats["base_color"][0], lin) and mats["roughness"][0] == 0.5 and mats["reflectance"][0] == 0.5
    assert mats["ior"][0] == 1.5 and mats["specular_transmission"][0] == 0.0 and mats["metallic"][0] == 0.0
    # the three big spheres: glass, diffuse (0.4,0.2,0.1), metal (0.7,0.6,0.5) roughness 0
    assert mats["specular_transmission"][-3] == 1.0 and mats["metallic"][-1] == 1.0 and mats["roughness"][-1] == 0.0
    glass = mats["specular_transmission"][1:-3] == 1.0
    metal = mats["metallic"][1:-3] == 1.0
    assert 0.01 < glass.mean() < 0.12 and 0.08 < metal.mean() < 0.25
    # deterministic in the seed, different across seeds
    s2, s3 = bvr.Scene.rtiow(1), bvr.Scene.rtiow(2)
    assert s.models.tobytes() == s2.models.tobytes() and s.nodes.tobytes() == s2.nodes.tobytes()
    assert s.models.tobytes() != s3.models.tobytes()


def test_camera_extract(bvr):
    """CameraExtract::extract_component (extract.rs:118-146) for the demo camera (main.rs:57-58)."""
    cam = bvr.make_camera(sample_count=4, bounces=4)
    assert list(cam.position) == [0.0, 0.0, 5.0] and list(cam.direction) == [0.0, 0.0, -1.0] and list(cam.up) == [0.0, 1.0, 0.0]
    assert cam.projection == 0 and abs(cam.fov - np.pi / 4) < 1e-7 and cam.near_plane == np.float32(0.1) and cam.far_plane == 1000.0
    assert cam.sample_count == 4 and cam.bounce_count == 4
    book = bvr.make_camera(position=(13, 2, 3), target=(0, 0, 0), fov=np.deg2rad(20))
    d, u = np.array(list(book.direction)), np.array(list(book.up))
    want = -np.array([13, 2, 3]) / np.linalg.norm([13, 2, 3])
    assert np.allclose(d, want, atol=1e-6) and abs(np.linalg.norm(u) - 1) < 1e-6 and abs(d @ u) < 1e-6


def test_plugin_surface_without_gpu(bvr):
    """RaytracePlugin::build (mod.rs:27-72): Msaa::Off, DepthPrepass auto-inserted on cameras, pipeline
    creation fails loudly without a device.  The frame is then skipped, never rendered on the CPU."""
    import torch
    lib = bvr.capi.lib
    app = C.c_void_p(lib.bvrh_app_create())
    try:
        assert lib.bvrh_app_msaa_off(app) == 0
        cam = lib.bvrh_app_setup_demo(app, 1)
        assert lib.bvrh_app_has_depth_prepass(app, cam) == 0
        st = lib.bvrh_app_add_raytrace_plugin(app, 0)
        assert lib.bvrh_app_msaa_off(app) == 1            # mod.rs:30
        if not torch.cuda.is_available():
            assert st == bvr.capi.BVR_ERR_NO_DEVICE
            lib.bvrh_app_set_seed(app, 0.37)
            assert lib.bvrh_app_update(app) == 0           # node returns early: pipeline not ready (pipeline.rs:82-85)
            assert lib.bvrh_app_has_depth_prepass(app, cam) == 1   # auto_add_camera_components, mod.rs:108-115
            models, mats, nodes, nn = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_size_t()
            n = lib.bvrh_app_buffers(app, C.byref(models), C.byref(mats), C.byref(nodes), C.byref(nn))
            ref = bvr.Scene.rtiow(1)
            assert n == len(ref.models) and nn.value == len(ref.nodes)   # prepare_buffers ran (extract.rs:280-337)
            got = np.frombuffer((C.c_char * (n * 32)).from_address(models.value), dtype=bvr.MODEL_DTYPE)
            assert got.tobytes() == ref.models.tobytes()
    finally:
        lib.bvrh_app_destroy(app)


def test_standard_material_defaults(bvr):
    m = bvr.capi.BvrhStandardMaterial()
    bvr.capi.lib.bvrh_app_standard_material_default(C.byref(m))
    assert list(m.base_color_srgb) == [1.0, 1.0, 1.0]
    assert (m.metallic, m.perceptual_roughness, m.reflectance, m.ior, m.specular_transmission) == (0.0, 0.5, 0.5, 1.5, 0.0)


def test_random_scene_and_animation(bvr):
    s = bvr.Scene.random(5, 2000, 40.0, 0.05, 0.25)
    assert len(s.models) == 2000 and len(s.nodes) == 3999 and bvr.validate_bvh(s.nodes, s.models) is None
    base = s.models.copy()
    s.animate(10)
    moved = np.any(s.models["position"] != base["position"], axis=1)
    assert moved.sum() == 500 and np.all(np.nonzero(moved)[0] % 4 == 0)
    assert bvr.validate_bvh(s.nodes, s.models) is None
    s.animate(0)
    assert np.allclose(s.models["position"][1], base["position"][1])


def _reference_leaf_order(nodes):
    """The order in which raytrace.wgsl:313-346 reaches the leaves when nothing is culled: the stack starts with
    node 0; an inner node pushes `index` and then `index + 1`, so `index + 1` is popped first."""
    order, stack = [], [0]
    while stack:
        i = stack.pop()
        if nodes["model_count"][i] > 0:
            order.extend(range(int(nodes["index"][i]), int(nodes["index"][i]) + int(nodes["model_count"][i])))
        else:
            stack.append(int(nodes["index"][i]))
            stack.append(int(nodes["index"][i]) + 1)
    return order


def test_traversal_ranks_follow_the_reference_order(bvr):
    """bvr_scene_traversal_ranks (host only): every model's position in the reference's traversal order, which
    the kernels use to resolve bit-exact ties in t the way the reference's strict `<` does."""
    for scene in (bvr.Scene.rtiow(1), bvr.Scene.random(3, 777, 20.0, 0.05, 0.4)):
        ranks, depth = bvr.traversal_ranks(scene.nodes, len(scene.models))
        order = _reference_leaf_order(scene.nodes)
        assert sorted(order) == list(range(len(scene.models)))
        want = np.empty(len(order), np.uint32)
        want[order] = np.arange(len(order), dtype=np.uint32)
        assert np.array_equal(ranks, want)
        assert 1 < depth < 64
    # multi-model leaves keep buffer order inside the leaf; models no leaf holds get 0xFFFFFFFF
    nodes = np.zeros(3, bvr.BVH_NODE_DTYPE)
    nodes["index"][0], nodes["model_count"][0] = 1, 0
    nodes["index"][1], nodes["model_count"][1] = 0, 3
    nodes["index"][2], nodes["model_count"][2] = 4, 2
    ranks, depth = bvr.traversal_ranks(nodes, 7)
    assert list(ranks) == [2, 3, 4, 0xFFFFFFFF, 0, 1, 0xFFFFFFFF] and depth == 2
    # a cycle is rejected like at upload time
    nodes["index"][0] = 0
    with pytest.raises(bvr.BvrError) as e:
        bvr.traversal_ranks(nodes, 7)
    assert e.value.status == bvr.capi.BVR_ERR_BAD_SCENE


def test_unreachable_nodes_are_validated_too(bvr, rtiow):
    """The derive kernels run over EVERY node of the array, so a node the root does not reach must still be in range
    (and the GPU validator, scene_validate.cu: gv_link, checks every node as well): same verdict on both paths."""
    n = len(rtiow.nodes)
    for mutate in ("child", "leaf", "leaf_count", "second_parent"):
        nodes = np.zeros(n + 3, bvr.BVH_NODE_DTYPE)
        nodes[:n] = rtiow.nodes
        # three extra nodes the root never reaches: one inner node with two leaf children — valid on their own
        nodes["index"][n], nodes["model_count"][n] = n + 1, 0
        nodes["index"][n + 1], nodes["model_count"][n + 1] = 0, 1
        nodes["index"][n + 2], nodes["model_count"][n + 2] = 1, 1
        ranks, depth = bvr.traversal_ranks(nodes, len(rtiow.models))
        want, want_depth = bvr.traversal_ranks(rtiow.nodes, len(rtiow.models))
        assert np.array_equal(ranks, want) and depth == want_depth
        if mutate == "child":
            nodes["index"][n] = 0x7ffffff0                      # child far outside the array
        elif mutate == "leaf":
            nodes["index"][n + 1] = len(rtiow.models) + 5        # model range outside the model buffer
        elif mutate == "leaf_count":
            nodes["model_count"][n + 2] = 129
        else:
            nodes["index"][n] = int(rtiow.nodes["index"][0])     # claims the root's children: two parents
        with pytest.raises(bvr.BvrError) as e:
            bvr.traversal_ranks(nodes, len(rtiow.models))
        assert e.value.status == bvr.capi.BVR_ERR_BAD_SCENE, mutate
