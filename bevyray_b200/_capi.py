"""ctypes binding of the C ABI (include/bevyray_b200.h, include/bevyray_b200_host.h).

The shared library is built in-tree by `make` / `__graft_entry__.build()`.  There is no Python or CPU
fallback: if the library is missing, importing this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BEVYRAY_B200_LIB: another build of the same library (A/B runs of compile-time variants); default = the in-tree build
LIB_PATH = os.environ.get("BEVYRAY_B200_LIB") or os.path.join(_HERE, "libbevyray_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
        "bevyray_b200 has no CPU fallback."
    )

lib = C.CDLL(LIB_PATH)

# ---- status codes (BvrStatus) ----
BVR_OK = 0
BVR_ERR_INVALID_ARGUMENT = 1
BVR_ERR_CUDA = 2
BVR_ERR_UNSUPPORTED_PROJECTION = 3
BVR_ERR_NO_SCENE = 4
BVR_ERR_BAD_SCENE = 5
BVR_ERR_OUT_OF_MEMORY = 6
BVR_ERR_NO_DEVICE = 7

# BvrRaytracing (src/raytracing/mod.rs:94-101)
RAYTRACING_SKIP, RAYTRACING_FALLBACK_RASTER, RAYTRACING_FALLBACK_RAYTRACED, RAYTRACING_PURE = 0, 1, 2, 3
# BvrKernel / BvrTraversal
KERNEL_AUTO, KERNEL_MEGAKERNEL, KERNEL_WAVEFRONT, KERNEL_CTA_WAVEFRONT = 0, 1, 2, 3
TRAVERSAL_AUTO, TRAVERSAL_REFERENCE_ORDER = 0, 1
ARRAY_MODELS, ARRAY_MATERIALS, ARRAY_BVH_NODES = 0, 1, 2
RENDER_DEFER_COMPOSITE = 1
RENDER_EXTRA_SAMPLE = 2


def render_extra_sample_bits(modulus, phase, count):
    """BVR_RENDER_EXTRA_SAMPLE_BITS of include/bevyray_b200.h"""
    return RENDER_EXTRA_SAMPLE | (int(modulus) << 8) | (int(phase) << 16) | (int(count) << 24)



class BvrModel(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("radius", C.c_float), ("material_id", C.c_uint32), ("_pad", C.c_uint32 * 3)]


class BvrMaterial(C.Structure):
    _fields_ = [("base_color", C.c_float * 3), ("metallic", C.c_float), ("roughness", C.c_float),
                ("reflectance", C.c_float), ("ior", C.c_float), ("specular_transmission", C.c_float)]


class BvrBvhNode(C.Structure):
    _fields_ = [("bounds_min", C.c_float * 3), ("_pad0", C.c_uint32), ("bounds_max", C.c_float * 3),
                ("index", C.c_uint32), ("model_count", C.c_uint32), ("_pad1", C.c_uint32 * 3)]


class BvrCamera(C.Structure):
    _fields_ = [("sample_count", C.c_uint32), ("bounce_count", C.c_uint32), ("projection", C.c_uint32),
                ("near_plane", C.c_float), ("far_plane", C.c_float), ("fov", C.c_float), ("aspect", C.c_float),
                ("_pad0", C.c_uint32), ("position", C.c_float * 3), ("_pad1", C.c_uint32),
                ("direction", C.c_float * 3), ("_pad2", C.c_uint32), ("up", C.c_float * 3), ("_pad3", C.c_uint32)]


class BvrRaytraceLevel(C.Structure):
    _fields_ = [("level", C.c_uint32), ("_pad0", C.c_uint32 * 3), ("_padding", C.c_float * 3), ("_pad1", C.c_uint32)]


class BvrWindow(C.Structure):
    _fields_ = [("random_seed", C.c_float), ("height", C.c_uint32), ("_padding", C.c_float * 2)]


class BvrDirtyRange(C.Structure):
    _fields_ = [("array", C.c_uint32), ("first", C.c_uint32), ("count", C.c_uint32)]


class BvrRenderOptions(C.Structure):
    _fields_ = [("width", C.c_uint32), ("kernel", C.c_uint32), ("traversal", C.c_uint32),
                ("shard_index", C.c_uint32), ("shard_count", C.c_uint32), ("strip_rows", C.c_uint32),
                ("flags", C.c_uint32), ("output_weight", C.c_float)]


class BvrOutputs(C.Structure):
    _fields_ = [("rgba", C.c_void_p), ("rt_depth", C.c_void_p), ("primary_id", C.c_void_p),
                ("primary_depth", C.c_void_p), ("srgb8", C.c_void_p)]


class BvrStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("paths", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("last_render_ms", C.c_float),
                ("last_upload_ms", C.c_float), ("selfcheck_rays", C.c_uint64), ("selfcheck_mismatches", C.c_uint64)]


class BvrhStandardMaterial(C.Structure):
    _fields_ = [("base_color_srgb", C.c_float * 3), ("metallic", C.c_float), ("perceptual_roughness", C.c_float),
                ("reflectance", C.c_float), ("ior", C.c_float), ("specular_transmission", C.c_float)]


assert C.sizeof(BvrModel) == 32 and C.sizeof(BvrMaterial) == 32 and C.sizeof(BvrBvhNode) == 48
assert C.sizeof(BvrCamera) == 80 and C.sizeof(BvrRaytraceLevel) == 32 and C.sizeof(BvrWindow) == 16

PEER_HANDLE_BYTES = 64

_P = C.POINTER
_vp, _u32, _u64, _sz, _f, _i = C.c_void_p, C.c_uint32, C.c_uint64, C.c_size_t, C.c_float, C.c_int

# name -> (restype, argtypes).  This table is also what tests/test_abi.py checks against the headers.
SIGNATURES = {
    # include/bevyray_b200.h
    "bvr_abi_version": (_u32, []),
    "bvr_status_string": (C.c_char_p, [_i]),
    "bvr_create": (_i, [_i, _P(_vp)]),
    "bvr_destroy": (None, [_vp]),
    "bvr_last_error": (C.c_char_p, [_vp]),
    "bvr_set_stream": (_i, [_vp, _vp]),
    "bvr_sync": (_i, [_vp]),
    "bvr_reload_tuning": (_i, [_vp]),
    "bvr_upload_scene": (_i, [_vp, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _sz]),
    "bvr_upload_scene_gpu_bvh": (_i, [_vp, _vp, _sz, _vp, _sz, _vp, _sz, _vp]),
    "bvr_refit_scene_gpu_bvh": (_i, [_vp, _vp, _sz, _vp, _sz, _vp, _sz, _vp]),
    "bvr_shard_rows": (_u32, [_u32, _P(BvrRenderOptions)]),
    "bvr_scene_traversal_ranks": (_i, [_vp, _sz, _sz, _vp, _vp]),
    "bvr_render": (_i, [_vp, _P(BvrCamera), _P(BvrRaytraceLevel), _P(BvrWindow), _P(BvrRenderOptions), _vp, _vp, _P(BvrOutputs)]),
    "bvr_render_async": (_i, [_vp, _P(BvrCamera), _P(BvrRaytraceLevel), _P(BvrWindow), _P(BvrRenderOptions), _vp, _vp, _P(BvrOutputs)]),
    "bvr_render_device": (_i, [_vp, _P(BvrCamera), _P(BvrRaytraceLevel), _P(BvrWindow), _P(BvrRenderOptions), _vp, _vp, _P(BvrOutputs)]),
    "bvr_axpby_device": (_i, [_vp, _vp, _f, _vp, _f, _sz]),
    "bvr_composite_device": (_i, [_vp, _P(BvrCamera), _P(BvrRaytraceLevel), _vp, _vp, _vp, _vp, _sz]),
    "bvr_unshard_device": (_i, [_vp, _vp, _sz, _vp, _u32, _u32, _u32, _u32, _u32]),
    "bvr_peer_alloc": (_i, [_vp, _sz, _P(_vp), _vp]),
    "bvr_peer_open": (_i, [_vp, _vp, _P(_vp)]),
    "bvr_peer_close": (_i, [_vp, _vp]),
    "bvr_peer_free": (_i, [_vp, _vp]),
    "bvr_sum_slots_device": (_i, [_vp, _vp, _sz, _u32, _u64, _vp, _sz]),
    "bvr_get_stats": (_i, [_vp, _P(BvrStats)]),
    "bvr_bench_fp32_peak": (_i, [_i, _P(_f)]),
    "bvr_bench_l2_bandwidth": (_i, [_i, _P(_f)]),
    # include/bevyray_b200_host.h
    "bvrh_scene_rtiow": (_vp, [_u64]),
    "bvrh_scene_random": (_vp, [_u64, _u32, _f, _f, _f]),
    "bvrh_scene_from_models": (_vp, [_vp, _sz, _vp, _sz]),
    "bvrh_scene_animate": (_i, [_vp, _u32]),
    "bvrh_scene_animate_models": (_i, [_vp, _u32]),
    "bvrh_scene_free": (None, [_vp]),
    "bvrh_scene_n_models": (_sz, [_vp]),
    "bvrh_scene_n_materials": (_sz, [_vp]),
    "bvrh_scene_n_nodes": (_sz, [_vp]),
    "bvrh_scene_models": (_vp, [_vp]),
    "bvrh_scene_materials": (_vp, [_vp]),
    "bvrh_scene_nodes": (_vp, [_vp]),
    "bvrh_build_ploc": (_sz, [_vp, _sz, _u32, _vp]),
    "bvrh_validate_bvh": (_i, [_vp, _sz, _vp, _sz, C.c_char_p, _sz]),
    "bvrh_camera_look_at": (None, [_P(_f), _P(_f), _P(_f), _f, _f, _f, _f, _u32, _u32, _P(BvrCamera)]),
    "bvrh_srgb_to_linear": (_f, [_f]),
    "bvrh_app_create": (_vp, []),
    "bvrh_app_destroy": (None, [_vp]),
    "bvrh_app_last_error": (C.c_char_p, [_vp]),
    "bvrh_app_add_raytrace_plugin": (_i, [_vp, _i]),
    "bvrh_app_setup_demo": (_u32, [_vp, _u64]),
    "bvrh_app_standard_material_default": (None, [_P(BvrhStandardMaterial)]),
    "bvrh_app_spawn_window": (_u32, [_vp, _u32, _u32]),
    "bvrh_app_spawn_sphere": (_u32, [_vp, _f, _f, _f, _f, _P(BvrhStandardMaterial)]),
    "bvrh_app_spawn_camera": (_u32, [_vp, _P(_f), _P(_f), _P(_f), _f, _f, _f, _f, _u32, _u32, _u32, _i]),
    "bvrh_app_set_raytraced_camera": (_i, [_vp, _u32, _u32, _u32, _u32]),
    "bvrh_app_set_translation": (_i, [_vp, _u32, _f, _f, _f]),
    "bvrh_app_set_material": (_i, [_vp, _u32, _P(BvrhStandardMaterial)]),
    "bvrh_app_set_window_size": (None, [_vp, _u32, _u32]),
    "bvrh_app_set_seed": (None, [_vp, _f]),
    "bvrh_app_set_render_options": (None, [_vp, _P(BvrRenderOptions)]),
    "bvrh_app_set_gpu_bvh": (None, [_vp, _i]),
    "bvrh_app_set_raster": (_i, [_vp, _u32, _vp, _vp, _sz]),
    "bvrh_app_update": (_i, [_vp]),
    "bvrh_app_frame": (_vp, [_vp, _u32, _P(_u32), _P(_u32)]),
    "bvrh_app_buffers": (_sz, [_vp, _P(_vp), _P(_vp), _P(_vp), _P(_sz)]),
    "bvrh_app_msaa_off": (_i, [_vp]),
    "bvrh_app_has_depth_prepass": (_i, [_vp, _u32]),
    "bvrh_app_get_stats": (_i, [_vp, _P(BvrStats)]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)   # AttributeError here = the library does not export a declared symbol
    _fn.restype = _res
    _fn.argtypes = _args
