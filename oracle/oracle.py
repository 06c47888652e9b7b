"""ctypes wrapper of the CPU oracle (oracle/libbvr_oracle.so).  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs — never by
bevyray_b200/."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libbvr_oracle.so")


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


if not os.path.exists(_LIB):
    build()
lib = C.CDLL(_LIB)


# The uniforms `fragment` reads, as the reference's encase layouts (extract.rs:56-61, 83-104; raytrace.wgsl:30-54).  The
# oracle's own mirrors: a caller may pass these or any other object with the same bytes (the product's ctypes structs).
class Camera(C.Structure):
    _fields_ = [("sample_count", C.c_uint32), ("bounce_count", C.c_uint32), ("projection", C.c_uint32),
                ("near_plane", C.c_float), ("far_plane", C.c_float), ("fov", C.c_float), ("aspect", C.c_float),
                ("_pad0", C.c_uint32), ("position", C.c_float * 3), ("_pad1", C.c_uint32),
                ("direction", C.c_float * 3), ("_pad2", C.c_uint32), ("up", C.c_float * 3), ("_pad3", C.c_uint32)]


class RaytraceLevel(C.Structure):
    _fields_ = [("level", C.c_uint32), ("_pad0", C.c_uint32 * 3), ("_padding", C.c_float * 3), ("_pad1", C.c_uint32)]


class Window(C.Structure):
    _fields_ = [("random_seed", C.c_float), ("height", C.c_uint32), ("_padding", C.c_float * 2)]


assert C.sizeof(Camera) == 80 and C.sizeof(RaytraceLevel) == 32 and C.sizeof(Window) == 16

MODEL_DTYPE = np.dtype({"names": ["position", "radius", "material_id"], "formats": [("<f4", 3), "<f4", "<u4"],
                        "offsets": [0, 12, 16], "itemsize": 32})
MATERIAL_DTYPE = np.dtype({"names": ["base_color", "metallic", "roughness", "reflectance", "ior", "specular_transmission"],
                           "formats": [("<f4", 3), "<f4", "<f4", "<f4", "<f4", "<f4"], "offsets": [0, 12, 16, 20, 24, 28],
                           "itemsize": 32})
BVH_NODE_DTYPE = np.dtype({"names": ["bounds_min", "bounds_max", "index", "model_count"],
                           "formats": [("<f4", 3), ("<f4", 3), "<u4", "<u4"], "offsets": [0, 16, 28, 32], "itemsize": 48})


def make_level(level):
    lv = RaytraceLevel()
    lv.level = int(level)
    return lv


def make_window(random_seed, height):
    w = Window()
    w.random_seed = float(random_seed)
    w.height = int(height)
    return w


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("rays", "paths", "node_pops", "inner_visits", "box_tests", "sphere_tests",
                                          "hits_shaded", "rng_draws", "stack_truncations", "max_stack")]


_vp = C.c_void_p
lib.bvro_rng_next_int.restype = C.c_uint32
lib.bvro_rng_next_int.argtypes = [C.c_uint32]
lib.bvro_rng_float_of_state.restype = C.c_float
lib.bvro_rng_float_of_state.argtypes = [C.c_uint32]
lib.bvro_pixel_seed.restype = C.c_uint32
lib.bvro_pixel_seed.argtypes = [C.c_float, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
lib.bvro_tan_half_fov.restype = C.c_float
lib.bvro_tan_half_fov.argtypes = [C.c_float]
lib.bvro_hit_sphere.restype = C.c_float
lib.bvro_hit_sphere.argtypes = [_vp, _vp, _vp]
lib.bvro_ray_bounding_dst.restype = C.c_float
lib.bvro_ray_bounding_dst.argtypes = [_vp, _vp, _vp, _vp]
lib.bvro_render.restype = C.c_int
lib.bvro_render.argtypes = [_vp, C.c_size_t, _vp, C.c_size_t, _vp, C.c_size_t, _vp, _vp, _vp,
                            C.c_uint32, C.c_uint32, C.c_uint32, _vp, _vp, _vp, _vp, _vp, _vp,
                            C.c_int, C.c_int, C.POINTER(Counters)]
lib.bvro_store_srgb8.restype = None
lib.bvro_store_srgb8.argtypes = [_vp, C.c_size_t, _vp]
lib.bvro_max_threads.restype = C.c_int
lib.bvro_reflect.restype = None
lib.bvro_reflect.argtypes = [_vp, _vp, _vp]
lib.bvro_refract.restype = None
lib.bvro_refract.argtypes = [_vp, _vp, C.c_float, _vp]
lib.bvro_reflectance.restype = C.c_float
lib.bvro_reflectance.argtypes = [C.c_float, C.c_float]
lib.bvro_background_gradient.restype = None
lib.bvro_background_gradient.argtypes = [_vp, _vp]
lib.bvro_random_ray_from_uv.restype = None
lib.bvro_random_ray_from_uv.argtypes = [_vp, _vp, C.c_float, C.c_float, C.POINTER(C.c_uint32), _vp, _vp]
lib.bvro_scatter.restype = C.c_int
lib.bvro_scatter.argtypes = [_vp, _vp, _vp, _vp, C.c_int, C.POINTER(C.c_uint32), _vp, _vp]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_vp)


def rng_sequence(state, n):
    out = []
    for _ in range(n):
        state = lib.bvro_rng_next_int(state)
        out.append(state)
    return out


def render(models, materials, nodes, camera, level, window, width, raster_rgba=None, raster_depth=None,
           rows=None, brute_force=False, threads=0):
    """One `fragment` invocation per pixel.  camera/level/window are the ctypes structs of the C ABI
    (any object exposing the same bytes).  Returns (planes dict, counters dict); planes are full-image."""
    height = window.height
    y0, y1 = (0, height) if rows is None else rows
    rgba = np.zeros((height, width, 4), np.float32)
    rt_depth = np.zeros((height, width), np.float32)
    pid = np.full((height, width), 0xFFFFFFFF, np.uint32)
    pdepth = np.zeros((height, width), np.float32)
    models = np.ascontiguousarray(models)
    materials = np.ascontiguousarray(materials)
    nodes = np.ascontiguousarray(nodes)
    # the reference's encase strides (extract.rs:181-237); numpy silently packs structured arrays in some operations
    assert models.dtype.itemsize == 32 and materials.dtype.itemsize == 32 and nodes.dtype.itemsize == 48, \
        "scene arrays must keep the 32/32/48-byte record layout"
    if raster_rgba is not None:
        raster_rgba = np.ascontiguousarray(raster_rgba, np.float32)
    if raster_depth is not None:
        raster_depth = np.ascontiguousarray(raster_depth, np.float32)
    cnt = Counters()
    lvl = level
    rc = lib.bvro_render(_ptr(models), len(models), _ptr(materials), len(materials), _ptr(nodes), len(nodes),
                         C.addressof(camera), C.addressof(lvl), C.addressof(window), width, y0, y1,
                         _ptr(raster_rgba), _ptr(raster_depth), _ptr(rgba), _ptr(rt_depth), _ptr(pid), _ptr(pdepth),
                         1 if brute_force else 0, threads, C.byref(cnt))
    if rc != 0:
        raise RuntimeError(f"bvro_render failed with status {rc}")
    planes = {"rgba": rgba, "rt_depth": rt_depth, "primary_id": pid, "primary_depth": pdepth}
    return planes, {k: getattr(cnt, k) for k, _ in cnt._fields_}


def store_srgb8(rgba):
    rgba = np.ascontiguousarray(rgba, np.float32)
    out = np.empty(rgba.shape, np.uint8)
    lib.bvro_store_srgb8(_ptr(rgba), rgba.size // 4, _ptr(out))
    return out


def max_threads():
    return lib.bvro_max_threads()


# ---- single shader functions (known-answer tests) ----
def _f3(v):
    return np.ascontiguousarray(v, np.float32)


def reflect(v, n):
    out = np.zeros(3, np.float32)
    lib.bvro_reflect(_ptr(_f3(v)), _ptr(_f3(n)), _ptr(out))
    return out


def refract(v, n, ratio):
    out = np.zeros(3, np.float32)
    lib.bvro_refract(_ptr(_f3(v)), _ptr(_f3(n)), float(ratio), _ptr(out))
    return out


def reflectance(cosine, ri):
    return np.float32(lib.bvro_reflectance(float(cosine), float(ri)))


def background_gradient(direction):
    out = np.zeros(3, np.float32)
    lib.bvro_background_gradient(_ptr(_f3(direction)), _ptr(out))
    return out


def random_ray_from_uv(camera, window, u, v, state):
    """Returns (origin, direction, state after the two jitter draws)."""
    st = C.c_uint32(int(state))
    o, d = np.zeros(3, np.float32), np.zeros(3, np.float32)
    lib.bvro_random_ray_from_uv(C.addressof(camera), C.addressof(window), float(u), float(v), C.byref(st), _ptr(o), _ptr(d))
    return o, d, st.value


def scatter(material, ray_dir, hit_pos, hit_normal, front_face, state):
    """material: one record of the 32-byte material layout.  Returns (absorbed, direction, attenuation, state)."""
    material = np.ascontiguousarray(material)
    assert material.dtype.itemsize == 32
    st = C.c_uint32(int(state))
    d, a = np.zeros(3, np.float32), np.zeros(3, np.float32)
    absorbed = lib.bvro_scatter(_ptr(material), _ptr(_f3(ray_dir)), _ptr(_f3(hit_pos)), _ptr(_f3(hit_normal)), int(bool(front_face)),
                                C.byref(st), _ptr(d), _ptr(a))
    return bool(absorbed), d, a, st.value
