"""Python convenience layer over the C ABI: numpy views of the reference's buffer layouts and a
`Context` wrapper.  All compute happens in libbevyray_b200.so on the GPU."""
import ctypes as C

import numpy as np

from . import _capi as capi
from ._capi import lib

# numpy mirrors of the encase layouts (src/raytracing/extract.rs:213-237, 181-189)
MODEL_DTYPE = np.dtype({"names": ["position", "radius", "material_id"],
                        "formats": [("<f4", 3), "<f4", "<u4"], "offsets": [0, 12, 16], "itemsize": 32})
MATERIAL_DTYPE = np.dtype({"names": ["base_color", "metallic", "roughness", "reflectance", "ior", "specular_transmission"],
                           "formats": [("<f4", 3), "<f4", "<f4", "<f4", "<f4", "<f4"],
                           "offsets": [0, 12, 16, 20, 24, 28], "itemsize": 32})
BVH_NODE_DTYPE = np.dtype({"names": ["bounds_min", "bounds_max", "index", "model_count"],
                           "formats": [("<f4", 3), ("<f4", 3), "<u4", "<u4"], "offsets": [0, 16, 28, 32], "itemsize": 48})

INF = np.float32(3.40282347e+38)   # assets/shaders/const.wgsl:2
MISS_ID = np.uint32(0xFFFFFFFF)


class BvrError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"[{lib.bvr_status_string(status).decode()}] {message}")
        self.status = status


def make_camera(position=(0.0, 0.0, 5.0), target=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), fov=np.pi / 4,
                aspect=16.0 / 9.0, near=0.1, far=1000.0, sample_count=1, bounces=4):
    """CameraExtract (extract.rs:118-146) of a `Transform::looking_at` camera, built by the host layer."""
    cam = capi.BvrCamera()
    f3 = C.c_float * 3
    lib.bvrh_camera_look_at(f3(*position), f3(*target), f3(*up), fov, aspect, near, far, sample_count, bounces, C.byref(cam))
    return cam


def make_level(level):
    lv = capi.BvrRaytraceLevel()
    lv.level = int(level)
    return lv


def make_window(random_seed, height):
    w = capi.BvrWindow()
    w.random_seed = float(random_seed)
    w.height = int(height)
    return w


def make_options(width, kernel=capi.KERNEL_AUTO, traversal=capi.TRAVERSAL_AUTO, shard_index=0, shard_count=0, strip_rows=0,
                 output_weight=0.0):
    o = capi.BvrRenderOptions()
    o.width, o.kernel, o.traversal = int(width), int(kernel), int(traversal)
    o.shard_index, o.shard_count, o.strip_rows = int(shard_index), int(shard_count), int(strip_rows)
    o.output_weight = float(output_weight)   # 0 = none
    return o


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Scene:
    """Scene buffers in the reference layout, produced by the C++ host layer."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("host layer failed to build the scene")
        self._h = C.c_void_p(handle)

    @classmethod
    def rtiow(cls, seed=1):
        return cls(lib.bvrh_scene_rtiow(seed))

    @classmethod
    def random(cls, seed, n, side, rmin, rmax):
        return cls(lib.bvrh_scene_random(seed, n, side, rmin, rmax))

    @classmethod
    def from_arrays(cls, models, materials):
        models = np.ascontiguousarray(models, dtype=MODEL_DTYPE)
        materials = np.ascontiguousarray(materials, dtype=MATERIAL_DTYPE)
        return cls(lib.bvrh_scene_from_models(_ptr(models), len(models), _ptr(materials), len(materials)))

    def animate(self, frame, rebuild_bvh=True):
        """Closed-form motion of every 4th sphere; rebuild_bvh=False leaves the (then stale) node array alone."""
        fn = lib.bvrh_scene_animate if rebuild_bvh else lib.bvrh_scene_animate_models
        if fn(self._h, frame) != 0:
            raise RuntimeError("bvrh_scene_animate failed")

    def _view(self, ptr, n, dtype):
        if n == 0:
            return np.zeros(0, dtype=dtype)
        buf = (C.c_char * (n * dtype.itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype, count=n)

    @property
    def models(self):
        return self._view(lib.bvrh_scene_models(self._h), lib.bvrh_scene_n_models(self._h), MODEL_DTYPE)

    @property
    def materials(self):
        return self._view(lib.bvrh_scene_materials(self._h), lib.bvrh_scene_n_materials(self._h), MATERIAL_DTYPE)

    @property
    def nodes(self):
        return self._view(lib.bvrh_scene_nodes(self._h), lib.bvrh_scene_n_nodes(self._h), BVH_NODE_DTYPE)

    def __del__(self):
        if getattr(self, "_h", None):
            lib.bvrh_scene_free(self._h)
            self._h = None


def build_ploc(models, search_distance=24):
    models = np.ascontiguousarray(models, dtype=MODEL_DTYPE)
    n = len(models)
    out = np.zeros(max(2 * n - 1, 0), dtype=BVH_NODE_DTYPE)
    got = lib.bvrh_build_ploc(_ptr(models), n, search_distance, _ptr(out))
    assert got == len(out)
    return out


def validate_bvh(nodes, models):
    nodes = np.ascontiguousarray(nodes, dtype=BVH_NODE_DTYPE)
    models = np.ascontiguousarray(models, dtype=MODEL_DTYPE)
    msg = C.create_string_buffer(256)
    rc = lib.bvrh_validate_bvh(_ptr(nodes), len(nodes), _ptr(models), len(models), msg, 256)
    return None if rc == 0 else msg.value.decode()


def traversal_ranks(nodes, n_models):
    """Position of every model in the reference's traversal order + the number of tree levels
    (bvr_scene_traversal_ranks: host only, what bvr_upload_scene derives for the tie rule)."""
    nodes = np.ascontiguousarray(nodes, dtype=BVH_NODE_DTYPE)
    ranks = np.empty(n_models, np.uint32)
    depth = C.c_uint32()
    st = lib.bvr_scene_traversal_ranks(_ptr(nodes), len(nodes), n_models, _ptr(ranks), C.byref(depth))
    if st != capi.BVR_OK:
        raise BvrError(st, "bvr_scene_traversal_ranks")
    return ranks, depth.value


class Context:
    """One GPU context (bvr_create / bvr_destroy)."""

    def __init__(self, device=0):
        h = C.c_void_p()
        st = lib.bvr_create(device, C.byref(h))
        if st != capi.BVR_OK:
            raise BvrError(st, "bvr_create failed")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib.bvr_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, st):
        if st != capi.BVR_OK:
            raise BvrError(st, lib.bvr_last_error(self._h).decode())

    def set_stream(self, cuda_stream_handle):
        self._check(lib.bvr_set_stream(self._h, C.c_void_p(cuda_stream_handle)))

    def sync(self):
        self._check(lib.bvr_sync(self._h))

    def reload_tuning(self):
        """Re-reads the BVR_* experiment knobs from the environment (they are otherwise read once, at creation)."""
        self._check(lib.bvr_reload_tuning(self._h))

    def upload_scene(self, models, materials, nodes, ranges=None):
        models = np.ascontiguousarray(models, dtype=MODEL_DTYPE)
        materials = np.ascontiguousarray(materials, dtype=MATERIAL_DTYPE)
        nodes = np.ascontiguousarray(nodes, dtype=BVH_NODE_DTYPE)
        if ranges is None:
            rp, rn = None, 0
        else:
            arr = (capi.BvrDirtyRange * max(len(ranges), 1))()
            for i, (a, f, c) in enumerate(ranges):
                arr[i].array, arr[i].first, arr[i].count = a, f, c
            rp, rn = C.cast(arr, C.c_void_p), len(ranges)
        self._check(lib.bvr_upload_scene(self._h, _ptr(models), len(models), _ptr(materials), len(materials),
                                         _ptr(nodes), len(nodes), rp, rn))

    def upload_scene_gpu_bvh(self, models, materials, ranges=None, want_nodes=False, refit=False):
        """bvr_upload_scene_gpu_bvh: the BVH is built on the GPU; optionally returns the nodes (reference layout).
        refit=True is bvr_refit_scene_gpu_bvh: the last GPU-built topology is kept, only its boxes are refitted."""
        models = np.ascontiguousarray(models, dtype=MODEL_DTYPE)
        materials = np.ascontiguousarray(materials, dtype=MATERIAL_DTYPE)
        if ranges is None:
            rp, rn = None, 0
        else:
            arr = (capi.BvrDirtyRange * max(len(ranges), 1))()
            for i, (a, f, c) in enumerate(ranges):
                arr[i].array, arr[i].first, arr[i].count = a, f, c
            rp, rn = C.cast(arr, C.c_void_p), len(ranges)
        nodes = np.zeros(max(2 * len(models) - 1, 0), dtype=BVH_NODE_DTYPE) if want_nodes else None
        fn = lib.bvr_refit_scene_gpu_bvh if refit else lib.bvr_upload_scene_gpu_bvh
        self._check(fn(self._h, _ptr(models), len(models), _ptr(materials), len(materials),
                       rp, rn, _ptr(nodes) if want_nodes and len(nodes) else None))
        return nodes

    def shard_rows(self, height, opts):
        return lib.bvr_shard_rows(height, C.byref(opts))

    def render(self, camera, level, window, opts, raster_rgba=None, raster_depth=None,
               want=("rgba", "rt_depth", "primary_id", "primary_depth"), out=None, asynchronous=False):
        """bvr_render with host (numpy) buffers.  Returns a dict of numpy planes of this shard's rows.
        asynchronous=True is bvr_render_async: every buffer must be page-locked, `out` must be given, and the planes
        are valid after sync()."""
        rows = self.shard_rows(window.height, opts)
        w = opts.width
        shapes = {"rgba": ((rows, w, 4), np.float32), "rt_depth": ((rows, w), np.float32),
                  "primary_id": ((rows, w), np.uint32), "primary_depth": ((rows, w), np.float32),
                  "srgb8": ((rows, w, 4), np.uint8)}
        res = {} if out is None else out
        o = capi.BvrOutputs()
        for k in want:
            if k not in res:
                res[k] = np.empty(shapes[k][0], dtype=shapes[k][1])
            setattr(o, k, res[k].ctypes.data)
        if raster_rgba is not None:
            raster_rgba = np.ascontiguousarray(raster_rgba, dtype=np.float32)
        if raster_depth is not None:
            raster_depth = np.ascontiguousarray(raster_depth, dtype=np.float32)
        lv = level if isinstance(level, capi.BvrRaytraceLevel) else make_level(level)
        fn = lib.bvr_render_async if asynchronous else lib.bvr_render
        self._check(fn(self._h, C.byref(camera), C.byref(lv), C.byref(window), C.byref(opts),
                       _ptr(raster_rgba), _ptr(raster_depth), C.byref(o)))
        return res

    def render_device(self, camera, level, window, opts, d_raster_rgba=0, d_raster_depth=0, **device_ptrs):
        """bvr_render_device: every pointer is a device address (int); enqueues on the context stream."""
        o = capi.BvrOutputs()
        for k, v in device_ptrs.items():
            setattr(o, k, v)
        lv = level if isinstance(level, capi.BvrRaytraceLevel) else make_level(level)
        self._check(lib.bvr_render_device(self._h, C.byref(camera), C.byref(lv), C.byref(window), C.byref(opts),
                                          C.c_void_p(d_raster_rgba or None), C.c_void_p(d_raster_depth or None), C.byref(o)))

    def axpby_device(self, d_dst, dst_weight, d_src, src_weight, n):
        self._check(lib.bvr_axpby_device(self._h, C.c_void_p(d_dst), dst_weight, C.c_void_p(d_src or None), src_weight, n))

    def composite_device(self, camera, level, d_rgba, d_rt_depth, d_raster_rgba, d_raster_depth, n_pixels):
        lv = level if isinstance(level, capi.BvrRaytraceLevel) else make_level(level)
        self._check(lib.bvr_composite_device(self._h, C.byref(camera), C.byref(lv), C.c_void_p(d_rgba), C.c_void_p(d_rt_depth),
                                             C.c_void_p(d_raster_rgba or None), C.c_void_p(d_raster_depth or None), n_pixels))

    # ---- peer memory for the fused sample-sharding exchange (bevyray_b200.h) ----
    def peer_alloc(self, nbytes):
        """(device pointer, 64-byte handle another process can open)"""
        ptr = C.c_void_p()
        handle = (C.c_uint8 * capi.PEER_HANDLE_BYTES)()
        self._check(lib.bvr_peer_alloc(self._h, nbytes, C.byref(ptr), C.cast(handle, C.c_void_p)))
        return int(ptr.value), bytes(handle)

    def peer_open(self, handle):
        ptr = C.c_void_p()
        buf = (C.c_uint8 * capi.PEER_HANDLE_BYTES).from_buffer_copy(handle)
        self._check(lib.bvr_peer_open(self._h, C.cast(buf, C.c_void_p), C.byref(ptr)))
        return int(ptr.value)

    def peer_close(self, ptr):
        self._check(lib.bvr_peer_close(self._h, C.c_void_p(ptr)))

    def peer_free(self, ptr):
        self._check(lib.bvr_peer_free(self._h, C.c_void_p(ptr)))

    def sum_slots_device(self, d_slots, slot_stride_floats, n_slots, mask, d_dst, n):
        self._check(lib.bvr_sum_slots_device(self._h, C.c_void_p(d_slots), slot_stride_floats, n_slots, mask, C.c_void_p(d_dst), n))

    def unshard_device(self, d_gathered, shard_stride_words, d_full, width, height, channels, shard_count, strip_rows):
        self._check(lib.bvr_unshard_device(self._h, C.c_void_p(d_gathered), shard_stride_words, C.c_void_p(d_full),
                                           width, height, channels, shard_count, strip_rows))

    def stats(self):
        s = capi.BvrStats()
        self._check(lib.bvr_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in s._fields_}
