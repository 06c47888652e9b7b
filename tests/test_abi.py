"""The C-ABI library loads and exports every symbol the headers declare; layouts match the reference's
encase layouts.  No compute calls: this runs without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//.*", "", text)
    text = text.split("#ifdef __cplusplus\n} /* extern")[0] if "static_assert" in text else text
    names = re.findall(r"\b(bvrh?_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


@pytest.mark.parametrize("header", ["bevyray_b200.h", "bevyray_b200_host.h"])
def test_every_declared_symbol_is_exported_and_bound(bvr, header):
    names = declared_functions(header)
    assert len(names) >= 10
    for n in names:
        assert hasattr(bvr.capi.lib, n), f"{n} declared in include/{header} but not exported"
        assert n in bvr.capi.SIGNATURES, f"{n} declared in include/{header} but not bound in _capi.py"


def test_no_undeclared_bindings(bvr):
    declared = set(declared_functions("bevyray_b200.h")) | set(declared_functions("bevyray_b200_host.h"))
    assert set(bvr.capi.SIGNATURES) <= declared


def test_layouts_match_reference_contract(bvr):
    cp = bvr.capi
    assert C.sizeof(cp.BvrModel) == 32 and cp.BvrModel.radius.offset == 12 and cp.BvrModel.material_id.offset == 16
    assert C.sizeof(cp.BvrMaterial) == 32 and cp.BvrMaterial.metallic.offset == 12
    assert cp.BvrMaterial.roughness.offset == 16 and cp.BvrMaterial.ior.offset == 24
    assert cp.BvrMaterial.specular_transmission.offset == 28
    assert C.sizeof(cp.BvrBvhNode) == 48 and cp.BvrBvhNode.bounds_max.offset == 16
    assert cp.BvrBvhNode.index.offset == 28 and cp.BvrBvhNode.model_count.offset == 32
    assert C.sizeof(cp.BvrCamera) == 80 and cp.BvrCamera.near_plane.offset == 12 and cp.BvrCamera.far_plane.offset == 16
    assert cp.BvrCamera.fov.offset == 20 and cp.BvrCamera.aspect.offset == 24 and cp.BvrCamera.position.offset == 32
    assert cp.BvrCamera.direction.offset == 48 and cp.BvrCamera.up.offset == 64
    assert C.sizeof(cp.BvrRaytraceLevel) == 32 and C.sizeof(cp.BvrWindow) == 16 and cp.BvrWindow.height.offset == 4
    assert bvr.MODEL_DTYPE.itemsize == 32 and bvr.MATERIAL_DTYPE.itemsize == 32 and bvr.BVH_NODE_DTYPE.itemsize == 48
    # enum Raytracing discriminants, src/raytracing/mod.rs:94-101
    assert (cp.RAYTRACING_SKIP, cp.RAYTRACING_FALLBACK_RASTER, cp.RAYTRACING_FALLBACK_RAYTRACED, cp.RAYTRACING_PURE) == (0, 1, 2, 3)


def test_abi_version_and_status_strings(bvr):
    lib = bvr.capi.lib
    assert lib.bvr_abi_version() == 1
    assert lib.bvr_status_string(0) == b"ok"
    assert b"CPU fallback" in lib.bvr_status_string(bvr.capi.BVR_ERR_NO_DEVICE)


def test_null_and_invalid_arguments_return_codes(bvr):
    lib = bvr.capi.lib
    assert lib.bvr_create(0, None) == bvr.capi.BVR_ERR_INVALID_ARGUMENT
    assert lib.bvr_sync(None) == bvr.capi.BVR_ERR_INVALID_ARGUMENT
    assert lib.bvr_get_stats(None, None) == bvr.capi.BVR_ERR_INVALID_ARGUMENT
    assert lib.bvr_last_error(None) == b"null context"
    lib.bvr_destroy(None)   # no-op, must not crash
    opts = bvr.make_options(64, shard_index=1, shard_count=4, strip_rows=4)
    assert lib.bvr_shard_rows(30, C.byref(opts)) == 8      # ceil(ceil(30/4)/4) strips x 4 rows
    assert lib.bvr_shard_rows(30, C.byref(bvr.make_options(64))) == 30


def test_product_fails_loudly_without_a_gpu(bvr):
    """There is no CPU fallback: on a box without CUDA bvr_create returns BVR_ERR_NO_DEVICE."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert bvr.capi.lib.bvr_create(0, C.byref(h)) == bvr.capi.BVR_ERR_NO_DEVICE
    assert not h.value
    with pytest.raises(bvr.BvrError):
        bvr.Context(0)
    t = C.c_float()
    assert bvr.capi.lib.bvr_bench_fp32_peak(0, C.byref(t)) != 0


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing under bevyray_b200/ or include/ may mention it."""
    bad = []
    for base in ("bevyray_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"bvr_oracle|libbvr_oracle|from oracle|import oracle|bvro_", text):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad
