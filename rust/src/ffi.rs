//! `extern "C"` view of include/bevyray_b200.h.  Layouts are `#[repr(C)]` twins of the encase
//! `ShaderType` structs the reference already serialises (src/raytracing/extract.rs:56-104, 181-237),
//! so the `Vec<u8>` inside each `StorageBuffer` can be handed over without conversion.
//! UNCOMPILED in this repository (no Rust toolchain in the build image).

use std::os::raw::{c_char, c_int, c_void};

pub const BVR_OK: c_int = 0;
pub const BVR_ERR_UNSUPPORTED_PROJECTION: c_int = 3;
/// BvrRenderFlags (include/bevyray_b200.h)
pub const BVR_RENDER_DEFER_COMPOSITE: u32 = 1;
pub const BVR_RENDER_EXTRA_SAMPLE: u32 = 2;
/// BVR_RENDER_EXTRA_SAMPLE_BITS: the 8x4 tiles with ((tx + ty + phase) % modulus) < count take one sample more
pub const fn bvr_render_extra_sample_bits(modulus: u32, phase: u32, count: u32) -> u32 {
    BVR_RENDER_EXTRA_SAMPLE | (modulus << 8) | (phase << 16) | (count << 24)
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct BvrCamera {
    pub sample_count: u32,
    pub bounce_count: u32,
    pub projection: u32,
    pub near_plane: f32,
    pub far_plane: f32,
    pub fov: f32,
    pub aspect: f32,
    pub _pad0: u32,
    pub position: [f32; 3],
    pub _pad1: u32,
    pub direction: [f32; 3],
    pub _pad2: u32,
    pub up: [f32; 3],
    pub _pad3: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct BvrRaytraceLevel {
    pub level: u32,
    pub _pad0: [u32; 3],
    pub _padding: [f32; 3],
    pub _pad1: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct BvrWindow {
    pub random_seed: f32,
    pub height: u32,
    pub _padding: [f32; 2],
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct BvrDirtyRange {
    pub array: u32,
    pub first: u32,
    pub count: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct BvrRenderOptions {
    pub width: u32,
    pub kernel: u32,
    pub traversal: u32,
    pub shard_index: u32,
    pub shard_count: u32,
    pub strip_rows: u32,
    pub flags: u32,
    /// 0 (or 1) = none; otherwise every rgba / rt_depth texel is multiplied by it as it is stored
    pub output_weight: f32,
}

#[repr(C)]
pub struct BvrOutputs {
    pub rgba: *mut f32,
    pub rt_depth: *mut f32,
    pub primary_id: *mut u32,
    pub primary_depth: *mut f32,
    pub srgb8: *mut u8,
}

#[repr(C)]
pub struct BvrContext {
    _opaque: [u8; 0],
}

extern "C" {
    pub fn bvr_create(device: c_int, out_ctx: *mut *mut BvrContext) -> c_int;
    pub fn bvr_destroy(ctx: *mut BvrContext);
    pub fn bvr_last_error(ctx: *const BvrContext) -> *const c_char;
    pub fn bvr_upload_scene(
        ctx: *mut BvrContext,
        models: *const c_void, n_models: usize,
        materials: *const c_void, n_materials: usize,
        nodes: *const c_void, n_nodes: usize,
        ranges: *const BvrDirtyRange, n_ranges: usize,
    ) -> c_int;
    pub fn bvr_render(
        ctx: *mut BvrContext,
        camera: *const BvrCamera, level: *const BvrRaytraceLevel, window: *const BvrWindow,
        opts: *const BvrRenderOptions,
        raster_rgba: *const f32, raster_depth: *const f32,
        host_out: *const BvrOutputs,
    ) -> c_int;
    /// bvr_render without the wait (page-locked buffers only); `bvr_sync` completes the frame.
    pub fn bvr_render_async(
        ctx: *mut BvrContext,
        camera: *const BvrCamera, level: *const BvrRaytraceLevel, window: *const BvrWindow,
        opts: *const BvrRenderOptions,
        raster_rgba: *const f32, raster_depth: *const f32,
        host_out: *const BvrOutputs,
    ) -> c_int;
    pub fn bvr_sync(ctx: *mut BvrContext) -> c_int;
    /// The BVH is built on the GPU (PLOC over the Morton order) instead of by obvhs on the host.
    pub fn bvr_upload_scene_gpu_bvh(
        ctx: *mut BvrContext,
        models: *const c_void, n_models: usize,
        materials: *const c_void, n_materials: usize,
        ranges: *const BvrDirtyRange, n_ranges: usize,
        out_nodes: *mut c_void,
    ) -> c_int;
    /// Same arguments: keeps the last GPU-built topology, refits the boxes (small motion).
    pub fn bvr_refit_scene_gpu_bvh(
        ctx: *mut BvrContext,
        models: *const c_void, n_models: usize,
        materials: *const c_void, n_materials: usize,
        ranges: *const BvrDirtyRange, n_ranges: usize,
        out_nodes: *mut c_void,
    ) -> c_int;
}
