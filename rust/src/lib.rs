//! Drop-in replacement of bevyray's `raytracing` module for the hot path, keeping its public surface:
//! `RaytracePlugin`, `RaytracedCamera { level, sample_count, bounces }`, `Raytracing`, `RaytracedSphere
//! { radius }`, `RaytraceLabel` (reference: src/raytracing/mod.rs:21-106).  What changes is the body of
//! the render-graph node: instead of binding three storage buffers and drawing a fullscreen triangle with
//! the WGSL shader (src/raytracing/pipeline.rs:132-217), it hands the same bytes to libbevyray_b200.so.
//!
//! UNCOMPILED in this repository: the build image has no Rust toolchain.  The C++ host layer under
//! bevyray_b200/csrc/host is the compiled, tested mirror of this file.

pub mod ffi;

use bevy::prelude::*;
use std::sync::Mutex;

/// Same discriminants as the reference enum (#[repr(u32)], mod.rs:94-101): they travel to the GPU.
#[repr(u32)]
#[derive(Reflect, Clone, Copy, PartialEq, Eq)]
pub enum Raytracing {
    Skip = 0,
    FallbackRaster = 1,
    FallbackRaytraced = 2,
    Pure = 3,
}

#[derive(Component, Reflect, Clone, Copy)]
pub struct RaytracedCamera {
    pub level: Raytracing,
    pub sample_count: u32,
    pub bounces: u32,
}

#[derive(Component, Reflect)]
pub struct RaytracedSphere {
    pub radius: f32,
}

/// Owns the CUDA context; `Send` but not `Sync`, hence the `Mutex` — the same arrangement the reference
/// uses for its `StorageBuffer`s (extract.rs:252-262).
#[derive(Resource)]
pub struct CudaRaytracer(pub Mutex<CudaContext>);

pub struct CudaContext {
    raw: *mut ffi::BvrContext,
}
unsafe impl Send for CudaContext {}

impl CudaContext {
    pub fn new(device: i32) -> Result<Self, i32> {
        let mut raw = std::ptr::null_mut();
        let status = unsafe { ffi::bvr_create(device, &mut raw) };
        if status == ffi::BVR_OK { Ok(Self { raw }) } else { Err(status) }
    }

    /// Replaces the three `write_buffer` calls of `RayTracingNode::run` (pipeline.rs:136-138).
    /// `models`, `materials`, `nodes` are the encase bytes of the reference's own `StorageBuffer`s.
    pub fn upload_scene(&mut self, models: &[u8], materials: &[u8], nodes: &[u8], dirty: Option<&[ffi::BvrDirtyRange]>) -> i32 {
        let (ranges, n_ranges) = match dirty {
            Some(d) => (d.as_ptr(), d.len()),
            None => (std::ptr::null(), 0),
        };
        unsafe {
            ffi::bvr_upload_scene(
                self.raw,
                models.as_ptr().cast(), models.len() / 32,
                materials.as_ptr().cast(), materials.len() / 32,
                nodes.as_ptr().cast(), nodes.len() / 48,
                ranges, n_ranges,
            )
        }
    }

    /// Replaces the render pass + `draw(0..3, 0..1)` (pipeline.rs:191-217): one `fragment` invocation
    /// per pixel.  `raster_rgba` / `raster_depth` are read-backs (or, with external-memory interop, the
    /// mapped textures) of `post_process.source` and the prepass depth view.
    pub fn render(
        &mut self,
        camera: &ffi::BvrCamera, level: u32, random_seed: f32, width: u32, height: u32,
        raster_rgba: &[f32], raster_depth: &[f32], out_rgba: &mut [f32],
    ) -> i32 {
        let lvl = ffi::BvrRaytraceLevel { level, ..Default::default() };
        let win = ffi::BvrWindow { random_seed, height, ..Default::default() };
        let opts = ffi::BvrRenderOptions { width, ..Default::default() };
        let out = ffi::BvrOutputs {
            rgba: out_rgba.as_mut_ptr(),
            rt_depth: std::ptr::null_mut(),
            primary_id: std::ptr::null_mut(),
            primary_depth: std::ptr::null_mut(),
            srgb8: std::ptr::null_mut(),
        };
        unsafe { ffi::bvr_render(self.raw, camera, &lvl, &win, &opts, raster_rgba.as_ptr(), raster_depth.as_ptr(), &out) }
    }
}

impl Drop for CudaContext {
    fn drop(&mut self) {
        unsafe { ffi::bvr_destroy(self.raw) }
    }
}

pub struct RaytracePlugin;

impl Plugin for RaytracePlugin {
    fn build(&self, app: &mut App) {
        // Same registrations as the reference plugin (mod.rs:27-34); the extract plugin, the render-graph
        // node and its edges (Tonemapping -> RaytraceLabel -> EndMainPassPostProcessing, mod.rs:56-71) stay
        // as they are in the reference — only RayTracingNode::run's body changes.
        app.insert_resource(Msaa::Off)
            .register_type::<RaytracedCamera>()
            .register_type::<Raytracing>()
            .register_type::<RaytracedSphere>();
    }

    fn finish(&self, app: &mut App) {
        // RaytracingPipeline::from_world (pipeline.rs:233-331) becomes: create the CUDA context once.
        let ctx = CudaContext::new(0).expect("bevyray_b200: no CUDA device (there is no CPU fallback)");
        app.insert_resource(CudaRaytracer(Mutex::new(ctx)));
    }
}
