/*
 * bevyray_b200_host.h — C view of the C++ host layer (bevyray_b200/csrc/host), for ctypes tests and
 * tools.  The host layer mirrors the reference's Rust plugin code for the hot path; the symbols here
 * are conveniences around it, not part of the drop-in boundary (that is bevyray_b200.h).
 *
 * Reference anchors: scene recipe src/main.rs:49-240; prepare_buffers src/raytracing/extract.rs:280-337;
 * plugin + frame schedule src/raytracing/mod.rs:24-115, src/raytracing/pipeline.rs:58-220.
 */
#ifndef BEVYRAY_B200_HOST_H
#define BEVYRAY_B200_HOST_H

#include "bevyray_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- scene buffers in the reference's storage-buffer layout -------------------------------- */
typedef struct BvrhScene BvrhScene;

BvrhScene* bvrh_scene_rtiow(uint64_t seed);   /* src/main.rs setup(), seeded */
BvrhScene* bvrh_scene_random(uint64_t seed, uint32_t n, float side, float rmin, float rmax);
/* models + materials supplied by the caller; the BVH is built like prepare_buffers does */
BvrhScene* bvrh_scene_from_models(const BvrModel* models, size_t n_models,
                                  const BvrMaterial* materials, size_t n_materials);
/* C5 animation: positions = closed-form function of `frame` applied to the scene's base positions;
 * rebuilds the BVH.  Returns 0 on success. */
int  bvrh_scene_animate(BvrhScene* scene, uint32_t frame);
/* the same motion without the host-side BVH rebuild (the node array goes stale): for callers that build the tree on the
 * GPU (bvr_upload_scene_gpu_bvh) */
int  bvrh_scene_animate_models(BvrhScene* scene, uint32_t frame);
void bvrh_scene_free(BvrhScene* scene);
size_t bvrh_scene_n_models(const BvrhScene* scene);
size_t bvrh_scene_n_materials(const BvrhScene* scene);
size_t bvrh_scene_n_nodes(const BvrhScene* scene);
const BvrModel*    bvrh_scene_models(const BvrhScene* scene);
const BvrMaterial* bvrh_scene_materials(const BvrhScene* scene);
const BvrBvhNode*  bvrh_scene_nodes(const BvrhScene* scene);

/* ---- BVH producer (replaces obvhs::ploc::build_ploc, extract.rs:316-332) -------------------- */
/* out_nodes must hold 2*n_models-1 nodes (0 when n_models == 0).  Returns the node count. */
size_t bvrh_build_ploc(const BvrModel* models, size_t n_models, uint32_t search_distance, BvrBvhNode* out_nodes);
/* 0 when the node array satisfies the reference contract; otherwise 1 and a message. */
int bvrh_validate_bvh(const BvrBvhNode* nodes, size_t n_nodes, const BvrModel* models, size_t n_models,
                      char* msg, size_t msg_len);

/* ---- CameraExtract from a look-at transform (extract.rs:118-146, Transform::looking_at) ------ */
void bvrh_camera_look_at(const float position[3], const float target[3], const float up[3],
                         float fov, float aspect, float near_plane, float far_plane,
                         uint32_t sample_count, uint32_t bounces, BvrCamera* out);
/* bevy_color sRGB -> linear, as RaytraceMaterial::prepare_asset applies it (extract.rs:201) */
float bvrh_srgb_to_linear(float v);

/* ---- App / plugin mirror ------------------------------------------------------------------- */
typedef struct BvrhApp BvrhApp;

typedef struct BvrhStandardMaterial {   /* the StandardMaterial fields extract.rs:200-207 reads */
    float base_color_srgb[3];
    float metallic;
    float perceptual_roughness;
    float reflectance;
    float ior;
    float specular_transmission;
} BvrhStandardMaterial;

BvrhApp* bvrh_app_create(void);
void     bvrh_app_destroy(BvrhApp* app);
const char* bvrh_app_last_error(const BvrhApp* app);
/* App::add_plugins(RaytracePlugin) — creates the pipeline (C-ABI context) on `device`.  Returns a BvrStatus. */
int      bvrh_app_add_raytrace_plugin(BvrhApp* app, int device);
/* main.rs setup(): window + camera + spheres.  Returns the camera entity. */
uint32_t bvrh_app_setup_demo(BvrhApp* app, uint64_t seed);
void     bvrh_app_standard_material_default(BvrhStandardMaterial* out);
uint32_t bvrh_app_spawn_window(BvrhApp* app, uint32_t physical_width, uint32_t physical_height);
uint32_t bvrh_app_spawn_sphere(BvrhApp* app, float x, float y, float z, float radius, const BvrhStandardMaterial* material);
/* orthographic != 0 spawns an orthographic camera (never extracted, extract.rs:148) */
uint32_t bvrh_app_spawn_camera(BvrhApp* app, const float position[3], const float target[3], const float up[3],
                               float fov, float aspect, float near_plane, float far_plane,
                               uint32_t level, uint32_t sample_count, uint32_t bounces, int orthographic);
int      bvrh_app_set_raytraced_camera(BvrhApp* app, uint32_t entity, uint32_t level, uint32_t sample_count, uint32_t bounces);
int      bvrh_app_set_translation(BvrhApp* app, uint32_t entity, float x, float y, float z);
int      bvrh_app_set_material(BvrhApp* app, uint32_t entity, const BvrhStandardMaterial* material);
void     bvrh_app_set_window_size(BvrhApp* app, uint32_t physical_width, uint32_t physical_height);
/* fixed value for WindowExtract.random_seed (extract.rs:72-73 draws a random one); negative = random again */
void     bvrh_app_set_seed(BvrhApp* app, float seed);
void     bvrh_app_set_render_options(BvrhApp* app, const BvrRenderOptions* opts);
/* non-zero: skip the host PLOC build in prepare_buffers and let the library build the BVH on the GPU */
void     bvrh_app_set_gpu_bvh(BvrhApp* app, int enabled);
int      bvrh_app_set_raster(BvrhApp* app, uint32_t camera, const float* rgba, const float* depth, size_t n_pixels);
/* One frame.  Returns the number of views rendered (0 = skipped like the reference's early returns), -1 on error. */
int      bvrh_app_update(BvrhApp* app);
const float* bvrh_app_frame(const BvrhApp* app, uint32_t camera, uint32_t* width, uint32_t* height);
/* the render world's storage buffers after the last update (what RayTracingNode::run uploads) */
size_t   bvrh_app_buffers(const BvrhApp* app, const BvrModel** models, const BvrMaterial** materials,
                          const BvrBvhNode** nodes, size_t* n_nodes);
int      bvrh_app_msaa_off(const BvrhApp* app);
int      bvrh_app_has_depth_prepass(const BvrhApp* app, uint32_t entity);
int      bvrh_app_get_stats(BvrhApp* app, BvrStats* out);

#ifdef __cplusplus
}
#endif
#endif
