// host_capi.cpp — extern "C" view of the C++ host layer (include/bevyray_b200_host.h).

#include "../../../include/bevyray_b200_host.h"
#include "bevyray_host.hpp"

#include <cstdio>
#include <cstring>
#include <exception>

using namespace bevyray;

struct BvrhScene {
    SceneBuffers buffers;
    std::vector<BvrModel> base;
};

struct BvrhApp {
    App app;
    std::string error;
    float fixed_seed = -1.0f;
};

static float fixed_seed_fn(void* user) { return static_cast<BvrhApp*>(user)->fixed_seed; }

static StandardMaterial to_material(const BvrhStandardMaterial* m) {
    StandardMaterial s;
    if (!m) return s;
    s.base_color = Color::srgb(m->base_color_srgb[0], m->base_color_srgb[1], m->base_color_srgb[2]);
    s.metallic = m->metallic;
    s.perceptual_roughness = m->perceptual_roughness;
    s.reflectance = m->reflectance;
    s.ior = m->ior;
    s.specular_transmission = m->specular_transmission;
    return s;
}

extern "C" {

BvrhScene* bvrh_scene_rtiow(uint64_t seed) {
    try {
        auto* s = new BvrhScene{scene_rtiow(seed), {}};
        s->base = s->buffers.models;
        return s;
    } catch (...) { return nullptr; }
}

BvrhScene* bvrh_scene_random(uint64_t seed, uint32_t n, float side, float rmin, float rmax) {
    try {
        auto* s = new BvrhScene{scene_random(seed, n, side, rmin, rmax), {}};
        s->base = s->buffers.models;
        return s;
    } catch (...) { return nullptr; }
}

BvrhScene* bvrh_scene_from_models(const BvrModel* models, size_t n_models,
                                  const BvrMaterial* materials, size_t n_materials) {
    try {
        auto* s = new BvrhScene{};
        if (models && n_models) s->buffers.models.assign(models, models + n_models);
        if (materials && n_materials) s->buffers.materials.assign(materials, materials + n_materials);
        s->buffers.nodes = build_ploc(s->buffers.models, 24);
        s->base = s->buffers.models;
        return s;
    } catch (...) { return nullptr; }
}

int bvrh_scene_animate(BvrhScene* scene, uint32_t frame) {
    if (!scene) return 1;
    try { animate_random(scene->buffers, scene->base, frame); } catch (...) { return 1; }
    return 0;
}

int bvrh_scene_animate_models(BvrhScene* scene, uint32_t frame) {
    if (!scene) return 1;
    try { animate_random(scene->buffers, scene->base, frame, false); } catch (...) { return 1; }
    return 0;
}

void bvrh_scene_free(BvrhScene* scene) { delete scene; }
size_t bvrh_scene_n_models(const BvrhScene* s) { return s ? s->buffers.models.size() : 0; }
size_t bvrh_scene_n_materials(const BvrhScene* s) { return s ? s->buffers.materials.size() : 0; }
size_t bvrh_scene_n_nodes(const BvrhScene* s) { return s ? s->buffers.nodes.size() : 0; }
const BvrModel* bvrh_scene_models(const BvrhScene* s) { return s ? s->buffers.models.data() : nullptr; }
const BvrMaterial* bvrh_scene_materials(const BvrhScene* s) { return s ? s->buffers.materials.data() : nullptr; }
const BvrBvhNode* bvrh_scene_nodes(const BvrhScene* s) { return s ? s->buffers.nodes.data() : nullptr; }

size_t bvrh_build_ploc(const BvrModel* models, size_t n_models, uint32_t search_distance, BvrBvhNode* out_nodes) {
    try {
        std::vector<BvrModel> m;
        if (models && n_models) m.assign(models, models + n_models);
        std::vector<BvrBvhNode> nodes = build_ploc(m, search_distance);
        if (out_nodes && !nodes.empty()) std::memcpy(out_nodes, nodes.data(), nodes.size() * sizeof(BvrBvhNode));
        return nodes.size();
    } catch (...) { return 0; }
}

int bvrh_validate_bvh(const BvrBvhNode* nodes, size_t n_nodes, const BvrModel* models, size_t n_models,
                      char* msg, size_t msg_len) {
    std::vector<BvrBvhNode> nv;
    std::vector<BvrModel> mv;
    if (nodes && n_nodes) nv.assign(nodes, nodes + n_nodes);
    if (models && n_models) mv.assign(models, models + n_models);
    std::string r = validate_bvh(nv, mv);
    if (msg && msg_len) std::snprintf(msg, msg_len, "%s", r.c_str());
    return r.empty() ? 0 : 1;
}

void bvrh_camera_look_at(const float position[3], const float target[3], const float up[3],
                         float fov, float aspect, float near_plane, float far_plane,
                         uint32_t sample_count, uint32_t bounces, BvrCamera* out) {
    if (!out) return;
    Transform t = Transform::from_translation(Vec3(position[0], position[1], position[2]))
                      .looking_at(Vec3(target[0], target[1], target[2]), Vec3(up[0], up[1], up[2]));
    PerspectiveProjection p;
    p.fov = fov; p.aspect_ratio = aspect; p.near = near_plane; p.far = far_plane;
    auto e = extract_camera(RaytracedCamera{Raytracing::Pure, sample_count, bounces}, t, Projection{p});
    *out = e->camera;
}

float bvrh_srgb_to_linear(float v) { return Color::srgb(v, v, v).to_linear_vec3().x; }

BvrhApp* bvrh_app_create(void) {
    try { return new BvrhApp(); } catch (...) { return nullptr; }
}
void bvrh_app_destroy(BvrhApp* app) { delete app; }
const char* bvrh_app_last_error(const BvrhApp* app) { return app ? app->error.c_str() : "null app"; }

int bvrh_app_add_raytrace_plugin(BvrhApp* app, int device) {
    if (!app) return BVR_ERR_INVALID_ARGUMENT;
    try {
        RaytracePlugin plugin;
        plugin.device = device;
        app->app.add_plugins(plugin);
        if (!app->app.render.pipeline || !app->app.render.pipeline->ready()) {
            app->error = app->app.render.pipeline ? app->app.render.pipeline->error() : "no pipeline";
            return BVR_ERR_NO_DEVICE;
        }
    } catch (const std::exception& e) { app->error = e.what(); return BVR_ERR_CUDA; }
    return BVR_OK;
}

uint32_t bvrh_app_setup_demo(BvrhApp* app, uint64_t seed) { return setup(app->app.world, seed); }

void bvrh_app_standard_material_default(BvrhStandardMaterial* out) {
    StandardMaterial s;
    out->base_color_srgb[0] = s.base_color.r; out->base_color_srgb[1] = s.base_color.g; out->base_color_srgb[2] = s.base_color.b;
    out->metallic = s.metallic; out->perceptual_roughness = s.perceptual_roughness; out->reflectance = s.reflectance;
    out->ior = s.ior; out->specular_transmission = s.specular_transmission;
}

uint32_t bvrh_app_spawn_window(BvrhApp* app, uint32_t w, uint32_t h) {
    EntityData e;
    e.name = "Window";
    e.window = Window{w, h};
    return app->app.world.spawn(std::move(e));
}

uint32_t bvrh_app_spawn_sphere(BvrhApp* app, float x, float y, float z, float radius, const BvrhStandardMaterial* material) {
    EntityData e;
    e.transform = Transform::from_xyz(x, y, z);
    e.material = app->app.world.materials.add(to_material(material));
    e.raytraced_sphere = RaytracedSphere{radius};
    return app->app.world.spawn(std::move(e));
}

uint32_t bvrh_app_spawn_camera(BvrhApp* app, const float position[3], const float target[3], const float up[3],
                               float fov, float aspect, float near_plane, float far_plane,
                               uint32_t level, uint32_t sample_count, uint32_t bounces, int orthographic) {
    EntityData e;
    e.name = "Raytraced Camera";
    e.transform = Transform::from_translation(Vec3(position[0], position[1], position[2]))
                      .looking_at(Vec3(target[0], target[1], target[2]), Vec3(up[0], up[1], up[2]));
    e.camera = Camera{};
    if (orthographic) {
        e.projection = Projection{OrthographicProjection{}};
    } else {
        PerspectiveProjection p;
        p.fov = fov; p.aspect_ratio = aspect; p.near = near_plane; p.far = far_plane;
        e.projection = Projection{p};
    }
    e.raytraced_camera = RaytracedCamera{(Raytracing)level, sample_count, bounces};
    return app->app.world.spawn(std::move(e));
}

int bvrh_app_set_raytraced_camera(BvrhApp* app, uint32_t entity, uint32_t level, uint32_t sample_count, uint32_t bounces) {
    if (!app || entity >= app->app.world.entities.size()) return 1;
    app->app.world.entity(entity).raytraced_camera = RaytracedCamera{(Raytracing)level, sample_count, bounces};
    return 0;
}

int bvrh_app_set_translation(BvrhApp* app, uint32_t entity, float x, float y, float z) {
    if (!app || entity >= app->app.world.entities.size() || !app->app.world.entity(entity).transform) return 1;
    app->app.world.entity(entity).transform->translation = Vec3(x, y, z);
    return 0;
}

int bvrh_app_set_material(BvrhApp* app, uint32_t entity, const BvrhStandardMaterial* material) {
    if (!app || entity >= app->app.world.entities.size() || !app->app.world.entity(entity).material) return 1;
    StandardMaterial* m = app->app.world.materials.get_mut(*app->app.world.entity(entity).material);
    if (!m) return 1;
    *m = to_material(material);
    return 0;
}

void bvrh_app_set_window_size(BvrhApp* app, uint32_t w, uint32_t h) {
    for (EntityData& e : app->app.world.entities) {
        if (e.window) { e.window->physical_width = w; e.window->physical_height = h; }
        // Bevy's camera_system keeps PerspectiveProjection::aspect_ratio in sync with the target size
        if (e.projection) if (auto* p = std::get_if<PerspectiveProjection>(&*e.projection)) p->aspect_ratio = (float)w / (float)h;
    }
}

void bvrh_app_set_seed(BvrhApp* app, float seed) {
    app->fixed_seed = seed;
    if (seed >= 0.0f) app->app.set_seed_source(&fixed_seed_fn, app);
    else app->app.set_seed_source(nullptr, nullptr);
}

void bvrh_app_set_render_options(BvrhApp* app, const BvrRenderOptions* opts) {
    if (app && opts) app->app.render.node.options = *opts;
}

void bvrh_app_set_gpu_bvh(BvrhApp* app, int enabled) {
    if (app) app->app.render.node.gpu_bvh = enabled != 0;
}

int bvrh_app_set_raster(BvrhApp* app, uint32_t camera, const float* rgba, const float* depth, size_t n_pixels) {
    if (!app || !rgba || !depth) return 1;
    app->app.set_raster(camera, std::vector<float>(rgba, rgba + 4 * n_pixels), std::vector<float>(depth, depth + n_pixels));
    return 0;
}

int bvrh_app_update(BvrhApp* app) {
    if (!app) return -1;
    try {
        return app->app.update();
    } catch (const std::exception& e) {
        app->error = e.what();
        return -1;
    } catch (...) {
        app->error = "unknown exception";
        return -1;
    }
}

const float* bvrh_app_frame(const BvrhApp* app, uint32_t camera, uint32_t* width, uint32_t* height) {
    if (!app) return nullptr;
    for (const auto& v : app->app.render.views) {
        if (v.entity == camera) {
            if (width) *width = v.target.width;
            if (height) *height = v.target.height;
            return v.target.main().data();
        }
    }
    return nullptr;
}

size_t bvrh_app_buffers(const BvrhApp* app, const BvrModel** models, const BvrMaterial** materials,
                        const BvrBvhNode** nodes, size_t* n_nodes) {
    const auto& r = app->app.render;
    if (models) *models = r.model_buffer.buffer.get().data();
    if (materials) *materials = r.material_buffer.buffer.get().data();
    if (nodes) *nodes = r.bvh_buffer.buffer.get().data();
    if (n_nodes) *n_nodes = r.bvh_buffer.buffer.get().size();
    return r.model_buffer.buffer.get().size();
}

int bvrh_app_msaa_off(const BvrhApp* app) { return app->app.world.msaa == Msaa::Off; }

int bvrh_app_has_depth_prepass(const BvrhApp* app, uint32_t entity) {
    return entity < app->app.world.entities.size() && app->app.world.entities[entity].depth_prepass;
}

int bvrh_app_get_stats(BvrhApp* app, BvrStats* out) {
    if (!app || !out || !app->app.render.pipeline || !app->app.render.pipeline->ready()) return BVR_ERR_INVALID_ARGUMENT;
    return bvr_get_stats(app->app.render.pipeline->context(), out);
}

}  // extern "C"
