// plugin.cpp — C++ mirror of the reference's plugin / extract / pipeline host code for the hot path.
// See bevyray_host.hpp for the mapping to src/raytracing/{mod,extract,pipeline}.rs.

#include "bevyray_host.hpp"

#include <chrono>
#include <cmath>
#include <cstring>
#include <stdexcept>

namespace bevyray {

// ------------------------------------------------------------------------------------------------
// maths
// ------------------------------------------------------------------------------------------------
const Vec3 Vec3::ZERO(0, 0, 0);
const Vec3 Vec3::X(1, 0, 0);
const Vec3 Vec3::Y(0, 1, 0);
const Vec3 Vec3::Z(0, 0, 1);
const Color Color::WHITE{1, 1, 1, 1, false};

float Vec3::length() const { return std::sqrt(dot(*this)); }
Vec3 Vec3::normalize() const {
    float inv = 1.0f / length();
    return *this * inv;
}

// glam::Quat::from_mat3 (rotation matrix with the given columns)
Quat Quat::from_mat3(Vec3 ax, Vec3 ay, Vec3 az) {
    const float m00 = ax.x, m01 = ax.y, m02 = ax.z;
    const float m10 = ay.x, m11 = ay.y, m12 = ay.z;
    const float m20 = az.x, m21 = az.y, m22 = az.z;
    Quat q;
    if (m22 <= 0.0f) {
        const float dif10 = m11 - m00, omm22 = 1.0f - m22;
        if (dif10 <= 0.0f) {
            const float four_xsq = omm22 - dif10, inv4x = 0.5f / std::sqrt(four_xsq);
            q = Quat{four_xsq * inv4x, (m01 + m10) * inv4x, (m02 + m20) * inv4x, (m12 - m21) * inv4x};
        } else {
            const float four_ysq = omm22 + dif10, inv4y = 0.5f / std::sqrt(four_ysq);
            q = Quat{(m01 + m10) * inv4y, four_ysq * inv4y, (m12 + m21) * inv4y, (m20 - m02) * inv4y};
        }
    } else {
        const float sum10 = m11 + m00, opm22 = 1.0f + m22;
        if (sum10 <= 0.0f) {
            const float four_zsq = opm22 - sum10, inv4z = 0.5f / std::sqrt(four_zsq);
            q = Quat{(m02 + m20) * inv4z, (m12 + m21) * inv4z, four_zsq * inv4z, (m01 - m10) * inv4z};
        } else {
            const float four_wsq = opm22 + sum10, inv4w = 0.5f / std::sqrt(four_wsq);
            q = Quat{(m12 - m21) * inv4w, (m20 - m02) * inv4w, (m01 - m10) * inv4w, four_wsq * inv4w};
        }
    }
    return q;
}

// glam::Quat * Vec3
Vec3 Quat::mul_vec3(Vec3 v) const {
    const Vec3 b(x, y, z);
    const float b2 = b.dot(b);
    return v * (w * w - b2) + b * (v.dot(b) * 2.0f) + b.cross(v) * (w * 2.0f);
}

// Transform::looking_at -> look_to(target - translation, up): back = -forward, right = up x back, up = back x right
Transform Transform::looking_at(Vec3 target, Vec3 up_dir) const {
    Transform t = *this;
    Vec3 dir = target - translation;
    Vec3 back = -(dir.length() > 0.0f ? dir.normalize() : Vec3(0, 0, -1));
    Vec3 upn = up_dir.length() > 0.0f ? up_dir.normalize() : Vec3::Y;
    Vec3 right_raw = upn.cross(back);
    Vec3 right = right_raw.length() > 0.0f ? right_raw.normalize() : Vec3::X;
    Vec3 up2 = back.cross(right);
    t.rotation = Quat::from_mat3(right, up2, back);
    return t;
}

// bevy_color Srgba -> LinearRgba gamma function
static float srgb_to_linear(float v) {
    if (v <= 0.0f) return v;
    if (v <= 0.04045f) return v / 12.92f;
    return std::pow((v + 0.055f) / 1.055f, 2.4f);
}

Vec3 Color::to_linear_vec3() const {
    if (linear) return Vec3(r, g, b);
    return Vec3(srgb_to_linear(r), srgb_to_linear(g), srgb_to_linear(b));
}

// ------------------------------------------------------------------------------------------------
// extract.rs
// ------------------------------------------------------------------------------------------------
BvrWindow WindowExtract::to_uniform() const {
    BvrWindow w;
    std::memset(&w, 0, sizeof w);
    w.random_seed = random_seed;
    w.height = height;
    return w;
}

BvrMaterial prepare_asset(const StandardMaterial& s) {
    BvrMaterial m;
    const Vec3 c = s.base_color.to_linear_vec3();
    m.base_color[0] = c.x; m.base_color[1] = c.y; m.base_color[2] = c.z;
    m.metallic = s.metallic;
    m.roughness = s.perceptual_roughness;
    m.reflectance = s.reflectance;
    m.ior = s.ior;
    m.specular_transmission = s.specular_transmission;
    return m;
}

std::optional<ExtractedCamera> extract_camera(const RaytracedCamera& camera, const GlobalTransform& transform,
                                              const Projection& projection) {
    const PerspectiveProjection* p = std::get_if<PerspectiveProjection>(&projection);
    if (!p) return std::nullopt;   // "Currently unsupported", extract.rs:148
    ExtractedCamera out;
    std::memset(&out, 0, sizeof out);
    const Vec3 position = transform.translation, direction = transform.forward(), up = transform.up();
    out.camera.sample_count = camera.sample_count;
    out.camera.bounce_count = camera.bounces;
    out.camera.projection = 0;
    out.camera.near_plane = p->near;
    out.camera.far_plane = p->far;
    out.camera.aspect = p->aspect_ratio;
    out.camera.fov = p->fov;
    out.camera.position[0] = position.x; out.camera.position[1] = position.y; out.camera.position[2] = position.z;
    out.camera.direction[0] = direction.x; out.camera.direction[1] = direction.y; out.camera.direction[2] = direction.z;
    out.camera.up[0] = up.x; out.camera.up[1] = up.y; out.camera.up[2] = up.z;
    out.level.level = (uint32_t)camera.level;
    return out;
}

template <class T>
void StorageBuffer<T>::set(std::vector<T> v) {
    if (v.size() != data_.size()) {
        resized_ = true;
        dirty_.clear();
        if (!v.empty()) dirty_.push_back({0u, (uint32_t)v.size()});
    } else {
        // coalesce differing elements into ranges; gaps shorter than 16 elements are bridged
        const uint32_t n = (uint32_t)v.size();
        const uint32_t bridge = 16;
        uint32_t i = 0;
        while (i < n) {
            if (std::memcmp(&v[i], &data_[i], sizeof(T)) == 0) { i++; continue; }
            uint32_t first = i, last = i;
            uint32_t j = i + 1;
            while (j < n && j - last <= bridge) {
                if (std::memcmp(&v[j], &data_[j], sizeof(T)) != 0) last = j;
                j++;
            }
            dirty_.push_back({first, last + 1 - first});
            i = last + 1;
        }
    }
    data_ = std::move(v);
}

template <class T>
std::vector<std::pair<uint32_t, uint32_t>> StorageBuffer<T>::take_dirty() {
    std::vector<std::pair<uint32_t, uint32_t>> out;
    out.swap(dirty_);
    resized_ = false;
    return out;
}

template class StorageBuffer<BvrModel>;
template class StorageBuffer<BvrMaterial>;
template class StorageBuffer<BvrBvhNode>;

void prepare_buffers(ModelBuffer& model_buffer, MaterialBuffer& material_buffer, BVHBuffer& bvh_buffer,
                     const std::vector<SphereQueryItem>& data,
                     const std::vector<std::optional<BvrMaterial>>& render_assets, bool build_bvh) {
    // `let Ok(..) = buffer.lock() else { return }` (extract.rs:287-297): lock() blocks and only fails on a
    // poisoned mutex, which std::mutex cannot be.
    std::lock_guard<std::mutex> l0(model_buffer.mutex), l1(material_buffer.mutex), l2(bvh_buffer.mutex);

    std::vector<BvrModel> all_spheres;
    std::vector<BvrMaterial> all_materials;
    all_spheres.reserve(data.size());
    all_materials.reserve(data.size());
    uint32_t index = 0;
    for (const SphereQueryItem& item : data) {
        if (item.material.id >= render_assets.size() || !render_assets[item.material.id])
            throw std::runtime_error("This should exist");   // extract.rs:302
        all_materials.push_back(*render_assets[item.material.id]);
        BvrModel m;
        std::memset(&m, 0, sizeof m);
        m.position[0] = item.sphere.position.x;
        m.position[1] = item.sphere.position.y;
        m.position[2] = item.sphere.position.z;
        m.radius = item.sphere.radius;
        m.material_id = index++;
        all_spheres.push_back(m);
    }
    if (build_bvh) bvh_buffer.buffer.set(build_ploc(all_spheres, 24));
    model_buffer.buffer.set(std::move(all_spheres));
    material_buffer.buffer.set(std::move(all_materials));
}

// ------------------------------------------------------------------------------------------------
// pipeline.rs
// ------------------------------------------------------------------------------------------------
void ViewTarget::resize(uint32_t w, uint32_t h, const Color& clear) {
    if (w == width && h == height && !main_texture[0].empty()) return;
    width = w; height = h;
    const Vec3 c = clear.to_linear_vec3();
    for (auto& t : main_texture) {
        t.resize((size_t)w * h * 4);
        for (size_t i = 0; i < (size_t)w * h; i++) { t[4 * i] = c.x; t[4 * i + 1] = c.y; t[4 * i + 2] = c.z; t[4 * i + 3] = 1.0f; }
    }
}

ViewTarget::PostProcessWrite ViewTarget::post_process_write() {
    PostProcessWrite w{&main_texture[current], &main_texture[1 - current]};
    current = 1 - current;
    return w;
}

RaytracingPipeline::RaytracingPipeline(int device) {
    int st = bvr_create(device, &ctx_);
    if (st != BVR_OK) {
        error_ = std::string("bvr_create: ") + bvr_status_string(st);
        ctx_ = nullptr;
    }
}

RaytracingPipeline::~RaytracingPipeline() {
    if (ctx_) bvr_destroy(ctx_);
}

template <class T>
static void append_ranges(std::vector<BvrDirtyRange>& out, uint32_t array, StorageBuffer<T>& buf, bool& full) {
    if (buf.resized()) full = true;
    for (auto& r : buf.take_dirty()) out.push_back(BvrDirtyRange{array, r.first, r.second});
}

bool RayTracingNode::run(RaytracingPipeline& pipeline, ViewTarget& view_target, const ViewPrepassTextures& prepass,
                         const BvrRaytraceLevel& level, const BvrCamera& camera, const WindowExtract& window,
                         ModelBuffer& model, MaterialBuffer& material, BVHBuffer& bvh) const {
    if (!pipeline.ready()) return false;                     // pipeline.rs:82-85
    ViewTarget::PostProcessWrite post = view_target.post_process_write();   // pipeline.rs:111
    if (!prepass.depth) return false;                        // pipeline.rs:113-115
    // pipeline.rs:117-130: a poisoned mutex panics there; std::mutex cannot be poisoned
    std::lock_guard<std::mutex> g0(model.mutex), g1(material.mutex), g2(bvh.mutex);
    const auto& models = model.buffer.get();
    const auto& materials = material.buffer.get();
    const auto& nodes = bvh.buffer.get();
    // an empty storage buffer has no binding -> the reference returns early (pipeline.rs:141-151)
    if (models.empty() || materials.empty() || (nodes.empty() && !gpu_bvh)) return false;

    BvrContext* ctx = pipeline.context();
    // pipeline.rs:136-138: three write_buffer calls; here only the dirty element ranges travel
    std::vector<BvrDirtyRange> ranges;
    bool full = false;
    append_ranges(ranges, BVR_ARRAY_MODELS, model.buffer, full);
    append_ranges(ranges, BVR_ARRAY_MATERIALS, material.buffer, full);
    if (!gpu_bvh) append_ranges(ranges, BVR_ARRAY_BVH_NODES, bvh.buffer, full);
    if (full || !ranges.empty()) {
        int st = gpu_bvh
                     ? bvr_upload_scene_gpu_bvh(ctx, models.data(), models.size(), materials.data(), materials.size(),
                                                full ? nullptr : ranges.data(), full ? 0 : ranges.size(), nullptr)
                     : bvr_upload_scene(ctx, models.data(), models.size(), materials.data(), materials.size(),
                                        nodes.data(), nodes.size(), full ? nullptr : ranges.data(),
                                        full ? 0 : ranges.size());
        if (st != BVR_OK) throw std::runtime_error(std::string("bvr_upload_scene: ") + bvr_last_error(ctx));
    }
    BvrRenderOptions opts = options;
    opts.width = view_target.width;
    BvrWindow win = window.to_uniform();
    BvrOutputs out;
    std::memset(&out, 0, sizeof out);
    out.rgba = post.destination->data();
    int st = bvr_render(ctx, &camera, &level, &win, &opts, post.source->data(), prepass.depth->data(), &out);
    if (st != BVR_OK) throw std::runtime_error(std::string("bvr_render: ") + bvr_last_error(ctx));
    return true;
}

// ------------------------------------------------------------------------------------------------
// mod.rs + the frame schedule
// ------------------------------------------------------------------------------------------------
App::App() {
    default_seed_state_ = (uint64_t)std::chrono::high_resolution_clock::now().time_since_epoch().count();
}
App::~App() = default;

void RaytracePlugin::build(App& app) const {
    app.world.msaa = Msaa::Off;              // mod.rs:30
    app.render.has_raytrace_node = true;     // mod.rs:56-71
}

void RaytracePlugin::finish(App& app) const {
    app.render.pipeline = std::make_unique<RaytracingPipeline>(device);   // mod.rs:74-83
}

App& App::add_plugins(const RaytracePlugin& plugin) {
    plugin.build(*this);
    plugin.finish(*this);
    plugin_added_ = true;
    return *this;
}

void App::auto_add_camera_components() {
    for (EntityData& e : world.entities)
        if (e.camera && e.projection && !e.depth_prepass) e.depth_prepass = true;
}

void App::set_raster(Entity camera, std::vector<float> rgba, std::vector<float> depth) {
    for (auto& r : rasters_)
        if (r.first == camera) { r.second.rgba = std::move(rgba); r.second.depth = std::move(depth); return; }
    rasters_.push_back({camera, Raster{std::move(rgba), std::move(depth)}});
}

const std::vector<float>* App::frame(Entity camera) const {
    for (const auto& v : render.views) if (v.entity == camera) return &v.target.main();
    return nullptr;
}

void App::extract() {
    // WindowExtract::extract_component, extract.rs:70-80 — a fresh seed in [0,1) every frame
    render.window.reset();
    const Window* window = nullptr;
    for (const EntityData& e : world.entities) if (e.window) { window = &*e.window; break; }
    if (window) {
        WindowExtract w;
        if (seed_fn_) {
            w.random_seed = seed_fn_(seed_user_);
        } else {
            SeededRng rng(default_seed_state_);
            w.random_seed = rng.next_f32();
            default_seed_state_ = rng.state;
        }
        w.height = window->physical_height;
        render.window = w;
    }
    // RaytracedSphereExtract + Handle<StandardMaterial>, extract.rs:160-179, 31
    render.spheres.clear();
    for (const EntityData& e : world.entities) {
        if (e.raytraced_sphere && e.transform && e.material)
            render.spheres.push_back(SphereQueryItem{RaytracedSphereExtract{e.transform->translation, e.raytraced_sphere->radius}, *e.material});
    }
    // RenderAssetPlugin<RaytraceMaterial>: only changed assets are re-prepared
    if (render.render_assets.size() < world.materials.len()) render.render_assets.resize(world.materials.len());
    for (uint32_t id : world.materials.drain_changed())
        render.render_assets[id] = prepare_asset(*world.materials.get(Handle<StandardMaterial>{id}));
    // CameraExtract, extract.rs:107-158
    std::vector<RenderWorld::View> views;
    for (Entity id = 0; id < world.entities.size(); id++) {
        const EntityData& e = world.entities[id];
        if (!(e.raytraced_camera && e.transform && e.projection && e.camera && window)) continue;
        auto extracted = extract_camera(*e.raytraced_camera, *e.transform, *e.projection);
        if (!extracted) continue;
        RenderWorld::View v;
        bool reused = false;
        for (auto& old : render.views) if (old.entity == id) { v = std::move(old); reused = true; break; }
        (void)reused;
        v.entity = id;
        v.extracted = *extracted;
        v.target.resize(window->physical_width, window->physical_height, e.camera->clear_color);
        const size_t npix = (size_t)window->physical_width * window->physical_height;
        const Raster* raster = nullptr;
        for (const auto& r : rasters_) if (r.first == id) raster = &r.second;
        if (raster && raster->rgba.size() == npix * 4) {
            v.target.main_texture[v.target.current] = raster->rgba;
        } else {
            const Vec3 c = e.camera->clear_color.to_linear_vec3();
            auto& t = v.target.main_texture[v.target.current];
            for (size_t i = 0; i < npix; i++) { t[4 * i] = c.x; t[4 * i + 1] = c.y; t[4 * i + 2] = c.z; t[4 * i + 3] = 1.0f; }
        }
        if (e.depth_prepass) {
            if (raster && raster->depth.size() == npix) v.prepass.depth = raster->depth;
            else v.prepass.depth = std::vector<float>(npix, 0.0f);
        } else {
            v.prepass.depth.reset();
        }
        views.push_back(std::move(v));
    }
    render.views = std::move(views);
}

int App::update() {
    auto_add_camera_components();          // Update schedule, mod.rs:34
    extract();                             // ExtractSchedule
    if (!plugin_added_) return 0;
    // RenderSet::PrepareResources, extract.rs:50
    prepare_buffers(render.model_buffer, render.material_buffer, render.bvh_buffer, render.spheres,
                    render.render_assets, !render.node.gpu_bvh);
    // Core3d graph: Tonemapping -> RaytraceLabel -> EndMainPassPostProcessing, one run per view
    int rendered = 0;
    if (!render.has_raytrace_node || !render.window) return 0;
    for (auto& v : render.views) {
        if (render.node.run(*render.pipeline, v.target, v.prepass, v.extracted.level, v.extracted.camera,
                            *render.window, render.model_buffer, render.material_buffer, render.bvh_buffer))
            rendered++;
    }
    return rendered;
}

}  // namespace bevyray
