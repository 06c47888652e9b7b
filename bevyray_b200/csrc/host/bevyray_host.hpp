// bevyray_host.hpp — C++ host side above the C ABI.
//
// The reference's host side is a Rust/Bevy crate (src/raytracing/{mod,extract,pipeline}.rs and
// src/main.rs).  No Rust toolchain exists in this image, so the host layer is written in C++ and
// mirrors the reference's plugin interface for the hot path: same type names, same field names,
// same defaults, same per-frame order (extract -> prepare_buffers -> RayTracingNode::run).
// The uncompiled Rust binding a maintainer would add is under rust/ and in INTEGRATION.md.
#pragma once

#include <cstddef>
#include <cstdint>
#include <memory>
#include <mutex>
#include <optional>
#include <string>
#include <variant>
#include <vector>

#include "../../../include/bevyray_b200.h"

namespace bevyray {

// ------------------------------------------------------------------------------------------------
// Small maths mirror of the glam / bevy_color pieces the path touches
// ------------------------------------------------------------------------------------------------
struct Vec3 {
    float x = 0, y = 0, z = 0;
    Vec3() = default;
    Vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    static Vec3 splat(float v) { return Vec3(v, v, v); }
    Vec3 operator+(Vec3 o) const { return Vec3(x + o.x, y + o.y, z + o.z); }
    Vec3 operator-(Vec3 o) const { return Vec3(x - o.x, y - o.y, z - o.z); }
    Vec3 operator*(Vec3 o) const { return Vec3(x * o.x, y * o.y, z * o.z); }
    Vec3 operator*(float s) const { return Vec3(x * s, y * s, z * s); }
    Vec3 operator-() const { return Vec3(-x, -y, -z); }
    float dot(Vec3 o) const { return x * o.x + y * o.y + z * o.z; }
    Vec3 cross(Vec3 o) const { return Vec3(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x); }
    float length() const;
    Vec3 normalize() const;
    static const Vec3 ZERO, X, Y, Z;
};

struct Quat {
    float x = 0, y = 0, z = 0, w = 1;
    static Quat from_mat3(Vec3 x_axis, Vec3 y_axis, Vec3 z_axis);
    Vec3 mul_vec3(Vec3 v) const;
};

// bevy_color::Color, restricted to the two spaces the demo uses (src/main.rs:79,88,120,135,...)
struct Color {
    float r = 1, g = 1, b = 1, a = 1;
    bool linear = false;
    static Color srgb(float r, float g, float b) { return Color{r, g, b, 1.0f, false}; }
    static Color srgb_from_array(const float (&v)[3]) { return srgb(v[0], v[1], v[2]); }
    static Color linear_rgb(float r, float g, float b) { return Color{r, g, b, 1.0f, true}; }
    static const Color WHITE;
    // Color::to_linear().to_vec3() — extract.rs:201
    Vec3 to_linear_vec3() const;
};

// ------------------------------------------------------------------------------------------------
// BVH producer (ploc.cpp) — replaces obvhs::ploc::build_ploc at extract.rs:316-332
// ------------------------------------------------------------------------------------------------
void model_aabb(const BvrModel& m, float mn[3], float mx[3]);
std::vector<BvrBvhNode> build_ploc(const std::vector<BvrModel>& models, uint32_t search_distance = 24);
std::string validate_bvh(const std::vector<BvrBvhNode>& nodes, const std::vector<BvrModel>& models);

// ------------------------------------------------------------------------------------------------
// Main-world ECS surface (src/raytracing/mod.rs) and the Bevy types the path reads
// ------------------------------------------------------------------------------------------------

// enum Raytracing — mod.rs:94-101, #[repr(u32)]
enum class Raytracing : uint32_t { Skip = 0, FallbackRaster = 1, FallbackRaytraced = 2, Pure = 3 };

// struct RaytracedCamera — mod.rs:86-91
struct RaytracedCamera {
    Raytracing level = Raytracing::Pure;
    uint32_t sample_count = 1;
    uint32_t bounces = 1;
};

// struct RaytracedSphere — mod.rs:103-106
struct RaytracedSphere {
    float radius = 1.0f;
};

// bevy_transform::Transform; GlobalTransform == Transform here (the demo has no hierarchy).
struct Transform {
    Vec3 translation;
    Quat rotation;
    Vec3 scale{1, 1, 1};
    static Transform from_xyz(float x, float y, float z) { Transform t; t.translation = Vec3(x, y, z); return t; }
    static Transform from_translation(Vec3 v) { Transform t; t.translation = v; return t; }
    Transform looking_at(Vec3 target, Vec3 up) const;   // src/main.rs:57-58
    Vec3 forward() const { return rotation.mul_vec3(Vec3(0, 0, -1)); }   // extract.rs:133
    Vec3 up() const { return rotation.mul_vec3(Vec3(0, 1, 0)); }         // extract.rs:134
};
using GlobalTransform = Transform;

// bevy_render::camera::{PerspectiveProjection, OrthographicProjection, Projection} with Bevy 0.14 defaults
struct PerspectiveProjection {
    float fov = 0.78539816339744830962f;  // pi/4
    float aspect_ratio = 1.0f;
    float near = 0.1f;
    float far = 1000.0f;
};
struct OrthographicProjection {};
using Projection = std::variant<PerspectiveProjection, OrthographicProjection>;

// bevy_pbr::StandardMaterial: the six fields RaytraceMaterial::prepare_asset reads (extract.rs:200-207),
// Bevy 0.14 defaults.
struct StandardMaterial {
    Color base_color = Color{1, 1, 1, 1, false};
    float metallic = 0.0f;
    float perceptual_roughness = 0.5f;
    float reflectance = 0.5f;
    float ior = 1.5f;
    float specular_transmission = 0.0f;
};

struct Window {
    uint32_t physical_width = 1280;   // Bevy's default window
    uint32_t physical_height = 720;
};

struct Camera {
    Color clear_color = Color{1, 1, 1, 1, false};
};

enum class Msaa { Off, Sample4 };

template <class T>
struct Handle {
    uint32_t id = 0xffffffffu;
    bool operator==(const Handle& o) const { return id == o.id; }
};

// bevy_asset::Assets<T> with the change tracking RenderAssetPlugin relies on
template <class T>
class Assets {
public:
    Handle<T> add(const T& v) {
        items_.push_back(v);
        changed_.push_back(1);
        return Handle<T>{(uint32_t)items_.size() - 1};
    }
    const T* get(Handle<T> h) const { return h.id < items_.size() ? &items_[h.id] : nullptr; }
    T* get_mut(Handle<T> h) {
        if (h.id >= items_.size()) return nullptr;
        changed_[h.id] = 1;
        return &items_[h.id];
    }
    size_t len() const { return items_.size(); }
    // ids changed since the last call
    std::vector<uint32_t> drain_changed() {
        std::vector<uint32_t> out;
        for (uint32_t i = 0; i < changed_.size(); i++) if (changed_[i]) { out.push_back(i); changed_[i] = 0; }
        return out;
    }
private:
    std::vector<T> items_;
    std::vector<uint8_t> changed_;
};

using Entity = uint32_t;

// One row of the (tiny) main world: every component the path's queries name.
struct EntityData {
    std::string name;
    std::optional<Transform> transform;
    std::optional<RaytracedSphere> raytraced_sphere;
    std::optional<Handle<StandardMaterial>> material;
    std::optional<RaytracedCamera> raytraced_camera;
    std::optional<Projection> projection;
    std::optional<Camera> camera;
    std::optional<Window> window;
    bool depth_prepass = false;   // DepthPrepass marker, auto-inserted (mod.rs:108-115)
};

class World {
public:
    Entity spawn(EntityData e) { entities.push_back(std::move(e)); return (Entity)entities.size() - 1; }
    EntityData& entity(Entity e) { return entities[e]; }
    std::vector<EntityData> entities;
    Assets<StandardMaterial> materials;
    Msaa msaa = Msaa::Sample4;   // Bevy 0.14 default; RaytracePlugin forces Off (mod.rs:30)
};

// ------------------------------------------------------------------------------------------------
// Render-world side (src/raytracing/extract.rs)
// ------------------------------------------------------------------------------------------------

// WindowExtract — extract.rs:56-81
struct WindowExtract {
    float random_seed = 0;
    uint32_t height = 0;
    BvrWindow to_uniform() const;
};

// RaytracedSphereExtract — extract.rs:160-179
struct RaytracedSphereExtract {
    Vec3 position;
    float radius = 0;
};

// RaytraceMaterial::prepare_asset — extract.rs:191-209
BvrMaterial prepare_asset(const StandardMaterial& source);

// CameraExtract::extract_component — extract.rs:107-158.  nullopt for orthographic (extract.rs:148).
struct ExtractedCamera {
    BvrRaytraceLevel level;
    BvrCamera camera;
};
std::optional<ExtractedCamera> extract_camera(const RaytracedCamera& camera, const GlobalTransform& transform,
                                              const Projection& projection);

// encase StorageBuffer<Vec<T>> + the Mutex wrappers at extract.rs:252-262, with the change
// tracking the reference lists as a TODO (extract.rs:303): set() records which element ranges differ
// from the previous contents so that the upload copies only those.
template <class T>
class StorageBuffer {
public:
    void set(std::vector<T> v);
    const std::vector<T>& get() const { return data_; }
    // element ranges changed since the last take_dirty(); {0,len} after a resize
    std::vector<std::pair<uint32_t, uint32_t>> take_dirty();
    bool resized() const { return resized_; }
private:
    std::vector<T> data_;
    std::vector<std::pair<uint32_t, uint32_t>> dirty_;
    bool resized_ = true;
    template <class U> friend struct LockedBuffer;
};

template <class T>
struct LockedBuffer {
    std::mutex mutex;
    StorageBuffer<T> buffer;
};
using ModelBuffer = LockedBuffer<BvrModel>;
using MaterialBuffer = LockedBuffer<BvrMaterial>;
using BVHBuffer = LockedBuffer<BvrBvhNode>;

struct SphereQueryItem {
    RaytracedSphereExtract sphere;
    Handle<StandardMaterial> material;
};

// prepare_buffers — extract.rs:280-337.  Throws std::runtime_error("This should exist") when a
// material handle has no prepared asset (the reference panics, extract.rs:302).
// build_bvh == false leaves the BVH buffer untouched (the library builds the tree on the GPU instead,
// RayTracingNode::gpu_bvh).
void prepare_buffers(ModelBuffer& model_buffer, MaterialBuffer& material_buffer, BVHBuffer& bvh_buffer,
                     const std::vector<SphereQueryItem>& data,
                     const std::vector<std::optional<BvrMaterial>>& render_assets, bool build_bvh = true);

// ------------------------------------------------------------------------------------------------
// Render graph node + pipeline (src/raytracing/pipeline.rs)
// ------------------------------------------------------------------------------------------------

// What Bevy's own raster passes hand to the node: the post-tonemap main texture (pipeline.rs:111,166)
// and the depth prepass (pipeline.rs:113,169).  fp32 RGBA / fp32 reverse-Z depth, row-major, top-left origin.
struct ViewTarget {
    uint32_t width = 0, height = 0;
    std::vector<float> main_texture[2];   // ping-pong pair behind post_process_write()
    int current = 0;
    struct PostProcessWrite { const std::vector<float>* source; std::vector<float>* destination; };
    PostProcessWrite post_process_write();   // flips the main texture like ViewTarget::post_process_write
    const std::vector<float>& main() const { return main_texture[current]; }
    void resize(uint32_t w, uint32_t h, const Color& clear);
};

struct ViewPrepassTextures {
    std::optional<std::vector<float>> depth;   // None => the node skips the frame (pipeline.rs:113-115)
};

// RaytracingPipeline — pipeline.rs:224-331: owns the GPU-side objects; here the C-ABI context.
class RaytracingPipeline {
public:
    explicit RaytracingPipeline(int device);
    ~RaytracingPipeline();
    RaytracingPipeline(const RaytracingPipeline&) = delete;
    RaytracingPipeline& operator=(const RaytracingPipeline&) = delete;
    BvrContext* context() const { return ctx_; }
    bool ready() const { return ctx_ != nullptr; }
    const std::string& error() const { return error_; }
private:
    BvrContext* ctx_ = nullptr;
    std::string error_;
};

// RayTracingNode::run — pipeline.rs:58-220.  Returns true when the frame was rendered; false when it
// was skipped the way the reference returns Ok(()) early (pipeline not ready, no depth view, empty
// storage buffer -> no binding).  Throws std::runtime_error on a C-ABI error.
struct RayTracingNode {
    BvrRenderOptions options{};   // kernel / traversal / sharding knobs the reference does not have
    bool gpu_bvh = false;         // build the BVH on the GPU (bvr_upload_scene_gpu_bvh) instead of uploading the host tree
    bool run(RaytracingPipeline& pipeline, ViewTarget& view_target, const ViewPrepassTextures& prepass,
             const BvrRaytraceLevel& level, const BvrCamera& camera, const WindowExtract& window,
             ModelBuffer& model, MaterialBuffer& material, BVHBuffer& bvh) const;
};

// The app shell: RaytracePlugin::build/finish (mod.rs:24-84) + one frame of the schedule
// (extract -> PrepareResources -> render graph), src/raytracing/extract.rs:20-51.
class App;
struct RaytracePlugin {
    int device = 0;
    void build(App& app) const;
    void finish(App& app) const;
};

struct RenderWorld {
    std::optional<WindowExtract> window;
    struct View {
        Entity entity;
        ExtractedCamera extracted;
        ViewTarget target;
        ViewPrepassTextures prepass;
    };
    std::vector<View> views;
    std::vector<SphereQueryItem> spheres;
    std::vector<std::optional<BvrMaterial>> render_assets;   // RenderAssets<RaytraceMaterial>
    ModelBuffer model_buffer;
    MaterialBuffer material_buffer;
    BVHBuffer bvh_buffer;
    std::unique_ptr<RaytracingPipeline> pipeline;
    RayTracingNode node;
    bool has_raytrace_node = false;
};

class App {
public:
    App();
    ~App();
    App& add_plugins(const RaytracePlugin& plugin);
    // Runs one frame: Update systems, extract, prepare_buffers, then the node for every view.
    // Returns the number of views rendered.
    int update();
    // Raster inputs for a camera entity (stand-in for Bevy's raster passes).  Without a call the main
    // texture holds the camera's clear colour and the depth prepass is all 0 (nothing rasterised).
    void set_raster(Entity camera, std::vector<float> rgba, std::vector<float> depth);
    // fp32 RGBA main texture of a camera after the last update()
    const std::vector<float>* frame(Entity camera) const;
    // Seed source replacing thread_rng().gen_range(0.0..1.0) at extract.rs:72-73
    void set_seed_source(float (*fn)(void*), void* user) { seed_fn_ = fn; seed_user_ = user; }

    World world;
    RenderWorld render;
private:
    void auto_add_camera_components();   // mod.rs:108-115
    void extract();
    float (*seed_fn_)(void*) = nullptr;
    void* seed_user_ = nullptr;
    uint64_t default_seed_state_;
    bool plugin_added_ = false;
    struct Raster { std::vector<float> rgba, depth; };
    std::vector<std::pair<Entity, Raster>> rasters_;
};

// src/main.rs:49-240 `setup`: the RTIOW-style demo scene and camera, with `rand::random` replaced by
// a seeded generator.  Returns the camera entity.  The rasterised cube (main.rs:76-85) has no
// raytraced component and is not spawned.
Entity setup(World& world, uint64_t seed);

// ------------------------------------------------------------------------------------------------
// Deterministic scene recipes (scene.cpp)
// ------------------------------------------------------------------------------------------------
struct SceneBuffers {
    std::vector<BvrModel> models;
    std::vector<BvrMaterial> materials;
    std::vector<BvrBvhNode> nodes;
};

// Seeded stand-in for rand 0.8's `random::<f32>()` (uniform in [0,1) with 24 bits).
struct SeededRng {
    uint64_t state;
    explicit SeededRng(uint64_t seed) : state(seed) {}
    uint64_t next_u64();
    float next_f32();
};

// Runs extract + prepare_buffers on a world and returns the three byte buffers.
SceneBuffers scene_from_world(World& world);
// RTIOW book-1 final scene via setup() (configs C1-C3)
SceneBuffers scene_rtiow(uint64_t seed);
// n random spheres, centres uniform in a cube of side `side`, radii U[rmin,rmax], 80/15/5 % material mix
// as src/main.rs:116-179 (configs C4, C5)
SceneBuffers scene_random(uint64_t seed, uint32_t n, float side, float rmin, float rmax);
// C5: centres of scene_random moved by a closed-form function of the frame index; rebuilds the BVH.
void animate_random(SceneBuffers& scene, const std::vector<BvrModel>& base, uint32_t frame, bool rebuild_bvh = true);

}  // namespace bevyray
