// scene.cpp — deterministic scene recipes on top of the ECS mirror.
//
// setup() restates src/main.rs:49-240 with `rand::random` replaced by a seeded generator (the
// reference's scene is unseeded and therefore not reproducible).  scene_random()/animate_random()
// are the synthetic stress scenes of BASELINE.json configs C4 / C5.

#include "bevyray_host.hpp"

#include <cmath>
#include <cstring>

namespace bevyray {

// splitmix64
uint64_t SeededRng::next_u64() {
    uint64_t z = (state += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

// rand 0.8 `Standard` for f32: 24 random bits scaled by 2^-24 -> [0,1)
float SeededRng::next_f32() { return (float)(next_u64() >> 40) * (1.0f / 16777216.0f); }

static Entity spawn_sphere(World& world, const StandardMaterial& material, Transform transform, float radius) {
    EntityData e;
    e.transform = transform;
    e.material = world.materials.add(material);
    e.raytraced_sphere = RaytracedSphere{radius};
    return world.spawn(std::move(e));
}

Entity setup(World& world, uint64_t seed) {
    SeededRng rng(seed);
    auto random = [&rng]() { return rng.next_f32(); };

    // window (Bevy's default primary window) and camera, src/main.rs:55-73
    {
        EntityData w;
        w.name = "Primary Window";
        w.window = Window{};
        world.spawn(std::move(w));
    }
    EntityData cam;
    cam.name = "Raytraced Camera";
    cam.transform = Transform::from_translation(Vec3(0.0f, 0.0f, 5.0f)).looking_at(Vec3(), Vec3::Y);
    cam.camera = Camera{Color::WHITE};
    PerspectiveProjection proj;
    proj.aspect_ratio = 1280.0f / 720.0f;   // Bevy's camera_system sets it from the render target
    cam.projection = Projection{proj};
    cam.raytraced_camera = RaytracedCamera{Raytracing::FallbackRaytraced, 4, 4};
    const Entity camera = world.spawn(std::move(cam));

    // ground, src/main.rs:87-103
    {
        StandardMaterial m;
        m.base_color = Color::srgb(0.5f, 0.5f, 0.5f);
        m.metallic = 0.0f;
        spawn_sphere(world, m, Transform::from_xyz(0.0f, -1000.0f, 0.0f), 1000.0f);
    }

    // the grid of small spheres, src/main.rs:105-186
    for (int a = -11; a <= 11; a++) {
        for (int b = -11; b < 11; b++) {
            const float choose_mat = random();
            const float cx = (float)a + 0.9f * random();
            const float cz = (float)b + 0.9f * random();
            const Vec3 center_v(cx, 0.2f, cz);
            const Transform center = Transform::from_xyz(center_v.x, center_v.y, center_v.z);
            if ((center_v - Vec3(4.0f, 0.2f, 0.0f)).length() > 0.9f) {
                StandardMaterial m;
                if (choose_mat < 0.8f) {            // diffuse
                    const float a0 = random(), a1 = random(), a2 = random();
                    const float b0 = random(), b1 = random(), b2 = random();
                    const float albedo[3] = {a0 * b0, a1 * b1, a2 * b2};
                    m.base_color = Color::srgb_from_array(albedo);
                    m.metallic = 0.0f;
                } else if (choose_mat < 0.95f) {    // metal
                    const float albedo[3] = {random(), random(), random()};
                    const float roughness = random();
                    m.base_color = Color::srgb_from_array(albedo);
                    m.metallic = 1.0f;
                    m.perceptual_roughness = roughness;
                } else {                            // glass
                    m.metallic = 0.0f;
                    m.ior = 1.5f;
                    m.specular_transmission = 1.0f;
                }
                spawn_sphere(world, m, center, 0.2f);
            }
        }
    }

    // the three big spheres, src/main.rs:188-239
    {
        StandardMaterial m;
        m.metallic = 0.0f; m.ior = 1.5f; m.specular_transmission = 1.0f;
        spawn_sphere(world, m, Transform::from_xyz(0.0f, 1.0f, 0.0f), 1.0f);
    }
    {
        StandardMaterial m;
        m.base_color = Color::srgb(0.4f, 0.2f, 0.1f); m.metallic = 0.0f;
        spawn_sphere(world, m, Transform::from_xyz(-4.0f, 1.0f, 0.0f), 1.0f);
    }
    {
        StandardMaterial m;
        m.base_color = Color::srgb(0.7f, 0.6f, 0.5f); m.metallic = 1.0f; m.perceptual_roughness = 0.0f;
        spawn_sphere(world, m, Transform::from_xyz(4.0f, 1.0f, 0.0f), 1.0f);
    }
    return camera;
}

SceneBuffers scene_from_world(World& world) {
    std::vector<SphereQueryItem> spheres;
    for (const EntityData& e : world.entities)
        if (e.raytraced_sphere && e.transform && e.material)
            spheres.push_back(SphereQueryItem{RaytracedSphereExtract{e.transform->translation, e.raytraced_sphere->radius}, *e.material});
    std::vector<std::optional<BvrMaterial>> assets(world.materials.len());
    for (uint32_t id = 0; id < world.materials.len(); id++)
        assets[id] = prepare_asset(*world.materials.get(Handle<StandardMaterial>{id}));
    ModelBuffer mb; MaterialBuffer tb; BVHBuffer bb;
    prepare_buffers(mb, tb, bb, spheres, assets);
    SceneBuffers out;
    out.models = mb.buffer.get();
    out.materials = tb.buffer.get();
    out.nodes = bb.buffer.get();
    return out;
}

SceneBuffers scene_rtiow(uint64_t seed) {
    World world;
    setup(world, seed);
    return scene_from_world(world);
}

SceneBuffers scene_random(uint64_t seed, uint32_t n, float side, float rmin, float rmax) {
    SeededRng rng(seed ^ 0x5ce9e5ull);
    SceneBuffers out;
    out.models.resize(n);
    out.materials.resize(n);
    for (uint32_t i = 0; i < n; i++) {
        BvrModel& m = out.models[i];
        std::memset(&m, 0, sizeof m);
        for (int k = 0; k < 3; k++) m.position[k] = (rng.next_f32() - 0.5f) * side;
        m.radius = rmin + (rmax - rmin) * rng.next_f32();
        m.material_id = i;
        StandardMaterial sm;
        const float choose_mat = rng.next_f32();
        if (choose_mat < 0.8f) {
            const float albedo[3] = {rng.next_f32() * rng.next_f32(), rng.next_f32() * rng.next_f32(),
                                     rng.next_f32() * rng.next_f32()};
            sm.base_color = Color::srgb_from_array(albedo);
        } else if (choose_mat < 0.95f) {
            const float albedo[3] = {rng.next_f32(), rng.next_f32(), rng.next_f32()};
            sm.base_color = Color::srgb_from_array(albedo);
            sm.metallic = 1.0f;
            sm.perceptual_roughness = rng.next_f32();
        } else {
            sm.specular_transmission = 1.0f;
            sm.ior = 1.5f;
        }
        out.materials[i] = prepare_asset(sm);
    }
    out.nodes = build_ploc(out.models, 24);
    return out;
}

void animate_random(SceneBuffers& scene, const std::vector<BvrModel>& base, uint32_t frame, bool rebuild_bvh) {
    const float t = (float)frame * (1.0f / 60.0f);
    scene.models = base;
    for (size_t i = 0; i < base.size(); i++) {
        // every 4th sphere bobs on a closed-form orbit; the rest stay put (so most elements are clean)
        if (i % 4 != 0) continue;
        const float phase = (float)(i % 97) * 0.0647f;
        scene.models[i].position[0] = base[i].position[0] + 0.5f * std::sin(2.0f * t + phase);
        scene.models[i].position[1] = base[i].position[1] + 0.5f * std::cos(3.0f * t + phase);
    }
    if (rebuild_bvh) scene.nodes = build_ploc(scene.models, 24);   // (callers that build the tree on the GPU skip this)
}

}  // namespace bevyray
