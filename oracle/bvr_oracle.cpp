// bvr_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Scalar, strict-IEEE CPU restatement of bevyray's WGSL fragment shader.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
// library; nothing under bevyray_b200/ links, imports or calls it.
//
// PARITY PINNING: the reference has no tests, golden vectors or fixtures for this path
//  and cannot be built or run in this image (no Rust toolchain, no Vulkan / lavapipe).
// The oracle is therefore pinned only against known-answer vectors DERIVED from the WGSL text
// (tests/test_oracle_kat.py) — "parity unpinned" against a running reference.
//
// Every function cites the reference lines it restates (paths relative to the reference tree).
// Arithmetic conventions the WGSL leaves to the implementation, fixed here and mirrored by the CUDA
// kernels:
//   * all f32 arithmetic is IEEE round-to-nearest-even, NO fused multiply-add
//     (build with -ffp-contract=off), division and sqrt correctly rounded;
//   * dot(a,b) = (a.x*b.x + a.y*b.y) + a.z*b.z ; length = sqrt(dot(v,v)); normalize(v) = v / length(v);
//   * min/max follow IEEE minNum/maxNum (a NaN operand is ignored) — fminf/fmaxf;
//   * pow(x, 5.0) (raytrace.wgsl:415) is evaluated as ((x*x)*(x*x))*x;
//   * tan(fov*0.5) (raytrace.wgsl:151) depends only on uniforms: evaluated in double and rounded to f32;
//   * u32(f32) truncates toward zero and saturates (negative / NaN -> 0);
//   * the fullscreen-triangle uv of pixel (x,y) is ((x+0.5)/W, (y+0.5)/H) in f32, (0,0) top-left;
//   * `a || b` short-circuits (WGSL spec) at raytrace.wgsl:269;
//   * textureSample of the raster colour / depth at a pixel centre returns that texel.
//
// SENSITIVITY VARIANTS (tools/oracle_sensitivity.py; never used as the checker): what a real WGSL toolchain may
// legally do differently.  Each macro switches ONE convention; the tool reports how far the image moves.
//   BVRO_VAR_POW_EXP2LOG2   pow(x, 5) = exp2(5 * log2(x)) (how SPIR-V / MSL back ends usually lower pow)
//   BVRO_VAR_RSQRT          normalize(v) = v * inversesqrt(dot(v, v))
//   BVRO_VAR_NAN_MINMAX     min / max propagate a NaN operand (the slab test at raytrace.wgsl:390-393)
//   BVRO_VAR_EAGER_OR       `||` evaluates both sides (older naga): one more RNG draw when cannot_refract
//   (FMA contraction is a compiler flag: -ffp-contract=fast -mfma)

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/bevyray_b200.h"

namespace {

// assets/shaders/const.wgsl:2
constexpr float INF = 3.40282347e+38f;

struct V3 { float x, y, z; };

inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator*(float s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
inline V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
inline V3 neg(V3 a) { return V3{-a.x, -a.y, -a.z}; }
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 a, V3 b) {
    return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
#ifdef BVRO_VAR_RSQRT
inline V3 normalize(V3 a) { return a * (1.0f / std::sqrt(dot(a, a))); }
#else
inline V3 normalize(V3 a) { return a / std::sqrt(dot(a, a)); }
#endif
#ifdef BVRO_VAR_NAN_MINMAX
inline float wmin(float a, float b) { return (a != a || b != b) ? NAN : (a < b ? a : b); }
inline float wmax(float a, float b) { return (a != a || b != b) ? NAN : (a > b ? a : b); }
#else
inline float wmin(float a, float b) { return fminf(a, b); }
inline float wmax(float a, float b) { return fmaxf(a, b); }
#endif
inline V3 ld3(const float* p) { return V3{p[0], p[1], p[2]}; }

struct Counters {
    uint64_t rays = 0, paths = 0, node_pops = 0, inner_visits = 0, box_tests = 0, sphere_tests = 0,
             hits_shaded = 0, rng_draws = 0, stack_truncations = 0, max_stack = 0;
    void merge(const Counters& o) {
        rays += o.rays; paths += o.paths; node_pops += o.node_pops; inner_visits += o.inner_visits;
        box_tests += o.box_tests; sphere_tests += o.sphere_tests; hits_shaded += o.hits_shaded;
        rng_draws += o.rng_draws; stack_truncations += o.stack_truncations;
        max_stack = std::max(max_stack, o.max_stack);
    }
};

// u32(f32) — saturating truncation
inline uint32_t f32_to_u32(float f) {
    if (!(f > 0.0f)) return 0u;
    if (f >= 4294967296.0f) return 0xffffffffu;
    return (uint32_t)f;
}

// assets/shaders/random.wgsl:8-15
inline void rng_next_int(uint32_t& state) {
    uint32_t old_state = state + 747796405u + 2891336453u;
    uint32_t word = ((old_state >> ((old_state >> 28u) + 4u)) ^ old_state) * 277803737u;
    state = (word >> 22u) ^ word;
}

// assets/shaders/random.wgsl:3-6.  f32(0xffffffffu) == 4294967296.0
inline float rng_next_float(uint32_t& state, Counters& c) {
    rng_next_int(state);
    c.rng_draws++;
    return (float)state / 4294967296.0f;
}

// assets/shaders/random.wgsl:17-30 (randomUnitVec3 is the un-normalised in-ball point)
inline V3 random_unit_vec3(uint32_t& state, Counters& c) {
    V3 p;
    for (;;) {
        float x = rng_next_float(state, c);
        float y = rng_next_float(state, c);
        float z = rng_next_float(state, c);
        p = 2.0f * v3(x, y, z) - v3(1.0f, 1.0f, 1.0f);
        if (dot(p, p) <= 1.0f) break;
    }
    return p;
}

struct Ray { V3 origin, direction; };

struct HitInfo {
    float distance;
    V3 position, normal;
    uint32_t material;
    bool front_face;
    uint32_t model;   // oracle-only bookkeeping: index of the model hit (not in the WGSL struct)
};

struct Scene {
    const BvrModel* models; size_t n_models;
    const BvrMaterial* materials; size_t n_materials;
    const BvrBvhNode* nodes; size_t n_nodes;
    BvrCamera camera; uint32_t level; BvrWindow window;
    float tan_half_fov;
    bool brute_force;
    bool diagnose = false;   // BVRO_DIAGNOSE=1: cross-check every ray against a near-first traversal (debug aid)
};

// raytrace.wgsl:371-383
inline float hit_sphere(const BvrModel& sphere, const Ray& ray) {
    V3 oc = ld3(sphere.position) - ray.origin;
    float a = dot(ray.direction, ray.direction);
    float h = dot(ray.direction, oc);
    float c = dot(oc, oc) - sphere.radius * sphere.radius;
    float discriminant = h * h - a * c;
    if (discriminant < 0.0f) return -1.0f;
    return (h - std::sqrt(discriminant)) / a;
}

// raytrace.wgsl:387-398
inline float ray_bounding_dst(const Ray& ray, V3 box_min, V3 box_max) {
    V3 inv = v3(1.0f / ray.direction.x, 1.0f / ray.direction.y, 1.0f / ray.direction.z);
    V3 t_min = (box_min - ray.origin) * inv;
    V3 t_max = (box_max - ray.origin) * inv;
    V3 t1 = v3(wmin(t_min.x, t_max.x), wmin(t_min.y, t_max.y), wmin(t_min.z, t_max.z));
    V3 t2 = v3(wmax(t_min.x, t_max.x), wmax(t_min.y, t_max.y), wmax(t_min.z, t_max.z));
    float t_near = wmax(wmax(t1.x, t1.y), t1.z);
    float t_far = wmin(wmin(t2.x, t2.y), t2.z);
    bool hit = t_far >= t_near && t_far > 0.0f;
    return hit ? (t_near > 0.0f ? t_near : 0.0f) : INF;
}

// raytrace.wgsl:348-362
inline void raycast_against_range(const Scene& s, const Ray& ray, uint32_t start, uint32_t amount,
                                  HitInfo& closest, Counters& c) {
    for (uint32_t i = start; i < start + amount; i++) {
        if (i >= s.n_models) break;  // WGSL robust access would clamp; never reached for a valid BVH
        const BvrModel& model = s.models[i];
        c.sphere_tests++;
        float d = hit_sphere(model, ray);
        if (d != -1.0f && d > 0.001f) {
            if (d < closest.distance) {
                V3 p = ray.origin + d * ray.direction;           // ray_at, raytrace.wgsl:130-132
                V3 n = normalize(p - ld3(model.position));
                closest = HitInfo{d, p, n, model.material_id, dot(ray.direction, n) < 0.0f, i};
            }
        }
    }
}

// Debug aid (BVRO_DIAGNOSE=1): the near-first, distance-culled traversal the CUDA kernels use, with the reference's
// own box arithmetic.  Prints every ray on which it and the reference-order traversal disagree.
inline void diagnose_near_first(const Scene& s, const Ray& ray, const HitInfo& ref) {
    HitInfo closest{INF, v3(0, 0, 0), v3(0, 0, 0), 0u, true, 0xffffffffu};
    Counters dummy;
    struct E { uint32_t node; float d; };
    std::vector<E> stack;
    uint32_t cur = 0;
    bool have = s.n_nodes > 0;
    while (have) {
        const BvrBvhNode& node = s.nodes[cur];
        have = false;
        if (node.model_count > 0) {
            raycast_against_range(s, ray, node.index, node.model_count, closest, dummy);
        } else {
            const BvrBvhNode& n1 = s.nodes[node.index];
            const BvrBvhNode& n2 = s.nodes[node.index + 1];
            float d1 = ray_bounding_dst(ray, ld3(n1.bounds_min), ld3(n1.bounds_max));
            float d2 = ray_bounding_dst(ray, ld3(n2.bounds_min), ld3(n2.bounds_max));
            bool h1 = d1 != INF && d1 < closest.distance, h2 = d2 != INF && d2 < closest.distance;
            if (h1 && h2) {
                bool first1 = d1 < d2;
                stack.push_back(first1 ? E{node.index + 1, d2} : E{node.index, d1});
                cur = first1 ? node.index : node.index + 1; have = true;
            } else if (h1) { cur = node.index; have = true; }
            else if (h2) { cur = node.index + 1; have = true; }
        }
        while (!have && !stack.empty()) {
            E e = stack.back(); stack.pop_back();
            if (e.d < closest.distance) { cur = e.node; have = true; }
        }
    }
    if (closest.model == ref.model && closest.distance == ref.distance) return;
    std::fprintf(stderr, "DIAG ray o=(%.9g %.9g %.9g) d=(%.9g %.9g %.9g): reference hit model %u t=%.9g | near-first hit model %u t=%.9g\n",
                 ray.origin.x, ray.origin.y, ray.origin.z, ray.direction.x, ray.direction.y, ray.direction.z,
                 ref.model, ref.distance, closest.model, closest.distance);
    // the chain of boxes above the reference's hit: node index, box distance
    const uint32_t want = ref.model != 0xffffffffu ? ref.model : closest.model;
    std::vector<uint32_t> path;
    std::vector<std::pair<uint32_t, int>> dfs;   // node, depth
    dfs.push_back({0u, 0});
    std::vector<uint32_t> cur_path;
    while (!dfs.empty()) {
        auto [n, depth] = dfs.back(); dfs.pop_back();
        cur_path.resize((size_t)depth); cur_path.push_back(n);
        const BvrBvhNode& nd = s.nodes[n];
        if (nd.model_count > 0) { if (want >= nd.index && want < nd.index + nd.model_count) { path = cur_path; break; } }
        else { dfs.push_back({nd.index, depth + 1}); dfs.push_back({nd.index + 1, depth + 1}); }
    }
    for (uint32_t n : path) {
        const BvrBvhNode& nd = s.nodes[n];
        std::fprintf(stderr, "   node %u box dst %.9g  min=(%.6g %.6g %.6g) max=(%.6g %.6g %.6g)%s\n", n,
                     ray_bounding_dst(ray, ld3(nd.bounds_min), ld3(nd.bounds_max)), nd.bounds_min[0], nd.bounds_min[1],
                     nd.bounds_min[2], nd.bounds_max[0], nd.bounds_max[1], nd.bounds_max[2], nd.model_count ? " (leaf)" : "");
    }
    if (want != 0xffffffffu) {
        const BvrModel& m = s.models[want];
        std::fprintf(stderr, "   sphere %u c=(%.9g %.9g %.9g) r=%.9g hit_sphere=%.9g\n", want, m.position[0], m.position[1], m.position[2],
                     m.radius, hit_sphere(m, ray));
    }
}

// raytrace.wgsl:313-346
inline HitInfo raycast_impl(const Scene& s, const Ray& ray, Counters& c);
inline HitInfo raycast(const Scene& s, const Ray& ray, Counters& c) {
    HitInfo h = raycast_impl(s, ray, c);
    if (s.diagnose && !s.brute_force) diagnose_near_first(s, ray, h);
    return h;
}
inline HitInfo raycast_impl(const Scene& s, const Ray& ray, Counters& c) {
    HitInfo closest{INF, v3(0, 0, 0), v3(0, 0, 0), 0u, true, 0xffffffffu};
    c.rays++;
    if (s.brute_force) {   // oracle-only: no BVH, every model tested in buffer order
        raycast_against_range(s, ray, 0u, (uint32_t)s.n_models, closest, c);
        return closest;
    }
    if (s.n_nodes == 0) return closest;
    constexpr int STACKSIZE = 32;                        // raytrace.wgsl:310
    uint32_t stack[STACKSIZE] = {0};
    int stack_index = 1;
    while (stack_index > 0 && stack_index < STACKSIZE) {
        stack_index--;
        uint32_t next = stack[stack_index];
        const BvrBvhNode& node = s.nodes[next];
        c.node_pops++;
        if (node.model_count > 0) {
            raycast_against_range(s, ray, node.index, node.model_count, closest, c);
        } else {
            c.inner_visits++;
            const BvrBvhNode& n1 = s.nodes[node.index];
            c.box_tests++;
            float d1 = ray_bounding_dst(ray, ld3(n1.bounds_min), ld3(n1.bounds_max));
            if (d1 != INF && d1 < closest.distance) { stack[stack_index] = node.index; stack_index++; }
            const BvrBvhNode& n2 = s.nodes[node.index + 1];
            c.box_tests++;
            float d2 = ray_bounding_dst(ray, ld3(n2.bounds_min), ld3(n2.bounds_max));
            if (d2 != INF && d2 < closest.distance) { stack[stack_index] = node.index + 1; stack_index++; }
            if ((uint64_t)stack_index > c.max_stack) c.max_stack = (uint64_t)stack_index;
        }
    }
    if (stack_index >= STACKSIZE) c.stack_truncations++;
    return closest;
}

// raytrace.wgsl:400-402
inline V3 reflect(V3 v, V3 n) { return v - (2.0f * dot(v, n)) * n; }

// raytrace.wgsl:404-409
inline V3 refract(V3 v, V3 n, float etai_over_etat) {
    float cos_theta = fminf(dot(neg(v), n), 1.0f);
    V3 r_out_perp = etai_over_etat * (v + cos_theta * n);
    V3 r_out_parallel = (-std::sqrt(std::fabs(1.0f - dot(r_out_perp, r_out_perp)))) * n;
    return r_out_perp + r_out_parallel;
}

// raytrace.wgsl:411-416
inline float reflectance(float cosine, float refraction_index) {
    float r0 = (1.0f - refraction_index) / (1.0f + refraction_index);
    r0 = r0 * r0;
    float x = 1.0f - cosine;
#ifdef BVRO_VAR_POW_EXP2LOG2
    float x5 = std::exp2(5.0f * std::log2(x));
#else
    float x2 = x * x;
    float x5 = (x2 * x2) * x;
#endif
    return r0 + (1.0f - r0) * x5;
}

// raytrace.wgsl:418-421
inline bool vec3_near_zero(V3 v) {
    const float s = 1e-8f;
    return std::fabs(v.x) < s && std::fabs(v.y) < s && std::fabs(v.z) < s;
}

// raytrace.wgsl:231-299 — returns whether the ray was absorbed
inline bool scatter(const Scene& s, Ray& scattered, V3& attenuation, const HitInfo& hit,
                    uint32_t& state, Counters& c) {
    const BvrMaterial& m = s.materials[hit.material < s.n_materials ? hit.material : s.n_materials - 1];
    c.hits_shaded++;
    if (rng_next_float(state, c) < m.metallic) {
        V3 reflected = normalize(reflect(scattered.direction, hit.normal)) +
                       (m.roughness * random_unit_vec3(state, c));
        scattered = Ray{hit.position, reflected};
        attenuation = ld3(m.base_color);
        return dot(scattered.direction, hit.normal) < 0.0f;
    } else {
        if (rng_next_float(state, c) < m.specular_transmission) {
            float ri = hit.front_face ? 1.0f / m.ior : m.ior;
            V3 unit_direction = normalize(scattered.direction);
            float cos_theta = fminf(dot(neg(unit_direction), hit.normal), 1.0f);
            float sin_theta = std::sqrt(1.0f - cos_theta * cos_theta);
            bool cannot_refract = ri * sin_theta > 1.0f;
            V3 direction;
#ifdef BVRO_VAR_EAGER_OR
            const bool schlick = reflectance(cos_theta, ri) > rng_next_float(state, c);
            if (cannot_refract | schlick) {
#else
            if (cannot_refract || reflectance(cos_theta, ri) > rng_next_float(state, c)) {
#endif
                direction = reflect(unit_direction, hit.normal);
            } else {
                direction = refract(unit_direction, hit.normal, ri);
            }
            scattered = Ray{hit.position, direction};
            attenuation = v3(1.0f, 1.0f, 1.0f);
            return false;
        } else {
            V3 b1 = random_unit_vec3(state, c);
            V3 b2 = random_unit_vec3(state, c);
            V3 scatter_direction = (hit.normal + b1) + (m.roughness * b2);
            if (vec3_near_zero(scatter_direction)) scatter_direction = hit.normal;
            scattered = Ray{hit.position, scatter_direction};
            attenuation = ld3(m.base_color);
            return dot(scattered.direction, hit.normal) < 0.0f;
        }
    }
}

// raytrace.wgsl:364-369
inline V3 background_gradient(const Ray& ray) {
    V3 unit = normalize(ray.direction);
    float a = 0.5f * (unit.y + 1.0f);
    return (1.0f - a) * v3(1.0f, 1.0f, 1.0f) + a * v3(0.5f, 0.7f, 1.0f);
}

struct RaytraceResult { V3 color; float depth; };

// raytrace.wgsl:174-224.  primary_id / primary_t report the first raycast of this path.
inline RaytraceResult raytrace(const Scene& s, Ray ray, uint32_t& state, Counters& c,
                               uint32_t* primary_id, float* primary_t) {
    float fallback_far = (s.level == 1u) ? s.camera.far_plane + 10.0f : s.camera.far_plane - 1.0f;
    float first_depth = INF;
    V3 ray_color = v3(1.0f, 1.0f, 1.0f);
    V3 light = v3(0.0f, 0.0f, 0.0f);
    uint32_t bounce = 0;
    for (; bounce <= s.camera.bounce_count; bounce++) {
        HitInfo hit = raycast(s, ray, c);
        if (bounce == 0) {
            first_depth = hit.distance;
            if (primary_id) *primary_id = (hit.distance == INF) ? 0xffffffffu : hit.model;
            if (primary_t) *primary_t = hit.distance;
        }
        if (hit.distance == INF) { light = background_gradient(ray); break; }
        V3 attenuation;
        bool absorbed = scatter(s, ray, attenuation, hit, state, c);
        if (absorbed) break;
        ray_color = ray_color * attenuation;
    }
    if (bounce == s.camera.bounce_count + 1u) ray_color = v3(0.0f, 0.0f, 0.0f);
    if (first_depth == INF) first_depth = fallback_far;
    V3 lin = ray_color * light;
    // linear_to_gamma_Vec3, raytrace.wgsl:226-228
    return RaytraceResult{v3(std::sqrt(lin.x), std::sqrt(lin.y), std::sqrt(lin.z)), first_depth};
}

// raytrace.wgsl:139-156
inline Ray random_ray_from_uv(const Scene& s, float u, float v, uint32_t& state, Counters& c) {
    float rx = rng_next_float(state, c) - 0.5f;
    float ry = rng_next_float(state, c) - 0.5f;
    float height = (float)s.window.height;
    float width = (float)s.window.height * s.camera.aspect;
    float delta_u = (1.0f / width) * rx;
    float delta_v = (1.0f / height) * ry;
    float ndc_x = (u * 2.0f - 1.0f) + delta_u;
    float ndc_y = (1.0f - v * 2.0f) + delta_v;
    V3 dir = ld3(s.camera.direction), up = ld3(s.camera.up);
    V3 right = cross(dir, up);
    float scale = s.tan_half_fov;
    V3 d = normalize((dir + (((ndc_x * s.camera.aspect) * scale) * right)) + ((ndc_y * scale) * up));
    return Ray{ld3(s.camera.position), d};
}

}  // namespace

extern "C" {

struct BvroCounters {
    uint64_t rays, paths, node_pops, inner_visits, box_tests, sphere_tests, hits_shaded, rng_draws,
             stack_truncations, max_stack;
};

// random.wgsl:8-15 — exposed for the known-answer tests
uint32_t bvro_rng_next_int(uint32_t state) { rng_next_int(state); return state; }
float bvro_rng_float_of_state(uint32_t state) { return (float)state / 4294967296.0f; }

// raytrace.wgsl:95
uint32_t bvro_pixel_seed(float random_seed, uint32_t x, uint32_t y, uint32_t width, uint32_t height) {
    float u = ((float)x + 0.5f) / (float)width;
    float v = ((float)y + 0.5f) / (float)height;
    return f32_to_u32(((random_seed * 10000.0f) * (u * 402.0f)) * (v * 31.5f));
}

float bvro_tan_half_fov(float fov) { return (float)std::tan((double)(fov * 0.5f)); }

float bvro_hit_sphere(const BvrModel* m, const float* origin, const float* dir) {
    return hit_sphere(*m, Ray{ld3(origin), ld3(dir)});
}
float bvro_ray_bounding_dst(const float* origin, const float* dir, const float* bmin, const float* bmax) {
    return ray_bounding_dst(Ray{ld3(origin), ld3(dir)}, ld3(bmin), ld3(bmax));
}

// ---- single functions of the shader, exposed for the known-answer tests (tests/test_oracle_kat.py) ----
// raytrace.wgsl:400-416, 364-369
void bvro_reflect(const float* v, const float* n, float* out) { V3 r = reflect(ld3(v), ld3(n)); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
void bvro_refract(const float* v, const float* n, float ratio, float* out) { V3 r = refract(ld3(v), ld3(n), ratio); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
float bvro_reflectance(float cosine, float ri) { return reflectance(cosine, ri); }
void bvro_background_gradient(const float* dir, float* out) {
    V3 r = background_gradient(Ray{v3(0, 0, 0), ld3(dir)});
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
// raytrace.wgsl:139-156: camera ray of the pixel whose uv is (u, v); *state advances by the two jitter draws
void bvro_random_ray_from_uv(const BvrCamera* camera, const BvrWindow* window, float u, float v, uint32_t* state,
                             float* origin, float* dir) {
    Scene s{};
    s.camera = *camera; s.window = *window; s.tan_half_fov = bvro_tan_half_fov(camera->fov);
    Counters c;
    Ray r = random_ray_from_uv(s, u, v, *state, c);
    origin[0] = r.origin.x; origin[1] = r.origin.y; origin[2] = r.origin.z;
    dir[0] = r.direction.x; dir[1] = r.direction.y; dir[2] = r.direction.z;
}
// raytrace.wgsl:231-299: one scatter event on `material` for a ray of direction ray_dir that hit at hit_pos with
// hit_normal / front_face.  Returns absorbed; *state advances by every draw taken.
int bvro_scatter(const BvrMaterial* material, const float* ray_dir, const float* hit_pos, const float* hit_normal,
                 int front_face, uint32_t* state, float* out_dir, float* out_attenuation) {
    Scene s{};
    s.materials = material; s.n_materials = 1;
    Counters c;
    Ray ray{v3(0, 0, 0), ld3(ray_dir)};
    HitInfo hit{1.0f, ld3(hit_pos), ld3(hit_normal), 0u, front_face != 0, 0u};
    V3 att = v3(0, 0, 0);
    const bool absorbed = scatter(s, ray, att, hit, *state, c);
    out_dir[0] = ray.direction.x; out_dir[1] = ray.direction.y; out_dir[2] = ray.direction.z;
    out_attenuation[0] = att.x; out_attenuation[1] = att.y; out_attenuation[2] = att.z;
    return absorbed ? 1 : 0;
}

// One `fragment` invocation per pixel (raytrace.wgsl:93-123) for rows [y0, y1).
// Output planes are FULL-IMAGE sized (width*height); only rows [y0,y1) are written.  Any may be NULL.
// threads <= 0: all cores.
int bvro_render(const BvrModel* models, size_t n_models,
                const BvrMaterial* materials, size_t n_materials,
                const BvrBvhNode* nodes, size_t n_nodes,
                const BvrCamera* camera, const BvrRaytraceLevel* level, const BvrWindow* window,
                uint32_t width, uint32_t y0, uint32_t y1,
                const float* raster_rgba, const float* raster_depth,
                float* rgba, float* rt_depth, uint32_t* primary_id, float* primary_depth,
                int brute_force, int threads, BvroCounters* counters_out) {
    if (!camera || !level || !window) return 1;
    if (camera->projection != 0u) return 3;
    Scene s;
    s.models = models; s.n_models = n_models;
    s.materials = materials; s.n_materials = n_materials;
    s.nodes = nodes; s.n_nodes = n_nodes;
    s.camera = *camera; s.level = level->level; s.window = *window;
    s.tan_half_fov = bvro_tan_half_fov(camera->fov);
    s.brute_force = brute_force != 0;
    { const char* e = std::getenv("BVRO_DIAGNOSE"); s.diagnose = e && *e == '1'; }
    const uint32_t height = window->height;
    if (y1 > height) y1 = height;
    if ((s.level <= 2u) && (!raster_rgba || (s.level != 0u && !raster_depth))) return 1;
#ifdef _OPENMP
    int nthreads = threads > 0 ? threads : omp_get_max_threads();
#else
    int nthreads = 1;
#endif
    std::vector<Counters> per_thread((size_t)nthreads);
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int64_t yy = (int64_t)y0; yy < (int64_t)y1; yy++) {
#ifdef _OPENMP
        Counters& c = per_thread[(size_t)omp_get_thread_num()];
#else
        Counters& c = per_thread[0];
#endif
        const uint32_t y = (uint32_t)yy;
        for (uint32_t x = 0; x < width; x++) {
            const size_t pix = (size_t)y * width + x;
            const float u = ((float)x + 0.5f) / (float)width;
            const float v = ((float)y + 0.5f) / (float)height;
            // raytrace.wgsl:95
            uint32_t state = f32_to_u32(((window->random_seed * 10000.0f) * (u * 402.0f)) * (v * 31.5f));
            float out[4];
            uint32_t pid = 0xffffffffu; float pt = INF; float depth_avg = 0.0f;
            if (s.level == 0u) {                                  // raytrace.wgsl:97-99
                std::memcpy(out, raster_rgba + 4 * pix, sizeof out);
            } else {
                // trace_multisampled, raytrace.wgsl:159-172
                V3 total = v3(0, 0, 0); float total_depth = 0.0f;
                for (uint32_t sidx = 0; sidx < camera->sample_count; sidx++) {
                    Ray ray = random_ray_from_uv(s, u, v, state, c);
                    c.paths++;
                    RaytraceResult r = raytrace(s, ray, state, c, sidx == 0 ? &pid : nullptr,
                                                sidx == 0 ? &pt : nullptr);
                    total = total + r.color;
                    total_depth += r.depth;
                }
                V3 color = total / (float)camera->sample_count;
                depth_avg = total_depth / (float)camera->sample_count;
                out[0] = color.x; out[1] = color.y; out[2] = color.z; out[3] = 1.0f;
                if (s.level == 1u || s.level == 2u) {             // raytrace.wgsl:104-120
                    float depth = raster_depth[pix];
                    float rd = depth_avg;
                    if (rd > camera->far_plane) rd = -1.0f; else rd = camera->near_plane / rd;
                    if (depth > rd) std::memcpy(out, raster_rgba + 4 * pix, sizeof out);
                }
            }
            if (rgba) std::memcpy(rgba + 4 * pix, out, sizeof out);
            if (rt_depth) rt_depth[pix] = depth_avg;
            if (primary_id) primary_id[pix] = pid;
            if (primary_depth) primary_depth[pix] = pt;
        }
    }
    if (counters_out) {
        Counters t;
        for (auto& c : per_thread) t.merge(c);
        *counters_out = BvroCounters{t.rays, t.paths, t.node_pops, t.inner_visits, t.box_tests,
                                     t.sphere_tests, t.hits_shaded, t.rng_draws, t.stack_truncations,
                                     t.max_stack};
    }
    return 0;
}

// Rgba8UnormSrgb store conversion of the colour attachment (pipeline.rs:311-315): clamp, linear->sRGB
// OETF on rgb, alpha linear, round to nearest.  Evaluated in double so it is libm-independent to 1 LSB.
void bvro_store_srgb8(const float* rgba, size_t n_pixels, uint8_t* out) {
    for (size_t i = 0; i < n_pixels; i++) {
        for (int ch = 0; ch < 4; ch++) {
            double x = (double)rgba[4 * i + ch];
            if (!(x > 0.0)) x = 0.0;
            if (x > 1.0) x = 1.0;
            if (ch < 3) x = (x <= 0.0031308) ? 12.92 * x : 1.055 * std::pow(x, 1.0 / 2.4) - 0.055;
            out[4 * i + ch] = (uint8_t)std::floor(x * 255.0 + 0.5);
        }
    }
}

int bvro_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
