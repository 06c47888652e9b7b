// trace.cuh — device-side path-tracing core shared by the megakernel and the wavefront kernels:
// camera-ray generation, BVH traversal with ray-sphere intersection, material scatter, background.
//
// Restates assets/shaders/raytrace.wgsl of the reference with the arithmetic conventions listed in
// device_math.cuh, so that results are bit-identical to the CPU oracle.  What is free to differ — and
// does — is everything that cannot change the closest hit: the node layout (child-pair SoA records
// instead of 48-byte AoS nodes), the traversal order (near child first) and where data is staged.
#pragma once

#include "device_math.cuh"

namespace bvr {

// ---- device scene layout -------------------------------------------------------------------------
// One 64-byte record per INNER node, holding both children (the reference re-reads the parent and
// both 48-byte children on every visit, raytrace.wgsl:323-335):
//   q0 = (c0.min.x, c0.min.y, c0.min.z, c0.max.x)
//   q1 = (c0.max.y, c0.max.z, c1.min.x, c1.min.y)
//   q2 = (c1.min.z, c1.max.x, c1.max.y, c1.max.z)
//   q3 = (bits ref0, bits ref1, 0, 0)
// A child ref is either the dense index of an inner record, or a leaf:
//   bit 31 set | (model_count-1) << 24 | first_model
#define BVR_LEAF_BIT 0x80000000u
#define BVR_LEAF_FIRST_MASK 0x00ffffffu
#define BVR_MAX_LEAF_COUNT 128u
#define BVR_REF_STACK 32   // raytrace.wgsl:310 STACKSIZE
#define BVR_FAST_STACK 128  // near-first stack of the fallback kernel: one entry per tree level (an LBVH over 63-bit
                            // keys plus index tie-break has at most 64 + log2(n) levels)

struct SceneView {
    const float4* __restrict__ pairs;            // 4 x float4 per inner node: child boxes as (min, max)
    const float4* __restrict__ pairs_ch;         // same records with the boxes as (centre, half extent):
                                                 //   q0 = (c0.xyz, h0.x) q1 = (h0.yz, c1.xy) q2 = (c1.z, h1.xyz) q3 = refs
    const float4* __restrict__ nodes4_ch;        // 7 x float4 per inner node: 4-wide fp32 records (small scenes), or null
    const float4* __restrict__ nodes4_tight;     // the same records with tight boxes (scene_kernels.cu), or null
    const float4* __restrict__ tight_groups;     // [0].x = group count (bits), then (C.xyz, D2) per radius group
    const uint4* __restrict__ pairs_q;           // 32-byte records on a 16-bit grid (scene_kernels.cu), or null
    const uint4* __restrict__ nodes4_q;          // 64-byte 4-wide records on the same grid, or null
    const float* __restrict__ qgrid;             // (base.xyz, -, step.xyz, -) of that grid
    const float4* __restrict__ spheres;          // (centre.xyz, radius) per model
    const uint32_t* __restrict__ sphere_material;  // Model::material_id per model
    const float4* __restrict__ materials;        // 2 x float4 per material (reference bytes)
    const uint32_t* __restrict__ model_rank;     // position of every model in the reference's traversal order, or null
    uint32_t root_ref;
    uint32_t n_materials;
    uint32_t has_scene;                          // 0: no nodes -> every ray misses
};

struct CameraParams {
    V3 position, direction, up, right;   // right = cross(direction, up), raytrace.wgsl:149
    float aspect, tan_half_fov;          // tan(fov*0.5), raytrace.wgsl:151 (host, double -> f32)
    float inv_width, inv_height;         // 1/(f32(height)*aspect), 1/f32(height), raytrace.wgsl:141-144
    float near_plane, far_plane, fallback_far;   // raytrace.wgsl:177-182
    float seed_scaled;                   // random_seed * 10000.0, raytrace.wgsl:95
    uint32_t sample_count, bounce_count, level;
    uint32_t width, height;
};

struct Ray {
    V3 o, d;
};

struct Hit {
    float t;          // INF = miss
    uint32_t model;
};

// raytrace.wgsl:95 — per-pixel RNG seed; uv of the pixel centre, (0,0) top-left
__device__ __forceinline__ uint32_t pixel_seed(const CameraParams& c, float u, float v) {
    return __float2uint_rz(fmul(fmul(c.seed_scaled, fmul(u, 402.0f)), fmul(v, 31.5f)));
}
__device__ __forceinline__ float pixel_u(const CameraParams& c, uint32_t x) {
    return fdiv(fadd((float)x, 0.5f), (float)c.width);
}
__device__ __forceinline__ float pixel_v(const CameraParams& c, uint32_t y) {
    return fdiv(fadd((float)y, 0.5f), (float)c.height);
}

// raytrace.wgsl:139-156
__device__ __forceinline__ Ray random_ray_from_uv(const CameraParams& c, float u, float v, uint32_t& rng) {
    const float rx = fsub(rng_next_float(rng), 0.5f);
    const float ry = fsub(rng_next_float(rng), 0.5f);
    const float delta_u = fmul(c.inv_width, rx);
    const float delta_v = fmul(c.inv_height, ry);
    const float ndc_x = fadd(fsub(fmul(u, 2.0f), 1.0f), delta_u);
    const float ndc_y = fadd(fsub(1.0f, fmul(v, 2.0f)), delta_v);
    const V3 a = vscale(fmul(fmul(ndc_x, c.aspect), c.tan_half_fov), c.right);
    const V3 b = vscale(fmul(ndc_y, c.tan_half_fov), c.up);
    Ray r;
    r.o = c.position;
    r.d = vnormalize(vadd(vadd(c.direction, a), b));
    return r;
}

// raytrace.wgsl:387-398 with 1/direction hoisted out of the node loop (same bits as recomputing it)
__device__ __forceinline__ float ray_bounding_dst(V3 o, V3 inv, float mnx, float mny, float mnz,
                                                  float mxx, float mxy, float mxz) {
    const float tminx = fmul(fsub(mnx, o.x), inv.x), tmaxx = fmul(fsub(mxx, o.x), inv.x);
    const float tminy = fmul(fsub(mny, o.y), inv.y), tmaxy = fmul(fsub(mxy, o.y), inv.y);
    const float tminz = fmul(fsub(mnz, o.z), inv.z), tmaxz = fmul(fsub(mxz, o.z), inv.z);
    const float t_near = fmaxf(fmaxf(fminf(tminx, tmaxx), fminf(tminy, tmaxy)), fminf(tminz, tmaxz));
    const float t_far = fminf(fminf(fmaxf(tminx, tmaxx), fmaxf(tminy, tmaxy)), fmaxf(tminz, tmaxz));
    const bool hit = (t_far >= t_near) && (t_far > 0.0f);
    return hit ? (t_near > 0.0f ? t_near : 0.0f) : BVR_INF;
}

// raytrace.wgsl:371-383 with a = dot(d,d) hoisted; raytrace.wgsl:353-354 acceptance test.  `sp` = (centre, radius) of
// model i, loaded by the caller (shared memory or global).
__device__ __forceinline__ void test_sphere(const SceneView& s, const Ray& ray, float a, uint32_t i, const float4 sp, Hit& closest) {
    const V3 oc = v3(fsub(sp.x, ray.o.x), fsub(sp.y, ray.o.y), fsub(sp.z, ray.o.z));
    const float h = vdot(ray.d, oc);
    const float c = fsub(vdot(oc, oc), fmul(sp.w, sp.w));
    const float disc = fsub(fmul(h, h), fmul(a, c));
    if (disc < 0.0f) return;                         // hit_sphere returns -1.0
    const float t = fdiv(fsub(h, fsqrt(disc)), a);
    if (t != -1.0f && t > 0.001f) {
        if (t < closest.t) { closest.t = t; closest.model = i; }
        // Bit-exact tie between two spheres (a small sphere resting on the ground sphere, duplicates): the
        // reference keeps the one it reaches FIRST (strict <, raytrace.wgsl:354), and its order over the
        // leaves is fixed by the tree — right subtree first.  Any other visiting order reproduces that choice
        // by preferring the lower rank.
        else if (t == closest.t && i != closest.model && s.model_rank &&
                 s.model_rank[i] < s.model_rank[closest.model]) closest.model = i;
    }
}

__device__ __forceinline__ void test_leaf(const SceneView& s, const Ray& ray, float a, uint32_t ref, Hit& closest) {
    const uint32_t first = ref & BVR_LEAF_FIRST_MASK;
    const uint32_t count = ((ref >> 24) & 0x7fu) + 1u;
    for (uint32_t i = first; i < first + count; i++) test_sphere(s, ray, a, i, s.spheres[i], closest);
}

// raytrace.wgsl:313-346 verbatim control flow: LIFO stack, first child pushed first, no ordering,
// traversal abandoned when the stack index reaches 32.
__device__ __forceinline__ Hit raycast_reference_order(const SceneView& s, const Ray& ray) {
    Hit closest{BVR_INF, 0xffffffffu};
    if (!s.has_scene) return closest;
    const V3 inv = v3(fdiv(1.0f, ray.d.x), fdiv(1.0f, ray.d.y), fdiv(1.0f, ray.d.z));
    const float a = vdot(ray.d, ray.d);
    uint32_t stack[BVR_REF_STACK];
    stack[0] = s.root_ref;
    int sp = 1;
    while (sp > 0 && sp < BVR_REF_STACK) {
        const uint32_t ref = stack[--sp];
        if (ref & BVR_LEAF_BIT) {
            test_leaf(s, ray, a, ref, closest);
        } else {
            const float4* p = s.pairs + 4u * ref;
            const float4 q0 = p[0], q1 = p[1], q2 = p[2], q3 = p[3];
            const float d0 = ray_bounding_dst(ray.o, inv, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y);
            if (d0 != BVR_INF && d0 < closest.t) stack[sp++] = __float_as_uint(q3.x);
            const float d1 = ray_bounding_dst(ray.o, inv, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w);
            if (d1 != BVR_INF && d1 < closest.t) stack[sp++] = __float_as_uint(q3.y);
        }
    }
    return closest;
}

// Same closest hit, fewer node visits: descend into the nearer child first and drop stacked
// subtrees that can no longer beat the current hit.  (Result equals the reference's as long as the
// reference's own 32-entry truncation does not trigger and no two spheres tie bit-exactly in t.)
template <class StackRef, class StackDst>
__device__ __forceinline__ Hit raycast_near_first(const SceneView& s, const Ray& ray, StackRef stack_ref,
                                                  StackDst stack_dst) {
    Hit closest{BVR_INF, 0xffffffffu};
    if (!s.has_scene) return closest;
    const V3 inv = v3(fdiv(1.0f, ray.d.x), fdiv(1.0f, ray.d.y), fdiv(1.0f, ray.d.z));
    const float a = vdot(ray.d, ray.d);
    uint32_t cur = s.root_ref;
    int sp = 0;
    for (;;) {
        if (cur & BVR_LEAF_BIT) {
            test_leaf(s, ray, a, cur, closest);
        } else {
            const float4* p = s.pairs + 4u * cur;
            const float4 q0 = p[0], q1 = p[1], q2 = p[2], q3 = p[3];
            const float d0 = ray_bounding_dst(ray.o, inv, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y);
            const float d1 = ray_bounding_dst(ray.o, inv, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w);
            const bool h0 = d0 != BVR_INF && d0 < closest.t;
            const bool h1 = d1 != BVR_INF && d1 < closest.t;
            const uint32_t r0 = __float_as_uint(q3.x), r1 = __float_as_uint(q3.y);
            if (h0 && h1) {
                const bool first0 = d0 < d1;   // ties go to the second child, like the reference's LIFO order
                stack_ref(sp) = first0 ? r1 : r0;
                stack_dst(sp) = first0 ? d1 : d0;
                sp++;
                cur = first0 ? r0 : r1;
                continue;
            }
            if (h0) { cur = r0; continue; }
            if (h1) { cur = r1; continue; }
        }
        // pop
        bool found = false;
        while (sp > 0) {
            --sp;
            if (stack_dst(sp) < closest.t) { cur = stack_ref(sp); found = true; break; }
        }
        if (!found) break;
    }
    return closest;
}

struct ShadeResult {
    bool terminated;   // path ended (absorbed)
    V3 attenuation;
};

// raytrace.wgsl:355-359 (hit record) + raytrace.wgsl:231-299 (scatter).  Updates `ray` in place and
// returns whether the ray was absorbed.
__device__ __forceinline__ bool scatter(const SceneView& s, Ray& ray, const Hit& hit, uint32_t& rng, V3& attenuation) {
    const float4 sp = s.spheres[hit.model];
    const V3 position = vadd(ray.o, vscale(hit.t, ray.d));                 // ray_at, raytrace.wgsl:130-132
    const V3 normal = vnormalize(vsub(position, v3(sp.x, sp.y, sp.z)));
    const bool front_face = vdot(ray.d, normal) < 0.0f;
    uint32_t mid = s.sphere_material[hit.model];
    if (mid >= s.n_materials) mid = s.n_materials - 1u;                    // robust buffer access clamps
    const float4 m0 = s.materials[2u * mid], m1 = s.materials[2u * mid + 1u];
    const V3 base_color = v3(m0.x, m0.y, m0.z);
    const float metallic = m0.w, roughness = m1.x, ior = m1.z, transmission = m1.w;

    if (rng_next_float(rng) < metallic) {
        const V3 reflected = vadd(vnormalize(reflect3(ray.d, normal)), vscale(roughness, random_unit_vec3(rng)));
        ray.o = position;
        ray.d = reflected;
        attenuation = base_color;
        return vdot(ray.d, normal) < 0.0f;
    }
    if (rng_next_float(rng) < transmission) {
        const float ri = front_face ? fdiv(1.0f, ior) : ior;
        const V3 unit_direction = vnormalize(ray.d);
        const float cos_theta = fminf(vdot(vneg(unit_direction), normal), 1.0f);
        const float sin_theta = fsqrt(fsub(1.0f, fmul(cos_theta, cos_theta)));
        const bool cannot_refract = fmul(ri, sin_theta) > 1.0f;
        V3 direction;
        // `||` short-circuits: no RNG draw when cannot_refract (raytrace.wgsl:269)
        if (cannot_refract || schlick_reflectance(cos_theta, ri) > rng_next_float(rng)) {
            direction = reflect3(unit_direction, normal);
        } else {
            direction = refract3(unit_direction, normal, ri);
        }
        ray.o = position;
        ray.d = direction;
        attenuation = v3(1.0f, 1.0f, 1.0f);
        return false;
    }
    const V3 b1 = random_unit_vec3(rng);
    const V3 b2 = random_unit_vec3(rng);
    V3 dir = vadd(vadd(normal, b1), vscale(roughness, b2));
    if (vec3_near_zero(dir)) dir = normal;
    ray.o = position;
    ray.d = dir;
    attenuation = base_color;
    return vdot(ray.d, normal) < 0.0f;
}

// raytrace.wgsl:364-369
__device__ __forceinline__ V3 background_gradient(const Ray& ray) {
    const V3 unit = vnormalize(ray.d);
    const float a = fmul(0.5f, fadd(unit.y, 1.0f));
    const float ia = fsub(1.0f, a);
    return v3(fadd(fmul(ia, 1.0f), fmul(a, 0.5f)), fadd(fmul(ia, 1.0f), fmul(a, 0.7f)),
              fadd(fmul(ia, 1.0f), fmul(a, 1.0f)));
}

// fragment's composite, raytrace.wgsl:104-120: true -> the raster texel wins
__device__ __forceinline__ bool raster_wins(const CameraParams& c, float raster_depth, float rt_depth_avg) {
    float rd = rt_depth_avg;
    if (rd > c.far_plane) rd = -1.0f; else rd = fdiv(c.near_plane, rd);
    return raster_depth > rd;
}

// Rgba8UnormSrgb store conversion of the colour attachment (pipeline.rs:311-315)
__device__ __forceinline__ uchar4 store_srgb8(float4 c) {
    auto enc = [](float x, bool oetf) -> unsigned char {
        x = !(x > 0.0f) ? 0.0f : (x > 1.0f ? 1.0f : x);
        if (oetf) x = (x <= 0.0031308f) ? 12.92f * x : 1.055f * powf(x, 1.0f / 2.4f) - 0.055f;
        return (unsigned char)floorf(x * 255.0f + 0.5f);
    };
    return make_uchar4(enc(c.x, true), enc(c.y, true), enc(c.z, true), enc(c.w, false));
}

}  // namespace bvr
