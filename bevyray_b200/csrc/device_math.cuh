// device_math.cuh — strict-IEEE f32 helpers for the path-tracing kernels.
//
// Every arithmetic helper uses the round-to-nearest intrinsics (__fmul_rn, __fadd_rn, ...), which the
// compiler never contracts into FMAs, so the operation order written here is the operation order
// executed, independent of compiler flags.  The order mirrors the WGSL expressions they restate
// (citations: assets/shaders/raytrace.wgsl, assets/shaders/random.wgsl of the reference).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace bvr {

// assets/shaders/const.wgsl:2
#define BVR_INF 3.40282347e+38f

struct V3 {
    float x, y, z;
};

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }

__device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 vadd(V3 a, V3 b) { return V3{fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z)}; }
__device__ __forceinline__ V3 vsub(V3 a, V3 b) { return V3{fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z)}; }
__device__ __forceinline__ V3 vmul(V3 a, V3 b) { return V3{fmul(a.x, b.x), fmul(a.y, b.y), fmul(a.z, b.z)}; }
__device__ __forceinline__ V3 vscale(float s, V3 a) { return V3{fmul(s, a.x), fmul(s, a.y), fmul(s, a.z)}; }
__device__ __forceinline__ V3 vdivs(V3 a, float s) { return V3{fdiv(a.x, s), fdiv(a.y, s), fdiv(a.z, s)}; }
__device__ __forceinline__ V3 vneg(V3 a) { return V3{-a.x, -a.y, -a.z}; }
// dot(a,b) = (a.x*b.x + a.y*b.y) + a.z*b.z
__device__ __forceinline__ float vdot(V3 a, V3 b) {
    return fadd(fadd(fmul(a.x, b.x), fmul(a.y, b.y)), fmul(a.z, b.z));
}
__device__ __forceinline__ V3 vcross(V3 a, V3 b) {
    return V3{fsub(fmul(a.y, b.z), fmul(a.z, b.y)), fsub(fmul(a.z, b.x), fmul(a.x, b.z)),
              fsub(fmul(a.x, b.y), fmul(a.y, b.x))};
}
// normalize(v) = v / sqrt(dot(v,v))
__device__ __forceinline__ V3 vnormalize(V3 a) { return vdivs(a, fsqrt(vdot(a, a))); }

// random.wgsl:8-15 (the state update is additive, not the multiplicative PCG step)
__device__ __forceinline__ void rng_next_int(uint32_t& state) {
    const uint32_t old_state = state + 747796405u + 2891336453u;
    const uint32_t word = ((old_state >> ((old_state >> 28u) + 4u)) ^ old_state) * 277803737u;
    state = (word >> 22u) ^ word;
}

// random.wgsl:3-6: f32(state) / f32(0xffffffffu) == f32(state) * 2^-32 exactly
__device__ __forceinline__ float rng_next_float(uint32_t& state) {
    rng_next_int(state);
    return fmul(__uint2float_rn(state), 2.3283064365386963e-10f);
}

// random.wgsl:17-30: rejection-sampled point in the unit ball, NOT normalised
__device__ __forceinline__ V3 random_unit_vec3(uint32_t& state) {
    V3 p;
    for (;;) {
        const float x = rng_next_float(state);
        const float y = rng_next_float(state);
        const float z = rng_next_float(state);
        p = v3(fsub(fmul(2.0f, x), 1.0f), fsub(fmul(2.0f, y), 1.0f), fsub(fmul(2.0f, z), 1.0f));
        if (vdot(p, p) <= 1.0f) break;
    }
    return p;
}

// raytrace.wgsl:400-402: v - 2*dot(v,n)*n
__device__ __forceinline__ V3 reflect3(V3 v, V3 n) { return vsub(v, vscale(fmul(2.0f, vdot(v, n)), n)); }

// raytrace.wgsl:404-409
__device__ __forceinline__ V3 refract3(V3 v, V3 n, float etai_over_etat) {
    const float cos_theta = fminf(vdot(vneg(v), n), 1.0f);
    const V3 r_out_perp = vscale(etai_over_etat, vadd(v, vscale(cos_theta, n)));
    const float k = -fsqrt(fabsf(fsub(1.0f, vdot(r_out_perp, r_out_perp))));
    return vadd(r_out_perp, vscale(k, n));
}

// raytrace.wgsl:411-416 with pow(x,5) = ((x*x)*(x*x))*x
__device__ __forceinline__ float schlick_reflectance(float cosine, float refraction_index) {
    float r0 = fdiv(fsub(1.0f, refraction_index), fadd(1.0f, refraction_index));
    r0 = fmul(r0, r0);
    const float x = fsub(1.0f, cosine);
    const float x2 = fmul(x, x);
    const float x5 = fmul(fmul(x2, x2), x);
    return fadd(r0, fmul(fsub(1.0f, r0), x5));
}

// raytrace.wgsl:418-421
__device__ __forceinline__ bool vec3_near_zero(V3 v) {
    const float s = 1e-8f;
    return fabsf(v.x) < s && fabsf(v.y) < s && fabsf(v.z) < s;
}

}  // namespace bvr
