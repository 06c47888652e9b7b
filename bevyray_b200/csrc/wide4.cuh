// wide4.cuh — the 4-wide fp32 record visit shared by the persistent megakernel (megakernel_v3.cu) and the wavefront
// extend stage (wavefront.cu), plus the small PTX helpers both use.  Everything here is culling-only arithmetic: it
// decides which boxes are entered and in which order, never a value that reaches the image (trace.cuh does that).
#pragma once

#include "kernels.cuh"

namespace bvr {

#ifndef BVR_SORT_MIDDLE
#define BVR_SORT_MIDDLE 1      // 0: the two middle children of a 4-wide visit are parked unordered (2 instructions less)
#endif

// Explicit 32-bit shared-window addressing: keeps the per-step address math at one IMAD instead of the
// generic-to-shared conversion the compiler re-derives every time (ncu r01_v3a: 18 instructions per push).
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
// 256-bit read-only global load (sm_100: LDG.E.256): one request per 32-byte sector instead of two 16-byte ones
__device__ __forceinline__ void ldg256(const float4* p, float4& a, float4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}
__device__ __forceinline__ void ldg256u(const uint4* p, uint4& a, uint4& b) {
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p));
}
// A whole 64-byte record as two 256-bit loads in ONE asm statement, so that they are issued back to back: left to itself
// ptxas sinks the second load below the box tests of the first half under the 64-register cap (profiles/r02_m3_c4_full_ncu:
// two dependent L2 round trips per visit instead of one).  POLICY: 0 = default, 1 = L1::evict_last, 2 = L1::no_allocate.
template <int POLICY>
__device__ __forceinline__ void ldg512u(const uint4* p, uint4& a, uint4& b, uint4& c, uint4& d) {
#define BVR_LDG512(HINT)                                                                                           \
    asm volatile("ld.global.nc" HINT ".v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%16];\n\t"                               \
                 "ld.global.nc" HINT ".v8.u32 {%8,%9,%10,%11,%12,%13,%14,%15}, [%16+32];"                          \
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w),         \
                   "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w), "=r"(d.x), "=r"(d.y), "=r"(d.z), "=r"(d.w)          \
                 : "l"(p))
    if (POLICY == 1) BVR_LDG512(".L1::evict_last");
    else if (POLICY == 2) BVR_LDG512(".L1::no_allocate");
    else BVR_LDG512("");
#undef BVR_LDG512
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t a) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}

// Packed fp32 pairs (sm_100: FFMA2 = two IEEE fp32 FMAs per issued instruction).  ptxas folds the negation and the
// absolute value of a pair operand into the instruction, and reads a (z, z) pair as one broadcast register.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// hides how a value was computed, so that the compiler keeps it in a register instead of re-deriving it in the
// traversal loop (the shared-window base costs an S2R + LEA per use otherwise: profiles/r01 SASS)
__device__ __forceinline__ uint32_t opaque(uint32_t v) { asm volatile("" : "+r"(v)); return v; }

// Four children as packed keys (entry distance in the high bits, ref in the low bits; 0xffffffff = not entered):
// sorts them with a 5-comparator network of integer min/max, parks the three farther ones (farthest first) and
// returns the nearest.
template <class Push>
__device__ __forceinline__ uint32_t sort4_park(uint32_t k0, uint32_t k1, uint32_t k2, uint32_t k3, Push push) {
    uint32_t t0;
    t0 = min(k0, k1); k1 = max(k0, k1); k0 = t0;
    t0 = min(k2, k3); k3 = max(k2, k3); k2 = t0;
    t0 = min(k0, k2); k2 = max(k0, k2); k0 = t0;
    t0 = min(k1, k3); k3 = max(k1, k3); k1 = t0;
#if BVR_SORT_MIDDLE
    t0 = min(k1, k2); k2 = max(k1, k2); k1 = t0;
#endif
    if (k3 != 0xffffffffu) push(k3);
    if (k2 != 0xffffffffu) push(k2);
    if (k1 != 0xffffffffu) push(k1);
    return k0;
}
// One visit of a 4-wide fp32 record (scene_kernels.cu: build_nodes4_ch_kernel):
//   q0..q3 = (c.x, c.y, h.x, h.y) of child 0..3, q4 / q5 = (c.z, c.z', h.z, h.z') of children 0,1 / 2,3, rr = the four refs.
// Per axis pair: tc = c/d - o/d, lo = tc - h |1/d|, hi = tc + h |1/d| — 18 FFMA2 for the four boxes (ptxas folds the
// negation, the absolute value and the (z, z) broadcast into the instruction).  Keys = 21 bits of entry distance | 11
// bits of ref; the three farther children are parked, the nearest ref is returned (S4_NONE when nothing was entered).
template <class Push>
__device__ __forceinline__ uint32_t visit4(const float4 q0, const float4 q1, const float4 q2, const float4 q3, const float4 q4,
                                           const float4 q5, const float4 rr, const u64 inv_xy, const u64 noi_xy, const float inv_z,
                                           const float noi_z, const float closest_t, Push push) {
    float ix, iy;
    upk2(inv_xy, ix, iy);
    const u64 a_xy = pk2(fabsf(ix), fabsf(iy));
    const u64 i_zz = pk2(inv_z, inv_z), n_zz = pk2(noi_z, noi_z), a_zz = pk2(fabsf(inv_z), fabsf(inv_z));
    const u64 t0 = ffma2(pk2(q0.x, q0.y), inv_xy, noi_xy), t1 = ffma2(pk2(q1.x, q1.y), inv_xy, noi_xy);
    const u64 t2 = ffma2(pk2(q2.x, q2.y), inv_xy, noi_xy), t3 = ffma2(pk2(q3.x, q3.y), inv_xy, noi_xy);
    const u64 tz01 = ffma2(pk2(q4.x, q4.y), i_zz, n_zz), tz23 = ffma2(pk2(q5.x, q5.y), i_zz, n_zz);
    float lx0, ly0, lx1, ly1, lx2, ly2, lx3, ly3, lz0, lz1, lz2, lz3;
    float hx0, hy0, hx1, hy1, hx2, hy2, hx3, hy3, hz0, hz1, hz2, hz3;
    upk2(ffma2(pk2(-q0.z, -q0.w), a_xy, t0), lx0, ly0); upk2(ffma2(pk2(q0.z, q0.w), a_xy, t0), hx0, hy0);
    upk2(ffma2(pk2(-q1.z, -q1.w), a_xy, t1), lx1, ly1); upk2(ffma2(pk2(q1.z, q1.w), a_xy, t1), hx1, hy1);
    upk2(ffma2(pk2(-q2.z, -q2.w), a_xy, t2), lx2, ly2); upk2(ffma2(pk2(q2.z, q2.w), a_xy, t2), hx2, hy2);
    upk2(ffma2(pk2(-q3.z, -q3.w), a_xy, t3), lx3, ly3); upk2(ffma2(pk2(q3.z, q3.w), a_xy, t3), hx3, hy3);
    upk2(ffma2(pk2(-q4.z, -q4.w), a_zz, tz01), lz0, lz1); upk2(ffma2(pk2(q4.z, q4.w), a_zz, tz01), hz0, hz1);
    upk2(ffma2(pk2(-q5.z, -q5.w), a_zz, tz23), lz2, lz3); upk2(ffma2(pk2(q5.z, q5.w), a_zz, tz23), hz2, hz3);
    const float e0 = fmaxf(fmaxf(lx0, ly0), fmaxf(lz0, 0.0f)), x0 = fminf(fminf(hx0, hy0), fminf(hz0, closest_t));
    const float e1 = fmaxf(fmaxf(lx1, ly1), fmaxf(lz1, 0.0f)), x1 = fminf(fminf(hx1, hy1), fminf(hz1, closest_t));
    const float e2 = fmaxf(fmaxf(lx2, ly2), fmaxf(lz2, 0.0f)), x2 = fminf(fminf(hx2, hy2), fminf(hz2, closest_t));
    const float e3 = fmaxf(fmaxf(lx3, ly3), fmaxf(lz3, 0.0f)), x3 = fminf(fminf(hx3, hy3), fminf(hz3, closest_t));
    const uint32_t k0 = e0 <= x0 ? ((__float_as_uint(e0) & ~0x7ffu) | __float_as_uint(rr.x)) : 0xffffffffu;
    const uint32_t k1 = e1 <= x1 ? ((__float_as_uint(e1) & ~0x7ffu) | __float_as_uint(rr.y)) : 0xffffffffu;
    const uint32_t k2 = e2 <= x2 ? ((__float_as_uint(e2) & ~0x7ffu) | __float_as_uint(rr.z)) : 0xffffffffu;
    const uint32_t k3 = e3 <= x3 ? ((__float_as_uint(e3) & ~0x7ffu) | __float_as_uint(rr.w)) : 0xffffffffu;
    return sort4_park(k0, k1, k2, k3, push) & 0x7ffu;   // 0xffffffff (nothing entered) decodes to NONE
}

#define S4_LEAF 0x400u
#define S4_NONE 0x7ffu      // == (0xffffffff & S4_REF_MASK): a key that was not entered decodes to NONE by itself
#define S4_REF_MASK 0x7ffu

}  // namespace bvr
