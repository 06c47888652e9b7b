// megakernel_v4.cu — persistent-lane megakernel with TWO PATHS PER LANE.
//
// ncu on v3 (profiles/r01_v3d_*): 15.4 of 32 lanes active per instruction.  Most of the loss is structural: a
// lane whose ray has finished idles until ~26 lanes of its warp wait for shading, because the shading stages
// (~450 warp instructions per pass) are only worth entering when most lanes take part.  Here every lane owns two
// pixel slots whose path state lives in shared memory; when the ray of one slot finishes, the lane parks the hit
// in the slot and continues with the ray of its other slot (a ~30-instruction switch), and the parked slot is
// shaded later, when most lanes of the warp have one.  Traversal registers hold only the ray being traversed;
// throughput, RNG state, sample/bounce counters and the pixel accumulators stay in the slot.
//
// A pixel's samples still run one after the other in one slot (the reference's RNG stream is sequential per
// pixel, raytrace.wgsl:89,161-167), so the image is bit-identical to v3's and to the oracle's.
//
// Shared memory (one CTA per SM):  scene (child-pair records, spheres, materials, material ids)
//                                | slots: 5 x float4 per slot, 2 slots per lane, interleaved by lane
//                                | traversal stacks: 4-byte entries (21 bits of distance | 11 bits of child ref)
// The 4-byte stack entry needs every child ref to fit 11 bits: at most 1024 inner nodes, 1024 spheres and one
// sphere per leaf (what the reference's PLOC build and the GPU builder produce); other scenes use v3.

#include "kernels.cuh"

namespace bvr {

namespace {

enum SlotStatus : int { S_EMPTY = 0, S_READY = 1, S_TRAV = 2, S_PENDING = 3, S_DONE = 4 };
enum LaneState : int { L_IDLE = 0, L_TRAVERSE = 1, L_FINISHED = 2 };
enum PathState : int { P_NONE = 0, P_SHADE = 1, P_NEW_PATH = 2, P_RAY_READY = 3, P_EMPTY = 4, P_DONE = 5 };
enum Kind : int { K_NONE = 0, K_MISS = 1, K_METAL = 2, K_GLASS = 3, K_DIFFUSE = 4 };

#define V4_LEAF 0x400u
#define V4_NONE 0x800u
#define V4_REF_MASK 0x7ffu
#ifndef BVR_V4_STEPS_PER_VOTE
#define BVR_V4_STEPS_PER_VOTE 2
#endif

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
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t a) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Culling-only slab test on (centre, half extent) boxes, as in v3 (megakernel_v3.cu: box_cull).
__device__ __forceinline__ bool box_cull(V3 inv, V3 noi, float closest_t, float cx, float cy, float cz, float hx,
                                         float hy, float hz, float& entry) {
    const float tcx = __fmaf_rn(cx, inv.x, noi.x), thx = hx * fabsf(inv.x);
    const float tcy = __fmaf_rn(cy, inv.y, noi.y), thy = hy * fabsf(inv.y);
    const float tcz = __fmaf_rn(cz, inv.z, noi.z), thz = hz * fabsf(inv.z);
    entry = fmaxf(fmaxf(tcx - thx, tcy - thy), fmaxf(tcz - thz, 0.0f));
    const float exit = fminf(fminf(tcx + thx, tcy + thy), fminf(tcz + thz, closest_t));
    return entry <= exit;
}

__device__ __forceinline__ uint32_t compact_ref(uint32_t ref) {
    return (ref & BVR_LEAF_BIT) ? (V4_LEAF | (ref & 0x3ffu)) : ref;
}

struct Tuning4 {
    uint32_t shade_lanes;    // shade when this many lanes hold a slot waiting for it ...
    uint32_t stuck_lanes;    // ... or when this many lanes have nothing else to do
    uint32_t switch_lanes;   // leave the traversal loop when this many lanes can switch to their other slot
    uint32_t leaf_batch_lanes;
};

template <int THREADS>
__global__ void __launch_bounds__(THREADS) megakernel_v4(const RenderParams p, unsigned int* __restrict__ pixel_counter,
                                                         const uint32_t n_inner, const uint32_t n_models,
                                                         const Tuning4 tune) {
    extern __shared__ float4 smem[];
    const CameraParams& cam = p.cam;
    const unsigned full = 0xffffffffu;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;

    // ---- stage the scene ----
    SceneView sv = p.scene;
    float4* sm_cursor = smem;
    float4* sm_pairs = sm_cursor;     sm_cursor += 4u * n_inner;
    float4* sm_spheres = sm_cursor;   sm_cursor += n_models;
    float4* sm_materials = sm_cursor; sm_cursor += 2u * sv.n_materials;
    uint32_t* sm_matid = reinterpret_cast<uint32_t*>(sm_cursor);
    sm_cursor += (n_models + 3u) / 4u;
    for (uint32_t i = tid; i < 4u * n_inner; i += THREADS) {
        float4 q = p.scene.pairs_ch[i];
        if ((i & 3u) == 3u) {   // child refs -> 11-bit form
            q.x = __uint_as_float(compact_ref(__float_as_uint(q.x)));
            q.y = __uint_as_float(compact_ref(__float_as_uint(q.y)));
        }
        sm_pairs[i] = q;
    }
    for (uint32_t i = tid; i < n_models; i += THREADS) sm_spheres[i] = p.scene.spheres[i];
    for (uint32_t i = tid; i < 2u * sv.n_materials; i += THREADS) sm_materials[i] = p.scene.materials[i];
    for (uint32_t i = tid; i < n_models; i += THREADS) sm_matid[i] = p.scene.sphere_material[i];
    sv.spheres = sm_spheres;
    sv.materials = sm_materials;
    sv.sphere_material = sm_matid;
    __syncthreads();
    const uint32_t s_pairs = smem_addr(sm_pairs);
    const uint32_t root = sv.has_scene ? compact_ref(sv.root_ref) : V4_NONE;

    // ---- slots: group g of slot s of this lane at slot0 + (g*2+s) * THREADS*16 ----
    //   G0 = (origin.xyz, hit t)   G1 = (direction.xyz, hit model)   G2 = (throughput.rgb, rng state)
    //   G3 = (px | ly<<16, sample index, bounce, first_depth)   G4 = (sum of sample colours, sum of first depths)
    constexpr uint32_t SLOT_STRIDE = THREADS * 16u;          // slot 0 -> slot 1
    constexpr uint32_t GROUP_STRIDE = 2u * SLOT_STRIDE;      // group g -> g+1
    const uint32_t slot0 = smem_addr(sm_cursor) + tid * 16u;
    sm_cursor += 5u * 2u * THREADS;
    // ---- stacks: entry k of this lane at s_stack0 + k * STACK_STRIDE ----
    constexpr uint32_t STACK_STRIDE = THREADS * 4u;
    const uint32_t s_stack0 = smem_addr(sm_cursor) + tid * 4u;

    const uint32_t tiles_x = (cam.width + 7u) / 8u, tiles_y = (p.shard.rows + 3u) / 4u;
    const uint32_t total_slots = tiles_x * tiles_y * 32u;

    int st0 = S_EMPTY, st1 = S_EMPTY;   // slot status
    int state = L_IDLE;                 // the ray in registers
    int cs = 0;                         // slot of the ray in registers
    Ray ray{v3(0, 0, 0), v3(0, 0, 1)};
    V3 inv = v3(0, 0, 0), noi = v3(0, 0, 0);
    float a = 1.0f;
    Hit closest{BVR_INF, 0xffffffffu};
    uint32_t cur = V4_NONE, pending = V4_NONE;
    uint32_t sp_addr = s_stack0;
    uint32_t rays = 0;

    for (;;) {
        // the slot this lane would shade (or refill) next
        int as = -1;
        if (st0 == S_PENDING || st0 == S_EMPTY) as = 0;
        else if (st1 == S_PENDING || st1 == S_EMPTY) as = 1;
        {
            const bool can_switch = state != L_TRAVERSE && (st0 == S_READY || st1 == S_READY);
            const unsigned m_want = __ballot_sync(full, as >= 0);
            const unsigned m_trav = __ballot_sync(full, state == L_TRAVERSE);
            const unsigned m_sw = __ballot_sync(full, can_switch);
            if ((m_want | m_trav | m_sw) == 0u) break;                       // every slot of the warp is S_DONE
            const unsigned m_stuck = __ballot_sync(full, state != L_TRAVERSE && !can_switch && as >= 0);
            const bool do_shade = (uint32_t)__popc(m_want) >= tune.shade_lanes ||
                                  (uint32_t)__popc(m_stuck) >= tune.stuck_lanes || (m_trav | m_sw) == 0u;
            if (!do_shade) as = -1;
        }

        // ======================= phase A: staged shading of one parked slot per lane =======================
        if (__any_sync(full, as >= 0)) {
            const uint32_t sb = slot0 + (as == 1 ? SLOT_STRIDE : 0u);
            const int sst = as == 0 ? st0 : (as == 1 ? st1 : S_DONE);
            int ps = sst == S_PENDING ? P_SHADE : (sst == S_EMPTY ? P_EMPTY : P_NONE);
            V3 so = v3(0, 0, 0), sd = v3(0, 0, 1), thr = v3(1, 1, 1);
            float ht = BVR_INF, first_depth = BVR_INF;
            uint32_t hmodel = 0xffffffffu, rng = 0u, pix = 0u, sidx = 0u, bounce = 0u;
            if (ps == P_SHADE) {
                const float4 g0 = lds128(sb), g1 = lds128(sb + GROUP_STRIDE), g2 = lds128(sb + 2u * GROUP_STRIDE),
                             g3 = lds128(sb + 3u * GROUP_STRIDE);
                so = v3(g0.x, g0.y, g0.z); ht = g0.w;
                sd = v3(g1.x, g1.y, g1.z); hmodel = __float_as_uint(g1.w);
                thr = v3(g2.x, g2.y, g2.z); rng = __float_as_uint(g2.w);
                pix = __float_as_uint(g3.x); sidx = __float_as_uint(g3.y); bounce = __float_as_uint(g3.z);
                first_depth = g3.w;
            }
            // --- A1: classification (raytrace.wgsl:193-201, 232-248) ---
            int kind = K_NONE;
            uint32_t mid = 0;
            if (ps == P_SHADE) {
                if (bounce == 0u) {
                    first_depth = ht;
                    if (sidx == 0u && (p.out_primary_id || p.out_primary_depth)) {
                        const size_t lpix = (size_t)(pix >> 16) * cam.width + (pix & 0xffffu);
                        if (p.out_primary_id) p.out_primary_id[lpix] = ht == BVR_INF ? 0xffffffffu : hmodel;
                        if (p.out_primary_depth) p.out_primary_depth[lpix] = ht;
                    }
                }
                if (ht == BVR_INF) {
                    kind = K_MISS;
                } else {
                    mid = sv.sphere_material[hmodel];
                    if (mid >= sv.n_materials) mid = sv.n_materials - 1u;
                    const float metallic = sv.materials[2u * mid].w;
                    const float transmission = sv.materials[2u * mid + 1u].w;
                    if (rng_next_float(rng) < metallic) kind = K_METAL;
                    else if (rng_next_float(rng) < transmission) kind = K_GLASS;
                    else kind = K_DIFFUSE;
                }
            }
            // --- A2: every unit-ball sample of this round in one rejection loop (random.wgsl:17-26) ---
            int need = kind == K_DIFFUSE ? 2 : (kind == K_METAL ? 1 : 0);
            V3 b1 = v3(0.0f, 0.0f, 0.0f), b2 = v3(0.0f, 0.0f, 0.0f);
            while (need > 0) {
                // fma(f32(state), 2^-31, -1) rounds once, exactly like the reference's (2*x) - 1 (both scalings exact)
                rng_next_int(rng); const float x = __uint2float_rn(rng);
                rng_next_int(rng); const float y = __uint2float_rn(rng);
                rng_next_int(rng); const float z = __uint2float_rn(rng);
                const float k = 4.6566128730773926e-10f;   // 2^-31
                const V3 c = v3(__fmaf_rn(x, k, -1.0f), __fmaf_rn(y, k, -1.0f), __fmaf_rn(z, k, -1.0f));
                if (vdot(c, c) <= 1.0f) {
                    if (need == 2) b1 = c; else b2 = c;   // diffuse: b1 then b2; metal: b2 only
                    need--;
                }
            }
            // --- A3: hit record / background share one normalize site ---
            if (ps == P_SHADE) {
                bool path_end = false;
                V3 sample_color = v3(0.0f, 0.0f, 0.0f);
                const float4 sph = kind == K_MISS ? make_float4(0.f, 0.f, 0.f, 0.f) : sv.spheres[hmodel];
                const V3 position = vadd(so, vscale(ht, sd));                          // ray_at, raytrace.wgsl:130-132
                const V3 nin = kind == K_MISS ? sd : vsub(position, v3(sph.x, sph.y, sph.z));
                const V3 unit = vnormalize(nin);
                if (kind == K_MISS) {
                    // background_gradient, raytrace.wgsl:364-369
                    const float aa = fmul(0.5f, fadd(unit.y, 1.0f));
                    const float ia = fsub(1.0f, aa);
                    const V3 bg = v3(fadd(fmul(ia, 1.0f), fmul(aa, 0.5f)), fadd(fmul(ia, 1.0f), fmul(aa, 0.7f)),
                                     fadd(fmul(ia, 1.0f), fmul(aa, 1.0f)));
                    const V3 lin = vmul(thr, bg);
                    sample_color = v3(fsqrt(lin.x), fsqrt(lin.y), fsqrt(lin.z));       // raytrace.wgsl:223
                    path_end = true;
                } else {
                    const V3 normal = unit;                                            // raytrace.wgsl:357
                    const float4 m0 = sv.materials[2u * mid], m1 = sv.materials[2u * mid + 1u];
                    V3 attenuation = v3(m0.x, m0.y, m0.z);
                    V3 dir;
                    bool absorbed;
                    if (kind == K_DIFFUSE) {                                           // raytrace.wgsl:283-298
                        dir = vadd(vadd(normal, b1), vscale(m1.x, b2));
                        if (vec3_near_zero(dir)) dir = normal;
                        absorbed = vdot(dir, normal) < 0.0f;
                    } else {
                        // metal and glass share the second normalize site
                        const V3 un = vnormalize(kind == K_METAL ? reflect3(sd, normal) : sd);
                        if (kind == K_METAL) {                                         // raytrace.wgsl:234-246
                            dir = vadd(un, vscale(m1.x, b2));
                            absorbed = vdot(dir, normal) < 0.0f;
                        } else {                                                       // raytrace.wgsl:248-282
                            const bool front_face = vdot(sd, normal) < 0.0f;
                            const float ri = front_face ? fdiv(1.0f, m1.z) : m1.z;
                            const float cos_theta = fminf(vdot(vneg(un), normal), 1.0f);
                            const float sin_theta = fsqrt(fsub(1.0f, fmul(cos_theta, cos_theta)));
                            const bool cannot_refract = fmul(ri, sin_theta) > 1.0f;
                            if (cannot_refract || schlick_reflectance(cos_theta, ri) > rng_next_float(rng)) dir = reflect3(un, normal);
                            else dir = refract3(un, normal, ri);
                            attenuation = v3(1.0f, 1.0f, 1.0f);
                            absorbed = false;
                        }
                    }
                    so = position;
                    sd = dir;
                    if (absorbed) {
                        path_end = true;                                               // raytrace.wgsl:207-209
                    } else {
                        thr = vmul(thr, attenuation);
                        bounce++;
                        if (bounce > cam.bounce_count) path_end = true;                // raytrace.wgsl:214-216
                    }
                }
                if (path_end) {
                    if (first_depth == BVR_INF) first_depth = cam.fallback_far;
                    float4 acc = lds128(sb + 4u * GROUP_STRIDE);                       // raytrace.wgsl:165-166
                    acc.x = fadd(acc.x, sample_color.x); acc.y = fadd(acc.y, sample_color.y);
                    acc.z = fadd(acc.z, sample_color.z); acc.w = fadd(acc.w, first_depth);
                    sts128(sb + 4u * GROUP_STRIDE, acc);
                    sidx++;
                    ps = P_NEW_PATH;
                } else {
                    ps = P_RAY_READY;
                }
            }
            for (;;) {
                // --- A4: pixel store (average, fused composite raytrace.wgsl:104-120) ---
                if (ps == P_NEW_PATH && sidx >= cam.sample_count) {
                    const uint32_t px = pix & 0xffffu, ly = pix >> 16;
                    const uint32_t gy = shard_global_row(p.shard, ly);
                    const float4 acc = lds128(sb + 4u * GROUP_STRIDE);
                    const float n = (float)cam.sample_count;
                    float4 out = make_float4(fdiv(acc.x, n), fdiv(acc.y, n), fdiv(acc.z, n), 1.0f);
                    const float depth_avg = fdiv(acc.w, n);
                    if (cam.level == 1u || cam.level == 2u) {
                        const size_t gpix = (size_t)gy * cam.width + px;
                        if (raster_wins(cam, p.raster_depth[gpix], depth_avg)) out = p.raster_rgba[gpix];
                    }
                    const size_t lpix = (size_t)ly * cam.width + px;
                    if (p.out_rgba) p.out_rgba[lpix] = out;
                    if (p.out_rt_depth) p.out_rt_depth[lpix] = depth_avg;
                    if (p.out_srgb8) p.out_srgb8[lpix] = store_srgb8(out);
                    ps = P_EMPTY;
                }
                // --- A5: pull new pixels from the tile-ordered queue (warp-convergent) ---
                const unsigned need_px = __ballot_sync(full, ps == P_EMPTY);
                if (need_px != 0u) {
                    const int leader = __ffs(need_px) - 1;
                    unsigned base = 0;
                    if ((int)lane == leader) base = atomicAdd(pixel_counter, (unsigned)__popc(need_px));
                    base = __shfl_sync(full, base, leader);
                    if (ps == P_EMPTY) {
                        const uint32_t slot = base + (uint32_t)__popc(need_px & ((1u << lane) - 1u));
                        ps = P_DONE;
                        if (slot < total_slots) {
                            const uint32_t tile = slot >> 5, within = slot & 31u;
                            const uint32_t px = (tile % tiles_x) * 8u + (within & 7u);
                            const uint32_t ly = (tile / tiles_x) * 4u + (within >> 3);
                            const uint32_t gy = shard_global_row(p.shard, ly);
                            ps = P_EMPTY;   // a padding slot of the tile grid: ask again
                            if (px < cam.width && ly < p.shard.rows && gy < cam.height) {
                                pix = px | (ly << 16);
                                rng = pixel_seed(cam, pixel_u(cam, px), pixel_v(cam, gy));
                                sidx = 0u;
                                sts128(sb + 4u * GROUP_STRIDE, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
                                if (p.out_primary_id && cam.sample_count == 0u) p.out_primary_id[(size_t)ly * cam.width + px] = 0xffffffffu;
                                if (p.out_primary_depth && cam.sample_count == 0u) p.out_primary_depth[(size_t)ly * cam.width + px] = BVR_INF;
                                ps = P_NEW_PATH;
                            }
                        }
                    }
                }
                if (!__any_sync(full, ps == P_EMPTY || (ps == P_NEW_PATH && sidx >= cam.sample_count))) break;
            }
            // --- A6: camera rays (raytrace.wgsl:139-156), one site ---
            if (ps == P_NEW_PATH) {
                const uint32_t px = pix & 0xffffu, ly = pix >> 16;
                const Ray r = random_ray_from_uv(cam, pixel_u(cam, px), pixel_v(cam, shard_global_row(p.shard, ly)), rng);
                so = r.o; sd = r.d;
                thr = v3(1.0f, 1.0f, 1.0f);
                bounce = 0u;
                first_depth = BVR_INF;
                ps = P_RAY_READY;
            }
            // --- write the slot back ---
            if (ps == P_RAY_READY) {
                sts128(sb, make_float4(so.x, so.y, so.z, BVR_INF));
                sts128(sb + GROUP_STRIDE, make_float4(sd.x, sd.y, sd.z, __uint_as_float(0xffffffffu)));
                sts128(sb + 2u * GROUP_STRIDE, make_float4(thr.x, thr.y, thr.z, __uint_as_float(rng)));
                sts128(sb + 3u * GROUP_STRIDE, make_float4(__uint_as_float(pix), __uint_as_float(sidx),
                                                           __uint_as_float(bounce), first_depth));
            }
            const int nst = ps == P_RAY_READY ? S_READY : S_DONE;
            if (as == 0) st0 = nst;
            if (as == 1) st1 = nst;
        }

        // ======================= phase S: switch to a slot whose ray is ready =======================
        if (state != L_TRAVERSE) {
            int rs = -1;
            if (st0 == S_READY) rs = 0; else if (st1 == S_READY) rs = 1;
            if (rs >= 0) {
                const uint32_t sb = slot0 + (rs == 1 ? SLOT_STRIDE : 0u);
                const float4 g0 = lds128(sb), g1 = lds128(sb + GROUP_STRIDE);
                ray.o = v3(g0.x, g0.y, g0.z);
                ray.d = v3(g1.x, g1.y, g1.z);
                // 1/d feeds the box tests only (culling): the approximate reciprocal is enough
                inv = v3(rcp_approx(ray.d.x), rcp_approx(ray.d.y), rcp_approx(ray.d.z));
                noi = v3(-(ray.o.x * inv.x), -(ray.o.y * inv.y), -(ray.o.z * inv.z));
                a = vdot(ray.d, ray.d);
                closest.t = BVR_INF;
                closest.model = 0xffffffffu;
                sp_addr = s_stack0;
                pending = V4_NONE;
                cur = root;
                rays++;
                cs = rs;
                if (rs == 0) st0 = S_TRAV; else st1 = S_TRAV;
                state = L_TRAVERSE;
            }
        }

        // ======================= phase B: traversal =======================
        for (;;) {
            bool blocked = false;
#pragma unroll
            for (int rep = 0; rep < BVR_V4_STEPS_PER_VOTE; rep++) {
                if (state == L_TRAVERSE) {
                    uint32_t c = cur;
                    if (c < V4_LEAF) {                       // inner node: test both children
                        const uint32_t na = s_pairs + c * 64u;
                        const float4 q0 = lds128(na), q1 = lds128(na + 16u), q2 = lds128(na + 32u);
                        const uint2 rr = lds64(na + 48u);
                        float d0, d1;
                        const bool h0 = box_cull(inv, noi, closest.t, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, d0);
                        const bool h1 = box_cull(inv, noi, closest.t, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, d1);
                        const bool first0 = d0 < d1;         // ties go to the second child (reference LIFO order)
                        if (h0 && h1) {
                            // far child: distance rounded towards zero (conservative for the cull at pop time)
                            sts32(sp_addr, (__float_as_uint(first0 ? d1 : d0) & ~V4_REF_MASK) | (first0 ? rr.y : rr.x));
                            sp_addr += STACK_STRIDE;
                            c = first0 ? rr.x : rr.y;
                        } else {
                            c = h0 ? rr.x : (h1 ? rr.y : V4_NONE);
                        }
                    }
                    if (c & V4_LEAF) {                       // leaf: park it, or wait for the batch test
                        if (pending == V4_NONE) { pending = c; c = V4_NONE; }
                        else blocked = true;
                    }
                    if (c == V4_NONE) {
                        while (sp_addr != s_stack0) {        // pop until an entry survives the cull
                            sp_addr -= STACK_STRIDE;
                            const uint32_t e = lds32(sp_addr);
                            if (__uint_as_float(e & ~V4_REF_MASK) < closest.t) { c = e & V4_REF_MASK; break; }
                        }
                        if (c == V4_NONE) {
                            if (pending == V4_NONE) state = L_FINISHED;   // traversal finished
                            else blocked = true;                          // only the parked leaf is left
                        }
                    }
                    cur = c;
                }
            }
            // batched sphere tests: once enough lanes cannot continue without theirs (or none can continue)
            const unsigned blk = __ballot_sync(full, blocked);
            const unsigned trav = __ballot_sync(full, state == L_TRAVERSE);
            if (trav == 0u) break;
            const uint32_t nblk = (uint32_t)__popc(blk), ntrav = (uint32_t)__popc(trav);
            if (nblk >= tune.leaf_batch_lanes || nblk == ntrav) {
                if (state == L_TRAVERSE && pending != V4_NONE) {
                    test_leaf(sv, ray, a, pending & 0x3ffu, closest);
                    pending = V4_NONE;
                }
            }
            if (32u - ntrav >= tune.switch_lanes) {
                // a lane whose ray just finished can continue with its other slot if that one holds a ready ray
                const int other = cs == 0 ? st1 : st0;
                const bool can_switch = state == L_FINISHED && other == S_READY;
                const unsigned sw = __ballot_sync(full, can_switch);
                if ((uint32_t)__popc(sw) >= tune.switch_lanes) break;
                const bool has_work = state == L_FINISHED || st0 == S_PENDING || st1 == S_PENDING || st0 == S_EMPTY || st1 == S_EMPTY;
                const unsigned stuck = __ballot_sync(full, state != L_TRAVERSE && !can_switch && has_work);
                if ((uint32_t)__popc(stuck) >= tune.stuck_lanes) break;
            }
        }
        // park the finished ray's hit in its slot
        if (state == L_FINISHED) {
            const uint32_t sb = slot0 + (cs == 1 ? SLOT_STRIDE : 0u);
            sts32(sb + 12u, __float_as_uint(closest.t));
            sts32(sb + GROUP_STRIDE + 12u, closest.model);
            if (cs == 0) st0 = S_PENDING; else st1 = S_PENDING;
            state = L_IDLE;
        }
    }

    unsigned long long sum = rays;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(full, sum, o);
    if (lane == 0u && p.ray_counter && sum) atomicAdd(p.ray_counter, sum);
}

template <int THREADS>
int launch_v4(const RenderParams& p, uint32_t n_inner, uint32_t n_models, uint32_t tree_depth,
              unsigned int* pixel_counter, Tuning4 tune, int sm_count, cudaStream_t stream) {
    const uint32_t stack_cap = tree_depth + 1u;
    const size_t scene_bytes = (size_t)(4u * n_inner + n_models + 2u * p.scene.n_materials + (n_models + 3u) / 4u) * 16u;
    const size_t slot_bytes = (size_t)THREADS * 2u * 5u * 16u;
    const size_t stack_bytes = (size_t)THREADS * stack_cap * sizeof(uint32_t);
    const size_t max_smem = 227u * 1024u;
    const size_t smem = scene_bytes + slot_bytes + stack_bytes;
    if (smem > max_smem) return -1;
    auto kern = megakernel_v4<THREADS>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    int blocks_per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, THREADS, smem) != cudaSuccess || blocks_per_sm < 1)
        return -1;
    const uint32_t tiles = ((p.cam.width + 7u) / 8u) * ((p.shard.rows + 3u) / 4u);
    uint32_t grid = (uint32_t)(sm_count * blocks_per_sm);
    const uint32_t max_useful = (tiles * 32u + 2u * THREADS - 1u) / (2u * THREADS);
    if (grid > max_useful) grid = max_useful;
    if (grid == 0) return 0;
    kern<<<grid, THREADS, smem, stream>>>(p, pixel_counter, n_inner, n_models, tune);
    return 1;
}

}  // namespace

// Returns -1 when the scene does not qualify (11-bit child refs, everything in shared memory): the caller then
// uses v3.
int launch_megakernel_v4(const RenderParams& p, uint32_t n_inner, uint32_t n_models, uint32_t tree_depth,
                         uint32_t max_leaf_models, unsigned int* pixel_counter, int threads, uint32_t shade_lanes,
                         uint32_t stuck_lanes, uint32_t switch_lanes, uint32_t leaf_batch_lanes, int sm_count,
                         cudaStream_t stream) {
    if (n_inner > 1024u || n_models > 1024u || max_leaf_models > 1u) return -1;
    if (p.cam.width > 0xffffu || p.shard.rows > 0xffffu) return -1;   // pixel packed as px | ly << 16
    Tuning4 t{shade_lanes, stuck_lanes, switch_lanes, leaf_batch_lanes};
    switch (threads) {
        case 512: return launch_v4<512>(p, n_inner, n_models, tree_depth, pixel_counter, t, sm_count, stream);
        case 640: return launch_v4<640>(p, n_inner, n_models, tree_depth, pixel_counter, t, sm_count, stream);
        case 768: return launch_v4<768>(p, n_inner, n_models, tree_depth, pixel_counter, t, sm_count, stream);
        default: return -1;
    }
}

}  // namespace bvr
