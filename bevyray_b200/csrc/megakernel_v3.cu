// megakernel_v3.cu — persistent-lane megakernel with STAGED shading and postponed leaf tests.
//
// Persistent lanes: every lane owns one pixel at a time, taken from a tile-ordered queue; scene and traversal stacks
// in shared memory.  How divergent work is regrouped inside the warp (ncu on the unstaged predecessor v2: 10 of 32
// lanes active per instruction, shading and sphere tests running at 4-5 lanes — profiles/r01_v2_*):
//
//   * phase A is cut into stages that every waiting lane walks through together, each stage having ONE
//     code site per expensive operation: classification draws (raytrace.wgsl:234,248) -> one shared
//     rejection loop for all unit-ball samples (random.wgsl:17-26; diffuse needs two, metal one) -> one
//     normalize site shared by the miss path (background_gradient) and the hit record -> one normalize
//     site shared by metal and glass -> per-material direction -> path bookkeeping -> one ray-generation
//     site -> one ray-setup site;
//   * phase B postpones sphere tests: a lane that reaches a leaf parks it and keeps walking; parked
//     leaves are tested together once enough lanes hold one (or a lane cannot continue without it).
//     Testing a sphere later never changes the closest hit; it only delays culling.
//
// Arithmetic of everything that reaches the image is the strict set of trace.cuh (bit-identical to the
// oracle); box tests use FMA + FMNMX3 (culling only).

#include <algorithm>

#include "kernels.cuh"
#include "wide4.cuh"

namespace bvr {

namespace {

enum LaneState : int { NEED_PIXEL = 0, NEW_PATH = 1, RAY_READY = 2, TRAVERSE = 3, SHADE = 4, DONE = 5 };
enum Kind : int { K_NONE = 0, K_MISS = 1, K_METAL = 2, K_GLASS = 3, K_DIFFUSE = 4 };

#define V3_NONE 0x7fffffffu
#ifndef BVR_FAR_GENERIC
#define BVR_FAR_GENERIC 1      // MODE 5: far rays pick their records through a generic pointer (no predicated second load path)
#endif
#ifndef BVR_LEAF_BATCH
#define BVR_LEAF_BATCH 1       // lean loop: parked leaves are tested once this many lanes are blocked
#endif
#ifndef BVR_POW2_ALL_MODES
#define BVR_POW2_ALL_MODES 1   // 0: only the shared-memory modes multiply by 1/n when n is a power of two
#endif
#ifndef BVR_RECONVERGE
#define BVR_RECONVERGE 1
#endif
#ifndef BVR_STEPS_PER_VOTE
#define BVR_STEPS_PER_VOTE 2   // traversal steps between two rounds of warp votes
#endif
#ifndef BVR_W4_STEPS_PER_VOTE
#define BVR_W4_STEPS_PER_VOTE 4   // ... on the records walked in HBM/L2 (C4: 1 / 2 / 3 / 4 steps = 384 / 369 / 367 / 363 ms)
#endif

// Culling-only slab test on (centre, half extent) boxes.  Per axis: tc = fma(c, 1/d, -o/d), th = h * |1/d|,
// lo = tc - th, hi = tc + th — four FMA-pipe instructions instead of two FFMA + two FMNMX: the traversal loop is
// bound by the ALU pipe (ncu r01_v3b: ALU 77 %, FMA 20 %), so the per-axis min/max is traded for arithmetic.
// entry = max(lo.x, lo.y, lo.z, 0), exit = min(hi.x, hi.y, hi.z, closest); worth visiting iff entry <= exit.
__device__ __forceinline__ bool box_cull(V3 inv, V3 ainv, V3 noi, float closest_t, float cx, float cy, float cz, float hx,
                                         float hy, float hz, float& entry) {
    // three FFMA per axis: tc = c/d - o/d, lo = tc - h*|1/d|, hi = tc + h*|1/d|
    const float tcx = __fmaf_rn(cx, inv.x, noi.x), tcy = __fmaf_rn(cy, inv.y, noi.y), tcz = __fmaf_rn(cz, inv.z, noi.z);
    const float lox = __fmaf_rn(-hx, ainv.x, tcx), loy = __fmaf_rn(-hy, ainv.y, tcy), loz = __fmaf_rn(-hz, ainv.z, tcz);
    const float hix = __fmaf_rn(hx, ainv.x, tcx), hiy = __fmaf_rn(hy, ainv.y, tcy), hiz = __fmaf_rn(hz, ainv.z, tcz);
    entry = fmaxf(fmaxf(lox, loy), fmaxf(loz, 0.0f));
    const float exit = fminf(fminf(hix, hiy), fminf(hiz, closest_t));
    return entry <= exit;
}

// Slab test on a 32-byte quantised record (scene_kernels.cu: build_pairs_q16_kernel).  A coordinate word holds
// lo | hi << 16 on the scene's 16-bit grid; PRMT builds the float 2^23 + q from the half the ray enters (selN)
// or leaves (selF) through, and ONE FFMA maps it to the ray parameter: t = (2^23 + q) * (step/d) + C with
// C = (base - o)/d - 2^23 * step/d.  C carries up to half a grid step of rounding error, which the extra step
// of outward rounding in the records absorbs.  No per-axis min/max: the ray's octant picks near and far.
// prmt.b32 without the `& 0x7777` that __byte_perm has to put in front of a selector it cannot see through
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}
__device__ __forceinline__ bool box_cull_q16(uint32_t wx, uint32_t wy, uint32_t wz, V3 sinv, V3 cq, uint32_t snx,
                                             uint32_t sny, uint32_t snz, uint32_t sfx, uint32_t sfy, uint32_t sfz,
                                             float closest_t, float& entry) {
    const uint32_t magic = 0x00004b00u;
    // (near, far) of one axis share the multiplier and the constant: one FFMA2 with broadcast operands per axis
    float tnx, tfx, tny, tfy, tnz, tfz;
    upk2(ffma2(pk2(__uint_as_float(prmt(wx, magic, snx)), __uint_as_float(prmt(wx, magic, sfx))),
               pk2(sinv.x, sinv.x), pk2(cq.x, cq.x)), tnx, tfx);
    upk2(ffma2(pk2(__uint_as_float(prmt(wy, magic, sny)), __uint_as_float(prmt(wy, magic, sfy))),
               pk2(sinv.y, sinv.y), pk2(cq.y, cq.y)), tny, tfy);
    upk2(ffma2(pk2(__uint_as_float(prmt(wz, magic, snz)), __uint_as_float(prmt(wz, magic, sfz))),
               pk2(sinv.z, sinv.z), pk2(cq.z, cq.z)), tnz, tfz);
    entry = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f));
    const float exit = fminf(fminf(tfx, tfy), fminf(tfz, closest_t));
    return entry <= exit;
}

#define Q16_LEAF 0x100000u
#define Q16_NONE 0x200000u
#define Q16_REF_MASK 0x1fffffu

// BVR_SELFCHECK=1: a runtime witness for the culling-only shortcuts (tight boxes, 16-bit grid, FMA slab test, near-first
// order, postponed leaf tests).  About one finished ray in 1024 is LOGGED by the render kernel (origin, direction, the
// closest hit it found); selfcheck_kernel then traces the logged rays again with the verbatim reference-order traversal
// on the reference's own boxes (trace.cuh: raycast_reference_order, strict arithmetic) and counts different closest hits.
__global__ void selfcheck_kernel(const SceneView sv, const float4* __restrict__ log, const unsigned int* __restrict__ n_logged,
                                 uint32_t cap, unsigned long long* __restrict__ counters) {
    const uint32_t n = min(*n_logged, cap);
    unsigned long long bad = 0, seen = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 a = log[2u * i], b = log[2u * i + 1u];
        const Ray r{v3(a.x, a.y, a.z), v3(b.x, b.y, b.z)};
        const Hit want = raycast_reference_order(sv, r);
        const bool same = __float_as_uint(want.t) == __float_as_uint(a.w) && (want.t == BVR_INF || want.model == __float_as_uint(b.w));
        seen++;
        if (!same) bad++;
    }
    if (seen) atomicAdd(&counters[0], seen);
    if (bad) atomicAdd(&counters[1], bad);
}

// BVR_RENDER_EXTRA_SAMPLE: does the pixel (px, global row gy) take one sample more than cam.sample_count?
__device__ __forceinline__ bool extra_sample(const RenderParams& p, uint32_t px, uint32_t gy) {
    return p.extra_modulus != 0u && ((px >> 3) + (gy >> 2) + p.extra_phase) % p.extra_modulus < p.extra_count;
}

struct Tuning {
    uint32_t shade_wait_lanes;   // leave phase B when this many lanes wait for shading
    uint32_t leaf_batch_lanes;   // test parked leaves when this many lanes hold one
    uint32_t n_top;              // MODE 3: records [0, n_top) of the array are staged in shared memory
    uint32_t n_hot;              // MODE 3: records [n_top, n_hot) are loaded L1::evict_last, the rest L1::no_allocate (0 = no hints)
    uint32_t lean_w4;            // MODE 3 -> MODE 7: the (cur, pending, stack) loop on the 16-bit-grid records
    const uint32_t* tile_order;  // tiles in the order they are handed out (heaviest first, tile_order.cu), or null = row-major
    uint32_t* tile_cost;         // += rays traced per 8x4 tile (feeds the next frame's order), or null
};

// MODE 0: scene in shared memory (64-byte fp32 records).  MODE 1: scene in HBM/L2, 64-byte fp32 records, 8-byte
// stack entries.  MODE 2: scene in HBM/L2, 32-byte quantised records (one 256-bit load per visit), 4-byte stack
// entries (11 bits of distance | 21 bits of child ref).  MODE 3: as 2, on the 64-byte 4-wide records (two 256-bit
// loads issued together): two levels of the tree per dependent fetch, the children sorted as packed 32-bit keys.
// MODE 4: scene in shared memory as 112-byte 4-wide fp32 records, refs in 11 bits, 4-byte stack entries (21 bits of
// distance | 11 bits of ref): half as many traversal steps as MODE 0, so half the per-step overhead.
// MODE 5: as 4, with the TIGHT-box records in shared memory (scene_kernels.cu): fewer node visits and sphere
// tests; a ray whose origin is too far from some radius group for the tight boxes to be safe reads the reference
// records from HBM/L2 instead (same code, different loads).
template <int THREADS, int MODE>
__global__ void __launch_bounds__(THREADS) megakernel_v3(const RenderParams p, unsigned int* __restrict__ pixel_counter,
                                                         const uint32_t n_inner, const uint32_t n_models,
                                                         const Tuning tune) {
    constexpr bool SMEM_SCENE = MODE == 0 || MODE == 4 || MODE == 5 || MODE == 6;
    constexpr bool TIGHT = MODE == 5 || MODE == 6;
    constexpr bool BOTH = MODE == 6;          // tight AND reference records staged in shared memory
    constexpr bool Q16 = MODE == 2 || MODE == 3 || MODE == 7;
    constexpr bool W4 = MODE == 3 || MODE == 7;
    constexpr bool LEAN = MODE == 4 || MODE == 5 || MODE == 6 || MODE == 7;   // the (cur, pending, stack) traversal loop
    constexpr bool S4 = MODE == 4 || MODE == 5 || MODE == 6;
    constexpr bool STACK4 = Q16 || S4;        // 4-byte stack entries
    constexpr uint32_t NONE = Q16 ? Q16_NONE : (S4 ? S4_NONE : V3_NONE);
    extern __shared__ float4 smem[];
    const CameraParams& cam = p.cam;
    const unsigned full = 0xffffffffu;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;

    SceneView sv = p.scene;
    float4* sm_cursor = smem;
    if (SMEM_SCENE) {
        const uint32_t rec4 = S4 ? 7u : 4u;   // float4 per inner node
        float4* sm_pairs = sm_cursor;     sm_cursor += rec4 * n_inner;
        float4* sm_ref = sm_cursor;       if (BOTH) sm_cursor += rec4 * n_inner;
        float4* sm_spheres = sm_cursor;   sm_cursor += n_models;
        float4* sm_materials = sm_cursor; sm_cursor += 2u * sv.n_materials;
        uint32_t* sm_matid = reinterpret_cast<uint32_t*>(sm_cursor);
        sm_cursor += (n_models + 3u) / 4u;
        for (uint32_t i = tid; i < rec4 * n_inner; i += THREADS) sm_pairs[i] = TIGHT ? p.scene.nodes4_tight[i] : (S4 ? p.scene.nodes4_ch[i] : p.scene.pairs_ch[i]);
        if (BOTH) for (uint32_t i = tid; i < rec4 * n_inner; i += THREADS) sm_ref[i] = p.scene.nodes4_ch[i];
        for (uint32_t i = tid; i < n_models; i += THREADS) sm_spheres[i] = p.scene.spheres[i];
        for (uint32_t i = tid; i < 2u * sv.n_materials; i += THREADS) sm_materials[i] = p.scene.materials[i];
        for (uint32_t i = tid; i < n_models; i += THREADS) sm_matid[i] = p.scene.sphere_material[i];
        sv.pairs_ch = sm_pairs;
        if (BOTH) sv.pairs = sm_ref;      // (MODE 6 only: the staged reference records)
        sv.spheres = sm_spheres;
        sv.materials = sm_materials;
        sv.sphere_material = sm_matid;
        __syncthreads();
    }
    // MODE 3: the hot top of the tree — the first n_top 64-byte records of the array, which the upload numbered breadth
    // first (scene_kernels.cu: top_bfs_kernel) — staged in shared memory: the levels every ray walks cost no L1/L2 traffic
    const uint32_t n_top = W4 ? tune.n_top : 0u, n_hot = W4 ? tune.n_hot : 0u;
    uint32_t s_top = 0u;
    if constexpr (W4) {
        uint4* sm_top = reinterpret_cast<uint4*>(sm_cursor);
        sm_cursor += 4u * n_top;
        for (uint32_t i = tid; i < 4u * n_top; i += THREADS) sm_top[i] = __ldg(p.scene.nodes4_q + i);
        s_top = opaque(smem_addr(sm_top));
        __syncthreads();
    }
    // per-lane stack: entry k of lane t at s_stack0 + k * STACK_STRIDE (interleaved: conflict-free)
    constexpr uint32_t STACK_STRIDE = THREADS * (STACK4 ? 4u : 8u);
    const uint32_t s_stack0 = opaque(smem_addr(sm_cursor) + tid * (STACK4 ? 4u : 8u));
    // MODE 2: the quantisation grid and the root in 21-bit form
    const float qbx = Q16 ? sv.qgrid[0] : 0.f, qby = Q16 ? sv.qgrid[1] : 0.f, qbz = Q16 ? sv.qgrid[2] : 0.f;
    const float qsx = Q16 ? sv.qgrid[4] : 0.f, qsy = Q16 ? sv.qgrid[5] : 0.f, qsz = Q16 ? sv.qgrid[6] : 0.f;
    const uint32_t root = !sv.has_scene ? NONE
                          : (Q16 ? ((sv.root_ref & BVR_LEAF_BIT) ? (Q16_LEAF | (sv.root_ref & 0xfffffu)) : sv.root_ref)
                             : S4 ? ((sv.root_ref & BVR_LEAF_BIT) ? (S4_LEAF | (sv.root_ref & 0x3ffu)) : sv.root_ref)
                                  : sv.root_ref);
    uint32_t snx = 0, sny = 0, snz = 0, sfx = 0, sfy = 0, sfz = 0;   // MODE 2: PRMT selectors of the near / far halves
    const uint32_t n_groups = TIGHT ? __float_as_uint(__ldg(&sv.tight_groups[0]).x) : 0u;
    bool far_ray = false;   // MODE 5: this ray walks the reference boxes
    const uint32_t s_pairs = SMEM_SCENE ? opaque(smem_addr(sv.pairs_ch)) : 0u;
    const float4* const g_pairs = sv.pairs_ch;   // MODE 5: the staged tight records through a generic pointer
    const uint32_t s_pairs_ref = BOTH ? opaque(smem_addr(sv.pairs)) : 0u;
    const uint32_t s_spheres = SMEM_SCENE ? opaque(smem_addr(sv.spheres)) : 0u;
    uint32_t s_rec = s_pairs;                    // S4: shared-window address of the records this ray walks

    const uint32_t tiles_x = (cam.width + 7u) / 8u, tiles_y = (p.shard.rows + 3u) / 4u;
    const uint32_t total_slots = tiles_x * tiles_y * 32u;

    int state = NEED_PIXEL;
    uint32_t pxy = 0;   // px | ly << 16
    float u = 0.0f, v = 0.0f;
    uint32_t rng = 0, bounce = 0;
    uint32_t sleft = 0;   // samples this pixel still has to take; bit 31 = its first sample has not been shaded yet
    V3 total = v3(0.0f, 0.0f, 0.0f);
    float total_depth = 0.0f, first_depth = BVR_INF;
    V3 throughput = v3(1.0f, 1.0f, 1.0f);
    Ray ray{v3(0, 0, 0), v3(0, 0, 1)};
    V3 inv = v3(0, 0, 0), ainv = v3(0, 0, 0), noi = v3(0, 0, 0);
    u64 inv_xy = 0, noi_xy = 0;                  // S4: (x, y) of 1/d and of -o/d as one register pair (FFMA2 operands)
    const float4* rec_base = g_pairs;            // MODE 5: records this ray walks (shared tight / global reference)
    float a = 1.0f;
    Hit closest{BVR_INF, 0xffffffffu};
    uint32_t cur = NONE, pending = NONE;
    uint32_t sp_addr = s_stack0;    // next free stack slot
    uint32_t rays = 0;

    for (;;) {
        // ======================= phase A: staged shading =======================
        if (__any_sync(full, state == SHADE)) {
            // --- A1: classification (raytrace.wgsl:193-201, 232-248) ---
            int kind = K_NONE;
            uint32_t mid = 0;
            if (state == SHADE) {
                if (p.selfcheck_log != nullptr && (rng & 1023u) == 0u) {
                    const uint32_t slot = atomicAdd(p.selfcheck_count, 1u);
                    if (slot < p.selfcheck_cap) {
                        p.selfcheck_log[2u * slot] = make_float4(ray.o.x, ray.o.y, ray.o.z, closest.t);
                        p.selfcheck_log[2u * slot + 1u] = make_float4(ray.d.x, ray.d.y, ray.d.z, __uint_as_float(closest.model));
                    }
                }
                if (bounce == 0u) {
                    first_depth = closest.t;
                    if ((int)sleft < 0) {   // first sample's primary hit
                        sleft &= 0x7fffffffu;
                        const size_t lpix = (size_t)(pxy >> 16) * cam.width + (pxy & 0xffffu);
                        if (p.out_primary_id) p.out_primary_id[lpix] = closest.t == BVR_INF ? 0xffffffffu : closest.model;
                        if (p.out_primary_depth) p.out_primary_depth[lpix] = closest.t;
                    }
                }
                if (closest.t == BVR_INF) {
                    kind = K_MISS;
                } else {
                    mid = sv.sphere_material[closest.model];
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
                // 2*x - 1 with x = f32(state) * 2^-32: both scalings are exact, so the single rounding of
                // fma(f32(state), 2^-31, -1) is the same rounding the reference's (2*x) - 1 performs
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
            bool path_end = false;
            V3 sample_color = v3(0.0f, 0.0f, 0.0f);
            if (state == SHADE) {
                const float4 sph = kind == K_MISS ? make_float4(0.f, 0.f, 0.f, 0.f) : sv.spheres[closest.model];
                const V3 position = vadd(ray.o, vscale(closest.t, ray.d));            // ray_at, raytrace.wgsl:130-132
                const V3 nin = kind == K_MISS ? ray.d : vsub(position, v3(sph.x, sph.y, sph.z));
                const V3 unit = vnormalize(nin);
                if (kind == K_MISS) {
                    // background_gradient, raytrace.wgsl:364-369
                    const float aa = fmul(0.5f, fadd(unit.y, 1.0f));
                    const float ia = fsub(1.0f, aa);
                    const V3 bg = v3(fadd(fmul(ia, 1.0f), fmul(aa, 0.5f)), fadd(fmul(ia, 1.0f), fmul(aa, 0.7f)),
                                     fadd(fmul(ia, 1.0f), fmul(aa, 1.0f)));
                    const V3 lin = vmul(throughput, bg);
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
                        const V3 un = vnormalize(kind == K_METAL ? reflect3(ray.d, normal) : ray.d);
                        if (kind == K_METAL) {                                         // raytrace.wgsl:234-246
                            dir = vadd(un, vscale(m1.x, b2));
                            absorbed = vdot(dir, normal) < 0.0f;
                        } else {                                                       // raytrace.wgsl:248-282
                            const bool front_face = vdot(ray.d, normal) < 0.0f;
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
                    ray.o = position;
                    ray.d = dir;
                    if (absorbed) {
                        path_end = true;                                               // raytrace.wgsl:207-209
                    } else {
                        throughput = vmul(throughput, attenuation);
                        bounce++;
                        if (bounce > cam.bounce_count) path_end = true;                // raytrace.wgsl:214-216
                    }
                }
                if (path_end) {
                    if (first_depth == BVR_INF) first_depth = cam.fallback_far;
                    total = vadd(total, sample_color);
                    total_depth = fadd(total_depth, first_depth);
                    sleft--;
                    state = NEW_PATH;
                } else {
                    state = RAY_READY;
                }
            }
        }
        // --- A4: pixel store (average, fused composite raytrace.wgsl:104-120) ---
        if (state == NEW_PATH && sleft == 0u) {
            const uint32_t px = pxy & 0xffffu, ly = pxy >> 16;
            const uint32_t gy = shard_global_row(p.shard, ly);
            // BVR_RENDER_EXTRA_SAMPLE: the pixels of some tiles took one sample more than cam.sample_count
            const float n = (float)(cam.sample_count + (extra_sample(p, px, gy) ? 1u : 0u));
            // total / n (raytrace.wgsl:170).  When n is a power of two, x * (1/n) IS x / n — both round the same exact value
            // x * 2^-k once — so the four IEEE divisions (some forty instructions at one or two active lanes) become four
            // multiplications: what a 1-spp or 4-spp frame mostly consists of is this per-pixel code.
            float4 out;
            float depth_avg;
            if ((BVR_POW2_ALL_MODES || S4) && p.inv_pow2_samples != 0.0f) {
                const float r = p.inv_pow2_samples;
                out = make_float4(fmul(total.x, r), fmul(total.y, r), fmul(total.z, r), 1.0f);
                depth_avg = fmul(total_depth, r);
            } else {
                out = make_float4(fdiv(total.x, n), fdiv(total.y, n), fdiv(total.z, n), 1.0f);
                depth_avg = fdiv(total_depth, n);
            }
            if (cam.level == 1u || cam.level == 2u) {
                const size_t gpix = (size_t)gy * cam.width + px;
                if (raster_wins(cam, p.raster_depth[gpix], depth_avg)) out = p.raster_rgba[gpix];
            }
            const size_t lpix = (size_t)ly * cam.width + px;
            // sample sharding: the partial frame leaves the kernel weighted by this rank's share of the samples, and
            // out_rgba may be this rank's slot in a buffer on ANOTHER GPU (the stores travel over NVLink as pixels finish)
            const float ow = p.extra_modulus ? fmul(n, p.out_weight) : p.out_weight;   // uneven samples: the weight of ONE sample x n
            if (p.out_rgba) p.out_rgba[lpix] = ow == 1.0f ? out : make_float4(fmul(out.x, ow), fmul(out.y, ow), fmul(out.z, ow), fmul(out.w, ow));
            if (p.out_rt_depth) p.out_rt_depth[lpix] = ow == 1.0f ? depth_avg : fmul(depth_avg, ow);
            if (cam.sample_count == 0u) {
                if (p.out_primary_id) p.out_primary_id[lpix] = 0xffffffffu;
                if (p.out_primary_depth) p.out_primary_depth[lpix] = BVR_INF;
            }
            if (p.out_srgb8) p.out_srgb8[lpix] = store_srgb8(out);
            // rays this pixel cost = the lane's count now minus its count when it took the pixel (subtracted in A5)
            if (tune.tile_cost) atomicAdd(tune.tile_cost + ((ly >> 2) * tiles_x + (px >> 3)), rays);
            state = NEED_PIXEL;
        }
        // --- A5: pull new pixels from the tile-ordered queue (warp-convergent) ---
        for (;;) {
            const unsigned need_px = __ballot_sync(full, state == NEED_PIXEL);
            if (need_px == 0u) break;
            const int leader = __ffs(need_px) - 1;
            unsigned base = 0;
            if ((int)lane == leader) base = atomicAdd(pixel_counter, (unsigned)__popc(need_px));
            base = __shfl_sync(full, base, leader);
            if (state == NEED_PIXEL) {
                const uint32_t slot = base + (uint32_t)__popc(need_px & ((1u << lane) - 1u));
                if (slot >= total_slots) {
                    state = DONE;
                } else {
                    uint32_t tile = slot >> 5;
                    const uint32_t within = slot & 31u;
                    if (tune.tile_order) tile = __ldg(tune.tile_order + tile);
                    const uint32_t px = (tile % tiles_x) * 8u + (within & 7u);
                    const uint32_t ly = (tile / tiles_x) * 4u + (within >> 3);
                    const uint32_t gy = shard_global_row(p.shard, ly);
                    if (px < cam.width && ly < p.shard.rows && gy < cam.height) {
                        if (tune.tile_cost) atomicSub(tune.tile_cost + tile, rays);
                        pxy = px | (ly << 16);
                        u = pixel_u(cam, px);
                        v = pixel_v(cam, gy);
                        rng = pixel_seed(cam, u, v);
                        sleft = cam.sample_count + (extra_sample(p, px, gy) ? 1u : 0u);
                        if (sleft) sleft |= 0x80000000u;
                        total = v3(0.0f, 0.0f, 0.0f);
                        total_depth = 0.0f;
                        state = NEW_PATH;
                    }
                }
            }
        }
        if (__all_sync(full, state == DONE)) break;
        // --- A6: camera rays (raytrace.wgsl:139-156), one site ---
        if (state == NEW_PATH && sleft != 0u) {
            ray = random_ray_from_uv(cam, u, v, rng);
            throughput = v3(1.0f, 1.0f, 1.0f);
            bounce = 0u;
            first_depth = BVR_INF;
            state = RAY_READY;
        }
        // --- A7: ray setup, one site for camera rays and scattered rays ---
        if (state == RAY_READY) {
            // 1/d feeds the box tests only (culling), so the approximate reciprocal (MUFU.RCP, 1 ulp) is
            // enough: boxes are padded by 0.1, rounding is ~1e-7 relative
            inv = v3(rcp_approx(ray.d.x), rcp_approx(ray.d.y), rcp_approx(ray.d.z));
            if (Q16) {
                // inv := step/d, noi := (base - o)/d - 2^23 * step/d (see box_cull_q16); selectors by octant
                snx = inv.x >= 0.0f ? 0x5410u : 0x5432u; sfx = snx ^ 0x0022u;
                sny = inv.y >= 0.0f ? 0x5410u : 0x5432u; sfy = sny ^ 0x0022u;
                snz = inv.z >= 0.0f ? 0x5410u : 0x5432u; sfz = snz ^ 0x0022u;
                const V3 bo = v3((qbx - ray.o.x) * inv.x, (qby - ray.o.y) * inv.y, (qbz - ray.o.z) * inv.z);
                inv = v3(qsx * inv.x, qsy * inv.y, qsz * inv.z);
                noi = v3(__fmaf_rn(-8388608.0f, inv.x, bo.x), __fmaf_rn(-8388608.0f, inv.y, bo.y), __fmaf_rn(-8388608.0f, inv.z, bo.z));
            } else if (S4) {
                inv_xy = pk2(inv.x, inv.y);
                noi_xy = pk2(-(ray.o.x * inv.x), -(ray.o.y * inv.y));
                noi.z = -(ray.o.z * inv.z);
            } else {
                ainv = v3(fabsf(inv.x), fabsf(inv.y), fabsf(inv.z));
                noi = v3(-(ray.o.x * inv.x), -(ray.o.y * inv.y), -(ray.o.z * inv.z));
            }
            if constexpr (TIGHT) {
                far_ray = false;
                for (uint32_t g = 0; g < n_groups; g++) {
                    const float4 gr = __ldg(&sv.tight_groups[1u + g]);
                    const float dx = ray.o.x - gr.x, dy = ray.o.y - gr.y, dz = ray.o.z - gr.z;
                    far_ray = far_ray || !(dx * dx + dy * dy + dz * dz <= gr.w);
                }
                rec_base = far_ray ? sv.nodes4_ch : g_pairs;
                if (BOTH) s_rec = far_ray ? s_pairs_ref : s_pairs;
            }
            a = vdot(ray.d, ray.d);
            closest.t = BVR_INF;
            closest.model = 0xffffffffu;
            sp_addr = s_stack0;
            pending = NONE;
            cur = root;
            rays++;
            state = TRAVERSE;
        }

        // ======================= phase B: traversal =======================
        if constexpr (LEAN) {
            // 4-wide records (staged in shared memory: S4; 16-bit-grid records in HBM/L2: MODE 7).  A lane's traversal state IS
            // (cur, pending, stack):
            //   cur < LEAFV inner record to visit, LEAFV <= cur < NONE a leaf, NONE nothing in hand;
            //   pending = a parked leaf (tested with the others' once a lane cannot go on without its test).
            // A lane with nothing in hand, nothing parked and an empty stack is idle: its ray is finished (or it has
            // none); idle lanes fall through every step without a state test.
            // (Parking a popped leaf on the spot and popping on until the lane holds an inner record again — so that every
            // step is a visit — was tried: C2 48.4 -> 53.3 ms, C4 371 -> 399 ms; the longer divergent loop costs more than
            // the saved step.  So was a hybrid stack with its deep part in local memory: profiles/r02_tuning_sweeps.txt.)
            constexpr uint32_t LEAFV = S4 ? S4_LEAF : Q16_LEAF;
            auto push = [&](uint32_t k) { sts32(sp_addr, k); sp_addr += STACK_STRIDE; };
            const bool had_ray = state == TRAVERSE;
            constexpr int STEPS = S4 ? BVR_STEPS_PER_VOTE : BVR_W4_STEPS_PER_VOTE;
            for (;;) {
#pragma unroll
                for (int rep = 0; rep < STEPS; rep++) {
                    uint32_t c = cur;
                    if (c < LEAFV) {
                        if constexpr (S4) {
                            float4 q0, q1, q2, q3, q4, q5, rr;
                            if constexpr (TIGHT && !BOTH && BVR_FAR_GENERIC) {
                                // one generic pointer per ray: LD resolves the shared / global window itself
                                const float4* nd = reinterpret_cast<const float4*>(reinterpret_cast<const char*>(rec_base) + c * 112u);
                                q0 = nd[0]; q1 = nd[1]; q2 = nd[2]; q3 = nd[3]; q4 = nd[4]; q5 = nd[5]; rr = nd[6];
                            } else if (TIGHT && !BOTH && far_ray) {
                                const float4* nd = sv.nodes4_ch + 7u * c;
                                q0 = __ldg(nd); q1 = __ldg(nd + 1); q2 = __ldg(nd + 2); q3 = __ldg(nd + 3);
                                q4 = __ldg(nd + 4); q5 = __ldg(nd + 5); rr = __ldg(nd + 6);
                            } else {
                                const uint32_t na = (BOTH ? s_rec : s_pairs) + c * 112u;
                                q0 = lds128(na); q1 = lds128(na + 16u); q2 = lds128(na + 32u);
                                q3 = lds128(na + 48u); q4 = lds128(na + 64u); q5 = lds128(na + 80u);
                                rr = lds128(na + 96u);
                            }
                            c = visit4(q0, q1, q2, q3, q4, q5, rr, inv_xy, noi_xy, inv.z, noi.z, closest.t, push);
                        } else {
                            uint4 qa, qb, qc, qd;
                            const uint4* np = sv.nodes4_q + 4u * c;
                            if (n_hot == 0u) ldg512u<0>(np, qa, qb, qc, qd);
                            else if (c < n_hot) ldg512u<1>(np, qa, qb, qc, qd);
                            else ldg512u<2>(np, qa, qb, qc, qd);
                            float e;
                            uint32_t k0 = box_cull_q16(qa.x, qa.y, qa.z, inv, noi, snx, sny, snz, sfx, sfy, sfz, closest.t, e)
                                              ? (((__float_as_uint(e) >> 20) << 21) | qa.w) : 0xffffffffu;
                            uint32_t k1 = box_cull_q16(qb.x, qb.y, qb.z, inv, noi, snx, sny, snz, sfx, sfy, sfz, closest.t, e)
                                              ? (((__float_as_uint(e) >> 20) << 21) | qb.w) : 0xffffffffu;
                            uint32_t k2 = box_cull_q16(qc.x, qc.y, qc.z, inv, noi, snx, sny, snz, sfx, sfy, sfz, closest.t, e)
                                              ? (((__float_as_uint(e) >> 20) << 21) | qc.w) : 0xffffffffu;
                            uint32_t k3 = box_cull_q16(qd.x, qd.y, qd.z, inv, noi, snx, sny, snz, sfx, sfy, sfz, closest.t, e)
                                              ? (((__float_as_uint(e) >> 20) << 21) | qd.w) : 0xffffffffu;
                            k0 = sort4_park(k0, k1, k2, k3, push);
                            c = k0 != 0xffffffffu ? (k0 & Q16_REF_MASK) : NONE;
                        }
                    }
#if BVR_RECONVERGE
                    __syncwarp();   // lanes that visited and lanes that did not park / pop together
#endif
                    if (c >= LEAFV) {
                        if (c != NONE && pending == NONE) { pending = c; c = NONE; }   // park the leaf, go on
                        if (c == NONE) {
                            // pop until an entry survives the cull (a culled entry costs ~5 instructions here)
                            while (sp_addr != s_stack0) {
                                sp_addr -= STACK_STRIDE;
                                const uint32_t e = lds32(sp_addr);
                                if constexpr (S4) {
                                    if (__uint_as_float(e & ~S4_REF_MASK) < closest.t) { c = e & S4_REF_MASK; break; }
                                } else {
                                    if (__uint_as_float((e >> 21) << 20) < closest.t) { c = e & Q16_REF_MASK; break; }
                                }
                            }
                        }
                    }
                    cur = c;
                }
                // a lane is blocked when it cannot go on without the sphere test of its parked leaf: it holds a second
                // leaf, or has nothing else left
                const bool parked = pending != NONE;
                const unsigned trav = __ballot_sync(full, cur != NONE || parked);
                if (trav == 0u) break;
#if BVR_LEAF_BATCH > 1
                const unsigned blk = __ballot_sync(full, parked && cur >= LEAFV);
                if (blk != 0u) {
#else
                if (__any_sync(full, parked && cur >= LEAFV)) {
#endif
                    // (A sphere test without early exits — one predicated update at the end, no branches, no register copies
                    // around them — was tried: 48.9 ms against 47.8 on C2; the exits skip the square root and the division
                    // often enough.)
                    // (waiting for 2, 3, 4 blocked lanes before testing measured 48.4 / 48.4 / 48.6 ms against 48.5 on C2 and
                    // 376 against 371 ms on C4: this loop tests as soon as one lane is blocked, BVR_LEAF_BATCH > 1 = the knob)
#if BVR_LEAF_BATCH > 1
                    const uint32_t nblk = (uint32_t)__popc(blk);
                    if (nblk >= (uint32_t)BVR_LEAF_BATCH || nblk == (uint32_t)__popc(trav))
#endif
                    {
                        if (parked) {
                            if constexpr (S4) {
                                const uint32_t m = pending & 0x3ffu;      // one sphere per leaf in these layouts
                                test_sphere(sv, ray, a, m, lds128(s_spheres + m * 16u), closest);
                            } else {
                                const uint32_t m = pending & 0xfffffu;
                                test_sphere(sv, ray, a, m, __ldg(sv.spheres + m), closest);
                            }
                            pending = NONE;
                        }
                    }
                }
                if (32u - (uint32_t)__popc(trav) >= tune.shade_wait_lanes) {
                    // idle lanes are waiting to shade or out of pixels; only the former justify phase A
                    const unsigned waiting = __ballot_sync(full, had_ray && cur == NONE && !parked);
                    if ((uint32_t)__popc(waiting) >= tune.shade_wait_lanes) break;
                }
            }
            if (had_ray && cur == NONE && pending == NONE) state = SHADE;   // traversal finished
        } else
        for (;;) {
            bool blocked = false;
#pragma unroll
            for (int rep = 0; rep < BVR_STEPS_PER_VOTE; rep++) {
                if (state == TRAVERSE) {
                    uint32_t c = cur;
                    if (Q16 ? c < Q16_LEAF : (S4 ? c < S4_LEAF : c < V3_NONE)) {  // inner node: test its children
                        if constexpr (S4) {
                            // (scenes staged as 4-wide records run the loop above)
                        } else if constexpr (W4) {
                            // four children: key = 11 bits of entry distance | 21 bits of ref, 0xffffffff = not entered
                            uint4 qa, qb, qc, qd;
                            if (c < n_top) {
                                const uint32_t na = s_top + c * 64u;
                                qa = lds128u(na); qb = lds128u(na + 16u); qc = lds128u(na + 32u); qd = lds128u(na + 48u);
                            } else {
                                const uint4* np = sv.nodes4_q + 4u * c;
                                if (n_hot == 0u) ldg512u<0>(np, qa, qb, qc, qd);
                                else if (c < n_hot) ldg512u<1>(np, qa, qb, qc, qd);
                                else ldg512u<2>(np, qa, qb, qc, qd);
                            }
                            float e;
                            uint32_t k0 = box_cull_q16(qa.x, qa.y, qa.z, inv, noi, snx, sny, snz, sfx, sfy, sfz, closest.t, e)
                                              ? (((__float_as_uint(e) >> 20) << 21) | qa.w) : 0xffffffffu;
                            uint32_t k1 = box_cull_q16(qb.x, qb.y, qb.z, inv, noi, snx, sny, snz, sfx, sfy, sfz, closest.t, e)
                                              ? (((__float_as_uint(e) >> 20) << 21) | qb.w) : 0xffffffffu;
                            uint32_t k2 = box_cull_q16(qc.x, qc.y, qc.z, inv, noi, snx, sny, snz, sfx, sfy, sfz, closest.t, e)
                                              ? (((__float_as_uint(e) >> 20) << 21) | qc.w) : 0xffffffffu;
                            uint32_t k3 = box_cull_q16(qd.x, qd.y, qd.z, inv, noi, snx, sny, snz, sfx, sfy, sfz, closest.t, e)
                                              ? (((__float_as_uint(e) >> 20) << 21) | qd.w) : 0xffffffffu;
                            k0 = sort4_park(k0, k1, k2, k3, [&](uint32_t k) { sts32(sp_addr, k); sp_addr += STACK_STRIDE; });
                            c = k0 != 0xffffffffu ? (k0 & Q16_REF_MASK) : NONE;
                        } else {
                          // two children
                          uint32_t r0, r1;
                          float d0, d1;
                          bool h0, h1;
                          if constexpr (Q16) {
                            uint4 qa, qb;
                            ldg256u(sv.pairs_q + 2u * c, qa, qb);
                            r0 = qa.w; r1 = qb.w;
                            h0 = box_cull_q16(qa.x, qa.y, qa.z, inv, noi, snx, sny, snz, sfx, sfy, sfz, closest.t, d0);
                            h1 = box_cull_q16(qb.x, qb.y, qb.z, inv, noi, snx, sny, snz, sfx, sfy, sfz, closest.t, d1);
                          } else {
                            float4 q0, q1, q2;
                            if (SMEM_SCENE) {
                                const uint32_t na = s_pairs + c * 64u;
                                q0 = lds128(na); q1 = lds128(na + 16u); q2 = lds128(na + 32u);
                                const uint2 rr = lds64(na + 48u);
                                r0 = rr.x; r1 = rr.y;
                            } else {
                                const float4* nd = sv.pairs_ch + 4u * c;
                                float4 q3;
                                ldg256(nd, q0, q1);
                                ldg256(nd + 2, q2, q3);
                                r0 = __float_as_uint(q3.x); r1 = __float_as_uint(q3.y);
                            }
                            h0 = box_cull(inv, ainv, noi, closest.t, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, d0);
                            h1 = box_cull(inv, ainv, noi, closest.t, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, d1);
                          }
                          const bool first0 = d0 < d1;         // ties go to the second child (reference LIFO order)
                          if (h0 && h1) {
                            if constexpr (Q16) {
                                // far child: 11 bits of distance, rounded towards zero (conservative at pop time)
                                sts32(sp_addr, ((__float_as_uint(first0 ? d1 : d0) >> 20) << 21) | (first0 ? r1 : r0));
                            } else {
                                sts64(sp_addr, first0 ? r1 : r0, __float_as_uint(first0 ? d1 : d0));
                            }
                            sp_addr += STACK_STRIDE;
                            c = first0 ? r0 : r1;
                          } else {
                            c = h0 ? r0 : (h1 ? r1 : NONE);
                          }
                        }
                    }
                    if (Q16 ? (c & Q16_LEAF) != 0u : (S4 ? (c >= S4_LEAF && c != S4_NONE) : (int)c < 0)) {   // leaf: park it, or wait for the batch test
                        if (pending == NONE) { pending = c; c = NONE; }
                        else blocked = true;
                    }
                    if (c == NONE) {
                        // pop until an entry survives the cull: a culled entry costs ~6 instructions here
                        // instead of a whole step (ncu r01_v3b: 18.5 steps per ray, half of them dead pops)
                        while (sp_addr != s_stack0) {
                            sp_addr -= STACK_STRIDE;
                            if (Q16) {
                                const uint32_t e = lds32(sp_addr);
                                if (__uint_as_float((e >> 21) << 20) < closest.t) { c = e & Q16_REF_MASK; break; }
                            } else if (S4) {
                                const uint32_t e = lds32(sp_addr);
                                if (__uint_as_float(e & ~S4_REF_MASK) < closest.t) { c = e & S4_REF_MASK; break; }
                            } else {
                                const uint2 e = lds64(sp_addr);
                                if (__uint_as_float(e.y) < closest.t) { c = e.x; break; }
                            }
                        }
                        if (c == NONE) {
                            if (pending == NONE) state = SHADE;   // traversal finished
                            else blocked = true;                  // only the parked leaf is left
                        }
                    }
                    cur = c;
                }
            }
            // batched sphere tests: once enough lanes cannot continue without theirs (or none can continue)
            const unsigned blk = __ballot_sync(full, blocked);
            const unsigned trav = __ballot_sync(full, state == TRAVERSE);
            if (trav == 0u) break;
            const uint32_t nblk = (uint32_t)__popc(blk), ntrav = (uint32_t)__popc(trav);
            if (nblk >= tune.leaf_batch_lanes || nblk == ntrav) {
                if (state == TRAVERSE && pending != NONE) {
                    test_leaf(sv, ray, a, Q16 ? (pending & 0xfffffu) : (S4 ? (pending & 0x3ffu) : pending), closest);
                    pending = NONE;
                }
            }
            if (32u - ntrav >= tune.shade_wait_lanes) {
                // lanes outside TRAVERSE are waiting to shade or out of pixels; only the former justify phase A
                const unsigned waiting = __ballot_sync(full, state == SHADE);
                if ((uint32_t)__popc(waiting) >= tune.shade_wait_lanes) break;
            }
        }
    }

    unsigned long long sum = rays;   // < 2^32 rays per lane
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(full, sum, o);
    if (lane == 0u && p.ray_counter && sum) atomicAdd(p.ray_counter, sum);
}

template <int THREADS>
int launch_v3(const RenderParams& p, uint32_t n_inner, uint32_t n_models, uint32_t tree_depth,
              unsigned int* pixel_counter, Tuning tune, bool no_both, int sm_count, cudaStream_t stream) {
    uint32_t stack_cap = tree_depth + 1u;
    size_t scene_bytes = (size_t)(4u * n_inner + n_models + 2u * p.scene.n_materials + (n_models + 3u) / 4u) * 16u;
    const size_t max_smem = 227u * 1024u;
    // 4-wide fp32 records in shared memory when they and their 4-byte stacks fit.  A 4-wide visit parks at most three
    // siblings and descends two levels of the uploaded tree; inner records sit on levels 1, 3, 5, ... <= depth - 1.
    const uint32_t cap4 = 3u * (tree_depth / 2u) + 1u;
    const size_t scene4_bytes = scene_bytes + (size_t)3u * n_inner * 16u;
    const bool s4 = p.scene.nodes4_ch != nullptr && scene4_bytes + (size_t)THREADS * cap4 * sizeof(uint32_t) <= max_smem;
    const bool tight = s4 && p.scene.nodes4_tight != nullptr;
    // ... and the reference records next to the tight ones when that fits too (far rays then stay in shared memory)
    const size_t both_bytes = scene4_bytes + (size_t)7u * n_inner * 16u;
    const bool both = tight && !no_both && both_bytes + (size_t)THREADS * cap4 * sizeof(uint32_t) <= max_smem;
    if (s4) {
        // tuned on C2 (profiles/r01_tuning_sweeps.txt): the 4-wide walk wants later shading and immediate leaf tests
        if (tune.shade_wait_lanes == 0u) tune.shade_wait_lanes = 22u;   // 20 / 22 / 24 / 26 / 29 lanes: 47.4 / 47.3 / 47.4 / 47.6 / 47.8 ms
        if (tune.leaf_batch_lanes == 0u) tune.leaf_batch_lanes = 1u;
        const size_t smem4 = (both ? both_bytes : scene4_bytes) + (size_t)THREADS * cap4 * sizeof(uint32_t);
        auto k4 = both ? megakernel_v3<THREADS, 6> : (tight ? megakernel_v3<THREADS, 5> : megakernel_v3<THREADS, 4>);
        if (cudaFuncSetAttribute(k4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4) != cudaSuccess) return -1;
        const uint32_t tiles4 = ((p.cam.width + 7u) / 8u) * ((p.shard.rows + 3u) / 4u);
        uint32_t grid4 = (uint32_t)sm_count;
        const uint32_t useful4 = (tiles4 * 32u + THREADS - 1u) / THREADS;
        if (grid4 > useful4) grid4 = useful4;
        if (grid4 == 0) return 0;
        k4<<<grid4, THREADS, smem4, stream>>>(p, pixel_counter, n_inner, n_models, tune);
        return 1;
    }
    const bool smem_scene = scene_bytes + (size_t)THREADS * stack_cap * sizeof(uint2) <= max_smem;
    const bool q16 = !smem_scene && p.scene.pairs_q != nullptr;   // quantised records qualify (bvr_api.cu)
    const bool w4 = q16 && p.scene.nodes4_q != nullptr;
    // scenes walked in HBM/L2 are latency bound: long rays, so lanes must not wait long for shading (C4: 26 -> 8
    // lanes is 622 -> 414 ms, profiles/r01_tuning_sweeps.txt)
    if (tune.shade_wait_lanes == 0u) tune.shade_wait_lanes = smem_scene ? 26u : 8u;
    if (tune.leaf_batch_lanes == 0u) tune.leaf_batch_lanes = smem_scene ? 4u : 1u;
    if (w4) stack_cap = 3u * (tree_depth / 2u) + 1u;              // up to three siblings parked per 4-wide level
    const size_t stack_bytes = (size_t)THREADS * stack_cap * (q16 ? sizeof(uint32_t) : sizeof(uint2));
    // MODE 3: what the stacks leave of shared memory holds the first records of the array (the top of the tree)
    if (w4 && stack_bytes <= max_smem) {
        const size_t room = (max_smem - stack_bytes) / 64u;
        tune.n_top = (uint32_t)std::min<size_t>(std::min<size_t>(room, n_inner), tune.n_top);
    } else {
        tune.n_top = 0u;
    }
    const size_t smem = (smem_scene ? scene_bytes : 0) + stack_bytes + (size_t)tune.n_top * 64u;
    if (smem > max_smem) return -1;
    auto kern = smem_scene ? megakernel_v3<THREADS, 0>
                           : (w4 ? (tune.lean_w4 ? megakernel_v3<THREADS, 7> : megakernel_v3<THREADS, 3>)
                                 : (q16 ? megakernel_v3<THREADS, 2> : megakernel_v3<THREADS, 1>));
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    int blocks_per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, THREADS, smem) != cudaSuccess || blocks_per_sm < 1)
        return -1;
    const uint32_t tiles = ((p.cam.width + 7u) / 8u) * ((p.shard.rows + 3u) / 4u);
    uint32_t grid = (uint32_t)(sm_count * blocks_per_sm);
    const uint32_t max_useful = (tiles * 32u + THREADS - 1u) / THREADS;
    if (grid > max_useful) grid = max_useful;
    if (grid == 0) return 0;
    kern<<<grid, THREADS, smem, stream>>>(p, pixel_counter, n_inner, n_models, tune);
    return 1;
}

}  // namespace

int launch_selfcheck(const RenderParams& p, unsigned long long* counters, cudaStream_t stream) {
    if (!p.selfcheck_log) return 0;
    selfcheck_kernel<<<296, 256, 0, stream>>>(p.scene, p.selfcheck_log, p.selfcheck_count, p.selfcheck_cap, counters);
    return 1;
}

int launch_megakernel_v3(const RenderParams& p, uint32_t n_inner, uint32_t n_models, uint32_t tree_depth,
                         unsigned int* pixel_counter, int threads, uint32_t shade_wait_lanes, uint32_t leaf_batch_lanes,
                         bool no_both, uint32_t max_top, uint32_t n_hot, bool lean_w4, const uint32_t* tile_order,
                         uint32_t* tile_cost, int sm_count, cudaStream_t stream) {
    if (p.cam.width > 0xffffu || p.shard.rows > 0xffffu) return -1;   // pixel packed as px | ly << 16
    Tuning t{shade_wait_lanes, leaf_batch_lanes, max_top, n_hot, lean_w4 ? 1u : 0u, tile_order, tile_cost};
    switch (threads) {
        case 256: return launch_v3<256>(p, n_inner, n_models, tree_depth, pixel_counter, t, no_both, sm_count, stream);
        case 512: return launch_v3<512>(p, n_inner, n_models, tree_depth, pixel_counter, t, no_both, sm_count, stream);
        case 768: return launch_v3<768>(p, n_inner, n_models, tree_depth, pixel_counter, t, no_both, sm_count, stream);
        case 896: return launch_v3<896>(p, n_inner, n_models, tree_depth, pixel_counter, t, no_both, sm_count, stream);
        case 1024: return launch_v3<1024>(p, n_inner, n_models, tree_depth, pixel_counter, t, no_both, sm_count, stream);
        default: return -1;
    }
}

}  // namespace bvr
