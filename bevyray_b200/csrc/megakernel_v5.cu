// megakernel_v5.cu — EXPERIMENT (BVR_MK_VARIANT=5): warp-specialised persistent kernel with a per-CTA ray pool.
//
// v3 keeps every path in one lane from its first camera ray to its last sample: a lane whose ray has finished idles
// until ~29 lanes of its warp wait for shading (ncu: 17.5 of 32 lanes inside the node visit).  Here the two halves
// of a path's life run in different warps of the CTA and meet in shared memory:
//   * traversal warps hold one ray per lane in registers and walk the tree (v3's MODE 5 walk: 4-wide tight records in
//     shared memory).  A lane whose ray finished SWAPS it against a ready ray from the pool — the finished path goes
//     into the very entry the new ray came from — so the warp stays full without ever waiting for shading;
//   * shader warps take 32 finished paths from the pool at a time and run v3's staged shading on them, all lanes
//     busy, write the next ray (or the next sample's camera ray, or the next pixel's) back into the entry and queue it
//     as ready.
// Pool entry = 64 bytes (origin|t, direction|model, throughput|rng, pixel|sample,bounce|first depth); the per-pixel
// accumulators live in HBM/L2 (one path per pixel is in flight, so the read-modify-write is race free).  Three
// multi-producer multi-consumer rings of entry ids (free, to-shade, ready) with a published-items counter each.
// A pixel's samples still run one after the other, so the image is bit-identical to v3's and the oracle's.

#include "kernels.cuh"

namespace bvr {

namespace {

enum LaneState5 : int { L5_IDLE = 0, L5_TRAVERSE = 1, L5_FINISHED = 2 };
enum PathState5 : int { P5_NONE = 0, P5_SHADE = 1, P5_NEW_PATH = 2, P5_RAY_READY = 3, P5_EMPTY = 4, P5_DEAD = 5 };
enum Kind5 : int { K5_NONE = 0, K5_MISS = 1, K5_METAL = 2, K5_GLASS = 3, K5_DIFFUSE = 4 };

#define V5_LEAF 0x400u
#define V5_NONE 0x800u
#define V5_REF_MASK 0x7ffu
#define V5_ENTRIES 512u
#define V5_RING 1024u
#define V5_KIND_HIT 0u
#define V5_KIND_NEW_PIXEL 1u

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

__device__ __forceinline__ bool box_cull(V3 inv, V3 ainv, V3 noi, float closest_t, float cx, float cy, float cz, float hx,
                                         float hy, float hz, float& entry) {
    const float tcx = __fmaf_rn(cx, inv.x, noi.x), tcy = __fmaf_rn(cy, inv.y, noi.y), tcz = __fmaf_rn(cz, inv.z, noi.z);
    const float lox = __fmaf_rn(-hx, ainv.x, tcx), loy = __fmaf_rn(-hy, ainv.y, tcy), loz = __fmaf_rn(-hz, ainv.z, tcz);
    const float hix = __fmaf_rn(hx, ainv.x, tcx), hiy = __fmaf_rn(hy, ainv.y, tcy), hiz = __fmaf_rn(hz, ainv.z, tcz);
    entry = fmaxf(fmaxf(lox, loy), fmaxf(loz, 0.0f));
    const float exit = fminf(fminf(hix, hiy), fminf(hiz, closest_t));
    return entry <= exit;
}

__device__ __forceinline__ uint32_t sort4_park(uint32_t k0, uint32_t k1, uint32_t k2, uint32_t k3, uint32_t& sp_addr,
                                               const uint32_t stride) {
    uint32_t t0;
    t0 = min(k0, k1); k1 = max(k0, k1); k0 = t0;
    t0 = min(k2, k3); k3 = max(k2, k3); k2 = t0;
    t0 = min(k0, k2); k2 = max(k0, k2); k0 = t0;
    t0 = min(k1, k3); k3 = max(k1, k3); k1 = t0;
    t0 = min(k1, k2); k2 = max(k1, k2); k1 = t0;
    if (k3 != 0xffffffffu) { sts32(sp_addr, k3); sp_addr += stride; }
    if (k2 != 0xffffffffu) { sts32(sp_addr, k2); sp_addr += stride; }
    if (k1 != 0xffffffffu) { sts32(sp_addr, k1); sp_addr += stride; }
    return k0;
}

// ---- rings of entry ids: many producers, many consumers, all of them whole warps ----
struct Ring {
    unsigned int count;   // items published and not yet claimed
    unsigned int head;    // next position to read
    unsigned int tail;    // next position to write
    unsigned int pad;
};
struct Control {
    Ring free_q, shade_q, ready_q;
    unsigned int live_paths;       // paths that exist (in a lane or in the pool)
    unsigned int no_more_pixels;   // the pixel queue ran dry
    unsigned int finished;
    unsigned int pad;
};

// Every lane with `has` pushes `id`.  Warp-converged.
__device__ __forceinline__ void ring_push(Ring* r, unsigned short* ring, bool has, uint32_t id, uint32_t lane) {
    const unsigned m = __ballot_sync(0xffffffffu, has);
    if (m == 0u) return;
    const uint32_t n = (uint32_t)__popc(m);
    unsigned int pos = 0;
    if (lane == 0u) pos = atomicAdd(&r->tail, n);
    pos = __shfl_sync(0xffffffffu, pos, 0);
    if (has) {
        const uint32_t k = (uint32_t)__popc(m & ((1u << lane) - 1u));
        reinterpret_cast<volatile unsigned short*>(ring)[(pos + k) & (V5_RING - 1u)] = (unsigned short)(id + 1u);
    }
    __threadfence_block();
    __syncwarp();
    if (lane == 0u) atomicAdd(&r->count, n);
}

// Claims up to `want` items; lane k < return value receives one in `id`.  Warp-converged.
__device__ __forceinline__ uint32_t ring_pop(Ring* r, unsigned short* ring, uint32_t want, uint32_t& id, uint32_t lane) {
    unsigned int pos = 0, take = 0;
    if (lane == 0u && want > 0u) {
        unsigned int c = *reinterpret_cast<volatile unsigned int*>(&r->count);
        while (c > 0u) {
            take = c < want ? c : want;
            const unsigned int old = atomicCAS(&r->count, c, c - take);
            if (old == c) break;
            c = old;
            take = 0;
        }
        if (take) pos = atomicAdd(&r->head, take);
    }
    take = __shfl_sync(0xffffffffu, take, 0);
    pos = __shfl_sync(0xffffffffu, pos, 0);
    if (lane < take) {
        volatile unsigned short* slot = reinterpret_cast<volatile unsigned short*>(ring) + ((pos + lane) & (V5_RING - 1u));
        unsigned short v;
        while ((v = *slot) == 0) { }          // its producer is between the reservation and the store
        *slot = 0;
        id = (uint32_t)v - 1u;
    }
    return take;
}

// A flag another warp may change under our feet, read ONCE per warp: lanes of a warp are not guaranteed to execute a
// plain load together (independent thread scheduling), and a loop exit that some lanes take and others do not ends in
// a collective executed by part of the warp (cuda-gdb: "Warp Illegal Instruction" at the final reduction).
__device__ __forceinline__ unsigned int warp_read(volatile unsigned int* ptr, uint32_t lane) {
    unsigned int v = 0;
    if (lane == 0u) v = *ptr;
    return __shfl_sync(0xffffffffu, v, 0);
}

struct Tuning5 {
    uint32_t shader_warps;     // warps that only shade
    uint32_t swap_lanes;       // traversal warps exchange finished rays when this many lanes need it
    uint32_t extra_paths;      // paths kept in the pool on top of one per traversal lane
    uint32_t min_batch;        // shader warps wait for this many finished paths (unless traversal is starving)
};

template <int THREADS>
__global__ void __launch_bounds__(THREADS) megakernel_v5(const RenderParams p, unsigned int* __restrict__ pixel_counter,
                                                         const uint32_t n_inner, const uint32_t n_models,
                                                         const Tuning5 tune, float4* __restrict__ px_acc,
                                                         const uint32_t stack_cap) {
    extern __shared__ float4 smem[];
    const CameraParams& cam = p.cam;
    const unsigned full = 0xffffffffu;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    // ---- stage the scene (tight 4-wide records) ----
    SceneView sv = p.scene;
    float4* sm_cursor = smem;
    float4* sm_nodes = sm_cursor;     sm_cursor += 7u * n_inner;
    float4* sm_spheres = sm_cursor;   sm_cursor += n_models;
    float4* sm_materials = sm_cursor; sm_cursor += 2u * sv.n_materials;
    uint32_t* sm_matid = reinterpret_cast<uint32_t*>(sm_cursor);
    sm_cursor += (n_models + 3u) / 4u;
    for (uint32_t i = tid; i < 7u * n_inner; i += THREADS) sm_nodes[i] = p.scene.nodes4_tight[i];
    for (uint32_t i = tid; i < n_models; i += THREADS) sm_spheres[i] = p.scene.spheres[i];
    for (uint32_t i = tid; i < 2u * sv.n_materials; i += THREADS) sm_materials[i] = p.scene.materials[i];
    for (uint32_t i = tid; i < n_models; i += THREADS) sm_matid[i] = p.scene.sphere_material[i];
    sv.spheres = sm_spheres;
    sv.materials = sm_materials;
    sv.sphere_material = sm_matid;
    const uint32_t s_nodes = smem_addr(sm_nodes);
    // ---- stacks ----
    constexpr uint32_t STACK_STRIDE = THREADS * 4u;
    const uint32_t s_stack0 = smem_addr(sm_cursor) + tid * 4u;
    sm_cursor += (size_t)THREADS * stack_cap / 4u;
    // ---- pool: group g of entry e at pool + (g * ENTRIES + e) * 16 ----
    const uint32_t s_pool = smem_addr(sm_cursor);
    sm_cursor += 4u * V5_ENTRIES;
    constexpr uint32_t GROUP = V5_ENTRIES * 16u;
    unsigned short* ring_free = reinterpret_cast<unsigned short*>(sm_cursor);
    unsigned short* ring_shade = ring_free + V5_RING;
    unsigned short* ring_ready = ring_shade + V5_RING;
    Control* ctl = reinterpret_cast<Control*>(ring_ready + V5_RING);
    for (uint32_t i = tid; i < V5_RING; i += THREADS) {
        ring_free[i] = i < V5_ENTRIES ? (unsigned short)(i + 1u) : 0;
        ring_shade[i] = 0;
        ring_ready[i] = 0;
    }
    if (tid == 0u) {
        ctl->free_q = Ring{V5_ENTRIES, 0u, V5_ENTRIES, 0u};
        ctl->shade_q = Ring{0u, 0u, 0u, 0u};
        ctl->ready_q = Ring{0u, 0u, 0u, 0u};
        ctl->live_paths = 0u; ctl->no_more_pixels = 0u; ctl->finished = 0u; ctl->pad = 0u;
    }
    __syncthreads();

    const uint32_t root = sv.has_scene ? ((sv.root_ref & BVR_LEAF_BIT) ? (V5_LEAF | (sv.root_ref & 0x3ffu)) : sv.root_ref) : V5_NONE;
    const uint32_t n_groups = __float_as_uint(__ldg(&sv.tight_groups[0]).x);
    const uint32_t tiles_x = (cam.width + 7u) / 8u, tiles_y = (p.shard.rows + 3u) / 4u;
    const uint32_t total_slots = tiles_x * tiles_y * 32u;
    const uint32_t n_warps = THREADS / 32u;
    const uint32_t target_paths = (n_warps - tune.shader_warps) * 32u + tune.extra_paths;
    volatile unsigned int* v_finished = &ctl->finished;
    volatile unsigned int* v_live = &ctl->live_paths;
    volatile unsigned int* v_nomore = &ctl->no_more_pixels;
    volatile unsigned int* v_ready = &ctl->ready_q.count;
    volatile unsigned int* v_shade = &ctl->shade_q.count;

    uint32_t rays = 0;

    // the ray a traversal lane holds, and the rest of its path (which travels with the ray)
    const bool dedicated_shader = warp < tune.shader_warps;
    int state = L5_IDLE;
    Ray ray{v3(0, 0, 0), v3(0, 0, 1)};
    V3 inv = v3(0, 0, 0), ainv = v3(0, 0, 0), noi = v3(0, 0, 0);
    float a = 1.0f;
    Hit closest{BVR_INF, 0xffffffffu};
    uint32_t cur = V5_NONE, pending = V5_NONE;
    uint32_t sp_addr = s_stack0;
    bool far_ray = false;
    float4 g2 = make_float4(1.f, 1.f, 1.f, 0.f);   // throughput, rng
    float4 g3 = make_float4(0.f, 0.f, 0.f, 0.f);   // pixel, sample|bounce, first depth, -
    uint32_t cooldown = 0;   // vote rounds to traverse before the next exchange attempt (after a fruitless one)
    uint32_t backoff = 64;   // ns to sleep when there is nothing to do (doubles up to 2 us)

    for (;;) {
        // ---- what does this warp do now?  Dedicated shader warps always shade; a traversal warp that holds no ray
        //      shades too when a whole batch waits or when there is nothing to traverse (a warp that only polls
        //      steals issue slots from the shaders it is waiting for: ncu on the first version) ----
        const unsigned m_fin = __ballot_sync(full, state == L5_FINISHED);
        const unsigned m_idle = __ballot_sync(full, state == L5_IDLE);
        const unsigned m_trav = __ballot_sync(full, state == L5_TRAVERSE);
        const bool empty = (m_trav | m_fin) == 0u;
        if (empty && warp_read(v_finished, lane)) break;
        const unsigned int waiting = warp_read(v_shade, lane);
        const unsigned int ready_now = warp_read(v_ready, lane);
        const bool shade_now = empty && (dedicated_shader || waiting >= 32u || (waiting > 0u && ready_now == 0u));
        if (shade_now) {
            // ============================== shading pass ==============================
            uint32_t eid = 0;
            const unsigned int nomore = warp_read(v_nomore, lane);
            uint32_t n = 0;
            // dedicated warps wait for a decent batch unless the traversal side is running out of rays
            if (!dedicated_shader || waiting >= tune.min_batch || (waiting > 0u && ready_now < 64u))
                n = ring_pop(&ctl->shade_q, ring_shade, 32u, eid, lane);
            // spawn new paths while the population is below target and pixels remain
            uint32_t m = 0;
            if (n < 32u && !nomore) {
                const unsigned int live = warp_read(v_live, lane);
                if (live < target_paths) {
                    uint32_t want = target_paths - live;
                    if (want > 32u - n) want = 32u - n;
                    uint32_t fid = 0;
                    m = ring_pop(&ctl->free_q, ring_free, want, fid, lane);
                    // lanes n .. n+m-1 take the fresh entries
                    const uint32_t src = lane - n;
                    const uint32_t got = __shfl_sync(full, fid, src & 31u);
                    if (lane >= n && lane < n + m) eid = got;
                    if (lane == 0u && m) atomicAdd(&ctl->live_paths, m);
                }
            }
            if (n + m == 0u) {
                if (lane == 0u && *v_nomore && *v_live == 0u) *v_finished = 1u;
                __nanosleep(backoff);
                if (backoff < 2048u) backoff *= 2u;
                continue;
            }
            backoff = 64;
            const bool mine = lane < n + m;
            const uint32_t eb = s_pool + eid * 16u;
            int ps = !mine ? P5_NONE : (lane < n ? P5_SHADE : P5_EMPTY);
            V3 so = v3(0, 0, 0), sd = v3(0, 0, 1), thr = v3(1, 1, 1);
            float ht = BVR_INF, first_depth = BVR_INF;
            uint32_t hmodel = 0xffffffffu, rng = 0u, pix = 0u, sidx = 0u, bounce = 0u;
            if (ps == P5_SHADE) {
                const float4 g0 = lds128(eb), g1 = lds128(eb + GROUP), g2 = lds128(eb + 2u * GROUP), g3 = lds128(eb + 3u * GROUP);
                so = v3(g0.x, g0.y, g0.z); ht = g0.w;
                sd = v3(g1.x, g1.y, g1.z); hmodel = __float_as_uint(g1.w);
                thr = v3(g2.x, g2.y, g2.z); rng = __float_as_uint(g2.w);
                pix = __float_as_uint(g3.x);
                sidx = __float_as_uint(g3.y) & 0xfffffu; bounce = __float_as_uint(g3.y) >> 20;
                first_depth = g3.z;
            }
            // --- A1: classification (raytrace.wgsl:193-201, 232-248) ---
            int kind = K5_NONE;
            uint32_t mid = 0;
            if (ps == P5_SHADE) {
                if (bounce == 0u) {
                    first_depth = ht;
                    if (sidx == 0u && (p.out_primary_id || p.out_primary_depth)) {
                        const size_t lpix = (size_t)(pix >> 16) * cam.width + (pix & 0xffffu);
                        if (p.out_primary_id) p.out_primary_id[lpix] = ht == BVR_INF ? 0xffffffffu : hmodel;
                        if (p.out_primary_depth) p.out_primary_depth[lpix] = ht;
                    }
                }
                if (ht == BVR_INF) {
                    kind = K5_MISS;
                } else {
                    mid = sv.sphere_material[hmodel];
                    if (mid >= sv.n_materials) mid = sv.n_materials - 1u;
                    const float metallic = sv.materials[2u * mid].w;
                    const float transmission = sv.materials[2u * mid + 1u].w;
                    if (rng_next_float(rng) < metallic) kind = K5_METAL;
                    else if (rng_next_float(rng) < transmission) kind = K5_GLASS;
                    else kind = K5_DIFFUSE;
                }
            }
            // --- A2: unit-ball samples in one rejection loop (random.wgsl:17-26) ---
            int need = kind == K5_DIFFUSE ? 2 : (kind == K5_METAL ? 1 : 0);
            V3 b1 = v3(0.0f, 0.0f, 0.0f), b2 = v3(0.0f, 0.0f, 0.0f);
            while (need > 0) {
                rng_next_int(rng); const float x = __uint2float_rn(rng);
                rng_next_int(rng); const float y = __uint2float_rn(rng);
                rng_next_int(rng); const float z = __uint2float_rn(rng);
                const float k = 4.6566128730773926e-10f;   // 2^-31: fma(x, 2^-31, -1) rounds like (2*x*2^-32) - 1
                const V3 c = v3(__fmaf_rn(x, k, -1.0f), __fmaf_rn(y, k, -1.0f), __fmaf_rn(z, k, -1.0f));
                if (vdot(c, c) <= 1.0f) {
                    if (need == 2) b1 = c; else b2 = c;
                    need--;
                }
            }
            // --- A3: hit record / background ---
            if (ps == P5_SHADE) {
                bool path_end = false;
                V3 sample_color = v3(0.0f, 0.0f, 0.0f);
                const float4 sph = kind == K5_MISS ? make_float4(0.f, 0.f, 0.f, 0.f) : sv.spheres[hmodel];
                const V3 position = vadd(so, vscale(ht, sd));
                const V3 nin = kind == K5_MISS ? sd : vsub(position, v3(sph.x, sph.y, sph.z));
                const V3 unit = vnormalize(nin);
                if (kind == K5_MISS) {
                    const float aa = fmul(0.5f, fadd(unit.y, 1.0f));
                    const float ia = fsub(1.0f, aa);
                    const V3 bg = v3(fadd(fmul(ia, 1.0f), fmul(aa, 0.5f)), fadd(fmul(ia, 1.0f), fmul(aa, 0.7f)),
                                     fadd(fmul(ia, 1.0f), fmul(aa, 1.0f)));
                    const V3 lin = vmul(thr, bg);
                    sample_color = v3(fsqrt(lin.x), fsqrt(lin.y), fsqrt(lin.z));
                    path_end = true;
                } else {
                    const V3 normal = unit;
                    const float4 m0 = sv.materials[2u * mid], m1 = sv.materials[2u * mid + 1u];
                    V3 attenuation = v3(m0.x, m0.y, m0.z);
                    V3 dir;
                    bool absorbed;
                    if (kind == K5_DIFFUSE) {
                        dir = vadd(vadd(normal, b1), vscale(m1.x, b2));
                        if (vec3_near_zero(dir)) dir = normal;
                        absorbed = vdot(dir, normal) < 0.0f;
                    } else {
                        const V3 un = vnormalize(kind == K5_METAL ? reflect3(sd, normal) : sd);
                        if (kind == K5_METAL) {
                            dir = vadd(un, vscale(m1.x, b2));
                            absorbed = vdot(dir, normal) < 0.0f;
                        } else {
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
                        path_end = true;
                    } else {
                        thr = vmul(thr, attenuation);
                        bounce++;
                        if (bounce > cam.bounce_count) path_end = true;
                    }
                }
                if (path_end) {
                    if (first_depth == BVR_INF) first_depth = cam.fallback_far;
                    const size_t lpix = (size_t)(pix >> 16) * cam.width + (pix & 0xffffu);
                    float4 acc = __ldcg(&px_acc[lpix]);                                // raytrace.wgsl:165-166
                    acc.x = fadd(acc.x, sample_color.x); acc.y = fadd(acc.y, sample_color.y);
                    acc.z = fadd(acc.z, sample_color.z); acc.w = fadd(acc.w, first_depth);
                    sidx++;
                    if (sidx >= cam.sample_count) {
                        // --- pixel store (average, fused composite raytrace.wgsl:104-120) ---
                        const uint32_t px = pix & 0xffffu, ly = pix >> 16;
                        const uint32_t gy = shard_global_row(p.shard, ly);
                        const float nn = (float)cam.sample_count;
                        float4 out = make_float4(fdiv(acc.x, nn), fdiv(acc.y, nn), fdiv(acc.z, nn), 1.0f);
                        const float depth_avg = fdiv(acc.w, nn);
                        if (cam.level == 1u || cam.level == 2u) {
                            const size_t gpix = (size_t)gy * cam.width + px;
                            if (raster_wins(cam, p.raster_depth[gpix], depth_avg)) out = p.raster_rgba[gpix];
                        }
                        if (p.out_rgba) p.out_rgba[lpix] = out;
                        if (p.out_rt_depth) p.out_rt_depth[lpix] = depth_avg;
                        if (p.out_srgb8) p.out_srgb8[lpix] = store_srgb8(out);
                        ps = P5_EMPTY;
                    } else {
                        __stcg(&px_acc[lpix], acc);
                        ps = P5_NEW_PATH;
                    }
                } else {
                    ps = P5_RAY_READY;
                }
            }
            // --- A5: pixels from the tile-ordered queue (warp-convergent) ---
            for (;;) {
                const unsigned need_px = __ballot_sync(full, ps == P5_EMPTY);
                if (need_px == 0u) break;
                const int leader = __ffs(need_px) - 1;
                unsigned base = 0;
                if ((int)lane == leader) base = atomicAdd(pixel_counter, (unsigned)__popc(need_px));
                base = __shfl_sync(full, base, leader);
                if (ps == P5_EMPTY) {
                    const uint32_t slot = base + (uint32_t)__popc(need_px & ((1u << lane) - 1u));
                    if (slot >= total_slots) {
                        ps = P5_DEAD;
                    } else {
                        const uint32_t tile = slot >> 5, within = slot & 31u;
                        const uint32_t px = (tile % tiles_x) * 8u + (within & 7u);
                        const uint32_t ly = (tile / tiles_x) * 4u + (within >> 3);
                        const uint32_t gy = shard_global_row(p.shard, ly);
                        if (px < cam.width && ly < p.shard.rows && gy < cam.height) {
                            pix = px | (ly << 16);
                            rng = pixel_seed(cam, pixel_u(cam, px), pixel_v(cam, gy));
                            sidx = 0u;
                            __stcg(&px_acc[(size_t)ly * cam.width + px], make_float4(0.0f, 0.0f, 0.0f, 0.0f));
                            ps = P5_NEW_PATH;
                        }
                    }
                }
            }
            // --- A6: camera rays (raytrace.wgsl:139-156) ---
            if (ps == P5_NEW_PATH) {
                const uint32_t px = pix & 0xffffu, ly = pix >> 16;
                const Ray r = random_ray_from_uv(cam, pixel_u(cam, px), pixel_v(cam, shard_global_row(p.shard, ly)), rng);
                so = r.o; sd = r.d;
                thr = v3(1.0f, 1.0f, 1.0f);
                bounce = 0u;
                first_depth = BVR_INF;
                ps = P5_RAY_READY;
            }
            // --- write back, hand over ---
            if (ps == P5_RAY_READY) {
                sts128(eb, make_float4(so.x, so.y, so.z, BVR_INF));
                sts128(eb + GROUP, make_float4(sd.x, sd.y, sd.z, __uint_as_float(0xffffffffu)));
                sts128(eb + 2u * GROUP, make_float4(thr.x, thr.y, thr.z, __uint_as_float(rng)));
                sts128(eb + 3u * GROUP, make_float4(__uint_as_float(pix), __uint_as_float(sidx | (bounce << 20)), first_depth, 0.0f));
            }
            ring_push(&ctl->ready_q, ring_ready, ps == P5_RAY_READY, eid, lane);
            const unsigned dead = __ballot_sync(full, ps == P5_DEAD);
            if (dead) {
                ring_push(&ctl->free_q, ring_free, ps == P5_DEAD, eid, lane);
                if (lane == 0u) { atomicSub(&ctl->live_paths, (unsigned)__popc(dead)); *v_nomore = 1u; }
            }
            continue;
        }
        if (dedicated_shader) {            // nothing to shade yet
            __nanosleep(backoff);
            if (backoff < 2048u) backoff *= 2u;
            continue;
        }
        {
            // ============================== traversal ==============================
            // ---- exchange: finished lanes swap against ready rays; idle lanes pull; leftovers deposit ----
            const uint32_t n_need = (uint32_t)__popc(m_fin | m_idle);
            uint32_t rid = 0;
            uint32_t got = ring_pop(&ctl->ready_q, ring_ready, n_need, rid, lane);
            // hand the claimed ready entries to finished lanes first, then to idle lanes
            const unsigned takers = m_fin | m_idle;
            // rank of this lane among takers, finished lanes first
            const uint32_t my_rank = (state == L5_FINISHED) ? (uint32_t)__popc(m_fin & ((1u << lane) - 1u))
                                                            : (uint32_t)__popc(m_fin) + (uint32_t)__popc(m_idle & ((1u << lane) - 1u));
            const bool taker = ((takers >> lane) & 1u) != 0u;
            const uint32_t my_entry = __shfl_sync(full, rid, my_rank & 31u);
            const bool gets = taker && my_rank < got;
            // finished lanes that get no ready ray deposit into a free entry
            const unsigned m_dep = __ballot_sync(full, state == L5_FINISHED && !gets);
            uint32_t fid = 0;
            uint32_t got_free = 0;
            if (m_dep) got_free = ring_pop(&ctl->free_q, ring_free, (uint32_t)__popc(m_dep), fid, lane);
            const uint32_t dep_rank = (uint32_t)__popc(m_dep & ((1u << lane) - 1u));
            const uint32_t dep_entry = __shfl_sync(full, fid, dep_rank & 31u);
            const bool deposits = state == L5_FINISHED && !gets && dep_rank < got_free;

            uint32_t push_shade_id = 0, push_free_id = 0;
            bool push_shade = false, push_free = false;
            if (gets) {
                const uint32_t eb = s_pool + my_entry * 16u;
                const float4 n0 = lds128(eb), n1 = lds128(eb + GROUP), n2 = lds128(eb + 2u * GROUP), n3 = lds128(eb + 3u * GROUP);
                if (state == L5_FINISHED) {
                    // swap: my finished path goes into the entry the new ray came from
                    sts128(eb, make_float4(ray.o.x, ray.o.y, ray.o.z, closest.t));
                    sts128(eb + GROUP, make_float4(ray.d.x, ray.d.y, ray.d.z, __uint_as_float(closest.model)));
                    sts128(eb + 2u * GROUP, g2);
                    sts128(eb + 3u * GROUP, g3);
                    push_shade = true; push_shade_id = my_entry;
                } else {
                    push_free = true; push_free_id = my_entry;
                }
                ray.o = v3(n0.x, n0.y, n0.z);
                ray.d = v3(n1.x, n1.y, n1.z);
                g2 = n2; g3 = n3;
                // ray setup (v3 A7)
                inv = v3(rcp_approx(ray.d.x), rcp_approx(ray.d.y), rcp_approx(ray.d.z));
                ainv = v3(fabsf(inv.x), fabsf(inv.y), fabsf(inv.z));
                noi = v3(-(ray.o.x * inv.x), -(ray.o.y * inv.y), -(ray.o.z * inv.z));
                far_ray = false;
                for (uint32_t g = 0; g < n_groups; g++) {
                    const float4 gr = __ldg(&sv.tight_groups[1u + g]);
                    const float dx = ray.o.x - gr.x, dy = ray.o.y - gr.y, dz = ray.o.z - gr.z;
                    far_ray = far_ray || !(dx * dx + dy * dy + dz * dz <= gr.w);
                }
                a = vdot(ray.d, ray.d);
                closest.t = BVR_INF;
                closest.model = 0xffffffffu;
                sp_addr = s_stack0;
                pending = V5_NONE;
                cur = root;
                rays++;
                state = L5_TRAVERSE;
            } else if (deposits) {
                const uint32_t eb = s_pool + dep_entry * 16u;
                sts128(eb, make_float4(ray.o.x, ray.o.y, ray.o.z, closest.t));
                sts128(eb + GROUP, make_float4(ray.d.x, ray.d.y, ray.d.z, __uint_as_float(closest.model)));
                sts128(eb + 2u * GROUP, g2);
                sts128(eb + 3u * GROUP, g3);
                push_shade = true; push_shade_id = dep_entry;
                state = L5_IDLE;
            }
            ring_push(&ctl->shade_q, ring_shade, push_shade, push_shade_id, lane);
            ring_push(&ctl->free_q, ring_free, push_free, push_free_id, lane);

            const unsigned trav_now = __ballot_sync(full, state == L5_TRAVERSE);
            if (trav_now == 0u) {
                __nanosleep(backoff);
                if (backoff < 2048u) backoff *= 2u;
                continue;
            }
            backoff = 64;
            // nothing moved although lanes wanted to: the pool had no ray / no room for them; do not come back at once
            cooldown = (__ballot_sync(full, gets || deposits) == 0u && n_need > 0u) ? 4u : 0u;

            // ---- traversal (v3 MODE 5) until enough lanes want an exchange ----
            for (;;) {
                bool blocked = false;
#pragma unroll
                for (int rep = 0; rep < 2; rep++) {
                    if (state == L5_TRAVERSE) {
                        uint32_t c = cur;
                        if (c < V5_LEAF) {
                            const uint32_t na = s_nodes + c * 112u;
                            float4 q0, q1, q2, q3, q4, q5, rr;
                            if (far_ray) {
                                const float4* nd = sv.nodes4_ch + 7u * c;
                                q0 = __ldg(nd); q1 = __ldg(nd + 1); q2 = __ldg(nd + 2); q3 = __ldg(nd + 3);
                                q4 = __ldg(nd + 4); q5 = __ldg(nd + 5); rr = __ldg(nd + 6);
                            } else {
                                q0 = lds128(na); q1 = lds128(na + 16u); q2 = lds128(na + 32u);
                                q3 = lds128(na + 48u); q4 = lds128(na + 64u); q5 = lds128(na + 80u);
                                rr = lds128(na + 96u);
                            }
                            float e;
                            uint32_t k0 = box_cull(inv, ainv, noi, closest.t, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, e)
                                              ? ((__float_as_uint(e) & ~V5_REF_MASK) | __float_as_uint(rr.x)) : 0xffffffffu;
                            uint32_t k1 = box_cull(inv, ainv, noi, closest.t, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, e)
                                              ? ((__float_as_uint(e) & ~V5_REF_MASK) | __float_as_uint(rr.y)) : 0xffffffffu;
                            uint32_t k2 = box_cull(inv, ainv, noi, closest.t, q3.x, q3.y, q3.z, q3.w, q4.x, q4.y, e)
                                              ? ((__float_as_uint(e) & ~V5_REF_MASK) | __float_as_uint(rr.z)) : 0xffffffffu;
                            uint32_t k3 = box_cull(inv, ainv, noi, closest.t, q4.z, q4.w, q5.x, q5.y, q5.z, q5.w, e)
                                              ? ((__float_as_uint(e) & ~V5_REF_MASK) | __float_as_uint(rr.w)) : 0xffffffffu;
                            k0 = sort4_park(k0, k1, k2, k3, sp_addr, STACK_STRIDE);
                            c = k0 != 0xffffffffu ? (k0 & V5_REF_MASK) : V5_NONE;
                        }
                        if (c & V5_LEAF) {
                            if (pending == V5_NONE) { pending = c; c = V5_NONE; }
                            else blocked = true;
                        }
                        if (c == V5_NONE) {
                            while (sp_addr != s_stack0) {
                                sp_addr -= STACK_STRIDE;
                                const uint32_t e = lds32(sp_addr);
                                if (__uint_as_float(e & ~V5_REF_MASK) < closest.t) { c = e & V5_REF_MASK; break; }
                            }
                            if (c == V5_NONE) {
                                if (pending == V5_NONE) state = L5_FINISHED;
                                else blocked = true;
                            }
                        }
                        cur = c;
                    }
                }
                const unsigned blk = __ballot_sync(full, blocked);
                const unsigned trav = __ballot_sync(full, state == L5_TRAVERSE);
                if (trav == 0u) break;
                if (blk) {
                    if (state == L5_TRAVERSE && pending != V5_NONE) {
                        test_leaf(sv, ray, a, pending & 0x3ffu, closest);
                        pending = V5_NONE;
                    }
                }
                if (cooldown) { cooldown--; continue; }
                if (32u - (uint32_t)__popc(trav) >= tune.swap_lanes) break;
            }
        }
    }

    unsigned long long sum = rays;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(full, sum, o);
    if (lane == 0u && p.ray_counter && sum) atomicAdd(p.ray_counter, sum);
}

}  // namespace

// Returns -1 when the scene does not qualify (tight 4-wide records in shared memory): the caller uses v3.
int launch_megakernel_v5(const RenderParams& p, uint32_t n_inner, uint32_t n_models, uint32_t tree_depth,
                         unsigned int* pixel_counter, float4* px_acc, uint32_t shader_warps, uint32_t swap_lanes,
                         uint32_t extra_paths, uint32_t min_batch, int sm_count, cudaStream_t stream) {
    constexpr int THREADS = 1024;
    if (!p.scene.nodes4_tight || !p.scene.tight_groups || !p.scene.nodes4_ch) return -1;
    if (p.cam.width > 0xffffu || p.shard.rows > 0xffffu || p.cam.sample_count >= (1u << 20) || p.cam.bounce_count >= 4000u) return -1;
    if (p.cam.sample_count == 0u) return -1;
    const uint32_t cap4 = 3u * ((tree_depth + 1u) / 2u) + 2u;
    const uint32_t stack_cap = cap4;                 // THREADS * 4 bytes per entry: the pool stays 16-byte aligned
    const size_t scene_bytes = (size_t)(7u * n_inner + n_models + 2u * p.scene.n_materials + (n_models + 3u) / 4u) * 16u;
    const size_t smem = scene_bytes + (size_t)THREADS * stack_cap * 4u + (size_t)V5_ENTRIES * 64u + 3u * V5_RING * 2u + sizeof(Control);
    if (smem > 227u * 1024u) return -1;
    if (shader_warps < 1u || shader_warps > 24u) return -1;
    auto kern = megakernel_v5<THREADS>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    Tuning5 t{shader_warps, swap_lanes, extra_paths > V5_ENTRIES - 64u ? V5_ENTRIES - 64u : extra_paths, min_batch};
    kern<<<sm_count, THREADS, smem, stream>>>(p, pixel_counter, n_inner, n_models, t, px_acc, stack_cap);
    return 1;
}

}  // namespace bvr
