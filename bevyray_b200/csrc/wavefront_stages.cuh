// wavefront_stages.cuh — the stages of the wavefront pipeline as device functions, shared by
//   * wavefront.cu      one kernel launch per stage per wave, queues + counters in HBM, slot == pixel;
//   * cta_wavefront.cu  one persistent CTA runs every stage on its own pool of path slots, separated by
//                       __syncthreads, counters in shared memory, finished pixels replaced from a global queue.
//
// A "group" is the set of threads that cooperatively runs a stage (whole grid, or one CTA).  Every stage is
// written so that all lanes of a participating warp execute the queue pushes together.
//
// Stage map (reference lines):
//   regen      next camera ray of a pixel (raytrace.wgsl:139-156) or average + composite + store
//              (raytrace.wgsl:166-171, 104-120)
//   extend     raycast() (raytrace.wgsl:313-346): persistent lanes pull rays, postponed sphere tests
//   classify   material-selection draws (raytrace.wgsl:234, 248) -> miss / metal / glass / diffuse queue
//   shade_*    background (364-369) or scatter (231-299) for one kind, throughput, termination
#pragma once

#include "kernels.cuh"

namespace bvr {

enum WfCounter : int {
    WC_RAY0 = 0, WC_RAY1 = 1, WC_MISS = 2, WC_METAL = 3, WC_GLASS = 4, WC_DIFFUSE = 5, WC_REGEN = 6, WC_HEAD = 7,
    WC_COUNT = 8
};

struct WfGroup {
    uint32_t tid;        // index of this thread in the group
    uint32_t nthreads;   // group size (multiple of 32)
};

// warp-aggregated append: lanes with `pred` get consecutive slots of queue `q` (all lanes of the warp call it)
__device__ __forceinline__ void wf_push(uint32_t* __restrict__ q, unsigned int* counter, bool pred, uint32_t value) {
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m == 0u) return;
    const uint32_t lane = threadIdx.x & 31u;
    const int leader = __ffs(m) - 1;
    unsigned base = 0;
    if ((int)lane == leader) base = atomicAdd(counter, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pred) q[base + (uint32_t)__popc(m & ((1u << lane) - 1u))] = value;
}

__device__ __forceinline__ void wf_pixel_coords(const WavefrontParams& w, uint32_t pixel, uint32_t& px, uint32_t& ly,
                                                uint32_t& gy) {
    px = pixel % w.r.cam.width;
    ly = pixel / w.r.cam.width;
    gy = shard_global_row(w.r.shard, ly);
}

// pixel index (shard-local, row-major) of entry `i` of the 8x4-tile-ordered pixel queue, or 0xffffffff
__device__ __forceinline__ uint32_t wf_tile_order_pixel(const WavefrontParams& w, uint32_t i) {
    const CameraParams& cam = w.r.cam;
    const uint32_t tiles_x = (cam.width + 7u) / 8u;
    const uint32_t tile = i >> 5, within = i & 31u;
    const uint32_t px = (tile % tiles_x) * 8u + (within & 7u);
    const uint32_t ly = (tile / tiles_x) * 4u + (within >> 3);
    const uint32_t gy = shard_global_row(w.r.shard, ly);
    if (px < cam.width && ly < w.r.shard.rows && gy < cam.height) return ly * cam.width + px;
    return 0xffffffffu;
}
__device__ __forceinline__ uint32_t wf_tile_order_count(const WavefrontParams& w) {
    return ((w.r.cam.width + 7u) / 8u) * ((w.r.shard.rows + 3u) / 4u) * 32u;
}

// fresh pixel in a slot: per-pixel seed (raytrace.wgsl:95), zero accumulators
__device__ __forceinline__ void wf_init_slot(const WavefrontParams& w, uint32_t slot, uint32_t pixel) {
    const CameraParams& cam = w.r.cam;
    uint32_t px, ly, gy;
    wf_pixel_coords(w, pixel, px, ly, gy);
    const float u = pixel_u(cam, px), v = pixel_v(cam, gy);
    w.slot_pixel[slot] = pixel;
    w.thr_rng[slot] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(pixel_seed(cam, u, v)));
    w.accum[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    w.misc[slot] = make_uint4(0u, 0u, __float_as_uint(BVR_INF), 0u);
    if (w.r.out_primary_id) w.r.out_primary_id[pixel] = 0xffffffffu;
    if (w.r.out_primary_depth) w.r.out_primary_depth[pixel] = BVR_INF;
}

// ---------------------------------------------------------------------------------------------------
// regen.  REFILL: a finished (or empty) slot takes the next pixel of the global tile-ordered queue.
// ---------------------------------------------------------------------------------------------------
template <bool REFILL>
__device__ __forceinline__ void wf_stage_regen(const WavefrontParams& w, WfGroup g, const uint32_t* __restrict__ q_regen,
                                               uint32_t n, uint32_t* __restrict__ q_ray_out, unsigned int* cnt_ray_out,
                                               unsigned int* pixel_queue_head) {
    const CameraParams& cam = w.r.cam;
    for (uint32_t base = 0; base < n; base += g.nthreads) {
        const uint32_t i = base + g.tid;
        bool push = false, want_pixel = false;
        uint32_t slot = 0;
        uint4 misc = make_uint4(0u, 0u, 0u, 0u);
        if (i < n) {
            slot = q_regen[i];
            misc = w.misc[slot];
            const uint32_t pixel = w.slot_pixel[slot];
            if (pixel == 0xffffffffu) {
                want_pixel = REFILL;
            } else if (misc.x >= cam.sample_count) {
                // trace_multisampled's average (raytrace.wgsl:169-171) + fragment's composite (104-120)
                uint32_t px, ly, gy;
                wf_pixel_coords(w, pixel, px, ly, gy);
                const float4 acc = w.accum[slot];
                const float nn = (float)cam.sample_count;
                float4 out = make_float4(fdiv(acc.x, nn), fdiv(acc.y, nn), fdiv(acc.z, nn), 1.0f);
                const float depth_avg = fdiv(acc.w, nn);
                if (cam.level == 1u || cam.level == 2u) {
                    const size_t gpix = (size_t)gy * cam.width + px;
                    if (raster_wins(cam, w.r.raster_depth[gpix], depth_avg)) out = w.r.raster_rgba[gpix];
                }
                if (w.r.out_rgba) w.r.out_rgba[pixel] = out;
                if (w.r.out_rt_depth) w.r.out_rt_depth[pixel] = depth_avg;
                if (w.r.out_srgb8) w.r.out_srgb8[pixel] = store_srgb8(out);
                want_pixel = REFILL;
            } else {
                push = true;
            }
        }
        if (REFILL) {
            // warp-aggregated pull from the global pixel queue; invalid (padding) entries are skipped
            for (;;) {
                const unsigned m = __ballot_sync(0xffffffffu, want_pixel);
                if (m == 0u) break;
                const uint32_t lane = threadIdx.x & 31u;
                const int leader = __ffs(m) - 1;
                unsigned qb = 0;
                if ((int)lane == leader) qb = atomicAdd(pixel_queue_head, (unsigned)__popc(m));
                qb = __shfl_sync(0xffffffffu, qb, leader);
                if (want_pixel) {
                    const uint32_t qi = qb + (uint32_t)__popc(m & ((1u << lane) - 1u));
                    if (qi >= wf_tile_order_count(w)) {
                        w.slot_pixel[slot] = 0xffffffffu;     // no pixels left: the slot retires
                        want_pixel = false;
                    } else {
                        const uint32_t pixel = wf_tile_order_pixel(w, qi);
                        if (pixel != 0xffffffffu) {
                            wf_init_slot(w, slot, pixel);
                            misc = make_uint4(0u, 0u, __float_as_uint(BVR_INF), 0u);
                            if (cam.sample_count > 0u) {
                                want_pixel = false;
                                push = true;
                            } else {
                                // 0 samples: the reference divides 0 by 0; store and keep pulling pixels
                                const float nanv = fdiv(0.0f, 0.0f);
                                if (w.r.out_rgba) w.r.out_rgba[pixel] = make_float4(nanv, nanv, nanv, 1.0f);
                                if (w.r.out_rt_depth) w.r.out_rt_depth[pixel] = nanv;
                                w.slot_pixel[slot] = 0xffffffffu;
                            }
                        }
                    }
                }
            }
        }
        if (push) {
            const uint32_t pixel = w.slot_pixel[slot];
            uint32_t px, ly, gy;
            wf_pixel_coords(w, pixel, px, ly, gy);
            float4 tr = w.thr_rng[slot];
            uint32_t rng = __float_as_uint(tr.w);
            const float u = pixel_u(cam, px), v = pixel_v(cam, gy);
            const Ray ray = random_ray_from_uv(cam, u, v, rng);
            w.ray_a[slot] = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.d.x);
            w.ray_b[slot] = make_float4(ray.d.y, ray.d.z, BVR_INF, __uint_as_float(0xffffffffu));
            w.thr_rng[slot] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(rng));
            misc.y = 0u;                               // bounce
            misc.z = __float_as_uint(BVR_INF);         // first_depth
            w.misc[slot] = misc;
        }
        wf_push(q_ray_out, cnt_ray_out, push, slot);
    }
}

// ---------------------------------------------------------------------------------------------------
// extend: persistent lanes; scene through `sv` (shared memory or global), stacks in shared memory.
// ---------------------------------------------------------------------------------------------------
#define WF_NONE 0x7fffffffu

__device__ __forceinline__ uint32_t wf_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float wf_rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float4 wf_lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 wf_lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void wf_sts64(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}

// culling-only slab test on (centre, half extent) boxes (see megakernel_v3.cu::box_cull)
__device__ __forceinline__ bool wf_box_cull(V3 inv, V3 noi, float closest_t, float cx, float cy, float cz, float hx,
                                            float hy, float hz, float& entry) {
    const float tcx = __fmaf_rn(cx, inv.x, noi.x), thx = hx * fabsf(inv.x);
    const float tcy = __fmaf_rn(cy, inv.y, noi.y), thy = hy * fabsf(inv.y);
    const float tcz = __fmaf_rn(cz, inv.z, noi.z), thz = hz * fabsf(inv.z);
    entry = fmaxf(fmaxf(tcx - thx, tcy - thy), fmaxf(tcz - thz, 0.0f));
    const float exit = fminf(fminf(tcx + thx, tcy + thy), fminf(tcz + thz, closest_t));
    return entry <= exit;
}

struct WfExtendTuning {
    uint32_t refill_idle_lanes;   // pull new rays once this many lanes of the warp are idle
    uint32_t leaf_blocked_lanes;  // run the batched sphere test once this many lanes are blocked on theirs
};

// STACK_STRIDE = bytes between consecutive stack entries of one lane (threads-per-CTA * 8).
// s_stack0 = this lane's first stack slot (32-bit shared address); s_pairs = shared address of the pair
// records when SMEM_SCENE.  `head` is the queue-head counter (shared or global).
template <uint32_t STACK_STRIDE, bool SMEM_SCENE>
__device__ __forceinline__ unsigned long long wf_stage_extend(const WavefrontParams& w, const SceneView& sv,
                                                              const uint32_t* __restrict__ q_ray_in, uint32_t n_rays,
                                                              unsigned int* head, uint32_t s_stack0, uint32_t s_pairs,
                                                              WfExtendTuning tune) {
    const unsigned full = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    bool active = false, exhausted = (n_rays == 0u);
    uint32_t slot = 0;
    Ray ray{v3(0, 0, 0), v3(0, 0, 1)};
    V3 inv = v3(0, 0, 0), noi = v3(0, 0, 0);
    float a = 1.0f;
    Hit closest{BVR_INF, 0xffffffffu};
    uint32_t cur = WF_NONE, pending = WF_NONE, sp_addr = s_stack0;
    float2 dyz = make_float2(0.0f, 0.0f);
    unsigned long long rays = 0;

    for (;;) {
        // ---- refill idle lanes (batched: one counter round trip per refill) ----
        if (!exhausted) {
            const unsigned idle = __ballot_sync(full, !active);
            const int leader = __ffs(idle) - 1;
            unsigned base = 0xffffffffu;
            if ((int)lane == leader) base = atomicAdd(head, (unsigned)__popc(idle));
            base = __shfl_sync(full, base, leader < 0 ? 0 : leader);
            if (!active && idle != 0u) {
                const uint32_t i = base + (uint32_t)__popc(idle & ((1u << lane) - 1u));
                if (i < n_rays) {
                    slot = q_ray_in[i];
                    const float4 ra = w.ray_a[slot], rb = w.ray_b[slot];
                    ray.o = v3(ra.x, ra.y, ra.z);
                    ray.d = v3(ra.w, rb.x, rb.y);
                    dyz = make_float2(rb.x, rb.y);
                    inv = v3(wf_rcp_approx(ray.d.x), wf_rcp_approx(ray.d.y), wf_rcp_approx(ray.d.z));
                    noi = v3(-(ray.o.x * inv.x), -(ray.o.y * inv.y), -(ray.o.z * inv.z));
                    a = vdot(ray.d, ray.d);
                    closest.t = BVR_INF;
                    closest.model = 0xffffffffu;
                    sp_addr = s_stack0;
                    pending = WF_NONE;
                    cur = sv.has_scene ? sv.root_ref : WF_NONE;
                    active = true;
                    rays++;
                }
            }
            if (idle != 0u && base + (uint32_t)__popc(idle) >= n_rays) exhausted = true;
        }
        if (!__any_sync(full, active)) break;

        // ---- traverse until enough lanes are idle ----
        for (;;) {
            bool blocked = false;
#pragma unroll
            for (int rep = 0; rep < 2; rep++) {
                if (active) {
                    uint32_t c = cur;
                    if (c < WF_NONE) {
                        float4 q0, q1, q2;
                        uint32_t r0, r1;
                        if (SMEM_SCENE) {
                            const uint32_t na = s_pairs + c * 64u;
                            q0 = wf_lds128(na); q1 = wf_lds128(na + 16u); q2 = wf_lds128(na + 32u);
                            const uint2 rr = wf_lds64(na + 48u);
                            r0 = rr.x; r1 = rr.y;
                        } else {
                            const float4* nd = sv.pairs_ch + 4u * c;
                            q0 = __ldg(nd); q1 = __ldg(nd + 1); q2 = __ldg(nd + 2);
                            const float4 q3 = __ldg(nd + 3);
                            r0 = __float_as_uint(q3.x); r1 = __float_as_uint(q3.y);
                        }
                        float d0, d1;
                        const bool h0 = wf_box_cull(inv, noi, closest.t, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, d0);
                        const bool h1 = wf_box_cull(inv, noi, closest.t, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, d1);
                        const bool first0 = d0 < d1;
                        if (h0 && h1) {
                            wf_sts64(sp_addr, first0 ? r1 : r0, __float_as_uint(first0 ? d1 : d0));
                            sp_addr += STACK_STRIDE;
                            c = first0 ? r0 : r1;
                        } else {
                            c = h0 ? r0 : (h1 ? r1 : WF_NONE);
                        }
                    }
                    if ((int)c < 0) {
                        if (pending == WF_NONE) { pending = c; c = WF_NONE; }
                        else blocked = true;
                    }
                    if (c == WF_NONE) {
                        if (sp_addr != s_stack0) {
                            sp_addr -= STACK_STRIDE;
                            const uint2 e = wf_lds64(sp_addr);
                            if (__uint_as_float(e.y) < closest.t) c = e.x;
                        } else if (pending == WF_NONE) {
                            // finished: hit record goes back to the path state; the lane is idle
                            w.ray_b[slot] = make_float4(dyz.x, dyz.y, closest.t, __uint_as_float(closest.model));
                            active = false;
                        } else {
                            blocked = true;
                        }
                    }
                    cur = c;
                }
            }
            const unsigned blk = __ballot_sync(full, blocked);
            const unsigned act = __ballot_sync(full, active);
            if (act == 0u) break;
            const uint32_t nblk = (uint32_t)__popc(blk), nact = (uint32_t)__popc(act);
            if (nblk >= tune.leaf_blocked_lanes || nblk == nact) {
                if (active && pending != WF_NONE) {
                    test_leaf(sv, ray, a, pending, closest);
                    pending = WF_NONE;
                }
            }
            if (!exhausted && 32u - nact >= tune.refill_idle_lanes) break;
        }
    }
    return rays;
}

// ---------------------------------------------------------------------------------------------------
// classify: material-selection draws of scatter() (raytrace.wgsl:232-248), one queue per kind
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void wf_stage_classify(const WavefrontParams& w, WfGroup g, const uint32_t* __restrict__ q_ray_in,
                                                  uint32_t n, uint32_t* q_miss, uint32_t* q_metal, uint32_t* q_glass,
                                                  uint32_t* q_diffuse, unsigned int* counters) {
    const SceneView& s = w.r.scene;
    for (uint32_t base = 0; base < n; base += g.nthreads) {
        const uint32_t i = base + g.tid;
        int kind = -1;
        uint32_t slot = 0;
        if (i < n) {
            slot = q_ray_in[i];
            const float4 rb = w.ray_b[slot];
            if (rb.z == BVR_INF) {
                kind = WC_MISS;
            } else {
                uint32_t mid = s.sphere_material[__float_as_uint(rb.w)];
                if (mid >= s.n_materials) mid = s.n_materials - 1u;
                const float metallic = s.materials[2u * mid].w, transmission = s.materials[2u * mid + 1u].w;
                float* rp = &w.thr_rng[slot].w;
                uint32_t rng = __float_as_uint(*rp);
                if (rng_next_float(rng) < metallic) kind = WC_METAL;
                else if (rng_next_float(rng) < transmission) kind = WC_GLASS;
                else kind = WC_DIFFUSE;
                *rp = __uint_as_float(rng);
            }
        }
        wf_push(q_miss, counters + WC_MISS, kind == WC_MISS, slot);
        wf_push(q_metal, counters + WC_METAL, kind == WC_METAL, slot);
        wf_push(q_glass, counters + WC_GLASS, kind == WC_GLASS, slot);
        wf_push(q_diffuse, counters + WC_DIFFUSE, kind == WC_DIFFUSE, slot);
    }
}

// ---------------------------------------------------------------------------------------------------
// shading
// ---------------------------------------------------------------------------------------------------
struct WfPath {
    uint32_t slot;
    Ray ray;
    Hit hit;
    V3 throughput;
    uint32_t rng;
    uint4 misc;   // (sample index, bounce, first_depth bits, -)
};

__device__ __forceinline__ WfPath wf_load_path(const WavefrontParams& w, uint32_t slot) {
    WfPath p;
    p.slot = slot;
    const float4 ra = w.ray_a[slot], rb = w.ray_b[slot], tr = w.thr_rng[slot];
    p.ray.o = v3(ra.x, ra.y, ra.z);
    p.ray.d = v3(ra.w, rb.x, rb.y);
    p.hit.t = rb.z;
    p.hit.model = __float_as_uint(rb.w);
    p.throughput = v3(tr.x, tr.y, tr.z);
    p.rng = __float_as_uint(tr.w);
    p.misc = w.misc[slot];
    return p;
}

// raytrace.wgsl:193-195: first_depth is the primary ray's distance; sample 0 also feeds the id/depth planes
__device__ __forceinline__ void wf_record_primary(const WavefrontParams& w, WfPath& p) {
    if (p.misc.y == 0u) {
        p.misc.z = __float_as_uint(p.hit.t);
        if (p.misc.x == 0u) {
            const uint32_t pixel = w.slot_pixel[p.slot];
            if (w.r.out_primary_id) w.r.out_primary_id[pixel] = p.hit.t == BVR_INF ? 0xffffffffu : p.hit.model;
            if (w.r.out_primary_depth) w.r.out_primary_depth[pixel] = p.hit.t;
        }
    }
}

// path ended with gamma-encoded sample colour `c` (raytrace.wgsl:219-223, 166-167)
__device__ __forceinline__ void wf_end_path(const WavefrontParams& w, WfPath& p, V3 c) {
    float first_depth = __uint_as_float(p.misc.z);
    if (first_depth == BVR_INF) first_depth = w.r.cam.fallback_far;
    float4 acc = w.accum[p.slot];
    acc.x = fadd(acc.x, c.x); acc.y = fadd(acc.y, c.y); acc.z = fadd(acc.z, c.z);
    acc.w = fadd(acc.w, first_depth);
    w.accum[p.slot] = acc;
    p.misc.x += 1u;
    w.misc[p.slot] = p.misc;
    w.thr_rng[p.slot] = make_float4(p.throughput.x, p.throughput.y, p.throughput.z, __uint_as_float(p.rng));
}

// path continues with the scattered ray; false when the bounce budget is exhausted (raytrace.wgsl:186, 214-216)
__device__ __forceinline__ bool wf_continue_path(const WavefrontParams& w, WfPath& p, V3 attenuation) {
    p.throughput = vmul(p.throughput, attenuation);
    p.misc.y += 1u;
    if (p.misc.y > w.r.cam.bounce_count) return false;
    w.ray_a[p.slot] = make_float4(p.ray.o.x, p.ray.o.y, p.ray.o.z, p.ray.d.x);
    w.ray_b[p.slot] = make_float4(p.ray.d.y, p.ray.d.z, BVR_INF, __uint_as_float(0xffffffffu));
    w.thr_rng[p.slot] = make_float4(p.throughput.x, p.throughput.y, p.throughput.z, __uint_as_float(p.rng));
    w.misc[p.slot] = p.misc;
    return true;
}

__device__ __forceinline__ void wf_stage_shade_miss(const WavefrontParams& w, WfGroup g, const uint32_t* __restrict__ q_miss,
                                                    uint32_t n, uint32_t* q_regen, unsigned int* cnt_regen) {
    for (uint32_t base = 0; base < n; base += g.nthreads) {
        const uint32_t i = base + g.tid;
        const bool valid = i < n;
        uint32_t slot = 0;
        if (valid) {
            slot = q_miss[i];
            WfPath p = wf_load_path(w, slot);
            wf_record_primary(w, p);
            const V3 lin = vmul(p.throughput, background_gradient(p.ray));
            wf_end_path(w, p, v3(fsqrt(lin.x), fsqrt(lin.y), fsqrt(lin.z)));
        }
        wf_push(q_regen, cnt_regen, valid, slot);
    }
}

// One material class; the class was chosen by the draws in wf_stage_classify, so scatter() is entered after them.
template <int KIND>
__device__ __forceinline__ void wf_stage_shade_hit(const WavefrontParams& w, const SceneView& s, WfGroup g,
                                                   const uint32_t* __restrict__ q, uint32_t n, uint32_t* q_ray_out,
                                                   unsigned int* cnt_ray_out, uint32_t* q_regen, unsigned int* cnt_regen) {
    for (uint32_t base = 0; base < n; base += g.nthreads) {
        const uint32_t i = base + g.tid;
        bool cont = false, ended = false;
        uint32_t slot = 0;
        if (i < n) {
            slot = q[i];
            WfPath p = wf_load_path(w, slot);
            wf_record_primary(w, p);
            // hit record, raytrace.wgsl:355-359
            const float4 sph = s.spheres[p.hit.model];
            const V3 position = vadd(p.ray.o, vscale(p.hit.t, p.ray.d));
            const V3 normal = vnormalize(vsub(position, v3(sph.x, sph.y, sph.z)));
            uint32_t mid = s.sphere_material[p.hit.model];
            if (mid >= s.n_materials) mid = s.n_materials - 1u;
            const float4 m0 = s.materials[2u * mid], m1 = s.materials[2u * mid + 1u];
            const V3 base_color = v3(m0.x, m0.y, m0.z);
            V3 attenuation;
            bool absorbed;
            if (KIND == WC_METAL) {                                  // raytrace.wgsl:234-246
                const V3 reflected = vadd(vnormalize(reflect3(p.ray.d, normal)), vscale(m1.x, random_unit_vec3(p.rng)));
                p.ray.o = position;
                p.ray.d = reflected;
                attenuation = base_color;
                absorbed = vdot(p.ray.d, normal) < 0.0f;
            } else if (KIND == WC_GLASS) {                           // raytrace.wgsl:248-282
                const bool front_face = vdot(p.ray.d, normal) < 0.0f;
                const float ri = front_face ? fdiv(1.0f, m1.z) : m1.z;
                const V3 unit_direction = vnormalize(p.ray.d);
                const float cos_theta = fminf(vdot(vneg(unit_direction), normal), 1.0f);
                const float sin_theta = fsqrt(fsub(1.0f, fmul(cos_theta, cos_theta)));
                const bool cannot_refract = fmul(ri, sin_theta) > 1.0f;
                V3 direction;
                if (cannot_refract || schlick_reflectance(cos_theta, ri) > rng_next_float(p.rng)) direction = reflect3(unit_direction, normal);
                else direction = refract3(unit_direction, normal, ri);
                p.ray.o = position;
                p.ray.d = direction;
                attenuation = v3(1.0f, 1.0f, 1.0f);
                absorbed = false;
            } else {                                                 // raytrace.wgsl:283-298
                const V3 b1 = random_unit_vec3(p.rng);
                const V3 b2 = random_unit_vec3(p.rng);
                V3 dir = vadd(vadd(normal, b1), vscale(m1.x, b2));
                if (vec3_near_zero(dir)) dir = normal;
                p.ray.o = position;
                p.ray.d = dir;
                attenuation = base_color;
                absorbed = vdot(p.ray.d, normal) < 0.0f;
            }
            if (absorbed) {
                wf_end_path(w, p, v3(0.0f, 0.0f, 0.0f));
                ended = true;
            } else if (wf_continue_path(w, p, attenuation)) {
                cont = true;
            } else {
                wf_end_path(w, p, v3(0.0f, 0.0f, 0.0f));
                ended = true;
            }
        }
        wf_push(q_ray_out, cnt_ray_out, cont, slot);
        wf_push(q_regen, cnt_regen, ended, slot);
    }
}

}  // namespace bvr
