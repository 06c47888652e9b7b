// wavefront.cu — the wavefront pipeline: raygen / extend / shade-by-material / compaction.
//
// The reference runs everything for a pixel inside one fragment invocation (raytrace.wgsl:93-123).  Here
// the same per-pixel sequence of operations is cut at every raycast() boundary and regrouped by KIND
// of work, so that a warp only ever runs one kind:
//
//   wf_init      seeds the per-pixel RNG (raytrace.wgsl:95) and queues every pixel for ray generation
//   wf_regen     next sample of a pixel: camera ray (raytrace.wgsl:139-156) or, after the last sample,
//                the average + fused depth composite + store (raytrace.wgsl:166-171, 104-120)
//   wf_extend    raycast() (raytrace.wgsl:313-346) for every queued ray: persistent warps pull rays from
//                the queue as lanes free up, then draw the material-selection randoms
//                (raytrace.wgsl:234,248) and append the pixel to the miss / metal / glass / diffuse queue
//   wf_shade_*   background (364-369) or scatter (231-299) of one material class, throughput update,
//                path termination; survivors go to the next wave's ray queue, ended paths to wf_regen
//
// Queues are compacted with warp-ballot + one atomicAdd per warp.  A pixel has exactly one path in
// flight at any time (its RNG stream is sequential, raytrace.wgsl:89,161-167), so path state is indexed
// by pixel slot and lives in HBM as SoA float4 arrays.  Arithmetic is the strict set of trace.cuh:
// results are bit-identical to the megakernel and to the oracle.

#include "kernels.cuh"

namespace bvr {

namespace {

constexpr int WF_THREADS = 256;

enum Counter : int {
    C_RAY0 = 0, C_RAY1 = 1,      // ray queue sizes (ping-pong)
    C_MISS = 2, C_METAL = 3, C_GLASS = 4, C_DIFFUSE = 5,
    C_REGEN = 6,
    C_HEAD = 7,                  // extend kernel's queue head
    C_COUNT = 8
};

// warp-aggregated append: lanes with `pred` get consecutive slots of queue `q`
__device__ __forceinline__ void queue_push(uint32_t* __restrict__ q, unsigned int* __restrict__ counter, bool pred,
                                           uint32_t value) {
    const unsigned active = __activemask();
    const unsigned m = __ballot_sync(active, pred);
    if (m == 0u) return;
    const uint32_t lane = threadIdx.x & 31u;
    const int leader = __ffs(m) - 1;
    unsigned base = 0;
    if ((int)lane == leader) base = atomicAdd(counter, (unsigned)__popc(m));
    base = __shfl_sync(active, base, leader);
    if (pred) q[base + (uint32_t)__popc(m & ((1u << lane) - 1u))] = value;
}

__device__ __forceinline__ void slot_to_pixel(const WavefrontParams& w, uint32_t slot, uint32_t& px, uint32_t& ly,
                                              uint32_t& gy) {
    px = slot % w.r.cam.width;
    ly = slot / w.r.cam.width;
    gy = shard_global_row(w.r.shard, ly);
}

// ---- init: per-pixel seed, zero accumulators, queue every valid pixel for wf_regen in 8x4-tile order ----
__global__ void __launch_bounds__(WF_THREADS) wf_init(const WavefrontParams w) {
    const CameraParams& cam = w.r.cam;
    const uint32_t tiles_x = (cam.width + 7u) / 8u, tiles_y = (w.r.shard.rows + 3u) / 4u;
    const uint32_t total = tiles_x * tiles_y * 32u;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t base = blockIdx.x * blockDim.x; base < total; base += stride) {
        const uint32_t i = base + threadIdx.x;
        bool valid = false;
        uint32_t slot = 0;
        if (i < total) {
            const uint32_t tile = i >> 5, within = i & 31u;
            const uint32_t px = (tile % tiles_x) * 8u + (within & 7u);
            const uint32_t ly = (tile / tiles_x) * 4u + (within >> 3);
            const uint32_t gy = shard_global_row(w.r.shard, ly);
            if (px < cam.width && ly < w.r.shard.rows && gy < cam.height) {
                valid = true;
                slot = ly * cam.width + px;
                const float u = pixel_u(cam, px), v = pixel_v(cam, gy);
                w.thr_rng[slot] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(pixel_seed(cam, u, v)));
                w.accum[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                w.misc[slot] = make_uint4(0u, 0u, __float_as_uint(BVR_INF), 0u);
                if (w.r.out_primary_id) w.r.out_primary_id[slot] = 0xffffffffu;
                if (w.r.out_primary_depth) w.r.out_primary_depth[slot] = BVR_INF;
            }
        }
        queue_push(w.q_regen, w.counters + C_REGEN, valid, slot);
    }
}

// ---- reset the counters the coming wave fills ----
__global__ void wf_reset(unsigned int* counters, int next_ray_counter) {
    if (threadIdx.x == 0) {
        counters[C_MISS] = 0u; counters[C_METAL] = 0u; counters[C_GLASS] = 0u; counters[C_DIFFUSE] = 0u;
        counters[C_REGEN] = 0u; counters[C_HEAD] = 0u;
        counters[next_ray_counter] = 0u;
    }
}

// ---- regen: next camera ray of a pixel, or finalize the pixel ----
__global__ void __launch_bounds__(WF_THREADS) wf_regen(const WavefrontParams w, uint32_t* __restrict__ q_ray_out,
                                                       int ray_counter_out) {
    const CameraParams& cam = w.r.cam;
    const uint32_t n = w.counters[C_REGEN];
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += stride) {
        const uint32_t i = base + threadIdx.x;
        bool push = false;
        uint32_t slot = 0;
        if (i < n) {
            slot = w.q_regen[i];
            uint4 misc = w.misc[slot];
            uint32_t px, ly, gy;
            slot_to_pixel(w, slot, px, ly, gy);
            if (misc.x >= cam.sample_count) {
                // trace_multisampled's average (raytrace.wgsl:169-171) + fragment's composite (104-120)
                const float4 acc = w.accum[slot];
                const float nn = (float)cam.sample_count;
                float4 out = make_float4(fdiv(acc.x, nn), fdiv(acc.y, nn), fdiv(acc.z, nn), 1.0f);
                const float depth_avg = fdiv(acc.w, nn);
                if (cam.level == 1u || cam.level == 2u) {
                    const size_t gpix = (size_t)gy * cam.width + px;
                    if (raster_wins(cam, w.r.raster_depth[gpix], depth_avg)) out = w.r.raster_rgba[gpix];
                }
                if (w.r.out_rgba) w.r.out_rgba[slot] = out;
                if (w.r.out_rt_depth) w.r.out_rt_depth[slot] = depth_avg;
                if (w.r.out_srgb8) w.r.out_srgb8[slot] = store_srgb8(out);
            } else {
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
                push = true;
            }
        }
        queue_push(q_ray_out, w.counters + ray_counter_out, push, slot);
    }
}

// ---- extend: persistent warps, one ray per lane, refilled from the queue as lanes finish ----
#define WF_NONE 0x7fffffffu

__device__ __forceinline__ float box_dst_fma(V3 inv, V3 noi, float mnx, float mny, float mnz, float mxx, float mxy,
                                             float mxz) {
    const float t0x = __fmaf_rn(mnx, inv.x, noi.x), t1x = __fmaf_rn(mxx, inv.x, noi.x);
    const float t0y = __fmaf_rn(mny, inv.y, noi.y), t1y = __fmaf_rn(mxy, inv.y, noi.y);
    const float t0z = __fmaf_rn(mnz, inv.z, noi.z), t1z = __fmaf_rn(mxz, inv.z, noi.z);
    const float t_near = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fminf(t0z, t1z));
    const float t_far = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
    const bool hit = (t_far >= t_near) && (t_far > 0.0f);
    return hit ? fmaxf(t_near, 0.0f) : BVR_INF;
}

template <bool SMEM_SCENE>
__global__ void __launch_bounds__(WF_THREADS) wf_extend(const WavefrontParams w, const uint32_t* __restrict__ q_ray_in,
                                                        int ray_counter_in, uint32_t stack_cap, uint32_t n_inner,
                                                        uint32_t n_models) {
    extern __shared__ float4 smem[];
    const unsigned full = 0xffffffffu;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint32_t n_rays = w.counters[ray_counter_in];
    if (n_rays == 0u) return;

    SceneView sv = w.r.scene;
    float4* sm_cursor = smem;
    if (SMEM_SCENE) {
        // only the traversal data (pairs + spheres); material parameters for the classification
        // draws are read through L1
        float4* sm_pairs = sm_cursor;   sm_cursor += 4u * n_inner;
        float4* sm_spheres = sm_cursor; sm_cursor += n_models;
        for (uint32_t i = tid; i < 4u * n_inner; i += WF_THREADS) sm_pairs[i] = w.r.scene.pairs[i];
        for (uint32_t i = tid; i < n_models; i += WF_THREADS) sm_spheres[i] = w.r.scene.spheres[i];
        sv.pairs = sm_pairs;
        sv.spheres = sm_spheres;
        __syncthreads();
    }
    uint2* const stack = reinterpret_cast<uint2*>(sm_cursor) + tid;

    bool active = false;
    uint32_t slot = 0;
    Ray ray{v3(0, 0, 0), v3(0, 0, 1)};
    V3 inv = v3(0, 0, 0), noi = v3(0, 0, 0);
    float a = 1.0f;
    Hit closest{BVR_INF, 0xffffffffu};
    uint32_t cur = WF_NONE;
    int sp = 0;
    bool exhausted = false;
    unsigned long long rays = 0;

    for (;;) {
        // refill idle lanes
        if (!exhausted) {
            const unsigned idle = __ballot_sync(full, !active);
            if (idle) {
                const int leader = __ffs(idle) - 1;
                unsigned base = 0;
                if ((int)lane == leader) base = atomicAdd(w.counters + C_HEAD, (unsigned)__popc(idle));
                base = __shfl_sync(full, base, leader);
                if (!active) {
                    const uint32_t i = base + (uint32_t)__popc(idle & ((1u << lane) - 1u));
                    if (i < n_rays) {
                        slot = q_ray_in[i];
                        const float4 ra = w.ray_a[slot], rb = w.ray_b[slot];
                        ray.o = v3(ra.x, ra.y, ra.z);
                        ray.d = v3(ra.w, rb.x, rb.y);
                        inv = v3(fdiv(1.0f, ray.d.x), fdiv(1.0f, ray.d.y), fdiv(1.0f, ray.d.z));
                        noi = v3(-fmul(ray.o.x, inv.x), -fmul(ray.o.y, inv.y), -fmul(ray.o.z, inv.z));
                        a = vdot(ray.d, ray.d);
                        closest.t = BVR_INF;
                        closest.model = 0xffffffffu;
                        sp = 0;
                        cur = sv.has_scene ? sv.root_ref : WF_NONE;
                        active = true;
                        rays++;
                    }
                }
                if (base + (uint32_t)__popc(idle) >= n_rays) exhausted = true;
            }
        }
        if (!__any_sync(full, active)) break;

        // traverse until enough lanes are idle again (or every lane, once the queue is exhausted)
        for (;;) {
            bool finished = false;
            if (active) {
                if (cur != WF_NONE) {
                    if (cur & BVR_LEAF_BIT) {
                        test_leaf(sv, ray, a, cur, closest);
                        cur = WF_NONE;
                    } else {
                        const float4* nd = sv.pairs + 4u * cur;
                        const float4 q0 = nd[0], q1 = nd[1], q2 = nd[2], q3 = nd[3];
                        const float d0 = box_dst_fma(inv, noi, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y);
                        const float d1 = box_dst_fma(inv, noi, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w);
                        const bool h0 = d0 < closest.t, h1 = d1 < closest.t;
                        const uint32_t r0 = __float_as_uint(q3.x), r1 = __float_as_uint(q3.y);
                        if (h0 && h1) {
                            const bool first0 = d0 < d1;
                            stack[(uint32_t)sp * WF_THREADS] = make_uint2(first0 ? r1 : r0, __float_as_uint(first0 ? d1 : d0));
                            sp++;
                            cur = first0 ? r0 : r1;
                        } else {
                            cur = h0 ? r0 : (h1 ? r1 : WF_NONE);
                        }
                    }
                }
                if (cur == WF_NONE) {
                    if (sp == 0) {
                        finished = true;
                    } else {
                        --sp;
                        const uint2 e = stack[(uint32_t)sp * WF_THREADS];
                        if (__uint_as_float(e.y) < closest.t) cur = e.x;
                    }
                }
            }
            const unsigned fin = __ballot_sync(full, finished);
            if (fin) {
                // classification draws (raytrace.wgsl:234, 248) + append to the per-kind queue
                int kind = -1;
                if (finished) {
                    w.ray_b[slot] = make_float4(ray.d.y, ray.d.z, closest.t, __uint_as_float(closest.model));
                    if (closest.t == BVR_INF) {
                        kind = C_MISS;
                    } else {
                        uint32_t mid = w.r.scene.sphere_material[closest.model];
                        if (mid >= sv.n_materials) mid = sv.n_materials - 1u;
                        const float4 m0 = w.r.scene.materials[2u * mid], m1 = w.r.scene.materials[2u * mid + 1u];
                        float4* trp = w.thr_rng + slot;
                        uint32_t rng = __float_as_uint(trp->w);
                        if (rng_next_float(rng) < m0.w) kind = C_METAL;
                        else if (rng_next_float(rng) < m1.w) kind = C_GLASS;
                        else kind = C_DIFFUSE;
                        trp->w = __uint_as_float(rng);
                    }
                    active = false;
                }
                // the four pushes are executed by the converged warp
                queue_push(w.q_miss, w.counters + C_MISS, kind == C_MISS, slot);
                queue_push(w.q_metal, w.counters + C_METAL, kind == C_METAL, slot);
                queue_push(w.q_glass, w.counters + C_GLASS, kind == C_GLASS, slot);
                queue_push(w.q_diffuse, w.counters + C_DIFFUSE, kind == C_DIFFUSE, slot);
            }
            const unsigned act = __ballot_sync(full, active);
            if (act == 0u) break;
            if (!exhausted && (uint32_t)__popc(act) <= w.refill_below) break;
        }
    }

    unsigned long long sum = rays;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(full, sum, o);
    if (lane == 0u && w.r.ray_counter && sum) atomicAdd(w.r.ray_counter, sum);
    (void)stack_cap;
}

// ---- shading helpers ----
struct PathIO {
    uint32_t slot;
    Ray ray;
    Hit hit;
    V3 throughput;
    uint32_t rng;
    uint4 misc;   // (sample index, bounce, first_depth bits, unused)
};

__device__ __forceinline__ PathIO load_path(const WavefrontParams& w, uint32_t slot) {
    PathIO p;
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

// raytrace.wgsl:193-195: first_depth is the primary ray's hit distance; sample 0 also feeds the id/depth planes
__device__ __forceinline__ void record_primary(const WavefrontParams& w, PathIO& p) {
    if (p.misc.y == 0u) {
        p.misc.z = __float_as_uint(p.hit.t);
        if (p.misc.x == 0u) {
            if (w.r.out_primary_id) w.r.out_primary_id[p.slot] = p.hit.t == BVR_INF ? 0xffffffffu : p.hit.model;
            if (w.r.out_primary_depth) w.r.out_primary_depth[p.slot] = p.hit.t;
        }
    }
}

// path ended with gamma-encoded sample colour `c` (raytrace.wgsl:219-223, 166-167)
__device__ __forceinline__ void end_path(const WavefrontParams& w, PathIO& p, V3 c) {
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

// path continues with the scattered ray; returns false when the bounce budget is exhausted
// (raytrace.wgsl:186, 214-216: the sample is black)
__device__ __forceinline__ bool continue_path(const WavefrontParams& w, PathIO& p, V3 attenuation) {
    p.throughput = vmul(p.throughput, attenuation);
    p.misc.y += 1u;
    if (p.misc.y > w.r.cam.bounce_count) return false;
    w.ray_a[p.slot] = make_float4(p.ray.o.x, p.ray.o.y, p.ray.o.z, p.ray.d.x);
    w.ray_b[p.slot] = make_float4(p.ray.d.y, p.ray.d.z, BVR_INF, __uint_as_float(0xffffffffu));
    w.thr_rng[p.slot] = make_float4(p.throughput.x, p.throughput.y, p.throughput.z, __uint_as_float(p.rng));
    w.misc[p.slot] = p.misc;
    return true;
}

__global__ void __launch_bounds__(WF_THREADS) wf_shade_miss(const WavefrontParams w) {
    const uint32_t n = w.counters[C_MISS];
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += stride) {
        const uint32_t i = base + threadIdx.x;
        const bool valid = i < n;
        uint32_t slot = 0;
        if (valid) {
            slot = w.q_miss[i];
            PathIO p = load_path(w, slot);
            record_primary(w, p);
            const V3 lin = vmul(p.throughput, background_gradient(p.ray));
            end_path(w, p, v3(fsqrt(lin.x), fsqrt(lin.y), fsqrt(lin.z)));
        }
        queue_push(w.q_regen, w.counters + C_REGEN, valid, slot);
    }
}

// One material class per launch (KIND = C_METAL / C_GLASS / C_DIFFUSE); the class was chosen by the
// draws in wf_extend, so scatter() is entered after them.
template <int KIND>
__global__ void __launch_bounds__(WF_THREADS) wf_shade_hit(const WavefrontParams w, uint32_t* __restrict__ q_ray_out,
                                                           int ray_counter_out) {
    const uint32_t n = w.counters[KIND];
    const uint32_t* __restrict__ q = KIND == C_METAL ? w.q_metal : (KIND == C_GLASS ? w.q_glass : w.q_diffuse);
    const SceneView& s = w.r.scene;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += stride) {
        const uint32_t i = base + threadIdx.x;
        bool cont = false, ended = false;
        uint32_t slot = 0;
        if (i < n) {
            slot = q[i];
            PathIO p = load_path(w, slot);
            record_primary(w, p);
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
            if (KIND == C_METAL) {                                   // raytrace.wgsl:234-246
                const V3 reflected = vadd(vnormalize(reflect3(p.ray.d, normal)), vscale(m1.x, random_unit_vec3(p.rng)));
                p.ray.o = position;
                p.ray.d = reflected;
                attenuation = base_color;
                absorbed = vdot(p.ray.d, normal) < 0.0f;
            } else if (KIND == C_GLASS) {                            // raytrace.wgsl:248-282
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
                end_path(w, p, v3(0.0f, 0.0f, 0.0f));
                ended = true;
            } else if (continue_path(w, p, attenuation)) {
                cont = true;
            } else {
                end_path(w, p, v3(0.0f, 0.0f, 0.0f));
                ended = true;
            }
        }
        queue_push(q_ray_out, w.counters + ray_counter_out, cont, slot);
        queue_push(w.q_regen, w.counters + C_REGEN, ended, slot);
    }
}

}  // namespace

size_t wavefront_state_bytes(size_t pixels) {
    // ray_a, ray_b, thr_rng, accum (float4) + misc (uint4) + 7 queues (u32) + counters
    return pixels * (5 * 16 + 7 * 4) + 256;
}

void wavefront_bind(WavefrontParams& w, void* state, size_t pixels) {
    char* p = static_cast<char*>(state);
    w.counters = reinterpret_cast<unsigned int*>(p); p += 256;
    w.ray_a = reinterpret_cast<float4*>(p); p += pixels * 16;
    w.ray_b = reinterpret_cast<float4*>(p); p += pixels * 16;
    w.thr_rng = reinterpret_cast<float4*>(p); p += pixels * 16;
    w.accum = reinterpret_cast<float4*>(p); p += pixels * 16;
    w.misc = reinterpret_cast<uint4*>(p); p += pixels * 16;
    w.q_ray[0] = reinterpret_cast<uint32_t*>(p); p += pixels * 4;
    w.q_ray[1] = reinterpret_cast<uint32_t*>(p); p += pixels * 4;
    w.q_miss = reinterpret_cast<uint32_t*>(p); p += pixels * 4;
    w.q_metal = reinterpret_cast<uint32_t*>(p); p += pixels * 4;
    w.q_glass = reinterpret_cast<uint32_t*>(p); p += pixels * 4;
    w.q_diffuse = reinterpret_cast<uint32_t*>(p); p += pixels * 4;
    w.q_regen = reinterpret_cast<uint32_t*>(p);
}

// Renders one frame.  `host_count` is a pinned word used to poll the ray-queue size.
// Returns the number of kernels launched, or -1 on a CUDA error / unsupported configuration.
int launch_wavefront(WavefrontParams w, uint32_t n_inner, uint32_t n_models, uint32_t tree_depth, int sm_count,
                     volatile unsigned int* host_counts, cudaStream_t stream) {
    const uint32_t pixels = w.r.cam.width * w.r.shard.rows;
    if (pixels == 0) return 0;
    int launches = 0;
    const int wide_grid = sm_count * 8;
    const uint32_t stack_cap = tree_depth + 1u;
    const size_t scene_bytes = (size_t)(4u * n_inner + n_models) * 16u;
    const size_t stack_bytes = (size_t)WF_THREADS * stack_cap * sizeof(uint2);
    const size_t max_smem = 227u * 1024u;
    const bool smem_scene = scene_bytes + stack_bytes <= max_smem;
    const size_t smem = (smem_scene ? scene_bytes : 0) + stack_bytes;
    if (smem > max_smem) return -1;
    auto extend = smem_scene ? wf_extend<true> : wf_extend<false>;
    if (cudaFuncSetAttribute(extend, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    int blocks_per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, extend, WF_THREADS, smem) != cudaSuccess || blocks_per_sm < 1)
        return -1;
    const int extend_grid_max = sm_count * blocks_per_sm;

    if (cudaMemsetAsync(w.counters, 0, C_COUNT * sizeof(unsigned int), stream) != cudaSuccess) return -1;
    wf_init<<<wide_grid, WF_THREADS, 0, stream>>>(w);
    launches++;
    // first regen turns every pixel into a camera ray in queue 0
    wf_regen<<<wide_grid, WF_THREADS, 0, stream>>>(w, w.q_ray[0], C_RAY0);
    launches++;

    int cur = 0;
    uint32_t bound = pixels;           // upper bound of the live ray count (refreshed every `poll` waves)
    const int poll = 16;
    for (int wave = 0;; wave++) {
        const int nxt = cur ^ 1;
        const int grid_small = (int)((bound + WF_THREADS - 1) / WF_THREADS);
        const int g_wide = grid_small < wide_grid ? (grid_small < 1 ? 1 : grid_small) : wide_grid;
        const int g_ext = grid_small < extend_grid_max ? (grid_small < 1 ? 1 : grid_small) : extend_grid_max;
        wf_reset<<<1, 32, 0, stream>>>(w.counters, nxt == 0 ? C_RAY0 : C_RAY1);
        extend<<<g_ext, WF_THREADS, smem, stream>>>(w, w.q_ray[cur], cur == 0 ? C_RAY0 : C_RAY1, stack_cap, n_inner, n_models);
        wf_shade_miss<<<g_wide, WF_THREADS, 0, stream>>>(w);
        wf_shade_hit<C_DIFFUSE><<<g_wide, WF_THREADS, 0, stream>>>(w, w.q_ray[nxt], nxt == 0 ? C_RAY0 : C_RAY1);
        wf_shade_hit<C_METAL><<<g_wide, WF_THREADS, 0, stream>>>(w, w.q_ray[nxt], nxt == 0 ? C_RAY0 : C_RAY1);
        wf_shade_hit<C_GLASS><<<g_wide, WF_THREADS, 0, stream>>>(w, w.q_ray[nxt], nxt == 0 ? C_RAY0 : C_RAY1);
        wf_regen<<<g_wide, WF_THREADS, 0, stream>>>(w, w.q_ray[nxt], nxt == 0 ? C_RAY0 : C_RAY1);
        launches += 7;
        cur = nxt;
        if ((wave + 1) % poll == 0) {
            if (cudaMemcpyAsync((void*)host_counts, w.counters, C_COUNT * sizeof(unsigned int), cudaMemcpyDeviceToHost,
                                stream) != cudaSuccess) return -1;
            if (cudaStreamSynchronize(stream) != cudaSuccess) return -1;
            const uint32_t live = host_counts[cur == 0 ? C_RAY0 : C_RAY1];
            if (live == 0u) break;
            bound = live;   // the live count never grows: one path per pixel, pixels only retire
        }
    }
    if (cudaGetLastError() != cudaSuccess) return -1;
    return launches;
}

}  // namespace bvr
