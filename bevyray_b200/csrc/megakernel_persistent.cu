// megakernel_persistent.cu — the production megakernel: persistent lanes with path regeneration.
//
// One lane owns one pixel at a time and runs that pixel's samples in order (the reference's RNG stream
// is sequential per pixel, assets/shaders/raytrace.wgsl:89,161-167), but the sample loop, the bounce
// loop and the BVH loop of raytrace.wgsl:159-224/313-346 are flattened into one state machine:
//
//   phase A (shade / regenerate): lanes whose traversal finished shade their hit, and either continue
//            the path, start the pixel's next sample, or — when the pixel is complete — write it and
//            pull the next pixel from a global tile-ordered queue;
//   phase B (traverse): all lanes with a live ray walk the BVH together; the loop is left as soon as
//            enough lanes are waiting for phase A, so no lane idles for a whole path or a whole pixel.
//
// The scene (child-pair nodes, spheres, materials) is staged in shared memory when it fits — the whole
// RTIOW scene is < 60 KB — and every lane's traversal stack lives in shared memory, interleaved so
// that lane accesses are bank-conflict free.  Sphere tests, shading and ray generation use the strict
// arithmetic of trace.cuh (bit-identical to the oracle).  Box tests only cull; they use one FMA per
// slab plane and 3-input min/max (FMNMX3), which cannot change the closest hit.

#include "kernels.cuh"

namespace bvr {

namespace {

enum LaneState : int { NEED_PIXEL = 0, NEW_PATH = 1, RAY_READY = 2, TRAVERSE = 3, SHADE = 4, DONE = 5 };

#define BVR_NONE 0x7fffffffu   // "no current node": pop on the next step

__device__ __forceinline__ float fmin3(float a, float b, float c) { return fminf(fminf(a, b), c); }
__device__ __forceinline__ float fmax3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// Slab test with t = fma(plane, 1/d, -o/d).  Returns the entry distance clamped to 0, or INF on a miss
// (same contract as ray_bounding_dst; culling only).
__device__ __forceinline__ float box_dst_fma(V3 inv, V3 noi, float mnx, float mny, float mnz, float mxx,
                                             float mxy, float mxz) {
    const float t0x = __fmaf_rn(mnx, inv.x, noi.x), t1x = __fmaf_rn(mxx, inv.x, noi.x);
    const float t0y = __fmaf_rn(mny, inv.y, noi.y), t1y = __fmaf_rn(mxy, inv.y, noi.y);
    const float t0z = __fmaf_rn(mnz, inv.z, noi.z), t1z = __fmaf_rn(mxz, inv.z, noi.z);
    const float t_near = fmax3(fminf(t0x, t1x), fminf(t0y, t1y), fminf(t0z, t1z));
    const float t_far = fmin3(fmaxf(t0x, t1x), fmaxf(t0y, t1y), fmaxf(t0z, t1z));
    const bool hit = (t_far >= t_near) && (t_far > 0.0f);
    return hit ? fmaxf(t_near, 0.0f) : BVR_INF;
}

template <int THREADS, bool SMEM_SCENE>
__global__ void __launch_bounds__(THREADS) megakernel_persistent(const RenderParams p, const uint32_t stack_cap,
                                                                 unsigned int* __restrict__ pixel_counter,
                                                                 const uint32_t n_inner, const uint32_t n_models,
                                                                 const uint32_t shade_wait_lanes) {
    extern __shared__ float4 smem[];
    const CameraParams& cam = p.cam;
    const unsigned full = 0xffffffffu;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;

    // ---- stage the scene in shared memory ----
    SceneView sv = p.scene;
    float4* sm_cursor = smem;
    if (SMEM_SCENE) {
        float4* sm_pairs = sm_cursor;                    sm_cursor += 4u * n_inner;
        float4* sm_spheres = sm_cursor;                  sm_cursor += n_models;
        float4* sm_materials = sm_cursor;                sm_cursor += 2u * sv.n_materials;
        uint32_t* sm_matid = reinterpret_cast<uint32_t*>(sm_cursor);
        sm_cursor += (n_models + 3u) / 4u;
        for (uint32_t i = tid; i < 4u * n_inner; i += THREADS) sm_pairs[i] = p.scene.pairs[i];
        for (uint32_t i = tid; i < n_models; i += THREADS) sm_spheres[i] = p.scene.spheres[i];
        for (uint32_t i = tid; i < 2u * sv.n_materials; i += THREADS) sm_materials[i] = p.scene.materials[i];
        for (uint32_t i = tid; i < n_models; i += THREADS) sm_matid[i] = p.scene.sphere_material[i];
        sv.pairs = sm_pairs;
        sv.spheres = sm_spheres;
        sv.materials = sm_materials;
        sv.sphere_material = sm_matid;
        __syncthreads();
    }
    // per-lane stack, entry k of lane t at stack[k * THREADS + t]  (ref, dst bits)
    uint2* const stack = reinterpret_cast<uint2*>(sm_cursor) + tid;

    // pixel queue geometry: 8x4 tiles in row-major tile order, 32 consecutive indices per tile
    const uint32_t tiles_x = (cam.width + 7u) / 8u, tiles_y = (p.shard.rows + 3u) / 4u;
    const uint32_t total_slots = tiles_x * tiles_y * 32u;

    // ---- lane state ----
    int state = NEED_PIXEL;
    uint32_t px = 0, ly = 0, gy = 0;
    float u = 0.0f, v = 0.0f;
    uint32_t rng = 0, sidx = 0, bounce = 0;
    V3 total = v3(0.0f, 0.0f, 0.0f);
    float total_depth = 0.0f, first_depth = BVR_INF;
    uint32_t primary_id = 0xffffffffu;
    float primary_t = BVR_INF;
    V3 throughput = v3(1.0f, 1.0f, 1.0f);
    Ray ray{v3(0, 0, 0), v3(0, 0, 1)};
    V3 inv = v3(0, 0, 0), noi = v3(0, 0, 0);
    float a = 1.0f;
    Hit closest{BVR_INF, 0xffffffffu};
    uint32_t cur = BVR_NONE;
    int sp = 0;
    unsigned long long rays = 0;

    for (;;) {
        // ================= phase A =================
        if (state == SHADE) {
            bool path_end = false;
            V3 sample_color = v3(0.0f, 0.0f, 0.0f);
            if (bounce == 0u) {
                first_depth = closest.t;
                if (sidx == 0u) { primary_id = closest.t == BVR_INF ? 0xffffffffu : closest.model; primary_t = closest.t; }
            }
            if (closest.t == BVR_INF) {                       // raytrace.wgsl:198-201
                const V3 lin = vmul(throughput, background_gradient(ray));
                sample_color = v3(fsqrt(lin.x), fsqrt(lin.y), fsqrt(lin.z));   // raytrace.wgsl:223
                path_end = true;
            } else {
                V3 attenuation;
                const bool absorbed = scatter(sv, ray, closest, rng, attenuation);
                if (absorbed) {
                    path_end = true;                          // light stays 0 -> black, raytrace.wgsl:207-209
                } else {
                    throughput = vmul(throughput, attenuation);
                    bounce++;
                    if (bounce > cam.bounce_count) path_end = true;   // raytrace.wgsl:214-216 -> black
                }
            }
            if (path_end) {
                if (first_depth == BVR_INF) first_depth = cam.fallback_far;
                total = vadd(total, sample_color);
                total_depth = fadd(total_depth, first_depth);
                sidx++;
                state = NEW_PATH;
            } else {
                state = RAY_READY;
            }
        }
        if (state == NEW_PATH && sidx >= cam.sample_count) {
            // pixel complete: average, fused composite (raytrace.wgsl:104-120), store
            const float n = (float)cam.sample_count;
            float4 out = make_float4(fdiv(total.x, n), fdiv(total.y, n), fdiv(total.z, n), 1.0f);
            const float depth_avg = fdiv(total_depth, n);
            if (cam.level == 1u || cam.level == 2u) {
                const size_t gpix = (size_t)gy * cam.width + px;
                if (raster_wins(cam, p.raster_depth[gpix], depth_avg)) out = p.raster_rgba[gpix];
            }
            const size_t lpix = (size_t)ly * cam.width + px;
            if (p.out_rgba) p.out_rgba[lpix] = out;
            if (p.out_rt_depth) p.out_rt_depth[lpix] = depth_avg;
            if (p.out_primary_id) p.out_primary_id[lpix] = primary_id;
            if (p.out_primary_depth) p.out_primary_depth[lpix] = primary_t;
            if (p.out_srgb8) p.out_srgb8[lpix] = store_srgb8(out);
            state = NEED_PIXEL;
        }
        // pull new pixels (warp-convergent: every lane executes the votes)
        for (;;) {
            const unsigned need = __ballot_sync(full, state == NEED_PIXEL);
            if (need == 0u) break;
            const int leader = __ffs(need) - 1;
            unsigned base = 0;
            if ((int)lane == leader) base = atomicAdd(pixel_counter, (unsigned)__popc(need));
            base = __shfl_sync(full, base, leader);
            if (state == NEED_PIXEL) {
                const uint32_t slot = base + (uint32_t)__popc(need & ((1u << lane) - 1u));
                if (slot >= total_slots) {
                    state = DONE;
                } else {
                    const uint32_t tile = slot >> 5, within = slot & 31u;
                    px = (tile % tiles_x) * 8u + (within & 7u);
                    ly = (tile / tiles_x) * 4u + (within >> 3);
                    gy = shard_global_row(p.shard, ly);
                    if (px < cam.width && ly < p.shard.rows && gy < cam.height) {
                        u = pixel_u(cam, px);
                        v = pixel_v(cam, gy);
                        rng = pixel_seed(cam, u, v);
                        sidx = 0u;
                        total = v3(0.0f, 0.0f, 0.0f);
                        total_depth = 0.0f;
                        primary_id = 0xffffffffu;
                        primary_t = BVR_INF;
                        state = NEW_PATH;   // sample_count == 0 falls through to the store next iteration
                    }
                }
            }
        }
        if (__all_sync(full, state == DONE)) break;
        if (state == NEW_PATH && sidx < cam.sample_count) {
            ray = random_ray_from_uv(cam, u, v, rng);          // raytrace.wgsl:162
            throughput = v3(1.0f, 1.0f, 1.0f);
            bounce = 0u;
            first_depth = BVR_INF;
            state = RAY_READY;
        }
        if (state == RAY_READY) {
            // a fresh ray (camera ray or scattered ray): set up its traversal
            inv = v3(fdiv(1.0f, ray.d.x), fdiv(1.0f, ray.d.y), fdiv(1.0f, ray.d.z));
            noi = v3(-fmul(ray.o.x, inv.x), -fmul(ray.o.y, inv.y), -fmul(ray.o.z, inv.z));
            a = vdot(ray.d, ray.d);
            closest.t = BVR_INF;
            closest.model = 0xffffffffu;
            sp = 0;
            cur = sv.has_scene ? sv.root_ref : BVR_NONE;
            rays++;
            state = TRAVERSE;
        }

        // ================= phase B =================
        for (;;) {
            if (state == TRAVERSE) {
                if (cur != BVR_NONE) {
                    if (cur & BVR_LEAF_BIT) {
                        test_leaf(sv, ray, a, cur, closest);
                        cur = BVR_NONE;
                    } else {
                        const float4* nd = sv.pairs + 4u * cur;
                        const float4 q0 = nd[0], q1 = nd[1], q2 = nd[2], q3 = nd[3];
                        const float d0 = box_dst_fma(inv, noi, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y);
                        const float d1 = box_dst_fma(inv, noi, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w);
                        const bool h0 = d0 < closest.t, h1 = d1 < closest.t;   // INF (miss) never passes
                        const uint32_t r0 = __float_as_uint(q3.x), r1 = __float_as_uint(q3.y);
                        if (h0 && h1) {
                            const bool first0 = d0 < d1;
                            stack[(uint32_t)sp * THREADS] = make_uint2(first0 ? r1 : r0, __float_as_uint(first0 ? d1 : d0));
                            sp++;
                            cur = first0 ? r0 : r1;
                        } else {
                            cur = h0 ? r0 : (h1 ? r1 : BVR_NONE);
                        }
                    }
                }
                if (cur == BVR_NONE) {
                    if (sp == 0) {
                        state = SHADE;
                    } else {
                        --sp;
                        const uint2 e = stack[(uint32_t)sp * THREADS];
                        if (__uint_as_float(e.y) < closest.t) cur = e.x;
                    }
                }
            }
            const unsigned trav = __ballot_sync(full, state == TRAVERSE);
            if (trav == 0u) break;
            const unsigned waiting = __ballot_sync(full, state == SHADE);
            if ((uint32_t)__popc(waiting) >= shade_wait_lanes) break;
        }
    }

    unsigned long long sum = rays;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(full, sum, o);
    if (lane == 0u && p.ray_counter && sum) atomicAdd(p.ray_counter, sum);
    (void)stack_cap;
}

template <int THREADS>
int launch_variant(const RenderParams& p, uint32_t n_inner, uint32_t n_models, uint32_t tree_depth,
                   unsigned int* pixel_counter, uint32_t shade_wait_lanes, int sm_count, cudaStream_t stream) {
    const uint32_t stack_cap = tree_depth + 1u;
    const size_t scene_bytes = (size_t)(4u * n_inner + n_models + 2u * p.scene.n_materials + (n_models + 3u) / 4u) * 16u;
    const size_t stack_bytes = (size_t)THREADS * stack_cap * sizeof(uint2);
    const size_t max_smem = 227u * 1024u;
    const bool smem_scene = scene_bytes + stack_bytes <= max_smem;
    const size_t smem = (smem_scene ? scene_bytes : 0) + stack_bytes;
    if (smem > max_smem) return -1;
    auto kern = smem_scene ? megakernel_persistent<THREADS, true> : megakernel_persistent<THREADS, false>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    int blocks_per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, THREADS, smem) != cudaSuccess || blocks_per_sm < 1)
        return -1;
    const uint32_t tiles = ((p.cam.width + 7u) / 8u) * ((p.shard.rows + 3u) / 4u);
    uint32_t grid = (uint32_t)(sm_count * blocks_per_sm);
    const uint32_t max_useful = (tiles * 32u + THREADS - 1u) / THREADS;
    if (grid > max_useful) grid = max_useful;
    if (grid == 0) return 0;
    kern<<<grid, THREADS, smem, stream>>>(p, stack_cap, pixel_counter, n_inner, n_models, shade_wait_lanes);
    return 1;
}

}  // namespace

// returns kernels launched, or -1 when the configuration does not fit (caller falls back to v1)
int launch_megakernel_persistent(const RenderParams& p, uint32_t n_inner, uint32_t n_models, uint32_t tree_depth,
                                 unsigned int* pixel_counter, int threads, uint32_t shade_wait_lanes, int sm_count,
                                 cudaStream_t stream) {
    switch (threads) {
        case 128: return launch_variant<128>(p, n_inner, n_models, tree_depth, pixel_counter, shade_wait_lanes, sm_count, stream);
        case 256: return launch_variant<256>(p, n_inner, n_models, tree_depth, pixel_counter, shade_wait_lanes, sm_count, stream);
        case 512: return launch_variant<512>(p, n_inner, n_models, tree_depth, pixel_counter, shade_wait_lanes, sm_count, stream);
        case 1024: return launch_variant<1024>(p, n_inner, n_models, tree_depth, pixel_counter, shade_wait_lanes, sm_count, stream);
        default: return -1;
    }
}

}  // namespace bvr
