// wavefront.cu — the wavefront pipeline with one kernel launch per stage per wave:
// raygen / extend / classify / shade-by-kind / regen, queues compacted with warp ballots, path state and
// queues in HBM (slot == pixel: every pixel has its one path in flight for the whole frame).
// The stage bodies are in wavefront_stages.cuh; results are bit-identical to the megakernel and the oracle.

#include "wavefront_stages.cuh"
#include "wide4.cuh"

namespace bvr {

namespace {

constexpr int WF_THREADS = 256;

__device__ __forceinline__ WfGroup grid_group() {
    return WfGroup{blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x};
}

// every valid pixel gets slot == pixel and is queued for regen in 8x4-tile order
__global__ void __launch_bounds__(WF_THREADS) wf_init(const WavefrontParams w) {
    const WfGroup g = grid_group();
    const uint32_t total = wf_tile_order_count(w);
    for (uint32_t base = 0; base < total; base += g.nthreads) {
        const uint32_t i = base + g.tid;
        uint32_t pixel = 0xffffffffu;
        if (i < total) pixel = wf_tile_order_pixel(w, i);
        const bool valid = pixel != 0xffffffffu;
        if (valid) wf_init_slot(w, pixel, pixel);
        wf_push(w.q_regen, w.counters + WC_REGEN, valid, pixel);
    }
}

__global__ void wf_reset(unsigned int* counters, int next_ray_counter) {
    if (threadIdx.x == 0) {
        counters[WC_MISS] = 0u; counters[WC_METAL] = 0u; counters[WC_GLASS] = 0u; counters[WC_DIFFUSE] = 0u;
        counters[WC_REGEN] = 0u; counters[WC_HEAD] = 0u;
        counters[next_ray_counter] = 0u;
    }
}

__global__ void __launch_bounds__(WF_THREADS) wf_regen(const WavefrontParams w, uint32_t* q_ray_out, int ray_counter_out) {
    wf_stage_regen<false>(w, grid_group(), w.q_regen, w.counters[WC_REGEN], q_ray_out, w.counters + ray_counter_out, nullptr);
}

template <bool SMEM_SCENE>
__global__ void __launch_bounds__(WF_THREADS) wf_extend(const WavefrontParams w, const uint32_t* __restrict__ q_ray_in,
                                                        int ray_counter_in, uint32_t n_inner, uint32_t n_models,
                                                        WfExtendTuning tune) {
    extern __shared__ float4 smem[];
    const uint32_t tid = threadIdx.x;
    const uint32_t n_rays = w.counters[ray_counter_in];
    if (n_rays == 0u) return;
    SceneView sv = w.r.scene;
    float4* sm_cursor = smem;
    if (SMEM_SCENE) {
        float4* sm_pairs = sm_cursor;   sm_cursor += 4u * n_inner;
        float4* sm_spheres = sm_cursor; sm_cursor += n_models;
        for (uint32_t i = tid; i < 4u * n_inner; i += WF_THREADS) sm_pairs[i] = w.r.scene.pairs_ch[i];
        for (uint32_t i = tid; i < n_models; i += WF_THREADS) sm_spheres[i] = w.r.scene.spheres[i];
        sv.pairs_ch = sm_pairs;
        sv.spheres = sm_spheres;
        __syncthreads();
    }
    const uint32_t s_stack0 = wf_smem_addr(sm_cursor) + tid * 8u;
    const uint32_t s_pairs = SMEM_SCENE ? wf_smem_addr(sv.pairs_ch) : 0u;
    const unsigned long long rays = wf_stage_extend<WF_THREADS * 8u, SMEM_SCENE>(w, sv, q_ray_in, n_rays, w.counters + WC_HEAD,
                                                                                s_stack0, s_pairs, tune);
    unsigned long long sum = rays;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((tid & 31u) == 0u && w.r.ray_counter && sum) atomicAdd(w.r.ray_counter, sum);
}

// extend on the records the megakernel walks: 4-wide fp32 records with tight boxes AND the reference
// records staged in shared memory (a ray too far from some radius group for the tight boxes walks the reference ones),
// FFMA2 slab test, packed-key sorting network, 4-byte stack entries, postponed sphere tests — the same visit4() and the
// same step as megakernel_v3's phase B, fed from the ray queue instead of from the lane's own path.
constexpr int WF4_THREADS = 1024;

__global__ void __launch_bounds__(WF4_THREADS) wf_extend4(const WavefrontParams w, const uint32_t* __restrict__ q_ray_in,
                                                          int ray_counter_in, uint32_t n_inner, uint32_t n_models,
                                                          WfExtendTuning tune) {
    extern __shared__ float4 smem[];
    const unsigned full = 0xffffffffu;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint32_t n_rays = w.counters[ray_counter_in];
    if (n_rays == 0u) return;
    const SceneView& sv = w.r.scene;
    float4* sm_tight = smem;
    float4* sm_ref = sm_tight + 7u * n_inner;
    float4* sm_spheres = sm_ref + 7u * n_inner;
    for (uint32_t i = tid; i < 7u * n_inner; i += WF4_THREADS) { sm_tight[i] = sv.nodes4_tight[i]; sm_ref[i] = sv.nodes4_ch[i]; }
    for (uint32_t i = tid; i < n_models; i += WF4_THREADS) sm_spheres[i] = sv.spheres[i];
    __syncthreads();
    constexpr uint32_t STRIDE = WF4_THREADS * 4u;
    const uint32_t s_stack0 = opaque(smem_addr(sm_spheres + n_models) + tid * 4u);
    const uint32_t s_tight = opaque(smem_addr(sm_tight)), s_ref = opaque(smem_addr(sm_ref)), s_spheres = opaque(smem_addr(sm_spheres));
    const uint32_t n_groups = __float_as_uint(__ldg(&sv.tight_groups[0]).x);
    const uint32_t root = !sv.has_scene ? S4_NONE : ((sv.root_ref & BVR_LEAF_BIT) ? (S4_LEAF | (sv.root_ref & 0x3ffu)) : sv.root_ref);
    unsigned int* head = w.counters + WC_HEAD;

    bool active = false, exhausted = false;
    uint32_t slot = 0, s_rec = s_tight;
    Ray ray{v3(0, 0, 0), v3(0, 0, 1)};
    u64 inv_xy = 0, noi_xy = 0;
    float inv_z = 0.0f, noi_z = 0.0f, a = 1.0f;
    Hit closest{BVR_INF, 0xffffffffu};
    uint32_t cur = S4_NONE, pending = S4_NONE, sp_addr = s_stack0;
    unsigned long long rays = 0;
    auto push = [&](uint32_t k) { sts32(sp_addr, k); sp_addr += STRIDE; };

    for (;;) {
        // ---- refill idle lanes (one counter round trip per refill) ----
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
                    const float ix = rcp_approx(ray.d.x), iy = rcp_approx(ray.d.y);
                    inv_z = rcp_approx(ray.d.z);
                    inv_xy = pk2(ix, iy);
                    noi_xy = pk2(-(ray.o.x * ix), -(ray.o.y * iy));
                    noi_z = -(ray.o.z * inv_z);
                    bool far_ray = false;
                    for (uint32_t g = 0; g < n_groups; g++) {
                        const float4 gr = __ldg(&sv.tight_groups[1u + g]);
                        const float dx = ray.o.x - gr.x, dy = ray.o.y - gr.y, dz = ray.o.z - gr.z;
                        far_ray = far_ray || !(dx * dx + dy * dy + dz * dz <= gr.w);
                    }
                    s_rec = far_ray ? s_ref : s_tight;
                    a = vdot(ray.d, ray.d);
                    closest.t = BVR_INF;
                    closest.model = 0xffffffffu;
                    sp_addr = s_stack0;
                    pending = S4_NONE;
                    cur = root;
                    active = true;
                    rays++;
                }
            }
            if (idle != 0u && base + (uint32_t)__popc(idle) >= n_rays) exhausted = true;
        }
        if (!__any_sync(full, active)) break;

        // ---- traverse until enough lanes are idle ----
        for (;;) {
#pragma unroll
            for (int rep = 0; rep < 2; rep++) {
                uint32_t c = cur;
                if (c < S4_LEAF) {
                    const uint32_t na = s_rec + c * 112u;
                    const float4 q0 = lds128(na), q1 = lds128(na + 16u), q2 = lds128(na + 32u), q3 = lds128(na + 48u);
                    const float4 q4 = lds128(na + 64u), q5 = lds128(na + 80u), rr = lds128(na + 96u);
                    c = visit4(q0, q1, q2, q3, q4, q5, rr, inv_xy, noi_xy, inv_z, noi_z, closest.t, push);
                }
                if (c >= S4_LEAF) {
                    if (c != S4_NONE && pending == S4_NONE) { pending = c; c = S4_NONE; }
                    if (c == S4_NONE) {
                        while (sp_addr != s_stack0) {
                            sp_addr -= STRIDE;
                            const uint32_t e = lds32(sp_addr);
                            if (__uint_as_float(e & ~S4_REF_MASK) < closest.t) { c = e & S4_REF_MASK; break; }
                        }
                    }
                }
                cur = c;
            }
            const bool parked = pending != S4_NONE;
            const unsigned blk = __ballot_sync(full, parked && cur >= S4_LEAF);
            if (blk != 0u) {
                const unsigned trav = __ballot_sync(full, cur != S4_NONE || parked);
                const uint32_t nblk = (uint32_t)__popc(blk);
                if (nblk >= tune.leaf_blocked_lanes || nblk == (uint32_t)__popc(trav)) {
                    if (parked) {
                        const uint32_t m = pending & 0x3ffu;
                        test_sphere(sv, ray, a, m, lds128(s_spheres + m * 16u), closest);
                        pending = S4_NONE;
                    }
                }
            }
            if (active && cur == S4_NONE && pending == S4_NONE) {
                // finished: the hit record goes back to the path state; the lane is idle
                w.ray_b[slot] = make_float4(ray.d.y, ray.d.z, closest.t, __uint_as_float(closest.model));
                active = false;
            }
            const unsigned act = __ballot_sync(full, active);
            if (act == 0u) break;
            if (!exhausted && 32u - (uint32_t)__popc(act) >= tune.refill_idle_lanes) break;
        }
    }
    unsigned long long sum = rays;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(full, sum, o);
    if (lane == 0u && w.r.ray_counter && sum) atomicAdd(w.r.ray_counter, sum);
}

__global__ void __launch_bounds__(WF_THREADS) wf_classify(const WavefrontParams w, const uint32_t* __restrict__ q_ray_in,
                                                          int ray_counter_in) {
    wf_stage_classify(w, grid_group(), q_ray_in, w.counters[ray_counter_in], w.q_miss, w.q_metal, w.q_glass, w.q_diffuse,
                      w.counters);
}

__global__ void __launch_bounds__(WF_THREADS) wf_shade_miss(const WavefrontParams w) {
    wf_stage_shade_miss(w, grid_group(), w.q_miss, w.counters[WC_MISS], w.q_regen, w.counters + WC_REGEN);
}

template <int KIND>
__global__ void __launch_bounds__(WF_THREADS) wf_shade_hit(const WavefrontParams w, uint32_t* q_ray_out, int ray_counter_out) {
    const uint32_t* q = KIND == WC_METAL ? w.q_metal : (KIND == WC_GLASS ? w.q_glass : w.q_diffuse);
    wf_stage_shade_hit<KIND>(w, w.r.scene, grid_group(), q, w.counters[KIND], q_ray_out, w.counters + ray_counter_out,
                             w.q_regen, w.counters + WC_REGEN);
}

}  // namespace

size_t wavefront_state_bytes(size_t slots) {
    // ray_a, ray_b, thr_rng, accum (float4) + misc (uint4) + slot_pixel (u32) + 7 queues (u32) + counters
    return slots * (5 * 16 + 4 + 7 * 4) + 256;
}

void wavefront_bind(WavefrontParams& w, void* state, size_t slots) {
    char* p = static_cast<char*>(state);
    w.counters = reinterpret_cast<unsigned int*>(p); p += 256;
    w.ray_a = reinterpret_cast<float4*>(p); p += slots * 16;
    w.ray_b = reinterpret_cast<float4*>(p); p += slots * 16;
    w.thr_rng = reinterpret_cast<float4*>(p); p += slots * 16;
    w.accum = reinterpret_cast<float4*>(p); p += slots * 16;
    w.misc = reinterpret_cast<uint4*>(p); p += slots * 16;
    w.slot_pixel = reinterpret_cast<uint32_t*>(p); p += slots * 4;
    w.q_ray[0] = reinterpret_cast<uint32_t*>(p); p += slots * 4;
    w.q_ray[1] = reinterpret_cast<uint32_t*>(p); p += slots * 4;
    w.q_miss = reinterpret_cast<uint32_t*>(p); p += slots * 4;
    w.q_metal = reinterpret_cast<uint32_t*>(p); p += slots * 4;
    w.q_glass = reinterpret_cast<uint32_t*>(p); p += slots * 4;
    w.q_diffuse = reinterpret_cast<uint32_t*>(p); p += slots * 4;
    w.q_regen = reinterpret_cast<uint32_t*>(p);
}

// Renders one frame.  `host_counts` = 8 pinned words used to poll the ray-queue size.
// The waves are launched as a CUDA graph of `poll` waves (8 kernels each) that is replayed until the polled ray count
// reaches zero: one graph launch per 128 kernels instead of 128 launches.  Grid sizes inside the graph are fixed (the
// kernels read their queue sizes from device counters and idle CTAs leave at once).
// Returns the number of kernels launched, or -1 on a CUDA error / unsupported configuration.
int launch_wavefront(WavefrontParams w, uint32_t n_inner, uint32_t n_models, uint32_t tree_depth, int sm_count,
                     volatile unsigned int* host_counts, cudaStream_t stream) {
    const uint32_t pixels = w.r.cam.width * w.r.shard.rows;
    if (pixels == 0) return 0;
    int launches = 0;
    const int wide_grid = sm_count * 8;
    const size_t max_smem = 227u * 1024u;
    // the megakernel's records (4-wide, tight + reference, in shared memory) when they fit ...
    const uint32_t cap4 = 3u * (tree_depth / 2u) + 1u;
    const size_t smem4 = ((size_t)14u * n_inner + n_models) * 16u + (size_t)WF4_THREADS * cap4 * 4u;
    const bool use4 = w.r.scene.nodes4_tight != nullptr && w.r.scene.nodes4_ch != nullptr && w.r.scene.tight_groups != nullptr &&
                      smem4 <= max_smem;
    // ... else child-pair records (shared memory or global)
    const uint32_t stack_cap = tree_depth + 1u;
    const size_t scene_bytes = (size_t)(4u * n_inner + n_models) * 16u;
    const size_t stack_bytes = (size_t)WF_THREADS * stack_cap * sizeof(uint2);
    const bool smem_scene = scene_bytes + stack_bytes <= max_smem;
    const size_t smem2 = (smem_scene ? scene_bytes : 0) + stack_bytes;
    if (!use4 && smem2 > max_smem) return -1;
    auto extend2 = smem_scene ? wf_extend<true> : wf_extend<false>;
    int extend_grid = sm_count;
    if (use4) {
        if (cudaFuncSetAttribute(wf_extend4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4) != cudaSuccess) return -1;
    } else {
        if (cudaFuncSetAttribute(extend2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2) != cudaSuccess) return -1;
        int blocks_per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, extend2, WF_THREADS, smem2) != cudaSuccess || blocks_per_sm < 1)
            return -1;
        extend_grid = sm_count * blocks_per_sm;
    }
    const WfExtendTuning tune{w.refill_below, use4 ? 1u : 4u};

    if (cudaMemsetAsync(w.counters, 0, WC_COUNT * sizeof(unsigned int), stream) != cudaSuccess) return -1;
    wf_init<<<wide_grid, WF_THREADS, 0, stream>>>(w);
    wf_regen<<<wide_grid, WF_THREADS, 0, stream>>>(w, w.q_ray[0], WC_RAY0);   // every pixel -> a camera ray
    launches += 2;

    // one chunk = `poll` waves, an even number: the ray queues are back in their starting roles afterwards
    const int poll = 16;
    auto enqueue_chunk = [&](cudaStream_t st) {
        int cur = 0;
        for (int wave = 0; wave < poll; wave++) {
            const int nxt = cur ^ 1;
            const int cc = cur == 0 ? WC_RAY0 : WC_RAY1, cn = nxt == 0 ? WC_RAY0 : WC_RAY1;
            wf_reset<<<1, 32, 0, st>>>(w.counters, cn);
            if (use4) wf_extend4<<<extend_grid, WF4_THREADS, smem4, st>>>(w, w.q_ray[cur], cc, n_inner, n_models, tune);
            else extend2<<<extend_grid, WF_THREADS, smem2, st>>>(w, w.q_ray[cur], cc, n_inner, n_models, tune);
            wf_classify<<<wide_grid, WF_THREADS, 0, st>>>(w, w.q_ray[cur], cc);
            wf_shade_miss<<<wide_grid, WF_THREADS, 0, st>>>(w);
            wf_shade_hit<WC_DIFFUSE><<<wide_grid, WF_THREADS, 0, st>>>(w, w.q_ray[nxt], cn);
            wf_shade_hit<WC_METAL><<<wide_grid, WF_THREADS, 0, st>>>(w, w.q_ray[nxt], cn);
            wf_shade_hit<WC_GLASS><<<wide_grid, WF_THREADS, 0, st>>>(w, w.q_ray[nxt], cn);
            wf_regen<<<wide_grid, WF_THREADS, 0, st>>>(w, w.q_ray[nxt], cn);
            cur = nxt;
        }
    };
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    bool graphed = false;
    if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        enqueue_chunk(stream);
        if (cudaStreamEndCapture(stream, &graph) == cudaSuccess && graph != nullptr &&
            cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess)
            graphed = true;
    }
    if (!graphed) cudaGetLastError();
    int status = 0;
    for (;;) {
        if (graphed) { if (cudaGraphLaunch(exec, stream) != cudaSuccess) { status = -1; break; } }
        else enqueue_chunk(stream);
        launches += 8 * poll;
        if (cudaMemcpyAsync((void*)host_counts, w.counters, WC_COUNT * sizeof(unsigned int), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
            cudaStreamSynchronize(stream) != cudaSuccess) { status = -1; break; }
        if (host_counts[WC_RAY0] == 0u) break;
    }
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    if (status < 0 || cudaGetLastError() != cudaSuccess) return -1;
    return launches;
}

}  // namespace bvr
