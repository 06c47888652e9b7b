// wavefront.cu — the wavefront pipeline with one kernel launch per stage per wave:
// raygen / extend / classify / shade-by-kind / regen, queues compacted with warp ballots, path state and
// queues in HBM (slot == pixel: every pixel has its one path in flight for the whole frame).
// The stage bodies are in wavefront_stages.cuh; results are bit-identical to the megakernel and the oracle.

#include "wavefront_stages.cuh"

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
    const WfExtendTuning tune{w.refill_below, 4u};

    if (cudaMemsetAsync(w.counters, 0, WC_COUNT * sizeof(unsigned int), stream) != cudaSuccess) return -1;
    wf_init<<<wide_grid, WF_THREADS, 0, stream>>>(w);
    wf_regen<<<wide_grid, WF_THREADS, 0, stream>>>(w, w.q_ray[0], WC_RAY0);   // every pixel -> a camera ray
    launches += 2;

    int cur = 0;
    uint32_t bound = pixels;   // upper bound of the live ray count (refreshed every `poll` waves; it never grows)
    const int poll = 16;
    for (int wave = 0;; wave++) {
        const int nxt = cur ^ 1;
        const int cc = cur == 0 ? WC_RAY0 : WC_RAY1, cn = nxt == 0 ? WC_RAY0 : WC_RAY1;
        const int grid_small = (int)((bound + WF_THREADS - 1) / WF_THREADS);
        const int g_wide = grid_small < wide_grid ? (grid_small < 1 ? 1 : grid_small) : wide_grid;
        const int g_ext = grid_small < extend_grid_max ? (grid_small < 1 ? 1 : grid_small) : extend_grid_max;
        wf_reset<<<1, 32, 0, stream>>>(w.counters, cn);
        extend<<<g_ext, WF_THREADS, smem, stream>>>(w, w.q_ray[cur], cc, n_inner, n_models, tune);
        wf_classify<<<g_wide, WF_THREADS, 0, stream>>>(w, w.q_ray[cur], cc);
        wf_shade_miss<<<g_wide, WF_THREADS, 0, stream>>>(w);
        wf_shade_hit<WC_DIFFUSE><<<g_wide, WF_THREADS, 0, stream>>>(w, w.q_ray[nxt], cn);
        wf_shade_hit<WC_METAL><<<g_wide, WF_THREADS, 0, stream>>>(w, w.q_ray[nxt], cn);
        wf_shade_hit<WC_GLASS><<<g_wide, WF_THREADS, 0, stream>>>(w, w.q_ray[nxt], cn);
        wf_regen<<<g_wide, WF_THREADS, 0, stream>>>(w, w.q_ray[nxt], cn);
        launches += 8;
        cur = nxt;
        if ((wave + 1) % poll == 0) {
            if (cudaMemcpyAsync((void*)host_counts, w.counters, WC_COUNT * sizeof(unsigned int), cudaMemcpyDeviceToHost,
                                stream) != cudaSuccess) return -1;
            if (cudaStreamSynchronize(stream) != cudaSuccess) return -1;
            const uint32_t live = host_counts[cur == 0 ? WC_RAY0 : WC_RAY1];
            if (live == 0u) break;
            bound = live;
        }
    }
    if (cudaGetLastError() != cudaSuccess) return -1;
    return launches;
}

}  // namespace bvr
