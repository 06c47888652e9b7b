// cta_wavefront.cu — the wavefront pipeline inside ONE persistent kernel.
//
// Every CTA owns a pool of POOL path slots and runs the stages of wavefront_stages.cuh on its own queues,
// separated by __syncthreads instead of kernel launches: regen -> extend -> classify -> shade x4 -> ...
// A slot whose pixel has all its samples takes the next pixel from a global tile-ordered queue, so the pool
// stays full until the frame runs out of pixels and there is no frame-wide tail of nearly empty waves (the
// launch-per-stage pipeline needs ~1056 waves x 8 launches on the RTIOW frame, most of them tiny).
// Queue counters live in shared memory; path state and queue entries live in HBM/L2 (84 B + 28 B per slot);
// the BVH and the spheres are staged in shared memory next to the traversal stacks.
// Work of one kind is always executed by full warps — the point of the exercise: ncu on the megakernel shows
// 16 of 32 lanes active per instruction (profiles/r01_v3b_*).

#include "wavefront_stages.cuh"

namespace bvr {

namespace {

constexpr uint32_t CW_POOL = 4096;   // path slots per CTA

template <int THREADS, bool SMEM_SCENE>
__global__ void __launch_bounds__(THREADS) cta_wavefront_kernel(const WavefrontParams w, unsigned int* pixel_queue_head,
                                                                uint32_t n_inner, uint32_t n_models, WfExtendTuning tune) {
    extern __shared__ float4 smem[];
    __shared__ unsigned int cnt[WC_COUNT];
    const uint32_t tid = threadIdx.x;
    const WfGroup g{tid, (uint32_t)THREADS};

    SceneView sv = w.r.scene;
    float4* sm_cursor = smem;
    if (SMEM_SCENE) {
        float4* sm_pairs = sm_cursor;   sm_cursor += 4u * n_inner;
        float4* sm_spheres = sm_cursor; sm_cursor += n_models;
        for (uint32_t i = tid; i < 4u * n_inner; i += THREADS) sm_pairs[i] = w.r.scene.pairs_ch[i];
        for (uint32_t i = tid; i < n_models; i += THREADS) sm_spheres[i] = w.r.scene.spheres[i];
        sv.pairs_ch = sm_pairs;
        sv.spheres = sm_spheres;
    }
    const uint32_t s_stack0 = wf_smem_addr(sm_cursor) + tid * 8u;
    const uint32_t s_pairs = SMEM_SCENE ? wf_smem_addr(sv.pairs_ch) : 0u;

    // this CTA's slice of the slot-indexed arrays and queues
    const uint32_t slot0 = blockIdx.x * CW_POOL;
    uint32_t* q_ray[2] = {w.q_ray[0] + slot0, w.q_ray[1] + slot0};
    uint32_t* q_miss = w.q_miss + slot0;
    uint32_t* q_metal = w.q_metal + slot0;
    uint32_t* q_glass = w.q_glass + slot0;
    uint32_t* q_diffuse = w.q_diffuse + slot0;
    uint32_t* q_regen = w.q_regen + slot0;

    // every slot starts empty and asks regen for a pixel
    if (tid < WC_COUNT) cnt[tid] = 0u;
    for (uint32_t s = tid; s < CW_POOL; s += THREADS) {
        w.slot_pixel[slot0 + s] = 0xffffffffu;
        w.misc[slot0 + s] = make_uint4(0u, 0u, 0u, 0u);
        q_regen[s] = slot0 + s;
    }
    __syncthreads();
    if (tid == 0) cnt[WC_REGEN] = CW_POOL;
    __syncthreads();

    unsigned long long rays = 0;
    int cur = 0;
    for (;;) {
        const int cc = cur == 0 ? WC_RAY0 : WC_RAY1, cn = cur == 0 ? WC_RAY1 : WC_RAY0;
        // regen: ended paths -> next sample / store + next pixel; appends to the CURRENT ray queue
        wf_stage_regen<true>(w, g, q_regen, cnt[WC_REGEN], q_ray[cur], cnt + cc, pixel_queue_head);
        __syncthreads();
        const uint32_t n_rays = cnt[cc];
        if (n_rays == 0u) break;
        if (tid == 0) {
            cnt[WC_REGEN] = 0u; cnt[WC_HEAD] = 0u; cnt[cn] = 0u;
            cnt[WC_MISS] = 0u; cnt[WC_METAL] = 0u; cnt[WC_GLASS] = 0u; cnt[WC_DIFFUSE] = 0u;
        }
        __syncthreads();
        rays += wf_stage_extend<THREADS * 8u, SMEM_SCENE>(w, sv, q_ray[cur], n_rays, cnt + WC_HEAD, s_stack0, s_pairs, tune);
        __syncthreads();
        wf_stage_classify(w, g, q_ray[cur], n_rays, q_miss, q_metal, q_glass, q_diffuse, cnt);
        __syncthreads();
        wf_stage_shade_miss(w, g, q_miss, cnt[WC_MISS], q_regen, cnt + WC_REGEN);
        wf_stage_shade_hit<WC_DIFFUSE>(w, w.r.scene, g, q_diffuse, cnt[WC_DIFFUSE], q_ray[cur ^ 1], cnt + cn, q_regen, cnt + WC_REGEN);
        wf_stage_shade_hit<WC_METAL>(w, w.r.scene, g, q_metal, cnt[WC_METAL], q_ray[cur ^ 1], cnt + cn, q_regen, cnt + WC_REGEN);
        wf_stage_shade_hit<WC_GLASS>(w, w.r.scene, g, q_glass, cnt[WC_GLASS], q_ray[cur ^ 1], cnt + cn, q_regen, cnt + WC_REGEN);
        __syncthreads();
        cur ^= 1;
    }

    unsigned long long sum = rays;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((tid & 31u) == 0u && w.r.ray_counter && sum) atomicAdd(w.r.ray_counter, sum);
}

constexpr int CW_THREADS = 512;

}  // namespace

size_t cta_wavefront_slots(int sm_count) { return (size_t)sm_count * 2u * CW_POOL; }

// returns kernels launched, -1 when the configuration does not fit
int launch_cta_wavefront(WavefrontParams w, uint32_t n_inner, uint32_t n_models, uint32_t tree_depth, int sm_count,
                         unsigned int* pixel_counter, cudaStream_t stream) {
    const uint32_t pixels = w.r.cam.width * w.r.shard.rows;
    if (pixels == 0) return 0;
    const uint32_t stack_cap = tree_depth + 1u;
    const size_t scene_bytes = (size_t)(4u * n_inner + n_models) * 16u;
    const size_t stack_bytes = (size_t)CW_THREADS * stack_cap * sizeof(uint2);
    const size_t max_smem = 227u * 1024u - 64u;
    const bool smem_scene = scene_bytes + stack_bytes <= max_smem / 2;   // two CTAs per SM
    const size_t smem = (smem_scene ? scene_bytes : 0) + stack_bytes;
    if (smem > max_smem) return -1;
    auto kern = smem_scene ? cta_wavefront_kernel<CW_THREADS, true> : cta_wavefront_kernel<CW_THREADS, false>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    int blocks_per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, CW_THREADS, smem) != cudaSuccess || blocks_per_sm < 1)
        return -1;
    if (blocks_per_sm > 2) blocks_per_sm = 2;
    uint32_t grid = (uint32_t)(sm_count * blocks_per_sm);
    const uint32_t max_useful = (pixels + CW_POOL - 1u) / CW_POOL;
    if (grid > max_useful) grid = max_useful;
    if (grid == 0) grid = 1;
    const WfExtendTuning tune{w.refill_below, 4u};
    kern<<<grid, CW_THREADS, smem, stream>>>(w, pixel_counter, n_inner, n_models, tune);
    return 1;
}

}  // namespace bvr
