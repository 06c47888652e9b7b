// kernels.cuh — launch-side declarations shared between bvr_api.cu and the kernel translation units.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "trace.cuh"

namespace bvr {

struct ShardParams {
    uint32_t index, count, strip_rows;
    uint32_t rows;   // rows of this shard's planes (padded to whole strips, equal on every shard)
};

struct RenderParams {
    SceneView scene;
    CameraParams cam;
    ShardParams shard;
    // inputs (full image), may be null for level 3
    const float4* raster_rgba;
    const float* raster_depth;
    // outputs (shard-local planes), any may be null
    float4* out_rgba;
    float* out_rt_depth;
    uint32_t* out_primary_id;
    float* out_primary_depth;
    uchar4* out_srgb8;
    unsigned long long* ray_counter;   // += raycast() invocations
    uint32_t reference_order;          // BvrTraversal
    uint32_t strict_slab;              // scene / camera too far from the origin for the culling-only slab arithmetic:
                                       // the one-thread-per-pixel kernel runs, with the reference's (box - o) * (1/d)
    float4* selfcheck_log;             // BVR_SELFCHECK: 2 x float4 per logged ray (o.xyz, t) (d.xyz, model bits), or null
    unsigned int* selfcheck_count;     // rays logged so far (may exceed the capacity: the excess is dropped)
    uint32_t selfcheck_cap;
    float out_weight;                  // BvrRenderOptions.output_weight (1 = none): applied to rgba and rt_depth as they are stored
    uint32_t extra_modulus, extra_phase, extra_count;   // BVR_RENDER_EXTRA_SAMPLE (modulus 0 = off): the pixels of the 8x4 tiles
                                       // with ((tx + ty + phase) % modulus) < count take one sample more (megakernel_v3 only)
    float inv_pow2_samples;            // 1 / sample_count when that is a power of two and every pixel takes sample_count samples
                                       // (x * 2^-k == x / 2^k bit for bit), else 0: megakernel_v3 divides
    uint32_t weight_in_kernel;         // set by the kernel that applies out_weight itself (others get a scale pass afterwards)
};

// Wavefront path state in HBM (SoA, indexed by shard-local pixel slot) and the work queues.
struct WavefrontParams {
    RenderParams r;
    unsigned int* counters;      // queue sizes + the extend kernel's queue head (wavefront.cu: enum Counter)
    float4* ray_a;               // (o.x, o.y, o.z, d.x)
    float4* ray_b;               // (d.y, d.z, hit t, hit model bits)
    float4* thr_rng;             // (throughput.rgb, rng state bits)
    float4* accum;               // (sum of gamma-encoded sample colours, sum of first depths)
    uint4* misc;                 // (sample index, bounce, first_depth bits, -)
    uint32_t* slot_pixel;        // shard-local pixel index rendered in this slot (0xffffffff = empty)
    uint32_t* q_ray[2];          // ping-pong queue of pixel slots with a live ray
    uint32_t* q_miss;
    uint32_t* q_metal;
    uint32_t* q_glass;
    uint32_t* q_diffuse;
    uint32_t* q_regen;           // paths that ended: next sample or pixel store
    uint32_t refill_below;       // extend: refill idle lanes when at most this many lanes are still traversing
};

// global row of shard-local row `ly`; >= height for padding rows
__host__ __device__ __forceinline__ uint32_t shard_global_row(const ShardParams& s, uint32_t ly) {
    const uint32_t strip = ly / s.strip_rows;
    return (strip * s.count + s.index) * s.strip_rows + (ly % s.strip_rows);
}

// ---- scene derivation (scene_kernels.cu) ----
struct RawNode {   // BvrBvhNode bytes
    float mn[3]; uint32_t pad0; float mx[3]; uint32_t index; uint32_t model_count; uint32_t pad1[3];
};
struct RawModel {  // BvrModel bytes
    float position[3]; float radius; uint32_t material_id; uint32_t pad[3];
};

// launches; each returns the number of kernels it launched
int launch_derive_spheres(const RawModel* models, uint32_t n, float4* spheres, uint32_t* sphere_material,
                          cudaStream_t stream);
// inner_id: scratch of n_nodes u32; block_sums: scratch of ceil(n_nodes/1024)+1 u32
int launch_derive_pairs(const RawNode* nodes, uint32_t n_nodes, uint32_t* inner_id, uint32_t* block_sums,
                        float4* pairs, float4* pairs_ch, uint32_t* root_ref_out, cudaStream_t stream);

// Record numbering of the quantised layouts: the first `max_top` records of a breadth-first walk over the 4-wide levels get
// the ids 0, 1, ..., every other inner node follows in array order (the render kernel stages a prefix of the record array
// in shared memory).  id_q: n_nodes + 1 u32 (the last word = breadth-first ids handed out); after launch_derive_pairs.
int launch_derive_top_order(const RawNode* nodes, uint32_t n_nodes, const uint32_t* inner_id, uint32_t* block_sums,
                            uint32_t max_top, uint32_t* id_q, cudaStream_t stream);

// 32-byte quantised records for scenes that are walked in HBM/L2 (after launch_derive_pairs: needs inner_id);
// grid = 8 floats (base.xyz, -, step.xyz, -); *bad != 0 afterwards -> the scene does not qualify
int launch_derive_pairs_q16(const RawNode* nodes, uint32_t n_nodes, const uint32_t* inner_id, uint4* pairs_q,
                            float* grid, uint32_t* bad, cudaStream_t stream);

// 64-byte 4-wide records (the children's children) from the 32-byte ones
int launch_derive_nodes4_q16(const RawNode* nodes, uint32_t n_nodes, const uint32_t* inner_id, const uint4* pairs_q,
                             uint4* nodes4, cudaStream_t stream);

// 112-byte 4-wide fp32 records for scenes staged in shared memory (needs <= 1024 inner nodes / models, one model
// per leaf); after launch_derive_pairs
int launch_derive_nodes4_ch(const RawNode* nodes, uint32_t n_nodes, const uint32_t* inner_id, const float4* pairs_ch,
                            float4* nodes4, cudaStream_t stream);

// tight boxes (small scenes): see scene_kernels.cu
int launch_derive_tight(const RawNode* nodes, uint32_t n_nodes, const float4* spheres, uint32_t n_models, float pad,
                        float4* groups, RawNode* nodes_tight, cudaStream_t stream);

// ---- structural validation of big node arrays on the GPU (scene_validate.cu) ----
struct ValidateOut { uint32_t bad, depth, max_leaf, n_inner; };
#define BVR_VALIDATE_LEAF_TOO_BIG 1u
#define BVR_VALIDATE_LEAF_RANGE 2u
#define BVR_VALIDATE_CHILD_RANGE 4u
#define BVR_VALIDATE_TWICE 8u
#define BVR_VALIDATE_MAX_DEPTH 2048u
size_t validate_scratch_bytes(uint32_t n_nodes);
int launch_validate_scene(const RawNode* nodes, uint32_t n_nodes, uint32_t n_models, void* scratch, uint32_t* model_rank,
                          ValidateOut* out, cudaStream_t stream);

// ---- GPU BVH builder (bvh_build.cu) ----
size_t bvh_build_scratch_bytes(uint32_t n_models);
#define BVH_BUILD_LBVH 0   // Karras radix tree over the Morton order (fastest build)
#define BVH_BUILD_PLOC 1   // parallel locally-ordered clustering over the Morton order (the reference's algorithm; better trees)
int launch_bvh_build(const RawModel* models, uint32_t n, RawNode* out_nodes, uint32_t* model_rank, void* scratch,
                     uint32_t** depth_out, int algorithm, int sm_count, cudaStream_t stream);
// same topology as the last launch_bvh_build on this scratch, boxes refitted to the current models
uint32_t* bvh_build_depth_word(void* scratch, uint32_t n_models);   // device word holding the tree depth of the last build / refit
int launch_bvh_refit(const RawModel* models, uint32_t n, RawNode* out_nodes, void* scratch, int algorithm, cudaStream_t stream);

// ---- render kernels ----
int launch_megakernel(const RenderParams& p, cudaStream_t stream);   // simple one-thread-per-pixel kernel (v1)
// staged-shading persistent-lane megakernel (v3); returns -1 when the configuration does not fit (the caller
// falls back to v1)
int launch_megakernel_v3(const RenderParams& p, uint32_t n_inner, uint32_t n_models, uint32_t tree_depth,
                         unsigned int* pixel_counter, int threads, uint32_t shade_wait_lanes, uint32_t leaf_batch_lanes,
                         bool no_both, uint32_t max_top, uint32_t n_hot, bool lean_w4, const uint32_t* tile_order,
                         uint32_t* tile_cost, int sm_count, cudaStream_t stream);
// ---- pixel-queue order (tile_order.cu) ----
// The megakernel hands out 8x4 tiles through a queue.  A pixel is a sequential chain (one RNG stream, raytrace.wgsl:89),
// so the frame ends when the slowest chain does: tiles are handed out heaviest first, judged by the rays each tile cost
// in the PREVIOUS frame of the same size (counting sort over logarithmic cost classes).  scratch:
// tile_order_scratch_bytes(n); cost: n u32 (zeroed by the update).
// mode: 2 = heaviest first, 3 = lightest first (experiment), 1 = reversed row-major (experiment)
size_t tile_order_scratch_bytes(uint32_t n_tiles);
int launch_tile_order_update(uint32_t* tile_cost, uint32_t* tile_order, void* scratch, uint32_t n_tiles, int mode, bool first,
                             cudaStream_t stream);
// BVR_SELFCHECK: re-traces the rays the render kernel logged, in reference order; counters = {checked, disagreements}
int launch_selfcheck(const RenderParams& p, unsigned long long* counters, cudaStream_t stream);
size_t wavefront_state_bytes(size_t slots);
void wavefront_bind(WavefrontParams& w, void* state, size_t slots);
// persistent per-CTA wavefront (cta_wavefront.cu): number of path slots it needs, and the launch
size_t cta_wavefront_slots(int sm_count);
int launch_cta_wavefront(WavefrontParams w, uint32_t n_inner, uint32_t n_models, uint32_t tree_depth, int sm_count,
                         unsigned int* pixel_counter, cudaStream_t stream);
// wavefront pipeline; `host_counts` = 8 pinned words for polling the queue sizes.  -1 = not supported / error
int launch_wavefront(WavefrontParams w, uint32_t n_inner, uint32_t n_models, uint32_t tree_depth, int sm_count,
                     volatile unsigned int* host_counts, cudaStream_t stream);
int launch_copy_raster(const RenderParams& p, cudaStream_t stream);   // level 0: raytrace.wgsl:97-99

// ---- multi-GPU helpers (shard_kernels.cu) ----
int launch_scale(float* dst, float w, size_t n, cudaStream_t stream);
int launch_sum_slots(const float* slots, size_t slot_stride, uint32_t n_slots, unsigned long long mask, float* dst, size_t n,
                     cudaStream_t stream);
int launch_axpby(float* dst, float dst_weight, const float* src, float src_weight, size_t n, cudaStream_t stream);
int launch_composite(float4* rgba, const float* rt_depth, const float4* raster_rgba, const float* raster_depth,
                     const CameraParams& cam, size_t n, cudaStream_t stream);
int launch_unshard(const uint32_t* gathered, size_t shard_stride_words, uint32_t* full, uint32_t width,
                   uint32_t height, uint32_t channels, uint32_t shard_count, uint32_t strip_rows,
                   cudaStream_t stream);

}  // namespace bvr
