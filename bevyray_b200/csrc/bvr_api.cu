// bvr_api.cu — the C ABI (include/bevyray_b200.h): context, scene upload, render dispatch.
//
// Replaces, for the hot path only, RaytracingPipeline::from_world and RayTracingNode::run of the
// reference (src/raytracing/pipeline.rs:58-331).  There is no CPU fallback: without a CUDA device
// bvr_create fails with BVR_ERR_NO_DEVICE.

#include "../../include/bevyray_b200.h"
#include "kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace bvr;

namespace {

struct DeviceBuffer {
    void* ptr = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (ptr) { cudaFree(ptr); ptr = nullptr; cap = 0; }
        size_t want = bytes + bytes / 4 + 256;   // growth slack so that small scene growth does not realloc
        cudaError_t e = cudaMalloc(&ptr, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (ptr) cudaFree(ptr); ptr = nullptr; cap = 0; }
    template <class T> T* as() const { return static_cast<T*>(ptr); }
};

struct PinnedBuffer {
    void* ptr = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (ptr) { cudaFreeHost(ptr); ptr = nullptr; cap = 0; }
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaHostAlloc(&ptr, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (ptr) cudaFreeHost(ptr); ptr = nullptr; cap = 0; }
};

}  // namespace

// Tuning knobs for experiments, read from the environment ONCE per context (bvr_create) and again
// only on request (bvr_reload_tuning); production uses the defaults.
struct EnvTuning {
    int no_tight = 0, tight_pad = 100, no_q16 = 0, no_bvh4 = 0, gpu_validate = -1, wf_refill = 8;
    int mk_v1 = 0, mk_threads = 0, mk_wait = 0, mk_leaf = 0, selfcheck = 0, no_top = 0, top_records = 0, hot_records = 512, w4_lean = 1, tile_order = -1, no_both = 0, gpu_lbvh = 0;
};

struct BvrContext {
    int device = 0;
    EnvTuning tune;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string error;

    // scene, reference layout (raw bytes in HBM) and the derived traversal layout
    DeviceBuffer raw_models, raw_materials, raw_nodes;
    DeviceBuffer spheres, sphere_material, pairs, pairs_ch, inner_id, block_sums, root_ref;
    DeviceBuffer model_rank;                  // position of every model in the reference's traversal order
    DeviceBuffer validate_scratch, validate_out;   // GPU-side validation of big node arrays (scene_validate.cu)
    ValidateOut* validate_host = nullptr;     // pinned
    DeviceBuffer nodes4_ch;                   // 4-wide fp32 records (scenes staged in shared memory)
    bool nodes4_ch_built = false;
    DeviceBuffer raw_nodes_tight, pairs_tight, pairs_ch_tight, nodes4_tight, tight_groups;   // tight-box variant
    bool tight_built = false;
    DeviceBuffer id_q;                                  // record numbering of the quantised layouts (hot top of the tree first)
    DeviceBuffer pairs_q, nodes4_q, qgrid;              // 32-byte quantised records + their grid (big scenes), flag at qgrid[8]
    unsigned int* q16_bad_host = nullptr;     // pinned copy of the 'does not qualify' flag
    bool q16_built = false, q16_pending = false;
    cudaEvent_t q16_done = nullptr;
    size_t n_models = 0, n_materials = 0, n_nodes = 0;
    bool scene_uploaded = false;
    bool has_scene = false;
    uint32_t root_ref_host = 0;
    uint32_t tree_depth = 0;
    float scene_extent = 0.0f;                // largest |coordinate| of the root box (guards the culling arithmetic)
    int gpu_bvh_algorithm = 0;                // BVH_BUILD_* of the last bvr_upload_scene_gpu_bvh (its topology is in bvh_scratch)
    bool tree_is_ours = false;                // built by bvr_upload_scene_gpu_bvh: the reference never saw this tree
    uint32_t n_inner = 0;
    uint32_t max_leaf_models = 0;
    int sm_count = 0;
    DeviceBuffer pixel_counter;
    DeviceBuffer tile_order, tile_cost, tile_scratch;   // pixel-queue order from the previous frame's per-tile ray counts
    uint32_t tile_geom[4] = {0, 0, 0, 0};               // (width, rows, tiles, mode) the order was made for; 0 = none yet
    bool tile_order_valid = false;
    DeviceBuffer wf_state;
    DeviceBuffer bvh_scratch;
    unsigned int* depth_host = nullptr;       // pinned
    BvrBvhNode* root_host = nullptr;          // pinned: node 0 of a GPU-built tree
    unsigned int* wf_host_counts = nullptr;   // pinned, 8 words
    PinnedBuffer upload_staging;
    cudaEvent_t upload_done = nullptr;
    bool upload_pending = false;
    cudaEvent_t build_done = nullptr;         // GPU BVH build finished (depth_host / root_host valid)
    bool build_pending = false;
    bool pinned_src_in_flight = false;        // an upload copies straight out of a pinned caller buffer

    // per-frame IO for the host-buffer entry point
    DeviceBuffer in_rgba, in_depth, out_rgba, out_rt_depth, out_id, out_pdepth, out_srgb8;
    PinnedBuffer io_staging;

    DeviceBuffer selfcheck_log;               // BVR_SELFCHECK: logged rays + their count
    DeviceBuffer ray_counter;                 // [0] rays, [1] self-check rays, [2] self-check disagreements
    unsigned long long* ray_counter_host = nullptr;   // pinned, 3 words
    cudaEvent_t ev_render0 = nullptr, ev_render1 = nullptr, ev_upload0 = nullptr, ev_upload1 = nullptr;
    bool render_timed = false, upload_timed = false;
    BvrStats stats{};
};

namespace {

int fail(BvrContext* ctx, int status, const std::string& msg) {
    if (ctx) ctx->error = msg;
    return status;
}

int fail_cuda(BvrContext* ctx, cudaError_t e, const char* what);

// bvr_upload_scene_gpu_bvh leaves the tree depth and the root box in pinned memory behind an event
int resolve_gpu_build(BvrContext* ctx) {
    if (!ctx->build_pending) return BVR_OK;
    cudaError_t e = cudaEventSynchronize(ctx->build_done);
    if (e != cudaSuccess) return fail_cuda(ctx, e, "GPU BVH build");
    ctx->build_pending = false;
    ctx->tree_depth = *ctx->depth_host;
    float m = 0.0f;
    for (int k = 0; k < 3; k++) {
        const float a = std::fabs(ctx->root_host->bounds_min[k]), b = std::fabs(ctx->root_host->bounds_max[k]);
        if (!(a <= m)) m = a;
        if (!(b <= m)) m = b;
    }
    ctx->scene_extent = m;
    return BVR_OK;
}

int fail_cuda(BvrContext* ctx, cudaError_t e, const char* what) {
    std::string msg = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    cudaGetLastError();   // clear the sticky-less error state
    return fail(ctx, e == cudaErrorMemoryAllocation ? BVR_ERR_OUT_OF_MEMORY : BVR_ERR_CUDA, msg);
}

#define BVR_CK(expr)                                                         \
    do {                                                                     \
        cudaError_t e__ = (expr);                                            \
        if (e__ != cudaSuccess) return fail_cuda(ctx, e__, #expr);           \
    } while (0)

bool is_pinned_host(const void* p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return attr.type == cudaMemoryTypeHost;
}

int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

EnvTuning read_env_tuning() {
    EnvTuning t;
    t.no_tight = env_int("BVR_NO_TIGHT", 0);
    t.tight_pad = env_int("BVR_TIGHT_PAD", 100);
    t.no_q16 = env_int("BVR_NO_Q16", 0);
    t.no_bvh4 = env_int("BVR_NO_BVH4", 0);
    t.gpu_validate = env_int("BVR_GPU_VALIDATE", -1);
    t.wf_refill = env_int("BVR_WF_REFILL", 8);
    t.mk_v1 = env_int("BVR_MK_V1", 0);
    t.mk_threads = env_int("BVR_MK_THREADS", 0);
    t.mk_wait = env_int("BVR_MK_WAIT", 0);
    t.mk_leaf = env_int("BVR_MK_LEAF", 0);
    t.selfcheck = env_int("BVR_SELFCHECK", 0);
    t.no_top = env_int("BVR_NO_TOP", 0);
    t.top_records = env_int("BVR_TOP_RECORDS", 0);
    t.w4_lean = env_int("BVR_W4_LEAN", 1);   // 0: the general traversal loop on the 4-wide 16-bit records (MODE 3 instead of 7)
    t.hot_records = env_int("BVR_HOT_RECORDS", 512);   // 32 KB of records: what the stacks leave of L1 (profiles/r02_tuning_sweeps.txt)
    t.no_both = env_int("BVR_NO_BOTH", 0);
    // -1 (default): heaviest tile of the previous frame first when the frame is big enough to repay the sort;
    // 0 row-major, 2 heaviest first whatever the size (1, 3: experiments — reversed row-major, lightest first)
    t.tile_order = env_int("BVR_TILE_ORDER", -1);
    t.gpu_lbvh = env_int("BVR_GPU_LBVH", 0);
    return t;
}

// The culling-only slab test computes t = c * (1/d) - o * (1/d): unlike the reference's (box - o) * (1/d) its rounding
// error grows with |c| + |o|, about (|c| + |o|) * 2^-22 in world units.  That must stay a small fraction of the pad that
// separates a box from its spheres — 0.01 for the tight boxes, 0.1 for the reference's — so scenes (or cameras) far from
// the origin drop the tight boxes first and the fast slab test altogether after that.
constexpr float BVR_TIGHT_MAX_EXTENT = 4096.0f;     // 2 * 4096 * 2^-22 = 0.002 = pad / 5
constexpr float BVR_FAST_MAX_EXTENT = 32768.0f;     // 2 * 32768 * 2^-22 = 0.016 = pad / 6

float root_extent(const BvrBvhNode& r) {
    float m = 0.0f;
    for (int k = 0; k < 3; k++) {
        const float a = std::fabs(r.bounds_min[k]), b = std::fabs(r.bounds_max[k]);
        if (!(a <= m)) m = a;     // NaN counts as "too far"
        if (!(b <= m)) m = b;
    }
    return m;
}

uint32_t effective_strip_rows(const BvrRenderOptions* o) { return (o && o->strip_rows) ? o->strip_rows : 8u; }

uint32_t shard_rows_impl(uint32_t height, const BvrRenderOptions* o) {
    const uint32_t count = (o && o->shard_count > 1) ? o->shard_count : 1u;
    if (count == 1u) return height;
    const uint32_t r = effective_strip_rows(o);
    const uint32_t strips = (height + r - 1) / r;
    const uint32_t per = (strips + count - 1) / count;
    return per * r;
}

// Host-side structural validation of the node array against the reference contract
// (raytrace.wgsl:80-87, 313-346).  EVERY node of the array is range-checked and may be referenced at most once —
// the derive kernels (scene_kernels.cu) run over all nodes, reachable or not, and the GPU validator
// (scene_validate.cu: gv_link) applies the same rule, so both paths accept and reject the same arrays.
// The walk from node 0 then yields the tree depth (stack bound) and the reference's order over the leaves.
// n_inner counts ALL inner nodes of the array: it sizes the derived record arrays.
int validate_scene(BvrContext* ctx, const BvrModel* models, size_t n_models, size_t n_materials,
                   const BvrBvhNode* nodes, size_t n_nodes, uint32_t* depth_out, uint32_t* n_inner_out,
                   uint32_t* max_leaf_out, std::vector<uint32_t>* rank_out) {
    (void)models;
    *depth_out = 0;
    *n_inner_out = 0;
    *max_leaf_out = 0;
    rank_out->assign(n_models, 0xffffffffu);
    uint32_t next_rank = 0;
    if (n_models > 0 && n_materials == 0) return fail(ctx, BVR_ERR_BAD_SCENE, "models without materials");
    if (n_models >= (size_t)BVR_LEAF_FIRST_MASK) return fail(ctx, BVR_ERR_BAD_SCENE, "more than 2^24-1 models");
    if (n_nodes == 0) return BVR_OK;
    if (n_nodes >= 0x7fffffffull) return fail(ctx, BVR_ERR_BAD_SCENE, "too many BVH nodes");
    std::vector<uint8_t> seen(n_nodes, 0);
    for (size_t i = 0; i < n_nodes; i++) {
        const BvrBvhNode& nd = nodes[i];
        if (nd.model_count > 0) {
            if (nd.model_count > BVR_MAX_LEAF_COUNT) return fail(ctx, BVR_ERR_BAD_SCENE, "leaf with more than 128 models");
            if ((size_t)nd.index + nd.model_count > n_models) return fail(ctx, BVR_ERR_BAD_SCENE, "leaf model range out of bounds");
        } else {
            (*n_inner_out)++;
            if ((size_t)nd.index + 1 >= n_nodes) return fail(ctx, BVR_ERR_BAD_SCENE, "child index out of bounds");
            for (uint32_t c = nd.index; c < nd.index + 2; c++) {
                // node 0 is the root: a reference to it closes a cycle
                if (c == 0u || seen[c]) return fail(ctx, BVR_ERR_BAD_SCENE, "BVH node referenced twice (cycle or DAG)");
                seen[c] = 1;
            }
        }
    }
    std::vector<std::pair<uint32_t, uint32_t>> stack;   // (node, depth)
    stack.push_back({0u, 1u});
    uint32_t max_depth = 0;
    while (!stack.empty()) {
        const auto [i, d] = stack.back();
        stack.pop_back();
        if (d > max_depth) max_depth = d;
        const BvrBvhNode& nd = nodes[i];
        if (nd.model_count > 0) {
            if (nd.model_count > *max_leaf_out) *max_leaf_out = nd.model_count;
            // this walk pops index+1 before index, like raytrace.wgsl:329-341: leaves come off in the reference's order
            for (uint32_t m = nd.index; m < nd.index + nd.model_count; m++)
                if ((*rank_out)[m] == 0xffffffffu) (*rank_out)[m] = next_rank++;
        } else {
            // (every node has at most one parent and node 0 none, so this walk is over a tree: it terminates)
            for (uint32_t c = nd.index; c < nd.index + 2; c++) stack.push_back({c, d + 1});
        }
    }
    *depth_out = max_depth;
    return BVR_OK;
}

// copy `bytes` from host memory to device, via pinned staging unless the source is already pinned
int h2d(BvrContext* ctx, void* dst, const void* src, size_t bytes, PinnedBuffer& staging, size_t& staging_off) {
    if (bytes == 0) return BVR_OK;
    if (is_pinned_host(src)) {
        BVR_CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        ctx->pinned_src_in_flight = true;     // the caller's buffer is read by the DMA engine until the copy completes
    } else {
        char* st = static_cast<char*>(staging.ptr) + staging_off;
        std::memcpy(st, src, bytes);
        BVR_CK(cudaMemcpyAsync(dst, st, bytes, cudaMemcpyHostToDevice, ctx->stream));
        staging_off += (bytes + 255) & ~(size_t)255;
    }
    ctx->stats.h2d_bytes += bytes;
    return BVR_OK;
}

// Scenes too large for shared memory also get 32-byte quantised records (one 256-bit load per node visit,
// megakernel_v3.cu MODE 2).  Returns the number of kernels launched; the verdict arrives through q16_done.
int derive_q16(BvrContext* ctx, size_t n_models, size_t n_nodes, uint32_t n_inner, uint32_t max_leaf, int* launches) {
    ctx->q16_built = false;
    ctx->q16_pending = false;
    ctx->nodes4_ch_built = false;
    if (n_inner >= 1u && n_inner <= 1024u && n_models <= 1023u && max_leaf <= 1u) {   // 11-bit refs, 0x7ff = none
        BVR_CK(ctx->nodes4_ch.ensure((size_t)n_inner * 112u));
        *launches += launch_derive_nodes4_ch(ctx->raw_nodes.as<RawNode>(), (uint32_t)n_nodes, ctx->inner_id.as<uint32_t>(),
                                             ctx->pairs_ch.as<float4>(), ctx->nodes4_ch.as<float4>(), ctx->stream);
        ctx->nodes4_ch_built = true;
        // tight boxes: pad 0.01 instead of the reference's 0.1; rays that start too far from the spheres for
        // the tight boxes to be safe walk the reference records (scene_kernels.cu)
        ctx->tight_built = false;
        if (!ctx->tune.no_tight && n_nodes <= 2047u) {
            BVR_CK(ctx->raw_nodes_tight.ensure(n_nodes * sizeof(RawNode)));
            BVR_CK(ctx->pairs_tight.ensure(n_nodes * 2 * sizeof(float4)));
            BVR_CK(ctx->pairs_ch_tight.ensure(n_nodes * 2 * sizeof(float4)));
            BVR_CK(ctx->nodes4_tight.ensure((size_t)n_inner * 112u));
            BVR_CK(ctx->tight_groups.ensure(33 * sizeof(float4)));
            *launches += launch_derive_tight(ctx->raw_nodes.as<RawNode>(), (uint32_t)n_nodes, ctx->spheres.as<float4>(),
                                             (uint32_t)n_models, 1e-4f * (float)ctx->tune.tight_pad, ctx->tight_groups.as<float4>(),
                                             ctx->raw_nodes_tight.as<RawNode>(), ctx->stream);
            *launches += launch_derive_pairs(ctx->raw_nodes_tight.as<RawNode>(), (uint32_t)n_nodes, ctx->inner_id.as<uint32_t>(),
                                             ctx->block_sums.as<uint32_t>(), ctx->pairs_tight.as<float4>(),
                                             ctx->pairs_ch_tight.as<float4>(), ctx->root_ref.as<uint32_t>(), ctx->stream);
            *launches += launch_derive_nodes4_ch(ctx->raw_nodes_tight.as<RawNode>(), (uint32_t)n_nodes, ctx->inner_id.as<uint32_t>(),
                                                 ctx->pairs_ch_tight.as<float4>(), ctx->nodes4_tight.as<float4>(), ctx->stream);
            ctx->tight_built = true;
        }
    }
    const size_t scene_bytes = (size_t)n_inner * 64u + n_models * 20u;
    if (ctx->tune.no_q16 || scene_bytes <= 160u * 1024u || n_inner > (1u << 20) || n_models > (1u << 20) || max_leaf > 1u)
        return BVR_OK;
    BVR_CK(ctx->pairs_q.ensure((size_t)n_inner * 32u + 32u));
    BVR_CK(ctx->qgrid.ensure(16 * sizeof(float)));
    uint32_t* bad = ctx->qgrid.as<uint32_t>() + 8;
    // the quantised records are numbered with the hot top of the tree first (the render kernel stages a prefix of the
    // array in shared memory); BVR_NO_TOP keeps the array order of the upload
    const uint32_t* id_q = ctx->inner_id.as<uint32_t>();
    if (!ctx->tune.no_top) {
        BVR_CK(ctx->id_q.ensure((n_nodes + 1u) * sizeof(uint32_t)));
        *launches += launch_derive_top_order(ctx->raw_nodes.as<RawNode>(), (uint32_t)n_nodes, ctx->inner_id.as<uint32_t>(),
                                             ctx->block_sums.as<uint32_t>(), 0xffffffffu, ctx->id_q.as<uint32_t>(), ctx->stream);
        id_q = ctx->id_q.as<uint32_t>();
    }
    *launches += launch_derive_pairs_q16(ctx->raw_nodes.as<RawNode>(), (uint32_t)n_nodes, id_q,
                                         ctx->pairs_q.as<uint4>(), ctx->qgrid.as<float>(), bad, ctx->stream);
    BVR_CK(ctx->nodes4_q.ensure((size_t)n_inner * 64u + 64u));
    *launches += launch_derive_nodes4_q16(ctx->raw_nodes.as<RawNode>(), (uint32_t)n_nodes, id_q,
                                          ctx->pairs_q.as<uint4>(), ctx->nodes4_q.as<uint4>(), ctx->stream);
    BVR_CK(cudaMemcpyAsync(ctx->q16_bad_host, bad, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    BVR_CK(cudaEventRecord(ctx->q16_done, ctx->stream));
    ctx->q16_built = true;
    ctx->q16_pending = true;
    return BVR_OK;
}

}  // namespace

extern "C" {

uint32_t bvr_abi_version(void) { return BVR_ABI_VERSION; }

const char* bvr_status_string(int status) {
    switch (status) {
        case BVR_OK: return "ok";
        case BVR_ERR_INVALID_ARGUMENT: return "invalid argument";
        case BVR_ERR_CUDA: return "CUDA error";
        case BVR_ERR_UNSUPPORTED_PROJECTION: return "unsupported projection (only perspective, projection == 0)";
        case BVR_ERR_NO_SCENE: return "no scene uploaded";
        case BVR_ERR_BAD_SCENE: return "malformed scene";
        case BVR_ERR_OUT_OF_MEMORY: return "out of device memory";
        case BVR_ERR_NO_DEVICE: return "no CUDA device (there is no CPU fallback)";
        default: return "unknown status";
    }
}

int bvr_create(int device, BvrContext** out_ctx) {
    if (!out_ctx) return BVR_ERR_INVALID_ARGUMENT;
    *out_ctx = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) { cudaGetLastError(); return BVR_ERR_NO_DEVICE; }
    if (device < 0 || device >= count) return BVR_ERR_INVALID_ARGUMENT;
    BvrContext* ctx = new (std::nothrow) BvrContext();
    if (!ctx) return BVR_ERR_OUT_OF_MEMORY;
    ctx->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_render0);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_render1);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_upload0);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_upload1);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->upload_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->build_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = ctx->ray_counter.ensure(3 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = ctx->root_ref.ensure(sizeof(uint32_t));
    if (e == cudaSuccess) e = ctx->pixel_counter.ensure(sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&ctx->ray_counter_host, 3 * sizeof(unsigned long long), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&ctx->wf_host_counts, 8 * sizeof(unsigned int), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&ctx->depth_host, sizeof(unsigned int), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&ctx->root_host, sizeof(BvrBvhNode), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&ctx->q16_bad_host, sizeof(unsigned int), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&ctx->validate_host, sizeof(ValidateOut), cudaHostAllocDefault);
    if (e == cudaSuccess) e = ctx->validate_out.ensure(sizeof(ValidateOut));
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->q16_done, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        cudaGetLastError();
        bvr_destroy(ctx);
        return e == cudaErrorMemoryAllocation ? BVR_ERR_OUT_OF_MEMORY : BVR_ERR_CUDA;
    }
    ctx->ray_counter_host[0] = ctx->ray_counter_host[1] = ctx->ray_counter_host[2] = 0;
    ctx->stream = ctx->own_stream;
    ctx->tune = read_env_tuning();
    *out_ctx = ctx;
    return BVR_OK;
}

void bvr_destroy(BvrContext* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    DeviceBuffer* bufs[] = {&ctx->raw_models, &ctx->raw_materials, &ctx->raw_nodes, &ctx->spheres,
                            &ctx->sphere_material, &ctx->pairs, &ctx->pairs_ch, &ctx->inner_id, &ctx->block_sums, &ctx->root_ref,
                            &ctx->in_rgba, &ctx->in_depth, &ctx->out_rgba, &ctx->out_rt_depth, &ctx->out_id,
                            &ctx->out_pdepth, &ctx->out_srgb8, &ctx->ray_counter, &ctx->pixel_counter, &ctx->tile_order, &ctx->tile_cost, &ctx->tile_scratch, &ctx->wf_state, &ctx->bvh_scratch, &ctx->pairs_q, &ctx->nodes4_q, &ctx->qgrid, &ctx->id_q, &ctx->model_rank, &ctx->selfcheck_log, &ctx->validate_scratch, &ctx->validate_out, &ctx->nodes4_ch, &ctx->raw_nodes_tight, &ctx->pairs_tight, &ctx->pairs_ch_tight, &ctx->nodes4_tight, &ctx->tight_groups};
    for (DeviceBuffer* b : bufs) b->release();
    ctx->upload_staging.release();
    ctx->io_staging.release();
    if (ctx->ray_counter_host) cudaFreeHost(ctx->ray_counter_host);
    if (ctx->wf_host_counts) cudaFreeHost(ctx->wf_host_counts);
    if (ctx->q16_bad_host) cudaFreeHost(ctx->q16_bad_host);
    if (ctx->validate_host) cudaFreeHost(ctx->validate_host);
    if (ctx->q16_done) cudaEventDestroy(ctx->q16_done);
    if (ctx->depth_host) cudaFreeHost(ctx->depth_host);
    if (ctx->root_host) cudaFreeHost(ctx->root_host);
    cudaEvent_t evs[] = {ctx->ev_render0, ctx->ev_render1, ctx->ev_upload0, ctx->ev_upload1, ctx->upload_done, ctx->build_done};
    for (cudaEvent_t ev : evs) if (ev) cudaEventDestroy(ev);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    cudaGetLastError();
    delete ctx;
}

const char* bvr_last_error(const BvrContext* ctx) { return ctx ? ctx->error.c_str() : "null context"; }

int bvr_set_stream(BvrContext* ctx, void* cuda_stream) {
    if (!ctx) return BVR_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    BVR_CK(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return BVR_OK;
}

int bvr_reload_tuning(BvrContext* ctx) {
    if (!ctx) return BVR_ERR_INVALID_ARGUMENT;
    ctx->tune = read_env_tuning();
    return BVR_OK;
}

int bvr_sync(BvrContext* ctx) {
    if (!ctx) return BVR_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    BVR_CK(cudaStreamSynchronize(ctx->stream));
    return BVR_OK;
}

int bvr_upload_scene(BvrContext* ctx,
                     const BvrModel* models, size_t n_models,
                     const BvrMaterial* materials, size_t n_materials,
                     const BvrBvhNode* nodes, size_t n_nodes,
                     const BvrDirtyRange* ranges, size_t n_ranges) {
    if (!ctx) return BVR_ERR_INVALID_ARGUMENT;
    ctx->error.clear();
    if ((n_models && !models) || (n_materials && !materials) || (n_nodes && !nodes))
        return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "null array with non-zero count");
    if (n_ranges && !ranges) return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "null ranges with non-zero count");
    cudaSetDevice(ctx->device);
    { const int rs = resolve_gpu_build(ctx); if (rs != BVR_OK) return rs; }

    const bool partial = ranges != nullptr && ctx->scene_uploaded && n_models == ctx->n_models &&
                         n_materials == ctx->n_materials && n_nodes == ctx->n_nodes;
    if (ranges && !partial && ctx->scene_uploaded)
        return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "dirty ranges given but the element counts changed");
    if (ranges && !ctx->scene_uploaded)
        return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "dirty ranges given before any full upload");

    // Small trees are validated on the host before anything is touched; big ones on the GPU, on the uploaded bytes
    // (scene_validate.cu: the host walk costs 80 ms for 2 M nodes).  BVR_GPU_VALIDATE=0/1 forces either.
    const int gv_env = ctx->tune.gpu_validate;
    bool gpu_validate = n_nodes > 0 && (gv_env >= 0 ? gv_env != 0 : n_nodes >= 32768u);
    uint32_t depth = ctx->tree_depth, n_inner = ctx->n_inner, max_leaf = ctx->max_leaf_models;
    std::vector<uint32_t> rank;
    int st = BVR_OK;
    if (!gpu_validate) {
        st = validate_scene(ctx, models, n_models, n_materials, nodes, n_nodes, &depth, &n_inner, &max_leaf, &rank);
        if (st != BVR_OK) return st;
    } else {
        if (n_models > 0 && n_materials == 0) return fail(ctx, BVR_ERR_BAD_SCENE, "models without materials");
        if (n_models >= (size_t)BVR_LEAF_FIRST_MASK) return fail(ctx, BVR_ERR_BAD_SCENE, "more than 2^24-1 models");
        if (n_nodes >= 0x7fffffffull) return fail(ctx, BVR_ERR_BAD_SCENE, "too many BVH nodes");
    }

    // the ranges to copy
    struct Copy { uint32_t array; size_t first, count; };
    std::vector<Copy> copies;
    const size_t counts[3] = {n_models, n_materials, n_nodes};
    const size_t strides[3] = {sizeof(BvrModel), sizeof(BvrMaterial), sizeof(BvrBvhNode)};
    if (partial) {
        for (size_t i = 0; i < n_ranges; i++) {
            const BvrDirtyRange& r = ranges[i];
            if (r.array > 2u) return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "dirty range: bad array id");
            if ((size_t)r.first + r.count > counts[r.array])
                return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "dirty range out of bounds");
            if (r.count) copies.push_back({r.array, r.first, r.count});
        }
    } else {
        for (uint32_t a = 0; a < 3; a++) if (counts[a]) copies.push_back({a, 0, counts[a]});
    }

    BVR_CK(ctx->raw_models.ensure(n_models * sizeof(BvrModel)));
    BVR_CK(ctx->raw_materials.ensure(n_materials * sizeof(BvrMaterial)));
    BVR_CK(ctx->raw_nodes.ensure(n_nodes * sizeof(BvrBvhNode)));
    BVR_CK(ctx->spheres.ensure(n_models * sizeof(float4)));
    BVR_CK(ctx->sphere_material.ensure(n_models * sizeof(uint32_t)));
    BVR_CK(ctx->model_rank.ensure(n_models * sizeof(uint32_t)));
    BVR_CK(ctx->pairs.ensure(n_nodes * 2 * sizeof(float4)));   // <= ceil(n/2) inner nodes x 64 B
    BVR_CK(ctx->pairs_ch.ensure(n_nodes * 2 * sizeof(float4)));
    BVR_CK(ctx->inner_id.ensure(n_nodes * sizeof(uint32_t)));
    BVR_CK(ctx->block_sums.ensure((n_nodes / 1024 + 2) * sizeof(uint32_t)));

    // stage + copy
    size_t staging_bytes = 0;
    for (const Copy& c : copies) staging_bytes += ((c.count * strides[c.array]) + 255) & ~(size_t)255;
    staging_bytes += (n_models * sizeof(uint32_t) + 255) & ~(size_t)255;   // model ranks
    if (ctx->upload_pending) { BVR_CK(cudaEventSynchronize(ctx->upload_done)); ctx->upload_pending = false; }
    BVR_CK(ctx->upload_staging.ensure(staging_bytes));
    BVR_CK(cudaEventRecord(ctx->ev_upload0, ctx->stream));
    const void* src_base[3] = {models, materials, nodes};
    void* dst_base[3] = {ctx->raw_models.ptr, ctx->raw_materials.ptr, ctx->raw_nodes.ptr};
    size_t off = 0;
    bool models_dirty = false, nodes_dirty = false;
    for (const Copy& c : copies) {
        const size_t stride = strides[c.array];
        st = h2d(ctx, static_cast<char*>(dst_base[c.array]) + c.first * stride,
                 static_cast<const char*>(src_base[c.array]) + c.first * stride, c.count * stride,
                 ctx->upload_staging, off);
        if (st != BVR_OK) return st;
        if (c.array == BVR_ARRAY_MODELS) models_dirty = true;
        if (c.array == BVR_ARRAY_BVH_NODES) nodes_dirty = true;
    }
    int launches = 0;
    if (gpu_validate && (nodes_dirty || !partial)) {
        BVR_CK(ctx->validate_scratch.ensure(validate_scratch_bytes((uint32_t)n_nodes)));
        launches += launch_validate_scene(ctx->raw_nodes.as<RawNode>(), (uint32_t)n_nodes, (uint32_t)n_models,
                                          ctx->validate_scratch.ptr, ctx->model_rank.as<uint32_t>(),
                                          ctx->validate_out.as<ValidateOut>(), ctx->stream);
        BVR_CK(cudaMemcpyAsync(ctx->validate_host, ctx->validate_out.ptr, sizeof(ValidateOut), cudaMemcpyDeviceToHost, ctx->stream));
        BVR_CK(cudaStreamSynchronize(ctx->stream));   // the verdict, the stack bound and the leaf size are needed now
        const ValidateOut v = *ctx->validate_host;
        if (v.bad) {
            ctx->scene_uploaded = false;              // the previous scene's bytes are already overwritten
            ctx->has_scene = false;
            return fail(ctx, BVR_ERR_BAD_SCENE,
                        (v.bad & BVR_VALIDATE_TWICE) ? "BVH node referenced twice (cycle or DAG)"
                        : (v.bad & BVR_VALIDATE_CHILD_RANGE) ? "child index out of bounds"
                        : (v.bad & BVR_VALIDATE_LEAF_RANGE) ? "leaf model range out of bounds"
                                                            : "leaf with more than 128 models");
        }
        depth = v.depth; n_inner = v.n_inner; max_leaf = v.max_leaf;
        if (depth > BVR_VALIDATE_MAX_DEPTH) {
            // a degenerate, chain-like tree: the ranks come from the host walk after all
            st = validate_scene(ctx, models, n_models, n_materials, nodes, n_nodes, &depth, &n_inner, &max_leaf, &rank);
            if (st != BVR_OK) return st;
            gpu_validate = false;
        }
    }
    if (!gpu_validate && (nodes_dirty || !partial) && n_models) {
        // derived on the host from the node array (the walk above): 4 bytes per model travel with every new tree
        st = h2d(ctx, ctx->model_rank.ptr, rank.data(), n_models * sizeof(uint32_t), ctx->upload_staging, off);
        if (st != BVR_OK) return st;
    }
    BVR_CK(cudaEventRecord(ctx->upload_done, ctx->stream));
    ctx->upload_pending = true;
    if (ctx->pinned_src_in_flight) {
        // the caller owns its arrays again as soon as the call returns (include/bevyray_b200.h): wait for the copies
        BVR_CK(cudaEventSynchronize(ctx->upload_done));
        ctx->upload_pending = false;
        ctx->pinned_src_in_flight = false;
    }

    // derive the traversal layout in HBM
    if (models_dirty || !partial)
        launches += launch_derive_spheres(ctx->raw_models.as<RawModel>(), (uint32_t)n_models,
                                          ctx->spheres.as<float4>(), ctx->sphere_material.as<uint32_t>(), ctx->stream);
    if (nodes_dirty || !partial) {
        launches += launch_derive_pairs(ctx->raw_nodes.as<RawNode>(), (uint32_t)n_nodes, ctx->inner_id.as<uint32_t>(),
                                        ctx->block_sums.as<uint32_t>(), ctx->pairs.as<float4>(), ctx->pairs_ch.as<float4>(),
                                        ctx->root_ref.as<uint32_t>(), ctx->stream);
        // root ref on the host (needed as a kernel parameter): a leaf root is encoded like any leaf ref
        if (n_nodes) {
            const BvrBvhNode& r = nodes[0];
            ctx->scene_extent = root_extent(r);
            ctx->root_ref_host = r.model_count > 0
                                     ? (BVR_LEAF_BIT | ((r.model_count - 1u) << 24) | (r.index & BVR_LEAF_FIRST_MASK))
                                     : 0u;   // node 0 is the first inner node -> dense id 0
        }
    }
    if (nodes_dirty || models_dirty || !partial) {
        st = derive_q16(ctx, n_models, n_nodes, n_inner, max_leaf, &launches);
        if (st != BVR_OK) return st;
    }
    BVR_CK(cudaGetLastError());
    BVR_CK(cudaEventRecord(ctx->ev_upload1, ctx->stream));
    ctx->upload_timed = true;
    ctx->stats.kernel_launches += (uint64_t)launches;
    ctx->n_models = n_models;
    ctx->n_materials = n_materials;
    ctx->n_nodes = n_nodes;
    ctx->tree_depth = depth;
    ctx->n_inner = n_inner;
    ctx->max_leaf_models = max_leaf;
    if (nodes_dirty || !partial) ctx->tree_is_ours = false;
    ctx->has_scene = n_nodes > 0 && n_models > 0;
    ctx->scene_uploaded = true;
    return BVR_OK;
}

static int upload_gpu_bvh_impl(BvrContext* ctx,
                               const BvrModel* models, size_t n_models,
                               const BvrMaterial* materials, size_t n_materials,
                               const BvrDirtyRange* ranges, size_t n_ranges,
                               BvrBvhNode* out_nodes, bool refit) {
    if (!ctx) return BVR_ERR_INVALID_ARGUMENT;
    ctx->error.clear();
    cudaSetDevice(ctx->device);
    { const int rs = resolve_gpu_build(ctx); if (rs != BVR_OK) return rs; }   // (the pinned read-back words are reused below)
    if (refit && !(ctx->scene_uploaded && ctx->tree_is_ours && n_models == ctx->n_models && n_materials == ctx->n_materials))
        return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "refit needs a tree built by bvr_upload_scene_gpu_bvh for the same counts");
    if ((n_models && !models) || (n_materials && !materials))
        return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "null array with non-zero count");
    if (n_ranges && !ranges) return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "null ranges with non-zero count");
    if (n_models > 0 && n_materials == 0) return fail(ctx, BVR_ERR_BAD_SCENE, "models without materials");
    if (n_models >= (size_t)BVR_LEAF_FIRST_MASK) return fail(ctx, BVR_ERR_BAD_SCENE, "more than 2^24-1 models");
    cudaSetDevice(ctx->device);
    const size_t n_nodes = n_models ? 2 * n_models - 1 : 0;
    const bool partial = ranges != nullptr && ctx->scene_uploaded && n_models == ctx->n_models &&
                         n_materials == ctx->n_materials;
    if (ranges && !partial) return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "dirty ranges need a previous upload with the same counts");

    struct Copy { uint32_t array; size_t first, count; };
    std::vector<Copy> copies;
    const size_t counts[2] = {n_models, n_materials};
    if (partial) {
        for (size_t i = 0; i < n_ranges; i++) {
            const BvrDirtyRange& r = ranges[i];
            if (r.array > 1u) return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "dirty range: only models / materials may be named");
            if ((size_t)r.first + r.count > counts[r.array]) return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "dirty range out of bounds");
            if (r.count) copies.push_back({r.array, r.first, r.count});
        }
    } else {
        for (uint32_t a = 0; a < 2; a++) if (counts[a]) copies.push_back({a, 0, counts[a]});
    }
    BVR_CK(ctx->raw_models.ensure(n_models * sizeof(BvrModel)));
    BVR_CK(ctx->raw_materials.ensure(n_materials * sizeof(BvrMaterial)));
    BVR_CK(ctx->raw_nodes.ensure(n_nodes * sizeof(BvrBvhNode)));
    BVR_CK(ctx->spheres.ensure(n_models * sizeof(float4)));
    BVR_CK(ctx->sphere_material.ensure(n_models * sizeof(uint32_t)));
    BVR_CK(ctx->model_rank.ensure(n_models * sizeof(uint32_t)));
    BVR_CK(ctx->pairs.ensure(n_nodes * 2 * sizeof(float4)));
    BVR_CK(ctx->pairs_ch.ensure(n_nodes * 2 * sizeof(float4)));
    BVR_CK(ctx->inner_id.ensure(n_nodes * sizeof(uint32_t)));
    BVR_CK(ctx->block_sums.ensure((n_nodes / 1024 + 2) * sizeof(uint32_t)));
    BVR_CK(ctx->bvh_scratch.ensure(bvh_build_scratch_bytes((uint32_t)n_models)));

    size_t staging_bytes = 0;
    for (const Copy& c : copies) staging_bytes += ((c.count * 32) + 255) & ~(size_t)255;
    if (ctx->upload_pending) { BVR_CK(cudaEventSynchronize(ctx->upload_done)); ctx->upload_pending = false; }
    BVR_CK(ctx->upload_staging.ensure(staging_bytes));
    BVR_CK(cudaEventRecord(ctx->ev_upload0, ctx->stream));
    const void* src_base[2] = {models, materials};
    void* dst_base[2] = {ctx->raw_models.ptr, ctx->raw_materials.ptr};
    size_t off = 0;
    for (const Copy& c : copies) {
        int st = h2d(ctx, static_cast<char*>(dst_base[c.array]) + c.first * 32,
                     static_cast<const char*>(src_base[c.array]) + c.first * 32, c.count * 32, ctx->upload_staging, off);
        if (st != BVR_OK) return st;
    }
    BVR_CK(cudaEventRecord(ctx->upload_done, ctx->stream));
    ctx->upload_pending = true;

    int launches = 0;
    uint32_t* d_depth = nullptr;
    *ctx->depth_host = 0;
    if (n_models) {
        if (!refit) ctx->gpu_bvh_algorithm = ctx->tune.gpu_lbvh ? BVH_BUILD_LBVH : BVH_BUILD_PLOC;
        const int nb = refit ? launch_bvh_refit(ctx->raw_models.as<RawModel>(), (uint32_t)n_models, ctx->raw_nodes.as<RawNode>(),
                                                ctx->bvh_scratch.ptr, ctx->gpu_bvh_algorithm, ctx->stream)
                             : launch_bvh_build(ctx->raw_models.as<RawModel>(), (uint32_t)n_models, ctx->raw_nodes.as<RawNode>(),
                                                ctx->model_rank.as<uint32_t>(), ctx->bvh_scratch.ptr, &d_depth,
                                                ctx->gpu_bvh_algorithm, ctx->sm_count, ctx->stream);
        if (nb < 0) return fail_cuda(ctx, cudaGetLastError(), "GPU BVH build");
        if (refit) d_depth = bvh_build_depth_word(ctx->bvh_scratch.ptr, (uint32_t)n_models);
        launches += nb;
        launches += launch_derive_spheres(ctx->raw_models.as<RawModel>(), (uint32_t)n_models, ctx->spheres.as<float4>(),
                                          ctx->sphere_material.as<uint32_t>(), ctx->stream);
        launches += launch_derive_pairs(ctx->raw_nodes.as<RawNode>(), (uint32_t)n_nodes, ctx->inner_id.as<uint32_t>(),
                                        ctx->block_sums.as<uint32_t>(), ctx->pairs.as<float4>(), ctx->pairs_ch.as<float4>(),
                                        ctx->root_ref.as<uint32_t>(), ctx->stream);
        {
            int st = derive_q16(ctx, n_models, n_nodes, (uint32_t)(n_models - 1), 1u, &launches);
            if (st != BVR_OK) return st;
        }
        BVR_CK(cudaMemcpyAsync(ctx->depth_host, d_depth, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
        BVR_CK(cudaMemcpyAsync(ctx->root_host, ctx->raw_nodes.ptr, sizeof(BvrBvhNode), cudaMemcpyDeviceToHost, ctx->stream));
        if (out_nodes) {
            BVR_CK(cudaMemcpyAsync(out_nodes, ctx->raw_nodes.ptr, n_nodes * sizeof(BvrBvhNode), cudaMemcpyDeviceToHost, ctx->stream));
            ctx->stats.d2h_bytes += n_nodes * sizeof(BvrBvhNode);
        }
    }
    BVR_CK(cudaGetLastError());
    BVR_CK(cudaEventRecord(ctx->ev_upload1, ctx->stream));
    // The stack bound (tree depth) and the root box are needed on the host before the next render is LAUNCHED, not
    // before this call returns: they are read back into pinned memory and picked up by resolve_gpu_build(), so that
    // the caller can go on enqueuing (e.g. the next frame's input copies) while the tree is being built.
    BVR_CK(cudaEventRecord(ctx->build_done, ctx->stream));
    ctx->build_pending = n_models > 0;
    ctx->upload_timed = true;
    ctx->stats.kernel_launches += (uint64_t)launches;
    ctx->n_models = n_models;
    ctx->n_materials = n_materials;
    ctx->n_nodes = n_nodes;
    ctx->tree_depth = 0;
    ctx->scene_extent = 0.0f;
    if (out_nodes || ctx->pinned_src_in_flight) {
        // the caller reads out_nodes / owns its pinned arrays again when the call returns
        const int rs = resolve_gpu_build(ctx);
        if (rs != BVR_OK) return rs;
        ctx->pinned_src_in_flight = false;
    }
    ctx->n_inner = n_models ? (uint32_t)(n_models - 1) : 0u;
    ctx->max_leaf_models = n_models ? 1u : 0u;   // the GPU builder emits one sphere per leaf
    ctx->tree_is_ours = true;
    ctx->root_ref_host = (n_models == 1) ? (BVR_LEAF_BIT | 0u) : 0u;
    ctx->has_scene = n_models > 0;
    ctx->scene_uploaded = true;
    return BVR_OK;
}

int bvr_upload_scene_gpu_bvh(BvrContext* ctx, const BvrModel* models, size_t n_models, const BvrMaterial* materials,
                             size_t n_materials, const BvrDirtyRange* ranges, size_t n_ranges, BvrBvhNode* out_nodes) {
    return upload_gpu_bvh_impl(ctx, models, n_models, materials, n_materials, ranges, n_ranges, out_nodes, false);
}

int bvr_refit_scene_gpu_bvh(BvrContext* ctx, const BvrModel* models, size_t n_models, const BvrMaterial* materials,
                            size_t n_materials, const BvrDirtyRange* ranges, size_t n_ranges, BvrBvhNode* out_nodes) {
    return upload_gpu_bvh_impl(ctx, models, n_models, materials, n_materials, ranges, n_ranges, out_nodes, true);
}

uint32_t bvr_shard_rows(uint32_t height, const BvrRenderOptions* opts) { return shard_rows_impl(height, opts); }

int bvr_scene_traversal_ranks(const BvrBvhNode* nodes, size_t n_nodes, size_t n_models, uint32_t* out_ranks,
                              uint32_t* out_depth) {
    if (n_nodes && !nodes) return BVR_ERR_INVALID_ARGUMENT;
    uint32_t depth = 0, n_inner = 0, max_leaf = 0;
    std::vector<uint32_t> rank;
    const int st = validate_scene(nullptr, nullptr, n_models, n_models ? 1 : 0, nodes, n_nodes, &depth, &n_inner, &max_leaf, &rank);
    if (st != BVR_OK) return st;
    if (out_ranks) std::memcpy(out_ranks, rank.data(), n_models * sizeof(uint32_t));
    if (out_depth) *out_depth = depth;
    return BVR_OK;
}

static int build_params(BvrContext* ctx, const BvrCamera* camera, const BvrRaytraceLevel* level,
                        const BvrWindow* window, const BvrRenderOptions* opts, RenderParams* out) {
    if (!camera || !level || !window || !opts) return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "null uniform / options pointer");
    if (!ctx->scene_uploaded) return fail(ctx, BVR_ERR_NO_SCENE, "bvr_render before bvr_upload_scene");
    { const int rs = resolve_gpu_build(ctx); if (rs != BVR_OK) return rs; }
    if (camera->projection != 0u) return fail(ctx, BVR_ERR_UNSUPPORTED_PROJECTION, "camera.projection != 0");
    if (opts->width == 0 || window->height == 0) return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "zero image size");
    if (level->level > 3u) return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "raytrace level > 3");
    if (opts->shard_count > 1 && opts->shard_index >= opts->shard_count)
        return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "shard_index >= shard_count");
    if (opts->kernel > BVR_KERNEL_CTA_WAVEFRONT || opts->traversal > BVR_TRAVERSAL_REFERENCE_ORDER)
        return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "unknown kernel / traversal");

    RenderParams p;
    std::memset(&p, 0, sizeof p);
    p.scene.pairs = ctx->pairs.as<float4>();
    p.scene.pairs_ch = ctx->pairs_ch.as<float4>();
    p.scene.spheres = ctx->spheres.as<float4>();
    p.scene.sphere_material = ctx->sphere_material.as<uint32_t>();
    p.scene.materials = ctx->raw_materials.as<float4>();
    p.scene.model_rank = ctx->n_models ? ctx->model_rank.as<uint32_t>() : nullptr;
    float far_out = ctx->scene_extent;
    for (int k = 0; k < 3; k++) {
        const float a = std::fabs(camera->position[k]);
        if (!(a <= far_out)) far_out = a;
    }
    if (ctx->nodes4_ch_built && !ctx->tune.no_bvh4) {
        p.scene.nodes4_ch = ctx->nodes4_ch.as<float4>();
        if (ctx->tight_built && !ctx->tune.no_tight && far_out <= BVR_TIGHT_MAX_EXTENT) {
            p.scene.nodes4_tight = ctx->nodes4_tight.as<float4>();
            p.scene.tight_groups = ctx->tight_groups.as<float4>();
        }
    }
    if (ctx->q16_built) {
        if (ctx->q16_pending) { BVR_CK(cudaEventSynchronize(ctx->q16_done)); ctx->q16_pending = false; }
        if (*ctx->q16_bad_host == 0u) {
            p.scene.pairs_q = ctx->pairs_q.as<uint4>();
            p.scene.nodes4_q = ctx->tune.no_bvh4 ? nullptr : ctx->nodes4_q.as<uint4>();
            p.scene.qgrid = ctx->qgrid.as<float>();
        }
    }
    p.scene.root_ref = ctx->root_ref_host;
    p.scene.n_materials = (uint32_t)ctx->n_materials;
    p.scene.has_scene = ctx->has_scene ? 1u : 0u;

    CameraParams& c = p.cam;
    c.position = V3{camera->position[0], camera->position[1], camera->position[2]};
    c.direction = V3{camera->direction[0], camera->direction[1], camera->direction[2]};
    c.up = V3{camera->up[0], camera->up[1], camera->up[2]};
    // cross(direction, up), raytrace.wgsl:149 — plain f32 mul/sub, compiled without FMA contraction
    {
        const volatile float dx = c.direction.x, dy = c.direction.y, dz = c.direction.z;
        const volatile float ux = c.up.x, uy = c.up.y, uz = c.up.z;
        const volatile float a0 = dy * uz, a1 = dz * uy, b0 = dz * ux, b1 = dx * uz, c0 = dx * uy, c1 = dy * ux;
        c.right = V3{a0 - a1, b0 - b1, c0 - c1};
    }
    c.aspect = camera->aspect;
    c.tan_half_fov = (float)std::tan((double)(camera->fov * 0.5f));       // raytrace.wgsl:151
    {
        const volatile float h = (float)window->height;
        const volatile float w = h * camera->aspect;                          // raytrace.wgsl:142
        c.inv_width = 1.0f / w;
        c.inv_height = 1.0f / h;
    }
    c.near_plane = camera->near_plane;
    c.far_plane = camera->far_plane;
    c.fallback_far = (level->level == 1u) ? camera->far_plane + 10.0f : camera->far_plane - 1.0f;
    c.seed_scaled = window->random_seed * 10000.0f;
    c.sample_count = camera->sample_count;
    c.bounce_count = camera->bounce_count;
    // BVR_RENDER_DEFER_COMPOSITE: the kernels skip the depth composite (they see level Pure) but keep the level's
    // fallback depth for misses (raytrace.wgsl:177-182); bvr_composite_device applies it later
    c.level = (opts->flags & BVR_RENDER_DEFER_COMPOSITE) && level->level != 0u ? 3u : level->level;
    c.width = opts->width;
    c.height = window->height;

    p.shard.count = opts->shard_count > 1 ? opts->shard_count : 1u;
    p.shard.index = opts->shard_count > 1 ? opts->shard_index : 0u;
    p.shard.strip_rows = p.shard.count > 1 ? effective_strip_rows(opts) : window->height;
    p.shard.rows = shard_rows_impl(window->height, opts);
    p.ray_counter = ctx->ray_counter.as<unsigned long long>();
    // The reference abandons a traversal when its stack index reaches 32 (raytrace.wgsl:320), which a tree of 32 or
    // more levels can trigger (an inner node on level k leaves k+1 entries when both children are entered).  For an
    // UPLOADED tree that deep the verbatim reference-order kernel runs, truncation included, so the image stays the
    // reference's.  A tree the library built itself (bvr_upload_scene_gpu_bvh) was never seen by the reference:
    // truncating there would only lose hits, so it is always walked completely (near-first, one stack entry per level).
    const bool may_truncate = ctx->tree_depth >= BVR_REF_STACK;
    p.reference_order = (opts->traversal == BVR_TRAVERSAL_REFERENCE_ORDER || (may_truncate && !ctx->tree_is_ours) ||
                         ctx->tree_depth + 1 > BVR_FAST_STACK) ? 1u : 0u;
    p.strict_slab = far_out <= BVR_FAST_MAX_EXTENT ? 0u : 1u;   // too far from the origin for the fast box arithmetic
    p.out_weight = opts->output_weight == 0.0f ? 1.0f : opts->output_weight;
    p.weight_in_kernel = 0u;
    p.extra_modulus = p.extra_phase = p.extra_count = 0u;
    {
        const uint32_t spp = camera->sample_count;
        p.inv_pow2_samples = (spp != 0u && (spp & (spp - 1u)) == 0u && !(opts->flags & BVR_RENDER_EXTRA_SAMPLE))
                                 ? 1.0f / (float)spp : 0.0f;   // exact: a power of two up to 2^31
    }
    if (opts->flags & BVR_RENDER_EXTRA_SAMPLE) {
        p.extra_modulus = (opts->flags >> 8) & 0xffu;
        p.extra_phase = (opts->flags >> 16) & 0xffu;
        p.extra_count = (opts->flags >> 24) & 0xffu;
        if (p.extra_modulus == 0u || p.extra_count > p.extra_modulus || camera->sample_count == 0u)
            return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "BVR_RENDER_EXTRA_SAMPLE: modulus 1..255, count <= modulus, sample_count >= 1");
        if (opts->kernel == BVR_KERNEL_WAVEFRONT || opts->kernel == BVR_KERNEL_CTA_WAVEFRONT || p.reference_order || p.strict_slab ||
            ctx->tune.mk_v1 || p.cam.level == 0u)
            return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "BVR_RENDER_EXTRA_SAMPLE needs the persistent megakernel (near-first traversal)");
    }
    *out = p;
    return BVR_OK;
}

static int render_device_impl(BvrContext* ctx, const BvrCamera* camera, const BvrRaytraceLevel* level,
                              const BvrWindow* window, const BvrRenderOptions* opts,
                              const float* d_raster_rgba, const float* d_raster_depth, const BvrOutputs* out) {
    RenderParams p;
    int st = build_params(ctx, camera, level, window, opts, &p);
    if (st != BVR_OK) return st;
    if (!out) return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "null outputs");
    if (p.cam.level <= 2u && !d_raster_rgba) return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "levels 0-2 need the raster colour");
    if ((p.cam.level == 1u || p.cam.level == 2u) && !d_raster_depth)
        return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "levels 1-2 need the raster depth");
    p.raster_rgba = reinterpret_cast<const float4*>(d_raster_rgba);
    p.raster_depth = d_raster_depth;
    p.out_rgba = reinterpret_cast<float4*>(out->rgba);
    p.out_rt_depth = out->rt_depth;
    p.out_primary_id = out->primary_id;
    p.out_primary_depth = out->primary_depth;
    p.out_srgb8 = reinterpret_cast<uchar4*>(out->srgb8);

    BVR_CK(cudaMemsetAsync(ctx->ray_counter.ptr, 0, 3 * sizeof(unsigned long long), ctx->stream));
    // the self-check walks the uploaded tree in reference order: pointless (and truncating) on trees the reference-order
    // kernel renders anyway
    p.selfcheck_log = nullptr;
    if (ctx->tune.selfcheck && !p.reference_order) {
        const uint32_t cap = 1u << 20;                       // 32 MB of log; a frame of 1 G rays logs about that many
        BVR_CK(ctx->selfcheck_log.ensure((size_t)cap * 32u + 16u));
        p.selfcheck_count = reinterpret_cast<unsigned int*>(ctx->selfcheck_log.as<char>() + (size_t)cap * 32u);
        BVR_CK(cudaMemsetAsync(p.selfcheck_count, 0, sizeof(unsigned int), ctx->stream));
        p.selfcheck_log = ctx->selfcheck_log.as<float4>();
        p.selfcheck_cap = cap;
    }
    BVR_CK(cudaEventRecord(ctx->ev_render0, ctx->stream));
    int launches = 0;
    if (p.cam.level == 0u) {
        launches += launch_copy_raster(p, ctx->stream);
    } else {
        int n = -1;
        if ((opts->kernel == BVR_KERNEL_WAVEFRONT || opts->kernel == BVR_KERNEL_CTA_WAVEFRONT) && !p.reference_order && !p.strict_slab) {
            const bool cta = opts->kernel == BVR_KERNEL_CTA_WAVEFRONT;
            const size_t slots = cta ? cta_wavefront_slots(ctx->sm_count) : (size_t)p.cam.width * p.shard.rows;
            BVR_CK(ctx->wf_state.ensure(wavefront_state_bytes(slots)));
            WavefrontParams w;
            std::memset(&w, 0, sizeof w);
            w.r = p;
            wavefront_bind(w, ctx->wf_state.ptr, slots);
            w.refill_below = (uint32_t)ctx->tune.wf_refill;
            if (cta) {
                BVR_CK(cudaMemsetAsync(ctx->pixel_counter.ptr, 0, sizeof(unsigned int), ctx->stream));
                n = launch_cta_wavefront(w, ctx->n_inner, (uint32_t)ctx->n_models, ctx->tree_depth, ctx->sm_count,
                                         ctx->pixel_counter.as<unsigned int>(), ctx->stream);
            } else {
                n = launch_wavefront(w, ctx->n_inner, (uint32_t)ctx->n_models, ctx->tree_depth, ctx->sm_count,
                                     ctx->wf_host_counts, ctx->stream);
            }
            if (n < 0) {
                cudaError_t e = cudaGetLastError();
                if (e != cudaSuccess) return fail_cuda(ctx, e, "wavefront pipeline");
            }
        }
        if (n < 0 && !p.reference_order && !p.strict_slab && !ctx->tune.mk_v1) {
            BVR_CK(cudaMemsetAsync(ctx->pixel_counter.ptr, 0, sizeof(unsigned int), ctx->stream));
            // largest CTA whose stacks (and, when it fits, the scene) fit in shared memory
            const int forced = ctx->tune.mk_threads;
            const int candidates[4] = {1024, 768, 512, 256};
            // pixel-queue order: heaviest tiles first, judged by the previous frame of the same size (tile_order.cu)
            const uint32_t n_tiles = ((p.cam.width + 7u) / 8u) * ((p.shard.rows + 3u) / 4u);
            // (three small kernels per frame: worth it from about 2^24 pixel samples, i.e. frames of a few milliseconds)
            const int order_mode = ctx->tune.tile_order >= 0 ? ctx->tune.tile_order
                                   : ((uint64_t)n_tiles * 32u * p.cam.sample_count >= (1ull << 24) ? 2 : 0);
            const bool ordered = order_mode != 0 && n_tiles > 0;
            bool tile_first = false;
            if (ordered) {
                if (ctx->tile_geom[0] != p.cam.width || ctx->tile_geom[1] != p.shard.rows || ctx->tile_geom[2] != n_tiles ||
                    ctx->tile_geom[3] != (uint32_t)order_mode) {
                    BVR_CK(ctx->tile_order.ensure((size_t)n_tiles * sizeof(uint32_t)));
                    BVR_CK(ctx->tile_cost.ensure((size_t)n_tiles * sizeof(uint32_t)));
                    BVR_CK(ctx->tile_scratch.ensure(tile_order_scratch_bytes(n_tiles)));
                    BVR_CK(cudaMemsetAsync(ctx->tile_cost.ptr, 0, (size_t)n_tiles * sizeof(uint32_t), ctx->stream));
                    ctx->tile_geom[0] = p.cam.width; ctx->tile_geom[1] = p.shard.rows; ctx->tile_geom[2] = n_tiles;
                    ctx->tile_geom[3] = (uint32_t)order_mode;
                    ctx->tile_order_valid = false;
                    tile_first = true;
                }
            }
            const uint32_t* tile_order = ordered && ctx->tile_order_valid ? ctx->tile_order.as<uint32_t>() : nullptr;
            uint32_t* tile_cost = ordered ? ctx->tile_cost.as<uint32_t>() : nullptr;
            for (int ci = 0; ci < 4 && n < 0; ci++) {
                const int threads = forced ? forced : candidates[ci];
                n = launch_megakernel_v3(p, ctx->n_inner, (uint32_t)ctx->n_models, ctx->tree_depth,
                                         ctx->pixel_counter.as<unsigned int>(), threads,
                                         (uint32_t)ctx->tune.mk_wait, (uint32_t)ctx->tune.mk_leaf,   // 0 = per-mode default
                                         ctx->tune.no_both != 0,
                                         ctx->tune.no_top ? 0u : (uint32_t)ctx->tune.top_records,
                                         ctx->tune.no_top ? 0u : (uint32_t)ctx->tune.hot_records, ctx->tune.w4_lean != 0,
                                         tile_order, tile_cost,
                                         ctx->sm_count, ctx->stream);
                if (forced) break;
            }
            if (n < 0) cudaGetLastError();
            else p.weight_in_kernel = 1u;   // megakernel_v3 applies output_weight as it stores
            if (n < 0 && p.extra_modulus) return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "BVR_RENDER_EXTRA_SAMPLE: the megakernel does not fit this scene");
            if (n >= 0 && ordered) {
                launches += launch_tile_order_update(tile_cost, ctx->tile_order.as<uint32_t>(), ctx->tile_scratch.ptr, n_tiles,
                                                     order_mode, tile_first, ctx->stream);
                ctx->tile_order_valid = true;
            }
        }
        if (n < 0) n = launch_megakernel(p, ctx->stream);   // reference-order traversal, or scene too deep
        launches += n;
        launches += launch_selfcheck(p, ctx->ray_counter.as<unsigned long long>() + 1, ctx->stream);
    }
    if (p.out_weight != 1.0f && !p.weight_in_kernel) {
        // the other kernels (level Skip's copy, one thread per pixel, the wavefront pipelines) get the weight as a pass of its own
        const size_t px = (size_t)p.cam.width * p.shard.rows;
        if (p.out_rgba) launches += launch_scale(reinterpret_cast<float*>(p.out_rgba), p.out_weight, px * 4u, ctx->stream);
        if (p.out_rt_depth) launches += launch_scale(p.out_rt_depth, p.out_weight, px, ctx->stream);
    }
    BVR_CK(cudaGetLastError());
    BVR_CK(cudaEventRecord(ctx->ev_render1, ctx->stream));
    BVR_CK(cudaMemcpyAsync(ctx->ray_counter_host, ctx->ray_counter.ptr, 3 * sizeof(unsigned long long),
                           cudaMemcpyDeviceToHost, ctx->stream));
    ctx->render_timed = true;
    ctx->stats.kernel_launches += (uint64_t)launches;
    // paths = pixels of this shard that exist x samples
    uint64_t rows = 0;
    for (uint32_t ly = 0; ly < p.shard.rows; ly++) if (shard_global_row(p.shard, ly) < p.cam.height) rows++;
    ctx->stats.paths = p.cam.level == 0u ? 0u : rows * (uint64_t)p.cam.width * p.cam.sample_count;
    if (p.extra_modulus) {   // + one path per pixel of the tiles that take an extra sample
        // pixels of one row that sit in such a tile, by the class (ty + phase) % modulus of the row's tile row
        std::vector<uint64_t> per_class(p.extra_modulus, 0);
        for (uint32_t r = 0; r < p.extra_modulus; r++)
            for (uint32_t tx = 0; tx * 8u < p.cam.width; tx++)
                if ((tx + r) % p.extra_modulus < p.extra_count) per_class[r] += std::min(8u, p.cam.width - tx * 8u);
        for (uint32_t ly = 0; ly < p.shard.rows; ly++) {
            const uint32_t gy = shard_global_row(p.shard, ly);
            if (gy < p.cam.height) ctx->stats.paths += per_class[((gy >> 2) + p.extra_phase) % p.extra_modulus];
        }
    }
    return BVR_OK;
}

int bvr_render_device(BvrContext* ctx, const BvrCamera* camera, const BvrRaytraceLevel* level,
                      const BvrWindow* window, const BvrRenderOptions* opts,
                      const float* d_raster_rgba, const float* d_raster_depth, const BvrOutputs* device_out) {
    if (!ctx) return BVR_ERR_INVALID_ARGUMENT;
    ctx->error.clear();
    cudaSetDevice(ctx->device);
    return render_device_impl(ctx, camera, level, window, opts, d_raster_rgba, d_raster_depth, device_out);
}

static int render_host_impl(BvrContext* ctx, const BvrCamera* camera, const BvrRaytraceLevel* level, const BvrWindow* window,
                            const BvrRenderOptions* opts, const float* raster_rgba, const float* raster_depth,
                            const BvrOutputs* host_out, bool async) {
    if (!ctx) return BVR_ERR_INVALID_ARGUMENT;
    ctx->error.clear();
    if (!camera || !level || !window || !opts || !host_out)
        return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "null uniform / options / outputs pointer");
    cudaSetDevice(ctx->device);
    const size_t full_px = (size_t)opts->width * window->height;
    const size_t shard_px = (size_t)opts->width * shard_rows_impl(window->height, opts);
    const uint32_t eff_level = ((opts->flags & BVR_RENDER_DEFER_COMPOSITE) && level->level != 0u) ? 3u : level->level;
    const bool need_rgba = eff_level <= 2u, need_depth = eff_level == 1u || eff_level == 2u;
    if (need_rgba && !raster_rgba) return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "levels 0-2 need the raster colour");
    if (need_depth && !raster_depth) return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "levels 1-2 need the raster depth");

    // inputs: pinned (or staged) host -> HBM
    size_t in_staging = 0;
    if (need_rgba && !is_pinned_host(raster_rgba)) in_staging += (full_px * 16 + 255) & ~(size_t)255;
    if (need_depth && !is_pinned_host(raster_depth)) in_staging += (full_px * 4 + 255) & ~(size_t)255;
    // outputs: HBM -> pinned (or staged) host
    struct Plane { void* host; DeviceBuffer* dev; size_t bytes; bool pinned; size_t off; };
    Plane planes[5] = {
        {host_out->rgba, &ctx->out_rgba, shard_px * 16, false, 0},
        {host_out->rt_depth, &ctx->out_rt_depth, shard_px * 4, false, 0},
        {host_out->primary_id, &ctx->out_id, shard_px * 4, false, 0},
        {host_out->primary_depth, &ctx->out_pdepth, shard_px * 4, false, 0},
        {host_out->srgb8, &ctx->out_srgb8, shard_px * 4, false, 0},
    };
    size_t out_staging = 0;
    for (Plane& pl : planes) {
        if (!pl.host) continue;
        pl.pinned = is_pinned_host(pl.host);
        if (!pl.pinned) { pl.off = out_staging; out_staging += (pl.bytes + 255) & ~(size_t)255; }
        BVR_CK(pl.dev->ensure(pl.bytes));
    }
    if (async) {
        // nothing may be staged: the call must not wait for the device, and the copies must outlive it
        if (in_staging || out_staging)
            return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "bvr_render_async needs page-locked (pinned) host buffers");
    } else {
        BVR_CK(cudaStreamSynchronize(ctx->stream));   // staging reuse
        BVR_CK(ctx->io_staging.ensure(in_staging > out_staging ? in_staging : out_staging));
    }

    size_t off = 0;
    int st;
    if (need_rgba) {
        BVR_CK(ctx->in_rgba.ensure(full_px * 16));
        if ((st = h2d(ctx, ctx->in_rgba.ptr, raster_rgba, full_px * 16, ctx->io_staging, off)) != BVR_OK) return st;
    }
    if (need_depth) {
        BVR_CK(ctx->in_depth.ensure(full_px * 4));
        if ((st = h2d(ctx, ctx->in_depth.ptr, raster_depth, full_px * 4, ctx->io_staging, off)) != BVR_OK) return st;
    }

    BvrOutputs dev_out;
    dev_out.rgba = host_out->rgba ? ctx->out_rgba.as<float>() : nullptr;
    dev_out.rt_depth = host_out->rt_depth ? ctx->out_rt_depth.as<float>() : nullptr;
    dev_out.primary_id = host_out->primary_id ? ctx->out_id.as<uint32_t>() : nullptr;
    dev_out.primary_depth = host_out->primary_depth ? ctx->out_pdepth.as<float>() : nullptr;
    dev_out.srgb8 = host_out->srgb8 ? ctx->out_srgb8.as<uint8_t>() : nullptr;
    st = render_device_impl(ctx, camera, level, window, opts, need_rgba ? ctx->in_rgba.as<float>() : nullptr,
                            need_depth ? ctx->in_depth.as<float>() : nullptr, &dev_out);
    if (st != BVR_OK) return st;

    bool staged = false;
    if (in_staging && out_staging) BVR_CK(cudaStreamSynchronize(ctx->stream));   // inputs consumed before reuse
    for (Plane& pl : planes) {
        if (!pl.host) continue;
        void* dst = pl.pinned ? pl.host : static_cast<char*>(ctx->io_staging.ptr) + pl.off;
        BVR_CK(cudaMemcpyAsync(dst, pl.dev->ptr, pl.bytes, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->stats.d2h_bytes += pl.bytes;
        staged |= !pl.pinned;
    }
    ctx->pinned_src_in_flight = false;            // (an upload-only notion: see bvr_upload_scene)
    if (async) return BVR_OK;                     // bvr_sync completes the frame
    BVR_CK(cudaStreamSynchronize(ctx->stream));
    if (staged)
        for (Plane& pl : planes)
            if (pl.host && !pl.pinned) std::memcpy(pl.host, static_cast<char*>(ctx->io_staging.ptr) + pl.off, pl.bytes);
    return BVR_OK;
}

int bvr_render(BvrContext* ctx, const BvrCamera* camera, const BvrRaytraceLevel* level, const BvrWindow* window,
               const BvrRenderOptions* opts, const float* raster_rgba, const float* raster_depth,
               const BvrOutputs* host_out) {
    return render_host_impl(ctx, camera, level, window, opts, raster_rgba, raster_depth, host_out, false);
}

int bvr_render_async(BvrContext* ctx, const BvrCamera* camera, const BvrRaytraceLevel* level, const BvrWindow* window,
                     const BvrRenderOptions* opts, const float* raster_rgba, const float* raster_depth,
                     const BvrOutputs* host_out) {
    return render_host_impl(ctx, camera, level, window, opts, raster_rgba, raster_depth, host_out, true);
}

int bvr_axpby_device(BvrContext* ctx, float* d_dst, float dst_weight, const float* d_src, float src_weight, size_t n) {
    if (!ctx) return BVR_ERR_INVALID_ARGUMENT;
    ctx->error.clear();
    if (n && (!d_dst || (!d_src && src_weight != 0.0f))) return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "null device pointer");
    cudaSetDevice(ctx->device);
    ctx->stats.kernel_launches += (uint64_t)launch_axpby(d_dst, dst_weight, d_src, src_weight, n, ctx->stream);
    BVR_CK(cudaGetLastError());
    return BVR_OK;
}

int bvr_composite_device(BvrContext* ctx, const BvrCamera* camera, const BvrRaytraceLevel* level, float* d_rgba,
                         const float* d_rt_depth, const float* d_raster_rgba, const float* d_raster_depth, size_t n_pixels) {
    if (!ctx) return BVR_ERR_INVALID_ARGUMENT;
    ctx->error.clear();
    if (!camera || !level) return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "null uniform pointer");
    if (level->level != 1u && level->level != 2u) return BVR_OK;   // levels 0 and 3 have no depth test (raytrace.wgsl:97-122)
    if (n_pixels && (!d_rgba || !d_rt_depth || !d_raster_rgba || !d_raster_depth))
        return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "null device pointer");
    cudaSetDevice(ctx->device);
    CameraParams c;
    std::memset(&c, 0, sizeof c);
    c.near_plane = camera->near_plane;
    c.far_plane = camera->far_plane;
    ctx->stats.kernel_launches += (uint64_t)launch_composite(reinterpret_cast<float4*>(d_rgba), d_rt_depth,
                                                             reinterpret_cast<const float4*>(d_raster_rgba), d_raster_depth,
                                                             c, n_pixels, ctx->stream);
    BVR_CK(cudaGetLastError());
    return BVR_OK;
}

int bvr_peer_alloc(BvrContext* ctx, size_t bytes, void** d_ptr, uint8_t* handle_out) {
    if (!ctx) return BVR_ERR_INVALID_ARGUMENT;
    ctx->error.clear();
    if (!d_ptr || !handle_out || bytes == 0) return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "bad peer allocation arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == BVR_PEER_HANDLE_BYTES, "peer handle size");
    cudaSetDevice(ctx->device);
    void* ptr = nullptr;
    BVR_CK(cudaMalloc(&ptr, bytes));
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
    if (e != cudaSuccess) { cudaFree(ptr); return fail_cuda(ctx, e, "cudaIpcGetMemHandle"); }
    std::memcpy(handle_out, &h, sizeof h);
    *d_ptr = ptr;
    return BVR_OK;
}

int bvr_peer_open(BvrContext* ctx, const uint8_t* handle, void** d_ptr) {
    if (!ctx) return BVR_ERR_INVALID_ARGUMENT;
    ctx->error.clear();
    if (!d_ptr || !handle) return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "bad peer handle arguments");
    cudaSetDevice(ctx->device);
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof h);
    void* ptr = nullptr;
    BVR_CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    *d_ptr = ptr;
    return BVR_OK;
}

int bvr_peer_close(BvrContext* ctx, void* d_ptr) {
    if (!ctx) return BVR_ERR_INVALID_ARGUMENT;
    ctx->error.clear();
    cudaSetDevice(ctx->device);
    if (d_ptr) { BVR_CK(cudaStreamSynchronize(ctx->stream)); BVR_CK(cudaIpcCloseMemHandle(d_ptr)); }
    return BVR_OK;
}

int bvr_peer_free(BvrContext* ctx, void* d_ptr) {
    if (!ctx) return BVR_ERR_INVALID_ARGUMENT;
    ctx->error.clear();
    cudaSetDevice(ctx->device);
    if (d_ptr) { BVR_CK(cudaStreamSynchronize(ctx->stream)); BVR_CK(cudaFree(d_ptr)); }
    return BVR_OK;
}

int bvr_sum_slots_device(BvrContext* ctx, const float* d_slots, size_t slot_stride_floats, uint32_t n_slots, uint64_t mask,
                         float* d_dst, size_t n) {
    if (!ctx) return BVR_ERR_INVALID_ARGUMENT;
    ctx->error.clear();
    if (n && (!d_slots || !d_dst)) return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "null device pointer");
    if (n_slots == 0 || n_slots > 64u || (n & 3u) || (slot_stride_floats & 3u) || slot_stride_floats < n)
        return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "bad slot geometry (1..64 slots, float4-aligned sizes)");
    cudaSetDevice(ctx->device);
    ctx->stats.kernel_launches += (uint64_t)launch_sum_slots(d_slots, slot_stride_floats, n_slots, mask, d_dst, n, ctx->stream);
    BVR_CK(cudaGetLastError());
    return BVR_OK;
}

int bvr_unshard_device(BvrContext* ctx, const void* d_gathered, size_t shard_stride_words, void* d_full,
                       uint32_t width, uint32_t height, uint32_t channels, uint32_t shard_count, uint32_t strip_rows) {
    if (!ctx) return BVR_ERR_INVALID_ARGUMENT;
    ctx->error.clear();
    if (!d_gathered || !d_full || shard_count == 0 || channels == 0)
        return fail(ctx, BVR_ERR_INVALID_ARGUMENT, "bad unshard arguments");
    cudaSetDevice(ctx->device);
    if (strip_rows == 0) strip_rows = 8;
    if (shard_count == 1) strip_rows = height ? height : 1;
    ctx->stats.kernel_launches += (uint64_t)launch_unshard(static_cast<const uint32_t*>(d_gathered), shard_stride_words,
                                                           static_cast<uint32_t*>(d_full), width, height, channels,
                                                           shard_count, strip_rows, ctx->stream);
    BVR_CK(cudaGetLastError());
    return BVR_OK;
}

int bvr_get_stats(BvrContext* ctx, BvrStats* out) {
    if (!ctx || !out) return BVR_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    BVR_CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->render_timed) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ctx->ev_render0, ctx->ev_render1) == cudaSuccess) ctx->stats.last_render_ms = ms;
        else cudaGetLastError();
        ctx->stats.rays = ctx->ray_counter_host[0];
        ctx->stats.selfcheck_rays = ctx->ray_counter_host[1];
        ctx->stats.selfcheck_mismatches = ctx->ray_counter_host[2];
    }
    if (ctx->upload_timed) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ctx->ev_upload0, ctx->ev_upload1) == cudaSuccess) ctx->stats.last_upload_ms = ms;
        else cudaGetLastError();
    }
    *out = ctx->stats;
    return BVR_OK;
}

}  // extern "C"
