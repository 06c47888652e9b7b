// scene_validate.cu — structural validation of a BIG node array on the GPU.
//
// bvr_upload_scene checks every uploaded tree against the reference's contract (raytrace.wgsl:80-87, 313-346): indices
// in range, nothing reachable twice, and it needs three facts about it — the number of levels (stack bound), the
// largest leaf, and every model's position in the reference's traversal order (the tie rule of trace.cuh).  On the
// host that is a depth-first walk: 80 ms of cache misses for the 2 M nodes of the 2^20-sphere scene, against 14 ms
// for the upload itself.  For arrays of 32 k nodes and more the same facts are computed here, on the bytes that are
// already in HBM:
//   gv_link   one thread per node: range checks; parent[child] = node with an exchange that catches a second reference
//   gv_fit    one thread per leaf, bottom-up: the second arrival at an inner node owns it (models below, height)
//   gv_rank   one thread per leaf, walking up: models in right siblings are reached earlier by the reference
// The verdict is read back once (one synchronisation per upload).

#include "kernels.cuh"

namespace bvr {

namespace {

constexpr uint32_t NO_PARENT = 0xffffffffu;

__global__ void gv_link(const RawNode* __restrict__ nodes, uint32_t n, uint32_t n_models, uint32_t* __restrict__ parent,
                        ValidateOut* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t bad = 0u, inner = 0u, leaf_models = 0u;
    if (i < n) {
        const RawNode nd = nodes[i];
        if (nd.model_count > 0u) {
            leaf_models = nd.model_count;
            if (nd.model_count > BVR_MAX_LEAF_COUNT) bad |= BVR_VALIDATE_LEAF_TOO_BIG;
            if ((unsigned long long)nd.index + nd.model_count > n_models) bad |= BVR_VALIDATE_LEAF_RANGE;
        } else {
            inner = 1u;
            if ((unsigned long long)nd.index + 1ull >= n) {
                bad |= BVR_VALIDATE_CHILD_RANGE;
            } else {
                for (uint32_t c = nd.index; c < nd.index + 2u; c++) {
                    // node 0 is the root: a reference to it closes a cycle
                    if (c == 0u || atomicExch(&parent[c], i) != NO_PARENT) bad |= BVR_VALIDATE_TWICE;
                }
            }
        }
    }
    // one atomic per warp and counter
    const unsigned full = 0xffffffffu;
    const uint32_t inner_warp = (uint32_t)__popc(__ballot_sync(full, inner != 0u));
    uint32_t bad_warp = bad, leaf_warp = leaf_models;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        bad_warp |= __shfl_xor_sync(full, bad_warp, o);
        leaf_warp = max(leaf_warp, __shfl_xor_sync(full, leaf_warp, o));
    }
    if ((threadIdx.x & 31u) == 0u) {
        if (inner_warp) atomicAdd(&out->n_inner, inner_warp);
        if (bad_warp) atomicOr(&out->bad, bad_warp);
        if (leaf_warp) atomicMax(&out->max_leaf, leaf_warp);
    }
}

__global__ void gv_fit(const RawNode* __restrict__ nodes, uint32_t n, const uint32_t* __restrict__ parent,
                       unsigned int* __restrict__ arrivals, uint32_t* __restrict__ below, uint32_t* __restrict__ height,
                       ValidateOut* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || out->bad != 0u) return;
    const RawNode leaf = nodes[i];
    if (leaf.model_count == 0u) return;
    below[i] = leaf.model_count;
    height[i] = 1u;
    if (i == 0u) out->depth = 1u;              // the root is a leaf
    __threadfence();
    uint32_t node = parent[i];
    while (node != NO_PARENT) {
        if (atomicAdd(&arrivals[node], 1u) != 1u) return;   // first arrival: the sibling is not ready (third: a stray cycle)
        __threadfence();
        const uint32_t c0 = nodes[node].index;
        const uint32_t h0 = __ldcg(&height[c0]), h1 = __ldcg(&height[c0 + 1u]);
        below[node] = __ldcg(&below[c0]) + __ldcg(&below[c0 + 1u]);
        const uint32_t h = 1u + (h0 > h1 ? h0 : h1);
        height[node] = h;
        __threadfence();
        if (node == 0u) out->depth = h;        // levels including the leaf level
        node = parent[node];
    }
}

__global__ void gv_rank(const RawNode* __restrict__ nodes, uint32_t n, const uint32_t* __restrict__ parent,
                        const uint32_t* __restrict__ below, uint32_t* __restrict__ model_rank,
                        const ValidateOut* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    // trees deeper than BVR_VALIDATE_MAX_DEPTH are walked on the host (this loop is O(depth) per leaf)
    const uint32_t depth = out->depth;
    if (i >= n || out->bad != 0u || depth > BVR_VALIDATE_MAX_DEPTH) return;
    const RawNode leaf = nodes[i];
    if (leaf.model_count == 0u) return;
    // the reference pops `index + 1` before `index` (raytrace.wgsl:329-341): whatever hangs off a right sibling
    // of one of my ancestors is reached before me
    uint32_t before = 0u, child = i, node = parent[i], steps = 0u;
    while (node != NO_PARENT) {
        if (++steps > depth) return;           // a stray cycle that the root does not reach
        const uint32_t c0 = nodes[node].index;
        if (child == c0) before += __ldcg(&below[c0 + 1u]);
        child = node;
        node = parent[node];
    }
    if (child != 0u) return;                   // not below the root: never visited
    for (uint32_t m = 0; m < leaf.model_count; m++) atomicMin(&model_rank[leaf.index + m], before + m);
}

}  // namespace

size_t validate_scratch_bytes(uint32_t n_nodes) { return (size_t)n_nodes * 4u * sizeof(uint32_t); }

// Enqueues the three kernels; `out` (device) must be read back by the caller.  model_rank receives the ranks
// (0xffffffff for models no reachable leaf holds).  Returns the number of kernels launched.
int launch_validate_scene(const RawNode* nodes, uint32_t n_nodes, uint32_t n_models, void* scratch, uint32_t* model_rank,
                          ValidateOut* out, cudaStream_t stream) {
    if (n_nodes == 0) return 0;
    uint32_t* parent = static_cast<uint32_t*>(scratch);
    unsigned int* arrivals = parent + n_nodes;
    uint32_t* below = arrivals + n_nodes;
    uint32_t* height = below + n_nodes;
    cudaMemsetAsync(parent, 0xff, (size_t)n_nodes * sizeof(uint32_t), stream);
    cudaMemsetAsync(arrivals, 0, (size_t)n_nodes * sizeof(uint32_t), stream);
    cudaMemsetAsync(out, 0, sizeof(ValidateOut), stream);
    if (n_models) cudaMemsetAsync(model_rank, 0xff, (size_t)n_models * sizeof(uint32_t), stream);
    const uint32_t T = 256, blocks = (n_nodes + T - 1) / T;
    gv_link<<<blocks, T, 0, stream>>>(nodes, n_nodes, n_models, parent, out);
    gv_fit<<<blocks, T, 0, stream>>>(nodes, n_nodes, parent, arrivals, below, height, out);
    gv_rank<<<blocks, T, 0, stream>>>(nodes, n_nodes, parent, below, model_rank, out);
    return 3;
}

}  // namespace bvr
