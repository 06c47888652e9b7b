// bvh_build.cu — GPU BVH builder emitting the reference's BVHNode contract.
//
// Replaces, on request, the per-frame host build at src/raytracing/extract.rs:316-332
// (`obvhs::ploc::build_ploc::<24>` + the node mapping) — the reference author's own TODO
// (extract.rs:264-267, 313).  Algorithm: LBVH — 63-bit Morton keys of the box centroids, radix sort,
// Karras' parallel radix tree ("Maximizing Parallelism in the Construction of BVHs, Octrees, and k-d
// Trees", HPG 2012), bottom-up box fitting with one atomic flag per inner node, then a re-layout into the
// contract of raytrace.wgsl:80-87 / 313-346:
//   node 0 = root; an inner node has model_count == 0 and its children at index, index+1;
//   a leaf has model_count == 1 and index = the model's position in the model buffer;
//   leaf bounds = centre -/+ (radius + 0.1)  (Model::aabb, extract.rs:220-227);  2n-1 nodes.
// Inner node k of the radix tree keeps its two children in slots 2k+1 and 2k+2, so the top of the tree is
// not contiguous as in the host PLOC builder, but every contract property holds (tests validate it with
// the host-side validator and render through the oracle with the downloaded nodes).
// The topology differs from PLOC's — as PLOC's differs from obvhs' — and the image does not depend on it.
//
// The key sort uses cub::DeviceRadixSort (library code, like a plain GEMM would use cuBLAS); every other
// step is a kernel in this file.

#include "kernels.cuh"

#include <cooperative_groups.h>
#include <cub/device/device_radix_sort.cuh>

namespace cg = cooperative_groups;

namespace bvr {

namespace {

__device__ __forceinline__ unsigned int float_to_ordered(float f) {
    const unsigned int b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(unsigned int o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// bounds[0..2] = min of centroids (ordered uint), bounds[3..5] = max
__global__ void bb_centroid_bounds(const RawModel* __restrict__ models, uint32_t n, unsigned int* __restrict__ bounds) {
    float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 a = reinterpret_cast<const float4*>(models)[2u * i];
        mn[0] = fminf(mn[0], a.x); mn[1] = fminf(mn[1], a.y); mn[2] = fminf(mn[2], a.z);
        mx[0] = fmaxf(mx[0], a.x); mx[1] = fmaxf(mx[1], a.y); mx[2] = fmaxf(mx[2], a.z);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
    }
    if ((threadIdx.x & 31u) == 0u) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            atomicMin(bounds + k, float_to_ordered(mn[k]));
            atomicMax(bounds + 3 + k, float_to_ordered(mx[k]));
        }
    }
}

__device__ __forceinline__ unsigned long long spread21(unsigned long long v) {
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x1f00000000ffffull;
    v = (v | (v << 16)) & 0x1f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}

__global__ void bb_morton(const RawModel* __restrict__ models, uint32_t n, const unsigned int* __restrict__ bounds,
                          unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = reinterpret_cast<const float4*>(models)[2u * i];
    const float c[3] = {a.x, a.y, a.z};
    unsigned long long q[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float lo = ordered_to_float(bounds[k]), hi = ordered_to_float(bounds[3 + k]);
        const float ext = hi - lo;
        const float s = ext > 0.0f ? 2097152.0f / ext : 0.0f;
        float v = floorf((c[k] - lo) * s);
        v = fminf(fmaxf(v, 0.0f), 2097151.0f);
        q[k] = (unsigned long long)v;
    }
    keys[i] = (spread21(q[0]) << 2) | (spread21(q[1]) << 1) | spread21(q[2]);
    vals[i] = i;
}

// common-prefix length of sorted keys i and j; duplicates are disambiguated by their position
__device__ __forceinline__ int delta(const unsigned long long* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const unsigned long long a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz((unsigned)i ^ (unsigned)j);
    return __clzll((long long)(a ^ b));
}

#define BB_LEAF 0x80000000u

// Karras 2012, Algorithm of section 4: one thread per inner node
__global__ void bb_radix_tree(const unsigned long long* __restrict__ keys, int n, uint2* __restrict__ children,
                              uint32_t* __restrict__ parent_inner, uint32_t* __restrict__ parent_leaf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) / 2;; t = (t + 1) / 2) {
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t <= 1) break;
    }
    const int gamma = i + s * d + (d < 0 ? -1 : 0);
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    uint2 ch;
    if (lo == gamma) { ch.x = BB_LEAF | (uint32_t)gamma; parent_leaf[gamma] = (uint32_t)i; }
    else { ch.x = (uint32_t)gamma; parent_inner[gamma] = (uint32_t)i; }
    if (hi == gamma + 1) { ch.y = BB_LEAF | (uint32_t)(gamma + 1); parent_leaf[gamma + 1] = (uint32_t)i; }
    else { ch.y = (uint32_t)(gamma + 1); parent_inner[gamma + 1] = (uint32_t)i; }
    children[i] = ch;
    if (i == 0) parent_inner[0] = 0xffffffffu;
}

// Model::aabb, extract.rs:220-227
__device__ __forceinline__ void leaf_box(const RawModel* __restrict__ models, uint32_t model, float4& lo, float4& hi) {
    const float4 a = reinterpret_cast<const float4*>(models)[2u * model];
    const float pad = __fadd_rn(a.w, 0.1f);
    lo = make_float4(__fsub_rn(a.x, pad), __fsub_rn(a.y, pad), __fsub_rn(a.z, pad), 0.0f);
    hi = make_float4(__fadd_rn(a.x, pad), __fadd_rn(a.y, pad), __fadd_rn(a.z, pad), 0.0f);
}

// bottom-up fit: the second thread to reach an inner node owns it
__global__ void bb_fit(const RawModel* __restrict__ models, const uint32_t* __restrict__ sorted_model, int n,
                       const uint2* __restrict__ children, const uint32_t* __restrict__ parent_inner,
                       const uint32_t* __restrict__ parent_leaf, unsigned int* __restrict__ flags,
                       float4* __restrict__ box_lo, float4* __restrict__ box_hi, uint32_t* __restrict__ height,
                       uint32_t* __restrict__ depth_out, uint32_t root) {
    const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n) return;
    uint32_t node = parent_leaf[leaf];
    while (node != 0xffffffffu) {
        if (atomicAdd(flags + node, 1u) == 0u) return;   // first arrival: the sibling subtree is not ready
        __threadfence();
        const uint2 ch = children[node];
        float4 lo0, hi0, lo1, hi1;
        uint32_t h0 = 0, h1 = 0;
        if (ch.x & BB_LEAF) leaf_box(models, sorted_model[ch.x & ~BB_LEAF], lo0, hi0);
        else { lo0 = __ldcg(box_lo + ch.x); hi0 = __ldcg(box_hi + ch.x); h0 = __ldcg(height + ch.x); }
        if (ch.y & BB_LEAF) leaf_box(models, sorted_model[ch.y & ~BB_LEAF], lo1, hi1);
        else { lo1 = __ldcg(box_lo + ch.y); hi1 = __ldcg(box_hi + ch.y); h1 = __ldcg(height + ch.y); }
        box_lo[node] = make_float4(fminf(lo0.x, lo1.x), fminf(lo0.y, lo1.y), fminf(lo0.z, lo1.z), 0.0f);
        box_hi[node] = make_float4(fmaxf(hi0.x, hi1.x), fmaxf(hi0.y, hi1.y), fmaxf(hi0.z, hi1.z), 0.0f);
        const uint32_t h = 1u + (h0 > h1 ? h0 : h1);
        height[node] = h;
        __threadfence();
        if (node == root) *depth_out = h + 1u;   // levels incl. the leaf level
        node = parent_inner[node];
    }
}

__device__ __forceinline__ void write_node(RawNode* __restrict__ out, uint32_t slot, float4 lo, float4 hi, uint32_t index,
                                           uint32_t count) {
    uint4* o = reinterpret_cast<uint4*>(out + slot);
    o[0] = make_uint4(__float_as_uint(lo.x), __float_as_uint(lo.y), __float_as_uint(lo.z), 0u);
    o[1] = make_uint4(__float_as_uint(hi.x), __float_as_uint(hi.y), __float_as_uint(hi.z), index);
    o[2] = make_uint4(count, 0u, 0u, 0u);
}

// reference layout: inner node k's children in slots 2k+1, 2k+2; root in slot 0
__global__ void bb_emit(const RawModel* __restrict__ models, const uint32_t* __restrict__ sorted_model, int n,
                        const uint2* __restrict__ children, const float4* __restrict__ box_lo,
                        const float4* __restrict__ box_hi, RawNode* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n - 1) return;
    if (k == 0) write_node(out, 0u, box_lo[0], box_hi[0], 1u, 0u);
    const uint2 ch = children[k];
    const uint32_t refs[2] = {ch.x, ch.y};
#pragma unroll
    for (int c = 0; c < 2; c++) {
        const uint32_t slot = 2u * (uint32_t)k + 1u + (uint32_t)c;
        if (refs[c] & BB_LEAF) {
            const uint32_t model = sorted_model[refs[c] & ~BB_LEAF];
            float4 lo, hi;
            leaf_box(models, model, lo, hi);
            write_node(out, slot, lo, hi, model, 1u);
        } else {
            write_node(out, slot, box_lo[refs[c]], box_hi[refs[c]], 2u * refs[c] + 1u, 0u);
        }
    }
}

// Position of every model in the reference's traversal order (raytrace.wgsl:329-341 pushes child `index` first and
// pops `index+1` first: right subtree before left).  Left children cover the lower part of the sorted range, so
// the reference reaches the leaves in DESCENDING sorted position.
__global__ void bb_rank(const uint32_t* __restrict__ sorted_model, uint32_t n, uint32_t* __restrict__ model_rank) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) model_rank[sorted_model[p]] = n - 1u - p;
}

__global__ void bb_single_leaf(const RawModel* __restrict__ models, RawNode* __restrict__ out, uint32_t* depth_out) {
    float4 lo, hi;
    leaf_box(models, 0u, lo, hi);
    write_node(out, 0u, lo, hi, 0u, 1u);
    *depth_out = 1u;
}

__global__ void bb_init_bounds(unsigned int* bounds, uint32_t* depth_out) {
    if (threadIdx.x < 3) bounds[threadIdx.x] = 0xffffffffu;
    else if (threadIdx.x < 6) bounds[threadIdx.x] = 0u;
    if (threadIdx.x == 0) *depth_out = 0u;
}

// ---- PLOC: parallel locally-ordered clustering (Meister & Bittner, TVCG 2018) over the Morton order -----------------
// The algorithm behind the reference's host build (`obvhs::ploc::build_ploc::<24>`, extract.rs:316-321), on the GPU: the
// clusters start as the leaves in Morton order; every round each cluster looks for the neighbour within PLOC_RADIUS
// positions whose union with it has the smallest surface area, mutual nearest neighbours merge into a new inner node
// that takes the place of the left one, the array is compacted in order, until one cluster is left.  One cooperative
// kernel runs all rounds (three grid barriers per round); ids of new nodes come from the compaction scan, so the tree is
// deterministic.  Inner node ids grow with the rounds: the root is the LAST one (n - 2); ploc_emit reverses them so that
// the root lands in slot 0 of the reference layout.
#define PLOC_RADIUS 24
#define PLOC_THREADS 512

struct PlocArrays {
    uint32_t n;                       // leaves
    const RawModel* models;
    const uint32_t* sorted_model;     // Morton order -> model
    uint32_t* cid[2];                 // cluster -> node ref (BB_LEAF | sorted position, or inner id), ping-pong
    float4* clo[2];                   // cluster boxes, ping-pong
    float4* chi[2];
    uint32_t* nn;                     // nearest neighbour of every cluster of the current round
    uint2* block_counts;              // per CTA: (clusters kept, merges) of its chunk
    uint2* children;                  // inner id -> (left ref, right ref)
    float4* box_lo;                   // inner id -> box
    float4* box_hi;
    uint32_t* height;                 // inner id -> levels below (a leaf child counts 0)
    uint32_t* below;                  // inner id -> leaves below
    uint32_t* parent_inner;           // inner id -> parent inner id (root: 0xffffffff)
    uint32_t* parent_leaf;            // sorted position -> parent inner id
    uint32_t* depth_out;
};

__device__ __forceinline__ float union_area(const float4 alo, const float4 ahi, const float4 blo, const float4 bhi) {
    const float dx = fmaxf(ahi.x, bhi.x) - fminf(alo.x, blo.x);
    const float dy = fmaxf(ahi.y, bhi.y) - fminf(alo.y, blo.y);
    const float dz = fmaxf(ahi.z, bhi.z) - fminf(alo.z, blo.z);
    return dx * dy + dy * dz + dz * dx;
}

// keep: this cluster survives the round (as itself, or as the merge of itself and its right partner)
__device__ __forceinline__ void ploc_decide(const uint32_t* __restrict__ nn, uint32_t i, bool& keep, bool& merge) {
    const uint32_t j = nn[i];
    const bool mutual = nn[j] == i;
    merge = mutual && i < j;
    keep = !mutual || i < j;
}

// SINGLE: one CTA does everything and the barriers are __syncthreads (small scenes: a grid barrier costs microseconds,
// a round of a 10 k-sphere scene much less)
template <bool SINGLE>
__global__ void __launch_bounds__(PLOC_THREADS) ploc_kernel(const PlocArrays p) {
    cg::grid_group grid = cg::this_grid();
    auto barrier = [&]() {
        if constexpr (SINGLE) { __threadfence_block(); __syncthreads(); }
        else grid.sync();
    };
    __shared__ uint32_t warp_keep[PLOC_THREADS / 32], warp_merge[PLOC_THREADS / 32];
    __shared__ uint32_t tile_base_keep, tile_base_merge;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    // leaves: cluster i = sorted position i
    for (uint32_t i = blockIdx.x * blockDim.x + tid; i < p.n; i += gridDim.x * blockDim.x) {
        float4 lo, hi;
        leaf_box(p.models, p.sorted_model[i], lo, hi);
        p.cid[0][i] = BB_LEAF | i;
        p.clo[0][i] = lo;
        p.chi[0][i] = hi;
    }
    barrier();
    uint32_t n = p.n, cur = 0, next_id = 0;
    while (n > 1u) {
        const uint32_t* cid = p.cid[cur];
        const float4* clo = p.clo[cur];
        const float4* chi = p.chi[cur];
        // --- 1: nearest neighbour within the window (ties go to the lower position: deterministic) ---
        for (uint32_t i = blockIdx.x * blockDim.x + tid; i < n; i += gridDim.x * blockDim.x) {
            const float4 lo = clo[i], hi = chi[i];
            const uint32_t j0 = i > PLOC_RADIUS ? i - PLOC_RADIUS : 0u;
            const uint32_t j1 = min(n - 1u, i + PLOC_RADIUS);
            float best = 3.4e38f;
            uint32_t bj = i == 0u ? 1u : i - 1u;
            for (uint32_t j = j0; j <= j1; j++) {
                if (j == i) continue;
                const float a = union_area(lo, hi, clo[j], chi[j]);
                if (a < best) { best = a; bj = j; }
            }
            p.nn[i] = bj;
        }
        barrier();
        // --- 2: every CTA owns a contiguous chunk: clusters kept / merged in it ---
        const uint32_t chunk = (n + gridDim.x - 1u) / gridDim.x;
        const uint32_t c0 = min(n, blockIdx.x * chunk), c1 = min(n, c0 + chunk);
        {
            uint32_t keeps = 0, merges = 0;
            for (uint32_t i = c0 + tid; i < c1; i += blockDim.x) {
                bool k, m;
                ploc_decide(p.nn, i, k, m);
                keeps += k ? 1u : 0u;
                merges += m ? 1u : 0u;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                keeps += __shfl_xor_sync(0xffffffffu, keeps, o);
                merges += __shfl_xor_sync(0xffffffffu, merges, o);
            }
            if (lane == 0u) { warp_keep[warp] = keeps; warp_merge[warp] = merges; }
            __syncthreads();
            if (tid == 0u) {
                uint32_t k = 0, m = 0;
                for (uint32_t w = 0; w < PLOC_THREADS / 32; w++) { k += warp_keep[w]; m += warp_merge[w]; }
                p.block_counts[blockIdx.x] = make_uint2(k, m);
            }
        }
        barrier();
        // --- 3: compact in order; merges create their inner node ---
        uint32_t total_keep = 0, total_merge = 0;
        {
            uint32_t bk = 0, bm = 0;      // everything before this CTA's chunk
            for (uint32_t b = 0; b < gridDim.x; b++) {
                const uint2 c = p.block_counts[b];
                if (b < blockIdx.x) { bk += c.x; bm += c.y; }
                total_keep += c.x;
                total_merge += c.y;
            }
            if (tid == 0u) { tile_base_keep = bk; tile_base_merge = bm; }
            __syncthreads();
        }
        uint32_t* ocid = p.cid[cur ^ 1u];
        float4* oclo = p.clo[cur ^ 1u];
        float4* ochi = p.chi[cur ^ 1u];
        for (uint32_t t0 = c0; t0 < c1; t0 += blockDim.x) {
            const uint32_t i = t0 + tid;
            bool k = false, m = false;
            if (i < c1) ploc_decide(p.nn, i, k, m);
            // block-wide exclusive scan of (keep, merge) over this tile
            const unsigned kb = __ballot_sync(0xffffffffu, k), mb = __ballot_sync(0xffffffffu, m);
            if (lane == 0u) { warp_keep[warp] = (uint32_t)__popc(kb); warp_merge[warp] = (uint32_t)__popc(mb); }
            __syncthreads();
            uint32_t wk = 0, wm = 0, tk = 0, tm = 0;
            for (uint32_t w = 0; w < PLOC_THREADS / 32; w++) {
                if (w < warp) { wk += warp_keep[w]; wm += warp_merge[w]; }
                tk += warp_keep[w];
                tm += warp_merge[w];
            }
            const uint32_t below_mask = (1u << lane) - 1u;
            const uint32_t pos = tile_base_keep + wk + (uint32_t)__popc(kb & below_mask);
            const uint32_t id = next_id + tile_base_merge + wm + (uint32_t)__popc(mb & below_mask);
            if (k) {
                if (m) {
                    const uint32_t j = p.nn[i];
                    const uint32_t ra = cid[i], rb = cid[j];
                    const float4 alo = clo[i], ahi = chi[i], blo = clo[j], bhi = chi[j];
                    const float4 lo = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), 0.0f);
                    const float4 hi = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0.0f);
                    uint32_t ha = 0, hb = 0, na = 1, nb = 1;
                    if (ra & BB_LEAF) p.parent_leaf[ra & ~BB_LEAF] = id;
                    else { p.parent_inner[ra] = id; ha = p.height[ra]; na = p.below[ra]; }
                    if (rb & BB_LEAF) p.parent_leaf[rb & ~BB_LEAF] = id;
                    else { p.parent_inner[rb] = id; hb = p.height[rb]; nb = p.below[rb]; }
                    p.children[id] = make_uint2(ra, rb);
                    p.box_lo[id] = lo;
                    p.box_hi[id] = hi;
                    p.height[id] = 1u + max(ha, hb);
                    p.below[id] = na + nb;
                    ocid[pos] = id;
                    oclo[pos] = lo;
                    ochi[pos] = hi;
                } else {
                    ocid[pos] = cid[i];
                    oclo[pos] = clo[i];
                    ochi[pos] = chi[i];
                }
            }
            __syncthreads();
            if (tid == 0u) { tile_base_keep += tk; tile_base_merge += tm; }
            __syncthreads();
        }
        next_id += total_merge;
        n = total_keep;
        cur ^= 1u;
        barrier();
    }
    if (blockIdx.x == 0 && tid == 0u) {
        const uint32_t root = p.n - 2u;
        p.parent_inner[root] = 0xffffffffu;
        *p.depth_out = p.height[root] + 1u;      // levels incl. the leaf level
    }
}

// reference layout from the PLOC tree: inner id k sits at reversed rank r = (n - 2) - k (root -> 0); the children of rank
// r occupy slots 2r + 1, 2r + 2, the root slot 0
__global__ void ploc_emit(const RawModel* __restrict__ models, const uint32_t* __restrict__ sorted_model, uint32_t n,
                          const uint2* __restrict__ children, const float4* __restrict__ box_lo,
                          const float4* __restrict__ box_hi, RawNode* __restrict__ out) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n - 1u) return;
    const uint32_t r = (n - 2u) - k;
    if (r == 0u) write_node(out, 0u, box_lo[k], box_hi[k], 1u, 0u);
    const uint2 ch = children[k];
    const uint32_t refs[2] = {ch.x, ch.y};
#pragma unroll
    for (int c = 0; c < 2; c++) {
        const uint32_t slot = 2u * r + 1u + (uint32_t)c;
        if (refs[c] & BB_LEAF) {
            const uint32_t model = sorted_model[refs[c] & ~BB_LEAF];
            float4 lo, hi;
            leaf_box(models, model, lo, hi);
            write_node(out, slot, lo, hi, model, 1u);
        } else {
            write_node(out, slot, box_lo[refs[c]], box_hi[refs[c]], 2u * ((n - 2u) - refs[c]) + 1u, 0u);
        }
    }
}

// Position of every model in the reference's traversal order for ANY binary tree: the reference pops `index + 1` (the
// right child) before `index`, so whatever hangs off a right sibling of one of my ancestors is reached before me.
__global__ void ploc_rank(const uint32_t* __restrict__ sorted_model, uint32_t n, const uint2* __restrict__ children,
                          const uint32_t* __restrict__ parent_inner, const uint32_t* __restrict__ parent_leaf,
                          const uint32_t* __restrict__ below, uint32_t* __restrict__ model_rank) {
    const uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= n) return;
    uint32_t before = 0u, child = BB_LEAF | pos, node = parent_leaf[pos];
    while (node != 0xffffffffu) {
        const uint2 ch = children[node];
        if (child == ch.x) before += (ch.y & BB_LEAF) ? 1u : below[ch.y];
        child = node;
        node = parent_inner[node];
    }
    model_rank[sorted_model[pos]] = before;
}

struct Layout {
    size_t keys, keys_sorted, vals, vals_sorted, children, parent_inner, parent_leaf, flags, box_lo, box_hi, height,
        bounds, depth, cub_temp, cid0, cid1, clo0, clo1, chi0, chi1, nn, below, block_counts, total;
};

size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

Layout make_layout(uint32_t n, size_t cub_bytes) {
    Layout L;
    size_t off = 0;
    auto take = [&off](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
    L.keys = take((size_t)n * 8);
    L.keys_sorted = take((size_t)n * 8);
    L.vals = take((size_t)n * 4);
    L.vals_sorted = take((size_t)n * 4);
    L.children = take((size_t)n * 8);
    L.parent_inner = take((size_t)n * 4);
    L.parent_leaf = take((size_t)n * 4);
    L.flags = take((size_t)n * 4);
    L.box_lo = take((size_t)n * 16);
    L.box_hi = take((size_t)n * 16);
    L.height = take((size_t)n * 4);
    L.bounds = take(32);
    L.depth = take(4);
    L.cub_temp = take(cub_bytes);
    L.cid0 = take((size_t)n * 4);
    L.cid1 = take((size_t)n * 4);
    L.clo0 = take((size_t)n * 16);
    L.clo1 = take((size_t)n * 16);
    L.chi0 = take((size_t)n * 16);
    L.chi1 = take((size_t)n * 16);
    L.nn = take((size_t)n * 4);
    L.below = take((size_t)n * 4);
    L.block_counts = take((size_t)4096 * 8);
    L.total = off;
    return L;
}

size_t cub_temp_bytes(uint32_t n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, 0, 63);
    return bytes;
}

}  // namespace

size_t bvh_build_scratch_bytes(uint32_t n_models) {
    if (n_models == 0) return 256;
    return make_layout(n_models, cub_temp_bytes(n_models)).total;
}

// Builds the node array for `n` models (device pointers) into `out_nodes` (2n-1 records).  `depth_out` is a
// device word that receives the number of tree levels; `model_rank` (n words) receives every model's position in
// the reference's traversal order (the tie-break of trace.cuh: test_leaf).  Returns the number of kernels launched, -1 on error.
int launch_bvh_build(const RawModel* models, uint32_t n, RawNode* out_nodes, uint32_t* model_rank, void* scratch,
                     uint32_t** depth_out, int algorithm, int sm_count, cudaStream_t stream) {
    if (n == 0) return 0;
    size_t cub_bytes = cub_temp_bytes(n);
    const Layout L = make_layout(n, cub_bytes);
    char* base = static_cast<char*>(scratch);
    auto* keys = reinterpret_cast<unsigned long long*>(base + L.keys);
    auto* keys_sorted = reinterpret_cast<unsigned long long*>(base + L.keys_sorted);
    auto* vals = reinterpret_cast<uint32_t*>(base + L.vals);
    auto* vals_sorted = reinterpret_cast<uint32_t*>(base + L.vals_sorted);
    auto* children = reinterpret_cast<uint2*>(base + L.children);
    auto* parent_inner = reinterpret_cast<uint32_t*>(base + L.parent_inner);
    auto* parent_leaf = reinterpret_cast<uint32_t*>(base + L.parent_leaf);
    auto* flags = reinterpret_cast<unsigned int*>(base + L.flags);
    auto* box_lo = reinterpret_cast<float4*>(base + L.box_lo);
    auto* box_hi = reinterpret_cast<float4*>(base + L.box_hi);
    auto* height = reinterpret_cast<uint32_t*>(base + L.height);
    auto* bounds = reinterpret_cast<unsigned int*>(base + L.bounds);
    auto* depth = reinterpret_cast<uint32_t*>(base + L.depth);
    *depth_out = depth;
    int launches = 0;
    if (n == 1) {
        bb_single_leaf<<<1, 1, 0, stream>>>(models, out_nodes, depth);
        cudaMemsetAsync(model_rank, 0, sizeof(uint32_t), stream);
        return 1;
    }
    const int T = 256;
    const int blocks = (int)((n + T - 1) / T);
    bb_init_bounds<<<1, 32, 0, stream>>>(bounds, depth);
    bb_centroid_bounds<<<blocks < 1184 ? blocks : 1184, T, 0, stream>>>(models, n, bounds);
    bb_morton<<<blocks, T, 0, stream>>>(models, n, bounds, keys, vals);
    launches += 3;
    if (cub::DeviceRadixSort::SortPairs(base + L.cub_temp, cub_bytes, keys, keys_sorted, vals, vals_sorted, (int)n, 0, 63,
                                        stream) != cudaSuccess)
        return -1;
    launches += 8;   // cub's onesweep passes (approximate; counted as library launches)
    if (algorithm == BVH_BUILD_PLOC) {
        PlocArrays pa;
        pa.n = n;
        pa.models = models;
        pa.sorted_model = vals_sorted;
        pa.cid[0] = reinterpret_cast<uint32_t*>(base + L.cid0);
        pa.cid[1] = reinterpret_cast<uint32_t*>(base + L.cid1);
        pa.clo[0] = reinterpret_cast<float4*>(base + L.clo0);
        pa.clo[1] = reinterpret_cast<float4*>(base + L.clo1);
        pa.chi[0] = reinterpret_cast<float4*>(base + L.chi0);
        pa.chi[1] = reinterpret_cast<float4*>(base + L.chi1);
        pa.nn = reinterpret_cast<uint32_t*>(base + L.nn);
        pa.block_counts = reinterpret_cast<uint2*>(base + L.block_counts);
        pa.children = children;
        pa.box_lo = box_lo;
        pa.box_hi = box_hi;
        pa.height = height;
        pa.below = reinterpret_cast<uint32_t*>(base + L.below);
        pa.parent_inner = parent_inner;
        pa.parent_leaf = parent_leaf;
        pa.depth_out = depth;
        // all CTAs must be resident at once (grid barriers): at most what the device holds, at most one CTA per 2048 leaves
        if (n <= 32768u) {
            ploc_kernel<true><<<1, PLOC_THREADS, 0, stream>>>(pa);
        } else {
            int per_sm = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ploc_kernel<false>, PLOC_THREADS, 0) != cudaSuccess || per_sm < 1) return -1;
            int grid = sm_count * (per_sm > 2 ? 2 : per_sm);
            const int useful = (int)((n + 2047u) / 2048u);
            if (grid > useful) grid = useful;
            if (grid > 4096) grid = 4096;
            void* args[] = {&pa};
            if (cudaLaunchCooperativeKernel((const void*)ploc_kernel<false>, dim3(grid), dim3(PLOC_THREADS), args, 0, stream) != cudaSuccess) return -1;
        }
        ploc_emit<<<blocks, T, 0, stream>>>(models, vals_sorted, n, children, box_lo, box_hi, out_nodes);
        ploc_rank<<<blocks, T, 0, stream>>>(vals_sorted, n, children, parent_inner, parent_leaf, pa.below, model_rank);
        launches += 3;
        return launches;
    }
    if (cudaMemsetAsync(flags, 0, (size_t)n * 4, stream) != cudaSuccess) return -1;
    bb_radix_tree<<<blocks, T, 0, stream>>>(keys_sorted, (int)n, children, parent_inner, parent_leaf);
    bb_fit<<<blocks, T, 0, stream>>>(models, vals_sorted, (int)n, children, parent_inner, parent_leaf, flags, box_lo, box_hi,
                                     height, depth, 0u);
    bb_emit<<<blocks, T, 0, stream>>>(models, vals_sorted, (int)n, children, box_lo, box_hi, out_nodes);
    bb_rank<<<blocks, T, 0, stream>>>(vals_sorted, n, model_rank);
    launches += 4;
    return launches;
}

uint32_t* bvh_build_depth_word(void* scratch, uint32_t n) {
    return reinterpret_cast<uint32_t*>(static_cast<char*>(scratch) + make_layout(n, cub_temp_bytes(n)).depth);
}

// Refit: same topology (the arrays of the last launch_bvh_build in `scratch`), new sphere positions / radii: boxes are
// recomputed bottom-up and the node array is emitted again.  Ranks and depth do not change.
int launch_bvh_refit(const RawModel* models, uint32_t n, RawNode* out_nodes, void* scratch, int algorithm, cudaStream_t stream) {
    if (n == 0) return 0;
    const Layout L = make_layout(n, cub_temp_bytes(n));
    char* base = static_cast<char*>(scratch);
    auto* vals_sorted = reinterpret_cast<uint32_t*>(base + L.vals_sorted);
    auto* children = reinterpret_cast<uint2*>(base + L.children);
    auto* parent_inner = reinterpret_cast<uint32_t*>(base + L.parent_inner);
    auto* parent_leaf = reinterpret_cast<uint32_t*>(base + L.parent_leaf);
    auto* flags = reinterpret_cast<unsigned int*>(base + L.flags);
    auto* box_lo = reinterpret_cast<float4*>(base + L.box_lo);
    auto* box_hi = reinterpret_cast<float4*>(base + L.box_hi);
    auto* height = reinterpret_cast<uint32_t*>(base + L.height);
    auto* depth = reinterpret_cast<uint32_t*>(base + L.depth);
    if (n == 1) {
        bb_single_leaf<<<1, 1, 0, stream>>>(models, out_nodes, depth);
        return 1;
    }
    const int T = 256;
    const int blocks = (int)((n + T - 1) / T);
    if (cudaMemsetAsync(flags, 0, (size_t)n * 4, stream) != cudaSuccess) return -1;
    const uint32_t root = algorithm == BVH_BUILD_PLOC ? n - 2u : 0u;
    bb_fit<<<blocks, T, 0, stream>>>(models, vals_sorted, (int)n, children, parent_inner, parent_leaf, flags, box_lo, box_hi,
                                     height, depth, root);
    if (algorithm == BVH_BUILD_PLOC) ploc_emit<<<blocks, T, 0, stream>>>(models, vals_sorted, n, children, box_lo, box_hi, out_nodes);
    else bb_emit<<<blocks, T, 0, stream>>>(models, vals_sorted, (int)n, children, box_lo, box_hi, out_nodes);
    return 2;
}

}  // namespace bvr
