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

#include <cub/device/device_radix_sort.cuh>

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
                       uint32_t* __restrict__ depth_out) {
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
        if (node == 0u) *depth_out = h + 1u;   // levels incl. the leaf level
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

struct Layout {
    size_t keys, keys_sorted, vals, vals_sorted, children, parent_inner, parent_leaf, flags, box_lo, box_hi, height,
        bounds, depth, cub_temp, total;
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
                     uint32_t** depth_out, cudaStream_t stream) {
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
    if (cudaMemsetAsync(flags, 0, (size_t)n * 4, stream) != cudaSuccess) return -1;
    bb_radix_tree<<<blocks, T, 0, stream>>>(keys_sorted, (int)n, children, parent_inner, parent_leaf);
    bb_fit<<<blocks, T, 0, stream>>>(models, vals_sorted, (int)n, children, parent_inner, parent_leaf, flags, box_lo, box_hi,
                                     height, depth);
    bb_emit<<<blocks, T, 0, stream>>>(models, vals_sorted, (int)n, children, box_lo, box_hi, out_nodes);
    bb_rank<<<blocks, T, 0, stream>>>(vals_sorted, n, model_rank);
    launches += 4;
    return launches;
}

}  // namespace bvr
