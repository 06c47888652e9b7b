// scene_kernels.cu — device-side re-layout of the reference's storage buffers into the traversal
// layout (trace.cuh).  The library accepts the reference bytes unchanged (Model 32 B, BVHNode 48 B,
// src/raytracing/extract.rs:213-237); everything below happens in HBM after the upload.

#include "kernels.cuh"

namespace bvr {

namespace {

constexpr int SCAN_BLOCK = 1024;

__global__ void derive_spheres_kernel(const RawModel* __restrict__ models, uint32_t n,
                                      float4* __restrict__ spheres, uint32_t* __restrict__ sphere_material) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // one 32-byte model = two 16-byte loads
    const float4 a = reinterpret_cast<const float4*>(models)[2u * i];
    const uint4 b = reinterpret_cast<const uint4*>(models)[2u * i + 1u];
    spheres[i] = a;                 // (position.xyz, radius)
    sphere_material[i] = b.x;       // material_id
}

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_sums, uint32_t& block_total) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= (uint32_t)o) x += y;
    }
    if (lane == 31u) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = (lane < (blockDim.x >> 5)) ? warp_sums[lane] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= (uint32_t)o) w += y;
        }
        warp_sums[lane] = w;   // inclusive
    }
    __syncthreads();
    const uint32_t warp_off = warp ? warp_sums[warp - 1] : 0u;
    block_total = warp_sums[(blockDim.x >> 5) - 1];
    return warp_off + x - v;
}

// pass 1: number of inner nodes per block of 1024
__global__ void count_inner_kernel(const RawNode* __restrict__ nodes, uint32_t n, uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t warp_sums[32];
    const uint32_t i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const uint32_t flag = (i < n && nodes[i].model_count == 0u) ? 1u : 0u;
    uint32_t total;
    (void)block_exclusive_scan(flag, warp_sums, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// pass 2: exclusive scan of the block sums (one block, serial over chunks)
__global__ void scan_block_sums_kernel(uint32_t* __restrict__ block_sums, uint32_t n_blocks) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0u;
    __syncthreads();
    for (uint32_t base = 0; base < n_blocks; base += SCAN_BLOCK) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n_blocks ? block_sums[i] : 0u;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, warp_sums, total);
        const uint32_t c = carry;
        if (i < n_blocks) block_sums[i] = c + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sums[n_blocks] = carry;   // total number of inner nodes
}

// pass 3: dense id of every inner node
__global__ void assign_inner_id_kernel(const RawNode* __restrict__ nodes, uint32_t n,
                                       const uint32_t* __restrict__ block_sums, uint32_t* __restrict__ inner_id) {
    __shared__ uint32_t warp_sums[32];
    const uint32_t i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const uint32_t flag = (i < n && nodes[i].model_count == 0u) ? 1u : 0u;
    uint32_t total;
    const uint32_t ex = block_exclusive_scan(flag, warp_sums, total);
    if (i < n) inner_id[i] = flag ? block_sums[blockIdx.x] + ex : 0xffffffffu;
}

// ---- record numbering for scenes walked in HBM/L2: the hot top of the tree first --------------------------------------
// The 4-wide walk of a big scene (megakernel_v3.cu MODE 3) stages the first records of its array in shared memory, so the
// array should START with the records every ray visits: the root's, then level by level (breadth first over the 4-wide
// levels: the record of node N refers to the inner children of N's inner children).  top_bfs_kernel gives the first
// `max_top` records reached that way the ids 0, 1, 2, ...; the other inner nodes follow in array order (rest_* below).
// Any prefix of a breadth-first order is closed under "parent of", and which prefix the render kernel stages cannot
// change an image: a staged record is a copy.
#define BVR_TOP_BFS_MAX 3072u
__global__ void __launch_bounds__(SCAN_BLOCK) top_bfs_kernel(const RawNode* __restrict__ nodes, uint32_t n, uint32_t max_top,
                                                             uint32_t* __restrict__ id_q) {
    __shared__ uint32_t frontier[2][BVR_TOP_BFS_MAX];
    __shared__ uint32_t warp_sums[32];
    const uint32_t tid = threadIdx.x;
    uint32_t done = 0, cur_n = 0;
    int cur = 0;
    if (max_top > BVR_TOP_BFS_MAX) max_top = BVR_TOP_BFS_MAX;
    if (n > 0u && nodes[0].model_count == 0u) {
        if (tid == 0u) frontier[0][0] = 0u;
        cur_n = 1u;
    }
    __syncthreads();
    while (cur_n > 0u && done < max_top) {
        const uint32_t take = min(cur_n, max_top - done);
        for (uint32_t e = tid; e < take; e += SCAN_BLOCK) id_q[frontier[cur][e]] = done + e;
        done += take;
        if (done >= max_top) break;
        uint32_t next_n = 0;
        for (uint32_t base = 0; base < cur_n; base += SCAN_BLOCK) {
            const uint32_t e = base + tid;
            uint32_t kids[4], nk = 0;
            if (e < cur_n) {
                const uint32_t first = nodes[frontier[cur][e]].index;
                for (uint32_t s = 0; s < 2u; s++) {
                    const uint32_t x = first + s;
                    if (x >= n || nodes[x].model_count != 0u) continue;
                    const uint32_t xf = nodes[x].index;
                    for (uint32_t t = 0; t < 2u; t++)
                        if (xf + t < n && nodes[xf + t].model_count == 0u) kids[nk++] = xf + t;
                }
            }
            uint32_t total;
            const uint32_t ex = block_exclusive_scan(nk, warp_sums, total);
            for (uint32_t k = 0; k < nk; k++)
                if (next_n + ex + k < BVR_TOP_BFS_MAX) frontier[cur ^ 1][next_n + ex + k] = kids[k];
            next_n += total;
            __syncthreads();
        }
        cur ^= 1;
        cur_n = min(next_n, BVR_TOP_BFS_MAX);
    }
    if (tid == 0u) id_q[n] = done;   // number of breadth-first ids handed out
}

// the other inner nodes: id = (breadth-first ids handed out) + position among them in array order
__global__ void rest_count_kernel(const uint32_t* __restrict__ inner_id, const uint32_t* __restrict__ id_q, uint32_t n,
                                  uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t warp_sums[32];
    const uint32_t i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const uint32_t flag = (i < n && inner_id[i] != 0xffffffffu && id_q[i] == 0xffffffffu) ? 1u : 0u;
    uint32_t total;
    (void)block_exclusive_scan(flag, warp_sums, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}
__global__ void rest_assign_kernel(const uint32_t* __restrict__ inner_id, uint32_t* __restrict__ id_q, uint32_t n,
                                   const uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t warp_sums[32];
    const uint32_t i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const uint32_t flag = (i < n && inner_id[i] != 0xffffffffu && id_q[i] == 0xffffffffu) ? 1u : 0u;
    uint32_t total;
    const uint32_t ex = block_exclusive_scan(flag, warp_sums, total);
    if (flag) id_q[i] = id_q[n] + block_sums[blockIdx.x] + ex;
}

__device__ __forceinline__ uint32_t make_ref(const RawNode& child, uint32_t child_inner_id) {
    if (child.model_count > 0u)
        return BVR_LEAF_BIT | ((child.model_count - 1u) << 24) | (child.index & BVR_LEAF_FIRST_MASK);
    return child_inner_id;
}

// pass 4: one 64-byte child-pair record per inner node
__global__ void build_pairs_kernel(const RawNode* __restrict__ nodes, uint32_t n,
                                   const uint32_t* __restrict__ inner_id, float4* __restrict__ pairs,
                                   float4* __restrict__ pairs_ch, uint32_t* __restrict__ root_ref_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const RawNode nd = nodes[i];
    if (i == 0u) *root_ref_out = make_ref(nd, nd.model_count == 0u ? inner_id[0] : 0u);
    if (nd.model_count != 0u) return;
    const uint32_t c0i = nd.index, c1i = nd.index + 1u;
    const RawNode c0 = nodes[c0i], c1 = nodes[c1i];
    const uint32_t r0 = make_ref(c0, inner_id[c0i]), r1 = make_ref(c1, inner_id[c1i]);
    float4* out = pairs + 4u * inner_id[i];
    out[0] = make_float4(c0.mn[0], c0.mn[1], c0.mn[2], c0.mx[0]);
    out[1] = make_float4(c0.mx[1], c0.mx[2], c1.mn[0], c1.mn[1]);
    out[2] = make_float4(c1.mn[2], c1.mx[0], c1.mx[1], c1.mx[2]);
    out[3] = make_float4(__uint_as_float(r0), __uint_as_float(r1), 0.0f, 0.0f);
    // centre / half-extent form for the culling-only slab test: the half extents are rounded UP, so that
    // [c - h, c + h] encloses [min, max] exactly — the box can only grow
    float c[2][3], h[2][3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        c[0][k] = __fmul_rn(0.5f, __fadd_rn(c0.mn[k], c0.mx[k]));
        h[0][k] = fmaxf(__fsub_ru(c0.mx[k], c[0][k]), __fsub_ru(c[0][k], c0.mn[k]));
        c[1][k] = __fmul_rn(0.5f, __fadd_rn(c1.mn[k], c1.mx[k]));
        h[1][k] = fmaxf(__fsub_ru(c1.mx[k], c[1][k]), __fsub_ru(c[1][k], c1.mn[k]));
    }
    float4* och = pairs_ch + 4u * inner_id[i];
    och[0] = make_float4(c[0][0], c[0][1], c[0][2], h[0][0]);
    och[1] = make_float4(h[0][1], h[0][2], c[1][0], c[1][1]);
    och[2] = make_float4(c[1][2], h[1][0], h[1][1], h[1][2]);
    och[3] = out[3];
}

// Quantisation grid of the 32-byte records: 65536 steps per axis across the root box plus two steps of margin on
// either side (so that the outward rounding below never has to clamp inside the root box).
struct QGrid { float base[3], step[3]; };
#define BVR_Q16_MAX_STEP 0.005f   // two steps = a tenth of the reference's 0.1 pad

__device__ __forceinline__ QGrid make_qgrid(const RawNode& root) {
    QGrid g;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        g.step[k] = fmaxf(__fdiv_ru(__fsub_ru(root.mx[k], root.mn[k]), 65531.0f), 1e-30f);
        g.base[k] = __fsub_rd(root.mn[k], __fmul_ru(2.0f, g.step[k]));
    }
    return g;
}

// pass 4b: one 32-byte record per inner node: child boxes as 16-bit grid coordinates, rounded OUTWARDS by one
// extra step (the decode in the kernel is one FFMA whose constant term carries up to half a step of rounding
// error, megakernel_v3.cu).  Word layout: (c0.x, c0.y, c0.z, ref0, c1.x, c1.y, c1.z, ref1), each coordinate word
// = lo | hi << 16; refs in 21-bit form: bit 20 = leaf, bits 0-19 = inner record / first (only) model.
// `bad` is raised when a box leaves the root box or a ref does not fit: the fp32 records are used then.
__global__ void build_pairs_q16_kernel(const RawNode* __restrict__ nodes, uint32_t n,
                                       const uint32_t* __restrict__ inner_id, uint4* __restrict__ pairs_q,
                                       float* __restrict__ grid_out, uint32_t* __restrict__ bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const QGrid g = make_qgrid(nodes[0]);
    if (i == 0u) {
#pragma unroll
        for (int k = 0; k < 3; k++) { grid_out[k] = g.base[k]; grid_out[4 + k] = g.step[k]; }
        grid_out[3] = 0.0f; grid_out[7] = 0.0f;
        // Outward rounding moves a box face by up to two grid steps.  That must stay small against the reference's
        // own pad of 0.1 (extract.rs:223-224): where the pad is all that separates "ray enters the box" from "f32
        // noise in hit_sphere", a visibly larger box finds hits the reference culls (tools/big_fuzz.py case 11:
        // extent 3400, step 0.05 -> 5 of 44150 rays differ).  Scenes wider than ~330 units keep the fp32 records.
        if (g.step[0] > BVR_Q16_MAX_STEP || g.step[1] > BVR_Q16_MAX_STEP || g.step[2] > BVR_Q16_MAX_STEP) *bad = 1u;
    }
    const RawNode nd = nodes[i];
    if (nd.model_count != 0u) {
        if (nd.model_count > 1u || nd.index >= (1u << 20)) *bad = 1u;
        return;
    }
    uint32_t w[8];
    bool is_bad = inner_id[i] >= (1u << 20);
#pragma unroll
    for (int c = 0; c < 2; c++) {
        const RawNode ch = nodes[nd.index + (uint32_t)c];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            // floor / ceil of the exact grid coordinate, then one step outwards
            const float lo = floorf(__fdiv_rd(__fsub_rd(ch.mn[k], g.base[k]), g.step[k])) - 1.0f;
            const float hi = ceilf(__fdiv_ru(__fsub_ru(ch.mx[k], g.base[k]), g.step[k])) + 1.0f;
            if (!(lo >= 0.0f) || !(hi <= 65535.0f) || !(lo <= hi)) is_bad = true;   // outside the grid, or NaN
            const uint32_t ql = (uint32_t)fminf(fmaxf(lo, 0.0f), 65535.0f), qh = (uint32_t)fminf(fmaxf(hi, 0.0f), 65535.0f);
            w[c * 4 + k] = ql | (qh << 16);
        }
        const uint32_t cid = inner_id[nd.index + (uint32_t)c];
        w[c * 4 + 3] = ch.model_count > 0u ? ((1u << 20) | (ch.index & 0xfffffu)) : cid;
    }
    if (is_bad) *bad = 1u;
    uint4* out = pairs_q + 2u * inner_id[i];
    out[0] = make_uint4(w[0], w[1], w[2], w[3]);
    out[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// pass 4c: 4-wide records for the latency-bound walk of big scenes.  The record of inner node N holds the
// children of N's two children (a leaf child stays as it is, next to an empty slot), i.e. two 32-byte pair
// records side by side: the walk descends two levels of the reference tree per dependent 64-byte fetch.  Only
// the records of even-depth nodes are ever reached from the root; the others are built and never read.
__global__ void build_nodes4_q16_kernel(const RawNode* __restrict__ nodes, uint32_t n,
                                        const uint32_t* __restrict__ inner_id, const uint4* __restrict__ pairs_q,
                                        uint4* __restrict__ nodes4) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const RawNode nd = nodes[i];
    if (nd.model_count != 0u) return;
    const uint32_t id = inner_id[i];
    uint4* out = nodes4 + 4u * id;
    const uint4 empty = make_uint4(0x0000ffffu, 0x0000ffffu, 0x0000ffffu, 0x200000u);   // lo > hi: never entered
#pragma unroll
    for (uint32_t s = 0; s < 2u; s++) {
        const uint32_t x = nd.index + s;
        if (nodes[x].model_count > 0u) {
            out[2u * s] = pairs_q[2u * id + s];
            out[2u * s + 1u] = empty;
        } else {
            const uint32_t xid = inner_id[x];
            out[2u * s] = pairs_q[2u * xid];
            out[2u * s + 1u] = pairs_q[2u * xid + 1u];
        }
    }
}

// pass 4d: 4-wide fp32 records for scenes staged in shared memory (megakernel_v3.cu MODE 4/5): 112 bytes per inner
// node = four child boxes as (centre, half extent) + four refs in 11-bit form (bit 10 = leaf, bits 0-9 = inner
// record / only model of the leaf; 0x7ff = empty slot, whose negative half extent is never entered).
// The floats are ordered for the packed FFMA2 slab test (two fp32 FMAs per instruction, sm_100):
//   q[i] = (c_i.x, c_i.y, h_i.x, h_i.y)  i = 0..3      one box, axes x and y side by side
//   q[4] = (c_0.z, c_1.z, h_0.z, h_1.z)                axis z of boxes 0 and 1 side by side
//   q[5] = (c_2.z, c_3.z, h_2.z, h_3.z)
//   q[6] = refs
__global__ void build_nodes4_ch_kernel(const RawNode* __restrict__ nodes, uint32_t n,
                                       const uint32_t* __restrict__ inner_id, const float4* __restrict__ pairs_ch,
                                       float4* __restrict__ nodes4) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const RawNode nd = nodes[i];
    if (nd.model_count != 0u) return;
    const uint32_t id = inner_id[i];
    float box[4][6];   // (c.xyz, h.xyz)
    uint32_t ref[4];
    auto take = [&](int slot, uint32_t rec, uint32_t which) {   // child `which` of pair record `rec`
        const float4 q0 = pairs_ch[4u * rec], q1 = pairs_ch[4u * rec + 1u], q2 = pairs_ch[4u * rec + 2u], q3 = pairs_ch[4u * rec + 3u];
        if (which == 0u) { box[slot][0] = q0.x; box[slot][1] = q0.y; box[slot][2] = q0.z; box[slot][3] = q0.w; box[slot][4] = q1.x; box[slot][5] = q1.y; }
        else { box[slot][0] = q1.z; box[slot][1] = q1.w; box[slot][2] = q2.x; box[slot][3] = q2.y; box[slot][4] = q2.z; box[slot][5] = q2.w; }
        const uint32_t r = __float_as_uint(which == 0u ? q3.x : q3.y);
        ref[slot] = (r & BVR_LEAF_BIT) ? (0x400u | (r & 0x3ffu)) : r;
    };
#pragma unroll
    for (uint32_t s = 0; s < 2u; s++) {
        const uint32_t x = nd.index + s;
        if (nodes[x].model_count > 0u) {
            take(2 * s, id, s);
            for (int k = 0; k < 3; k++) { box[2 * s + 1][k] = 0.0f; box[2 * s + 1][3 + k] = -1.0f; }
            ref[2 * s + 1] = 0x7ffu;
        } else {
            take(2 * s, inner_id[x], 0u);
            take(2 * s + 1, inner_id[x], 1u);
        }
    }
    float4* out = nodes4 + 7u * id;
#pragma unroll
    for (int k = 0; k < 4; k++) out[k] = make_float4(box[k][0], box[k][1], box[k][3], box[k][4]);
    out[4] = make_float4(box[0][2], box[1][2], box[0][5], box[1][5]);
    out[5] = make_float4(box[2][2], box[3][2], box[2][5], box[3][5]);
    out[6] = make_float4(__uint_as_float(ref[0]), __uint_as_float(ref[1]), __uint_as_float(ref[2]), __uint_as_float(ref[3]));
}

// ---- tight boxes for scenes staged in shared memory ----------------------------------------------
// The reference pads every sphere box by 0.1 (extract.rs:220-227).  A hit needs disc >= 0 in hit_sphere's f32
// arithmetic; with L = |origin - centre| the rounding error of disc is below ERR * L^2 * |d|^2, so such a ray
// passes within sqrt(r^2 + ERR * L^2) of the centre: a box padded by `pad` cannot lose the hit as long as
// ERR * L^2 <= 2 r pad + pad^2.  Spheres are grouped by radius octave; per group the kernel emits the bounding
// ball (C, R) of the centres and the squared distance D2 = (sqrt((2 rmin pad + pad^2) / ERR) - R)^2 inside which a
// ray origin satisfies that bound for every sphere of the group.  A ray whose origin lies outside any group's
// ball ("far") walks the reference boxes instead.  ERR = 4 x 18 x 2^-24 (18 roundings on the way to disc, x4).
__device__ __forceinline__ uint32_t ordered_bits(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__global__ void tight_groups_kernel(const float4* __restrict__ spheres, uint32_t n, float pad, float err,
                                    float4* __restrict__ groups_out) {
    __shared__ uint32_t lo[32][3], hi[32][3], rmin[32];
    __shared__ uint32_t bad;
    const uint32_t tid = threadIdx.x;
    if (tid < 32u) {
        for (int k = 0; k < 3; k++) { lo[tid][k] = 0xffffffffu; hi[tid][k] = 0u; }
        rmin[tid] = 0x7f800000u;
    }
    if (tid == 0u) bad = 0u;
    __syncthreads();
    for (uint32_t i = tid; i < n; i += blockDim.x) {
        const float4 sp = spheres[i];
        const float r = fabsf(sp.w);
        if (!(r < 3.0e38f) || !(fabsf(sp.x) < 3.0e38f) || !(fabsf(sp.y) < 3.0e38f) || !(fabsf(sp.z) < 3.0e38f)) { bad = 1u; continue; }
        int g = (int)(__float_as_uint(r) >> 23) - 127 + 16;
        g = g < 0 ? 0 : (g > 31 ? 31 : g);
        atomicMin(&lo[g][0], ordered_bits(sp.x)); atomicMax(&hi[g][0], ordered_bits(sp.x));
        atomicMin(&lo[g][1], ordered_bits(sp.y)); atomicMax(&hi[g][1], ordered_bits(sp.y));
        atomicMin(&lo[g][2], ordered_bits(sp.z)); atomicMax(&hi[g][2], ordered_bits(sp.z));
        atomicMin(&rmin[g], __float_as_uint(r));
    }
    __syncthreads();
    if (tid < 32u) {
        float4 out = make_float4(0.f, 0.f, 0.f, -1.0f);
        const bool used = rmin[tid] != 0x7f800000u;
        if (used) {
            float c[3], ext2 = 0.0f;
            for (int k = 0; k < 3; k++) {
                const float a = from_ordered_bits(lo[tid][k]), b = from_ordered_bits(hi[tid][k]);
                c[k] = 0.5f * (a + b);
                const float h = 0.5f * (b - a);
                ext2 += h * h;
            }
            const float R = sqrtf(ext2) * 1.0001f + 1e-6f;
            const float rm = __uint_as_float(rmin[tid]);
            const float D = sqrtf((2.0f * rm * pad + pad * pad) / err) * 0.999f - R;
            out = make_float4(c[0], c[1], c[2], (bad == 0u && D > 0.0f) ? D * D : -1.0f);
        }
        const unsigned m = __ballot_sync(0xffffffffu, used);
        if (used) groups_out[1 + __popc(m & ((1u << tid) - 1u))] = out;
        if (tid == 0u) groups_out[0] = make_float4(__uint_as_float((uint32_t)__popc(m)), 0.f, 0.f, 0.f);
    }
}

// Copy of the node array with every box shrunk to (union of its spheres padded by `pad`) ∩ (uploaded box): one CTA,
// bottom-up by rounds (at most 2047 nodes).
__global__ void tight_refit_kernel(const RawNode* __restrict__ nodes, uint32_t n, const float4* __restrict__ spheres,
                                   float pad, RawNode* __restrict__ out) {
    __shared__ unsigned char done[2048];
    const uint32_t tid = threadIdx.x;
    for (uint32_t i = tid; i < n; i += blockDim.x) {
        RawNode nd = nodes[i];
        done[i] = 0;
        if (nd.model_count > 0u) {
            float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
            for (uint32_t m = nd.index; m < nd.index + nd.model_count; m++) {
                const float4 sp = spheres[m];
                const float e = __fadd_ru(fabsf(sp.w), pad);
                const float c[3] = {sp.x, sp.y, sp.z};
                for (int k = 0; k < 3; k++) { mn[k] = fminf(mn[k], __fsub_rd(c[k], e)); mx[k] = fmaxf(mx[k], __fadd_ru(c[k], e)); }
            }
            for (int k = 0; k < 3; k++) { nd.mn[k] = fmaxf(nd.mn[k], mn[k]); nd.mx[k] = fminf(nd.mx[k], mx[k]); }
            done[i] = 1;
        }
        out[i] = nd;
    }
    __syncthreads();
    for (int round = 0; round < 2048; round++) {
        bool progress = false;
        uint32_t mine[2] = {0xffffffffu, 0xffffffffu};
        int k2 = 0;
        for (uint32_t i = tid; i < n; i += blockDim.x, k2++) {
            if (done[i]) continue;
            const RawNode nd = nodes[i];
            const uint32_t c0 = nd.index, c1 = nd.index + 1u;
            if (c1 < n && done[c0] && done[c1]) {
                const RawNode a = out[c0], b = out[c1];
                RawNode o = nd;
                for (int k = 0; k < 3; k++) {
                    o.mn[k] = fmaxf(nd.mn[k], fminf(a.mn[k], b.mn[k]));
                    o.mx[k] = fminf(nd.mx[k], fmaxf(a.mx[k], b.mx[k]));
                }
                out[i] = o;
                if (k2 < 2) mine[k2] = i;
                progress = true;
            }
        }
        const int any = __syncthreads_or(progress ? 1 : 0);
        for (int k = 0; k < 2; k++) if (mine[k] != 0xffffffffu) done[mine[k]] = 1;
        __syncthreads();
        if (!any) break;
    }
}

}  // namespace

// Tight variant of the small-scene records: groups (33 float4: count, then (C.xyz, D2) per group), the refitted node
// copy, and from it the usual derived records.  Caller guarantees n_nodes <= 2047.
int launch_derive_tight(const RawNode* nodes, uint32_t n_nodes, const float4* spheres, uint32_t n_models, float pad,
                        float4* groups, RawNode* nodes_tight, cudaStream_t stream) {
    if (n_nodes == 0 || n_nodes > 2047u) return 0;
    const float err = 4.0f * 18.0f * 5.9604645e-8f;
    tight_groups_kernel<<<1, 1024, 0, stream>>>(spheres, n_models, pad, err, groups);
    tight_refit_kernel<<<1, 1024, 0, stream>>>(nodes, n_nodes, spheres, pad, nodes_tight);
    return 2;
}

int launch_derive_nodes4_ch(const RawNode* nodes, uint32_t n_nodes, const uint32_t* inner_id, const float4* pairs_ch,
                            float4* nodes4, cudaStream_t stream) {
    if (n_nodes == 0) return 0;
    build_nodes4_ch_kernel<<<(n_nodes + 255) / 256, 256, 0, stream>>>(nodes, n_nodes, inner_id, pairs_ch, nodes4);
    return 1;
}

int launch_derive_nodes4_q16(const RawNode* nodes, uint32_t n_nodes, const uint32_t* inner_id, const uint4* pairs_q,
                             uint4* nodes4, cudaStream_t stream) {
    if (n_nodes == 0) return 0;
    build_nodes4_q16_kernel<<<(n_nodes + 255) / 256, 256, 0, stream>>>(nodes, n_nodes, inner_id, pairs_q, nodes4);
    return 1;
}

int launch_derive_pairs_q16(const RawNode* nodes, uint32_t n_nodes, const uint32_t* inner_id, uint4* pairs_q,
                            float* grid, uint32_t* bad, cudaStream_t stream) {
    if (n_nodes == 0) return 0;
    cudaMemsetAsync(bad, 0, sizeof(uint32_t), stream);
    build_pairs_q16_kernel<<<(n_nodes + 255) / 256, 256, 0, stream>>>(nodes, n_nodes, inner_id, pairs_q, grid, bad);
    return 1;
}

int launch_derive_top_order(const RawNode* nodes, uint32_t n_nodes, const uint32_t* inner_id, uint32_t* block_sums,
                            uint32_t max_top, uint32_t* id_q, cudaStream_t stream) {
    if (n_nodes == 0) return 0;
    const uint32_t n_blocks = (n_nodes + SCAN_BLOCK - 1) / SCAN_BLOCK;
    cudaMemsetAsync(id_q, 0xff, ((size_t)n_nodes + 1u) * sizeof(uint32_t), stream);
    top_bfs_kernel<<<1, SCAN_BLOCK, 0, stream>>>(nodes, n_nodes, max_top, id_q);
    rest_count_kernel<<<n_blocks, SCAN_BLOCK, 0, stream>>>(inner_id, id_q, n_nodes, block_sums);
    scan_block_sums_kernel<<<1, SCAN_BLOCK, 0, stream>>>(block_sums, n_blocks);
    rest_assign_kernel<<<n_blocks, SCAN_BLOCK, 0, stream>>>(inner_id, id_q, n_nodes, block_sums);
    return 4;
}

int launch_derive_spheres(const RawModel* models, uint32_t n, float4* spheres, uint32_t* sphere_material,
                          cudaStream_t stream) {
    if (n == 0) return 0;
    derive_spheres_kernel<<<(n + 255) / 256, 256, 0, stream>>>(models, n, spheres, sphere_material);
    return 1;
}

int launch_derive_pairs(const RawNode* nodes, uint32_t n_nodes, uint32_t* inner_id, uint32_t* block_sums,
                        float4* pairs, float4* pairs_ch, uint32_t* root_ref_out, cudaStream_t stream) {
    if (n_nodes == 0) return 0;
    const uint32_t n_blocks = (n_nodes + SCAN_BLOCK - 1) / SCAN_BLOCK;
    count_inner_kernel<<<n_blocks, SCAN_BLOCK, 0, stream>>>(nodes, n_nodes, block_sums);
    scan_block_sums_kernel<<<1, SCAN_BLOCK, 0, stream>>>(block_sums, n_blocks);
    assign_inner_id_kernel<<<n_blocks, SCAN_BLOCK, 0, stream>>>(nodes, n_nodes, block_sums, inner_id);
    build_pairs_kernel<<<(n_nodes + 255) / 256, 256, 0, stream>>>(nodes, n_nodes, inner_id, pairs, pairs_ch, root_ref_out);
    return 4;
}

}  // namespace bvr
