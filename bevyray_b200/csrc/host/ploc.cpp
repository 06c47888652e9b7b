// ploc.cpp — host BVH2 builder emitting the reference's BVHNode contract.
//
// Replaces the call `obvhs::ploc::build_ploc::<24>(&aabbs, 0..n, SortPrecision::U64, 0)` at
// src/raytracing/extract.rs:316-321 and the node mapping at extract.rs:323-332.  obvhs 0.1.x is a
// third-party crate whose source is not in the reference tree, so this is a restatement of the
// published PLOC algorithm (Meister & Bittner, "Parallel Locally-Ordered Clustering for Bounding
// Volume Hierarchy Construction", TVCG 2018) with the same parameters (search radius 24, 64-bit
// Morton keys = 21 bits per axis, one primitive per leaf).  Topology parity with obvhs is UNPINNED;
// the closest hit does not depend on topology.
//
// Node contract (raytrace.wgsl:80-87, 313-346): node 0 is the root; an inner node has
// model_count == 0 and its children at index, index+1; a leaf has model_count >= 1 and index = first
// model in model-buffer order.  n models -> 2n-1 nodes; n == 1 -> a single leaf root; n == 0 -> no nodes.

#include "bevyray_host.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace bevyray {

namespace {

struct Box {
    float mn[3], mx[3];
};

inline Box box_union(const Box& a, const Box& b) {
    Box r;
    for (int k = 0; k < 3; k++) {
        r.mn[k] = std::fmin(a.mn[k], b.mn[k]);
        r.mx[k] = std::fmax(a.mx[k], b.mx[k]);
    }
    return r;
}

inline float half_area(const Box& b) {
    float dx = b.mx[0] - b.mn[0], dy = b.mx[1] - b.mn[1], dz = b.mx[2] - b.mn[2];
    return dx * dy + dy * dz + dz * dx;
}

// spread the low 21 bits of v so that there are two zero bits between each
inline uint64_t spread21(uint64_t v) {
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x1f00000000ffffull;
    v = (v | (v << 16)) & 0x1f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}

// LSD radix sort of (key, value) pairs by 64-bit key, stable.
void radix_sort_pairs(std::vector<uint64_t>& keys, std::vector<uint32_t>& vals) {
    const size_t n = keys.size();
    std::vector<uint64_t> k2(n);
    std::vector<uint32_t> v2(n);
    for (int pass = 0; pass < 8; pass++) {
        const int shift = pass * 8;
        size_t hist[257] = {0};
        for (size_t i = 0; i < n; i++) hist[((keys[i] >> shift) & 0xff) + 1]++;
        bool trivial = false;
        for (int b = 0; b < 256; b++) if (hist[b + 1] == n) trivial = true;
        if (trivial) continue;
        for (int b = 0; b < 256; b++) hist[b + 1] += hist[b];
        for (size_t i = 0; i < n; i++) {
            size_t d = hist[(keys[i] >> shift) & 0xff]++;
            k2[d] = keys[i];
            v2[d] = vals[i];
        }
        keys.swap(k2);
        vals.swap(v2);
    }
}

// A cluster in the working list: its bounds and, once merged, where its two children live.
struct Cluster {
    Box box;
    uint32_t index;  // leaf: model index; inner: slot of the first child in the output array
    uint32_t count;  // leaf: 1; inner: 0
};

inline void write_node(BvrBvhNode& out, const Cluster& c) {
    std::memset(&out, 0, sizeof out);
    for (int k = 0; k < 3; k++) { out.bounds_min[k] = c.box.mn[k]; out.bounds_max[k] = c.box.mx[k]; }
    out.index = c.index;
    out.model_count = c.count;
}

}  // namespace

// Model::aabb — src/raytracing/extract.rs:220-227: centre -/+ (radius + 0.1)
void model_aabb(const BvrModel& m, float mn[3], float mx[3]) {
    const float pad = m.radius + 0.1f;
    for (int k = 0; k < 3; k++) {
        mn[k] = m.position[k] - pad;
        mx[k] = m.position[k] + pad;
    }
}

std::vector<BvrBvhNode> build_ploc(const std::vector<BvrModel>& models, uint32_t search_distance) {
    const size_t n = models.size();
    std::vector<BvrBvhNode> nodes;
    if (n == 0) return nodes;
    if (search_distance == 0) search_distance = 1;

    // 1. leaf boxes + scene bounds of the centroids
    std::vector<Cluster> cur(n);
    Box cb;
    for (int k = 0; k < 3; k++) { cb.mn[k] = std::numeric_limits<float>::infinity(); cb.mx[k] = -cb.mn[k]; }
    for (size_t i = 0; i < n; i++) {
        model_aabb(models[i], cur[i].box.mn, cur[i].box.mx);
        cur[i].index = (uint32_t)i;
        cur[i].count = 1;
        for (int k = 0; k < 3; k++) {
            float c = 0.5f * (cur[i].box.mn[k] + cur[i].box.mx[k]);
            cb.mn[k] = std::fmin(cb.mn[k], c);
            cb.mx[k] = std::fmax(cb.mx[k], c);
        }
    }
    nodes.resize(2 * n - 1);
    if (n == 1) { write_node(nodes[0], cur[0]); return nodes; }

    // 2. 63-bit Morton keys (21 bits / axis) of the centroids, sort
    std::vector<uint64_t> keys(n);
    std::vector<uint32_t> order(n);
    double scale[3];
    for (int k = 0; k < 3; k++) {
        double ext = (double)cb.mx[k] - (double)cb.mn[k];
        scale[k] = ext > 0.0 ? 2097152.0 / ext : 0.0;
    }
    for (size_t i = 0; i < n; i++) {
        uint64_t q[3];
        for (int k = 0; k < 3; k++) {
            double c = 0.5 * ((double)cur[i].box.mn[k] + (double)cur[i].box.mx[k]);
            double v = (c - (double)cb.mn[k]) * scale[k];
            q[k] = (uint64_t)std::min(2097151.0, std::max(0.0, std::floor(v)));
        }
        keys[i] = (spread21(q[0]) << 2) | (spread21(q[1]) << 1) | spread21(q[2]);
        order[i] = (uint32_t)i;
    }
    radix_sort_pairs(keys, order);
    {
        std::vector<Cluster> sorted(n);
        for (size_t i = 0; i < n; i++) sorted[i] = cur[order[i]];
        cur.swap(sorted);
    }

    // 3. PLOC iterations: nearest neighbour within +-search_distance, merge mutual pairs, compact.
    // Child pairs are written from the back of the array so that the root lands in slot 0 and the
    // top of the tree is contiguous at the front.
    size_t insert = 2 * n - 1;  // one past the last free slot
    std::vector<uint32_t> nn(n);
    std::vector<Cluster> next(n);
    size_t m = n;
    const size_t r = (size_t)search_distance;
    // dist[i*r + (k-1)] = merged half-area of clusters i and i+k: every pair is evaluated once and read
    // from both sides (the backward neighbours of i are the forward entries of i-k)
    std::vector<float> dist(n * r);
    while (m > 1) {
        const int64_t mm = (int64_t)m;
#pragma omp parallel for schedule(static) if (m > 2048)
        for (int64_t i = 0; i < mm; i++) {
            float* di = &dist[(size_t)i * r];
            const Box& bi = cur[(size_t)i].box;
            const size_t kmax = std::min(r, (size_t)(mm - 1 - i));
            for (size_t k = 1; k <= kmax; k++) di[k - 1] = half_area(box_union(bi, cur[(size_t)i + k].box));
        }
#pragma omp parallel for schedule(static) if (m > 2048)
        for (int64_t i = 0; i < mm; i++) {
            float best = std::numeric_limits<float>::infinity();
            uint32_t best_j = 0xffffffffu;
            // backward neighbours first (lower index wins ties), then forward
            const size_t kback = std::min(r, (size_t)i);
            for (size_t k = kback; k >= 1; k--) {
                const float a = dist[((size_t)i - k) * r + (k - 1)];
                if (a < best) { best = a; best_j = (uint32_t)((size_t)i - k); }
            }
            const size_t kmax = std::min(r, (size_t)(mm - 1 - i));
            for (size_t k = 1; k <= kmax; k++) {
                const float a = dist[(size_t)i * r + (k - 1)];
                if (a < best) { best = a; best_j = (uint32_t)((size_t)i + k); }
            }
            nn[(size_t)i] = best_j;
        }
        size_t out = 0;
        for (size_t i = 0; i < m; i++) {
            const uint32_t j = nn[i];
            if (nn[j] == i) {
                if (i < j) {  // merge; the pair's nodes get the two highest free slots
                    insert -= 2;
                    write_node(nodes[insert], cur[i]);
                    write_node(nodes[insert + 1], cur[j]);
                    Cluster c;
                    c.box = box_union(cur[i].box, cur[j].box);
                    c.index = (uint32_t)insert;
                    c.count = 0;
                    next[out++] = c;
                }  // else: absorbed into the merge at position j
            } else {
                next[out++] = cur[i];
            }
        }
        cur.swap(next);
        m = out;
    }
    write_node(nodes[0], cur[0]);
    return nodes;
}

// Structural validation of a node array against the contract above.  Returns an empty string when
// valid, otherwise a description of the first violation.
std::string validate_bvh(const std::vector<BvrBvhNode>& nodes, const std::vector<BvrModel>& models) {
    const size_t n = models.size();
    if (n == 0) return nodes.empty() ? "" : "nodes without models";
    if (nodes.empty()) return "no nodes";
    std::vector<uint8_t> seen_model(n, 0), seen_node(nodes.size(), 0);
    std::vector<uint32_t> stack{0};
    seen_node[0] = 1;
    while (!stack.empty()) {
        uint32_t i = stack.back(); stack.pop_back();
        const BvrBvhNode& nd = nodes[i];
        if (nd.model_count > 0) {
            for (uint32_t k = nd.index; k < nd.index + nd.model_count; k++) {
                if (k >= n) return "leaf model index out of range";
                if (seen_model[k]) return "model referenced twice";
                seen_model[k] = 1;
                float mn[3], mx[3];
                model_aabb(models[k], mn, mx);
                for (int a = 0; a < 3; a++)
                    if (nd.bounds_min[a] > mn[a] || nd.bounds_max[a] < mx[a]) return "leaf bounds do not enclose model";
            }
        } else {
            for (uint32_t c = nd.index; c < nd.index + 2; c++) {
                if (c >= nodes.size()) return "child index out of range";
                if (seen_node[c]) return "node referenced twice";
                seen_node[c] = 1;
                for (int a = 0; a < 3; a++)
                    if (nd.bounds_min[a] > nodes[c].bounds_min[a] || nd.bounds_max[a] < nodes[c].bounds_max[a])
                        return "inner bounds do not enclose child";
                stack.push_back(c);
            }
        }
    }
    for (size_t k = 0; k < n; k++) if (!seen_model[k]) return "model not referenced";
    for (size_t k = 0; k < nodes.size(); k++) if (!seen_node[k]) return "unreachable node";
    return "";
}

}  // namespace bevyray
