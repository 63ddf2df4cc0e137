// vt_bvh_ploc.cpp — the reference's OWN hierarchy, rebuilt from its published algorithm: the build sequence of
// source/objects/AccelStruct.cpp:762-770 = bvh::LocallyOrderedClusteringBuilder<BVH, uint32_t> (PLOC, Meister & Bittner;
// libs/bvh/include/bvh/locally_ordered_clustering_builder.hpp) followed by bvh::LeafCollapser
// (libs/bvh/include/bvh/leaf_collapser.hpp).
//
// Why: the product's default builder (vt_bvh_build.cpp, binned SAH) gives a better tree (42 vs 58 pair visits per bounce ray
// at 5 M triangles), but WHICH of two exactly tied candidates a ray reports depends on the tree's visit order.  With
// VT_BUILDER=ploc (or vt_build_bvh_ploc) the engine builds node for node the array the shipped module would build — same
// boxes, same child order, same primitive order — so the exact layout reproduces the reference's tie winners without the
// caller having to hand its tree over (vt_accel_populate_with_bvh).  Checked against the compiled reference:
// tests/test_host.py::test_ploc_builder_reproduces_the_reference_hierarchy (nodes and primitive_indices equal bit for bit).
//
// What has to be reproduced exactly, all of it deterministic and independent of the thread count:
//   * Triangle::bounding_box / center (source/objects/Primitives.h:107-118) and their union (utilities.hpp:163-175);
//   * MortonEncoder (morton.hpp:37-60): world_to_grid = 1024 * (1 / diagonal), grid_offset = -min * world_to_grid,
//     cell = min(1023, uint32(max(p * world_to_grid + grid_offset, 0))) — two roundings, float -> uint32 truncation;
//     morton_split's mask ladder (morton.hpp:14-26), x | y << 1 | z << 2;
//   * a STABLE sort by code (radix_sort.hpp is a stable LSD radix sort);
//   * cluster() (locally_ordered_clustering_builder.hpp:39-165): search radius 14, distance = half_area of the union,
//     backward candidates first and strict `<` (the lowest index wins ties), mutual nearest neighbours merge, the parent
//     lands at the slot derived from the HIGHER index, children in (lower, higher) order, output region [begin - m, end);
//   * LeafCollapser::collapse (leaf_collapser.hpp:36-148): collapse where half_area * (count - 1) <= the children's
//     half_area * count sums, new indices by inclusive prefix sums over the node ARRAY order, primitives gathered left to right.
// Outside the claim: scenes with an axis of zero extent (every centre on one plane or line) — the reference converts a NaN to
// an unsigned integer there, which is undefined behaviour (see the Morton loop below); 250 random scenes of every other kind
// (soups, duplicates, degenerate triangles, far-apart clusters, coordinates of 1e7) match bit for bit.
// half_area is (d0 + d1) * d2 + d0 * d1 (bounding_box.hpp:43-46), uncontracted (-ffp-contract=off, like the reference build).
#include <algorithm>
#include <cstring>
#include <limits>
#include <numeric>
#include <parallel/algorithm>

#include "vt_host.h"

namespace vt {

namespace {

struct BBox {
    float lo[3], hi[3];
};

inline BBox node_box(const vt_node &n) { return BBox{{n.bounds[0], n.bounds[2], n.bounds[4]}, {n.bounds[1], n.bounds[3], n.bounds[5]}}; }
inline void set_box(vt_node &n, const BBox &b) {
    for (int a = 0; a < 3; a++) n.bounds[2 * a] = b.lo[a], n.bounds[2 * a + 1] = b.hi[a];
}
inline BBox extend(BBox a, const BBox &b) {  // BoundingBox::extend: per-component std::min / std::max (bounding_box.hpp:23-27)
    for (int k = 0; k < 3; k++) a.lo[k] = std::min(a.lo[k], b.lo[k]), a.hi[k] = std::max(a.hi[k], b.hi[k]);
    return a;
}
inline float half_area(const BBox &b) {
    const float d0 = b.hi[0] - b.lo[0], d1 = b.hi[1] - b.lo[1], d2 = b.hi[2] - b.lo[2];
    return (d0 + d1) * d2 + d0 * d1;
}

uint32_t morton_split(uint32_t x) {  // morton.hpp:14-26 for a 32-bit code
    uint32_t mask = 0xFFFFFFFFu >> 16;
    x &= mask;
    for (uint32_t i = 4, n = 16; i > 0; --i, n >>= 1) {
        mask = (mask | (mask << n)) & ~(mask << (n / 2));
        x = (x | (x << n)) & mask;
    }
    return x;
}

constexpr size_t kSearchRadius = 14;  // LocallyOrderedClusteringBuilder::search_radius

}  // namespace

void build_bvh_ploc(const TriangleVec &tris, HostBvh &out) {
    const size_t n = tris.size();
    out.nodes.clear();
    out.prim_indices.clear();
    if (n == 0) return;
    // ---- compute_bounding_boxes_and_centers + union
    std::vector<BBox> boxes(n);
    std::vector<float> centers(3 * n);
    BBox global{{std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()},
                {-std::numeric_limits<float>::max(), -std::numeric_limits<float>::max(), -std::numeric_limits<float>::max()}};
#pragma omp parallel
    {
        BBox local = global;
#pragma omp for nowait
        for (int64_t i = 0; i < (int64_t)n; i++) {
            const Triangle &t = tris[i];
            BBox b;
            for (int k = 0; k < 3; k++) {
                const float p0 = t.p0[k], p1 = t.p0[k] - t.e1[k], p2 = t.p0[k] + t.e2[k];
                b.lo[k] = std::min(std::min(p0, p1), p2);  // bbox(p0).extend(p1()).extend(p2())
                b.hi[k] = std::max(std::max(p0, p1), p2);
                centers[3 * i + k] = ((p0 + p1) + p2) * (1.0f / 3.0f);
            }
            boxes[i] = b;
            local = extend(local, b);
        }
#pragma omp critical
        global = extend(global, local);
    }
    // ---- Morton codes (10 bits per axis) and the stable sort
    const float grid_dim = 1024.0f;
    float w2g[3], off[3];
    for (int k = 0; k < 3; k++) {
        w2g[k] = grid_dim * (1.0f / (global.hi[k] - global.lo[k]));
        off[k] = -global.lo[k] * w2g[k];
    }
    std::vector<uint32_t> codes(n);
#pragma omp parallel for
    for (int64_t i = 0; i < (int64_t)n; i++) {
        uint32_t c[3];
        for (int k = 0; k < 3; k++) {
            // morton.hpp:52-57 converts max(g, 0) to unsigned and clamps to 1023.  On an axis of zero extent g is NaN (c * inf - inf)
            // and that conversion is undefined behaviour in the reference: whatever its compiler made of it is not reproducible,
            // so such scenes are outside the bit-identity claim; here NaN (and anything <= 0) maps to cell 0, deterministically.
            const float g = centers[3 * i + k] * w2g[k] + off[k];
            c[k] = g > 0.0f ? (g >= 1023.0f ? 1023u : (uint32_t)g) : 0u;
        }
        codes[i] = morton_split(c[0]) | (morton_split(c[1]) << 1) | (morton_split(c[2]) << 2);
    }
    // stable by code = plain sort of the unique keys (code << 32 | original index); libstdc++'s OpenMP parallel sort
    std::vector<uint64_t> order(n);
#pragma omp parallel for
    for (int64_t i = 0; i < (int64_t)n; i++) order[i] = ((uint64_t)codes[i] << 32) | (uint64_t)i;
    __gnu_parallel::sort(order.begin(), order.end());
#pragma omp parallel for
    for (int64_t i = 0; i < (int64_t)n; i++) order[i] &= 0xFFFFFFFFull;

    // ---- leaves at the end of the array, then level after level of clustering towards index 0
    const size_t node_count = 2 * n - 1;
    std::vector<vt_node> nodes(node_count);
    size_t begin = node_count - n, end = node_count;
#pragma omp parallel for
    for (int64_t i = 0; i < (int64_t)n; i++) {
        vt_node &nd = nodes[begin + i];
        set_box(nd, boxes[order[i]]);
        nd.prim_count = 1;
        nd.first = (uint32_t)i;
    }
    std::vector<vt_node> active;
    std::vector<size_t> neighbors, merged;
    while (end - begin > 1) {
        const size_t count = end - begin;
        active.assign(nodes.begin() + begin, nodes.begin() + end);
        neighbors.resize(count);
        merged.resize(count);
        // nearest neighbour inside the search window; the lowest index wins ties (backward candidates are visited first)
#pragma omp parallel for schedule(static) if (count > 256)
        for (int64_t ii = 0; ii < (int64_t)count; ii++) {
            const size_t i = (size_t)ii;
            const size_t lo = i > kSearchRadius ? i - kSearchRadius : 0, hi = std::min(i + kSearchRadius + 1, count);
            const BBox bi = node_box(active[i]);
            float best = std::numeric_limits<float>::max();
            size_t best_j = (size_t)-1;
            for (size_t j = lo; j < hi; j++) {
                if (j == i) continue;
                const float d = half_area(extend(bi, node_box(active[j])));
                if (d < best) best = d, best_j = j;
            }
            // every distance infinite / NaN (the reference asserts here): keep going with the adjacent node
            neighbors[i] = best_j != (size_t)-1 ? best_j : (i + 1 < count ? i + 1 : i - 1);
        }
        // pairs of mutual nearest neighbours merge; the lower index leads
        size_t merged_count = 0;
        for (int attempt = 0; attempt < 2 && merged_count == 0; attempt++) {
            // Finite boxes always contain a mutual pair (the closest two).  Non-finite ones do not have to: NaN planes make the
            // distance asymmetric (std::min / std::max keep their FIRST argument on an unordered compare), and a cycle a -> b -> c -> a
            // without any mutual pair would repeat this level for ever.  Then the first two nodes are merged, so every level shrinks.
            if (attempt == 1) neighbors[0] = 1, neighbors[1] = 0;
            for (size_t i = 0; i < count; i++) {
                const size_t j = neighbors[i];
                merged_count += (i < j && neighbors[j] == i) ? 1 : 0;
                merged[i] = merged_count;  // inclusive prefix sum
            }
        }
        const size_t children_begin = end - 2 * merged_count;
        const size_t unmerged_begin = begin - merged_count;  // = end - (2 m + (count - m))
#pragma omp parallel for schedule(static) if (count > 256)
        for (int64_t ii = 0; ii < (int64_t)count; ii++) {
            const size_t i = (size_t)ii, j = neighbors[i];
            if (neighbors[j] == i) {
                if (i < j) {
                    vt_node &parent = nodes[unmerged_begin + j - merged[j]];
                    const size_t first_child = children_begin + (merged[i] - 1) * 2;
                    set_box(parent, extend(node_box(active[j]), node_box(active[i])));
                    parent.prim_count = 0;
                    parent.first = (uint32_t)first_child;
                    nodes[first_child] = active[i];
                    nodes[first_child + 1] = active[j];
                }
            } else {
                nodes[unmerged_begin + i - merged[i]] = active[i];
            }
        }
        begin = unmerged_begin;
        end = children_begin;
    }
    out.nodes.assign(nodes.begin(), nodes.end());
    out.prim_indices.assign(order.begin(), order.end());
}

// bvh::LeafCollapser::collapse.  When the root itself becomes a leaf the reference turns node 0 into a leaf over all
// primitives in their Morton order and keeps one node (leaf_collapser.hpp:87-94: the two array swaps there cancel with the two
// at :142-143, node_count ends up as the surviving count, 1); same here.  Always returns true.
bool collapse_leaves(HostBvh &bvh) {
    const size_t node_count = bvh.nodes.size();
    if (node_count == 0 || bvh.nodes[0].prim_count != 0) return true;
    std::vector<size_t> node_counts(node_count, 1), prim_counts(node_count, 0), parents(node_count, 0);
    std::vector<uint32_t> order;
    order.reserve(node_count);
    std::vector<uint32_t> stack{0u};
    while (!stack.empty()) {  // pre-order; its reverse visits children before parents (any bottom-up order gives the same decisions)
        const uint32_t i = stack.back();
        stack.pop_back();
        order.push_back(i);
        const vt_node &nd = bvh.nodes[i];
        if (nd.prim_count == 0) {
            parents[nd.first] = parents[nd.first + 1] = i;
            stack.push_back(nd.first);
            stack.push_back(nd.first + 1);
        }
    }
    const float traversal_cost = 1.0f;  // SahBasedAlgorithm::traversal_cost (sah_based_algorithm.hpp:16)
    for (size_t k = order.size(); k-- > 0;) {
        const uint32_t i = order[k];
        const vt_node &nd = bvh.nodes[i];
        if (nd.prim_count != 0) {
            prim_counts[i] = nd.prim_count;
            continue;
        }
        const size_t l = nd.first, r = l + 1;
        const size_t lc = prim_counts[l], rc = prim_counts[r], total = lc + rc;
        if (lc > 0 && rc > 0) {
            const float collapse_cost = half_area(node_box(nd)) * ((float)total - traversal_cost);
            const float base_cost = half_area(node_box(bvh.nodes[l])) * (float)lc + half_area(node_box(bvh.nodes[r])) * (float)rc;
            if (collapse_cost <= base_cost) {
                prim_counts[i] = total;
                prim_counts[l] = prim_counts[r] = 0;
                node_counts[l] = node_counts[r] = 0;
            }
        }
    }
    if (prim_counts[0] > 0) {
        vt_node root = bvh.nodes[0];
        root.first = 0;
        root.prim_count = (uint32_t)prim_counts[0];
        bvh.nodes.assign(1, root);
        return true;
    }
    for (size_t i = 1; i < node_count; i++) node_counts[i] += node_counts[i - 1], prim_counts[i] += prim_counts[i - 1];  // inclusive, array order
    prim_counts[0] = 0;  // the root is inner: nothing before node 1
    const size_t new_count = node_counts[node_count - 1];
    std::vector<vt_node> nodes(new_count);
    std::vector<uint64_t> prims(prim_counts[node_count - 1]);
    nodes[0] = bvh.nodes[0];
    nodes[0].first = (uint32_t)node_counts[bvh.nodes[0].first - 1];
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t ii = 1; ii < (int64_t)node_count; ii++) {
        const size_t i = (size_t)ii;
        const size_t node_index = node_counts[i - 1];
        if (node_index == node_counts[i]) continue;  // swallowed by a collapsed ancestor
        vt_node nd = bvh.nodes[i];
        size_t first_primitive = prim_counts[i - 1];
        if (first_primitive != prim_counts[i]) {
            nd.prim_count = (uint32_t)(prim_counts[i] - first_primitive);
            nd.first = (uint32_t)first_primitive;
            size_t j = i;  // left-to-right walk over the original leaves below i
            for (;;) {
                const vt_node &c = bvh.nodes[j];
                if (c.prim_count != 0) {
                    std::copy(bvh.prim_indices.begin() + c.first, bvh.prim_indices.begin() + c.first + c.prim_count, prims.begin() + first_primitive);
                    first_primitive += c.prim_count;
                    // is_left_sibling(j) = j is odd (children sit at first, first + 1 with `first` odd: bvh.hpp:68-78)
                    while (j % 2 == 0 && j != i) j = parents[j];
                    if (j == i) break;
                    j = j + 1;  // sibling of a left child
                } else {
                    j = c.first;
                }
            }
        } else {
            nd.first = (uint32_t)node_counts[nd.first - 1];
        }
        nodes[node_index] = nd;
    }
    bvh.nodes.assign(nodes.begin(), nodes.end());
    bvh.prim_indices.assign(prims.begin(), prims.end());
    return true;
}

}  // namespace vt
