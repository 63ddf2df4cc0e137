// vt_bvh_collapse.cpp — choosing the children of the wide (4- / 8-wide) nodes: SAH-optimal collapse of the binary hierarchy by
// dynamic programming (Ylitie, Karras, Laine, HPG 2017, section 3.1; the published algorithm restated, no code of theirs).
//
//   cost(n, 1) = A(n) * c_node + distribute(n, W)              n becomes a wide node with up to W children        (inner n)
//   cost(n, i) = min(distribute(n, i), cost(n, i - 1))         n's subtree presented as at most i roots, 1 < i < W
//   distribute(n, j) = min over 0 < k < j of cost(left, k) + cost(right, j - k)
//   cost(leaf, i) = A(leaf) * triangles * c_prim
//
// Which plan is used (VT_COLLAPSE = auto | dp | greedy, default auto): the SAH-optimal plan lowers the number of wide nodes a ray
// visits on surfaces and separate objects (quad visits per bounce ray 21.28 -> 21.02 on the 5 M-triangle terrain) but on a volume of
// heavily overlapping alpha-tested cards it ruins the front-to-back order of camera rays (config 4, primary K1: 3.10 ms greedy,
// 5.72 ms SAH-optimal; 78.8 -> 109.7 visits, 83 -> 124 triangle tests per ray — session r4c).  auto measures the sibling overlap
// of the binary tree (sibling_overlap below) and keeps the round-1 rule "adopt the children of the largest child" above 0.225 (VT_COLLAPSE_OVERLAP) — the same figure
// switches the kernel's child order to entry + exit (vt_traverse.cu: VT_KEY_MID), which removes most of that loss by itself.
//
// Leaves are kept as the builder made them (no triangles are merged or moved), so the set of triangles a ray can reach through a
// given box is unchanged; only how many box tests and dependent fetches it takes to get there.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include "vt_host.h"

namespace vt {

namespace {
float half_area(const vt_node &n) {
    const float dx = n.bounds[1] - n.bounds[0], dy = n.bounds[3] - n.bounds[2], dz = n.bounds[5] - n.bounds[4];
    return dx * dy + dy * dz + dz * dx;
}
}  // namespace

// Area-weighted overlap of sibling boxes: sum over inner nodes of A(left box ^ right box) / sum of A(node).  The SAH prices a tree for
// rays that pass through everything; how well an ORDERED traversal terminates early depends on how much siblings overlap (Aila,
// Karras, Laine, "On Quality Metrics of Bounding Volume Hierarchies", HPG 2013).  Measured on the product builder's trees: 0.15 - 0.20
// for surfaces and separate objects (terrain, props), 0.35 - 0.44 for the alpha-tested foliage volume (profiles/r2_child_order.md).
double sibling_overlap(const HostBvh &bvh) {
    const size_t n = bvh.nodes.size();
    double overlap = 0.0, total = 0.0;
#pragma omp parallel for reduction(+ : overlap, total) schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        const vt_node &nd = bvh.nodes[i];
        if (nd.prim_count != 0 || nd.first == 0 || (size_t)nd.first + 1 >= n) continue;
        const vt_node &l = bvh.nodes[nd.first], &r = bvh.nodes[nd.first + 1];
        double e[3];
        bool apart = false;
        for (int a = 0; a < 3; a++) {
            e[a] = (double)std::min(l.bounds[2 * a + 1], r.bounds[2 * a + 1]) - (double)std::max(l.bounds[2 * a], r.bounds[2 * a]);
            apart = apart || !(e[a] >= 0.0);
        }
        if (!apart) overlap += e[0] * e[1] + e[1] * e[2] + e[2] * e[0];
        total += (double)half_area(nd);
    }
    return total > 0.0 ? overlap / total : 0.0;
}

bool plan_collapse(const HostBvh &bvh, int width, CollapsePlan &plan, std::string &err) {
    plan.width = width;
    plan.split.clear();
    const char *mode = std::getenv("VT_COLLAPSE");
    plan.greedy = mode && std::string(mode) == "greedy";
    const size_t n = bvh.nodes.size();
    // The array may be a caller's (vt_accel_populate_with_bvh, vt_build_quads): before anything follows a child reference, every
    // inner node must name an odd-indexed pair inside the array and no pair may have two parents — then whatever the root reaches is
    // a tree and every walk below terminates.  (CollapsePlan::children follows grandchildren without further checks.)
    if (n > 1 || (n == 1 && bvh.nodes[0].prim_count == 0)) {
        std::vector<uint8_t> referenced((n + 1) / 2, 0);
        for (size_t i = 0; i < n; i++) {
            const vt_node &nd = bvh.nodes[i];
            if (nd.prim_count != 0) continue;
            if (nd.first == 0 || (nd.first & 1u) == 0 || (size_t)nd.first + 1 >= n) {
                err = "BVH child index invalid (children must be an adjacent pair at an odd index)";
                return false;
            }
            if (referenced[nd.first / 2]++) {
                err = "BVH is not a tree (a node pair is referenced twice)";
                return false;
            }
        }
    }
    plan.sibling_overlap = sibling_overlap(bvh);
    if (!(mode && *mode) || std::string(mode) == "auto") {
        const char *th = std::getenv("VT_COLLAPSE_OVERLAP");
        plan.greedy = plan.sibling_overlap > ((th && *th) ? std::atof(th) : 0.225);
    }
    if (plan.greedy || n == 0 || bvh.nodes[0].prim_count != 0) return true;
    if (width < 2 || width > 8) {
        err = "collapse: width must be 2..8";
        return false;
    }
    const float c_node = 1.0f;
    const char *cp = std::getenv("VT_COLLAPSE_CPRIM");
    const float c_prim = (cp && *cp) ? (float)std::atof(cp) : 0.6f;
    const int W = width;
    std::vector<float> cost(n * (size_t)W);  // cost[node * W + (i - 1)], i = 1 .. W - 1 used
    plan.split.assign(n * (size_t)W, 0);
    std::atomic<int> bad{0};
    // post-order over the binary tree as OpenMP tasks (children before parents); the depth bound also stops a malformed "tree"
    struct Dp {
        const HostBvh &bvh;
        int W;
        float c_node, c_prim;
        float *cost;
        uint8_t *split;
        std::atomic<int> &bad;
        void run(uint32_t ni, int depth) const {
            const size_t n = bvh.nodes.size();
            const vt_node &nd = bvh.nodes[ni];
            float *c = cost + (size_t)ni * W;
            uint8_t *sp = split + (size_t)ni * W;
            if (nd.prim_count != 0) {
                const float leaf = half_area(nd) * (float)nd.prim_count * c_prim;
                for (int i = 0; i < W; i++) c[i] = leaf;
                return;
            }
            if (nd.first == 0 || (size_t)nd.first + 1 >= n || depth > 128) {
                bad = 1;
                for (int i = 0; i < W; i++) c[i] = 0.f;
                return;
            }
            if (depth < 12) {
#pragma omp task
                run(nd.first, depth + 1);
#pragma omp task
                run(nd.first + 1, depth + 1);
#pragma omp taskwait
            } else {
                run(nd.first, depth + 1);
                run(nd.first + 1, depth + 1);
            }
            const float *cl = cost + (size_t)nd.first * W, *cr = cost + ((size_t)nd.first + 1) * W;
            // distribute(n, j) for j = 2 .. W; a child may take 1 .. W - 1 roots
            float dist[9];
            uint8_t dk[9];
            for (int j = 2; j <= W; j++) {
                float best = std::numeric_limits<float>::max();
                int bk = 1;
                for (int k = 1; k < j; k++) {
                    const float v = cl[k - 1] + cr[j - k - 1];
                    if (v < best) best = v, bk = k;
                }
                dist[j] = best, dk[j] = (uint8_t)bk;
            }
            c[0] = half_area(nd) * c_node + dist[W];
            sp[0] = dk[W];  // the wide node's own children: forest of W roots split dk[W] : W - dk[W]
            for (int i = 2; i < W; i++) {
                if (dist[i] < c[i - 2]) c[i - 1] = dist[i], sp[i - 1] = dk[i];
                else c[i - 1] = c[i - 2], sp[i - 1] = 0;  // one root fewer is cheaper
            }
        }
    } dp{bvh, W, c_node, c_prim, cost.data(), plan.split.data(), bad};
#pragma omp parallel
#pragma omp single
    dp.run(0, 0);
    if (bad.load()) {
        err = "collapse: malformed hierarchy";
        return false;
    }
    return true;
}

int CollapsePlan::children(const HostBvh &bvh, uint32_t ni, uint32_t *kids) const {
    const uint32_t first = bvh.nodes[ni].first;
    if (greedy || split.empty()) {  // adopt the children of the inner child with the largest box until the node is full
        int nk = 2;
        kids[0] = first, kids[1] = first + 1;
        while (nk < width) {
            int best = -1;
            float best_area = -1.f;
            for (int i = 0; i < nk; i++) {
                const vt_node &c = bvh.nodes[kids[i]];
                if (c.prim_count == 0 && half_area(c) > best_area) best_area = half_area(c), best = i;
            }
            if (best < 0) break;
            const uint32_t f = bvh.nodes[kids[best]].first;
            for (int i = nk; i > best + 1; i--) kids[i] = kids[i - 1];  // keep left-to-right order
            kids[best] = f, kids[best + 1] = f + 1;
            nk++;
        }
        return nk;
    }
    // forest(node, i): the roots that present `node`'s subtree with a budget of i
    struct Item {
        uint32_t node;
        int budget;
    };
    Item todo[16];
    int nt = 0, nk = 0;
    const int k0 = split[(size_t)ni * width];
    todo[nt++] = {first + 1, width - k0};
    todo[nt++] = {first, k0};
    while (nt) {
        const Item it = todo[--nt];
        const vt_node &nd = bvh.nodes[it.node];
        int b = it.budget;
        if (nd.prim_count != 0) {
            kids[nk++] = it.node;
            continue;
        }
        int k = 0;
        while (b > 1 && (k = split[(size_t)it.node * width + (b - 1)]) == 0) b--;  // "one root fewer" chain
        if (b <= 1) {
            kids[nk++] = it.node;  // stays one root: a wide node of its own (or it would have been split)
            continue;
        }
        todo[nt++] = {nd.first + 1, b - k};
        todo[nt++] = {nd.first, k};
    }
    return nk;
}

}  // namespace vt
