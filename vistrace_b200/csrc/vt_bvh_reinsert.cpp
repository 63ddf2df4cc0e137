// vt_bvh_reinsert.cpp — builder-quality option: reinsertion optimisation of a finished hierarchy.
//
// The reference's bvh library offers bvh::ParallelReinsertionOptimizer (libs/bvh/include/bvh/parallel_reinsertion_optimizer.hpp,
// after Meister & Bittner, "Parallel Reinsertion for Bounding Volume Hierarchy Optimization", 2018) as a post-pass for any
// builder; the reference itself does not call it (source/objects/AccelStruct.cpp:762-775).  This file is the product's own
// version of that pass over the bvh::Bvh<float>-form arrays (vt_node: sibling pairs adjacent, root at 0), off by default
// (VT_REINSERT=<iterations>, vt_optimize_bvh): a node with a large box is taken out of the tree together with its parent and put
// back where the sum of the inner-node areas — the SAH's traversal term, sah_based_algorithm.hpp:16-41 — grows least.
//
// One iteration:
//   1. candidates = the `fraction` of the nodes (not the root or its children) with the largest half-areas;
//   2. every candidate searches its best new position IN PARALLEL on the unchanged tree (branch and bound: walking up from the
//      parent, the sub-tree hanging off each ancestor is searched with the area already saved below as credit);
//   3. the moves are applied one after the other, best gain first; a move whose nodes another move of this iteration has
//      touched is dropped (its gain was computed for a tree that no longer exists);
//   4. all inner boxes are recomputed bottom-up.
// Leaves keep their primitive ranges, so prim_indices is untouched and every triangle stays in its leaf.  The result is laid out
// depth-first again (children behind their parents); a pass that would make the tree deeper than the 64-entry traversal stack of
// single_ray_traverser.hpp:14 allows (the builder's own bound is 60) ends the optimisation with the result of the pass before.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

#include <omp.h>

#include "vt_host.h"

namespace vt {

namespace {

constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr int kMaxDepth = 60;

struct Bx {
    float lo[3], hi[3];
};

inline Bx box_of(const vt_node &n) { return Bx{{n.bounds[0], n.bounds[2], n.bounds[4]}, {n.bounds[1], n.bounds[3], n.bounds[5]}}; }
inline void store_box(vt_node &n, const Bx &b) {
    for (int a = 0; a < 3; a++) n.bounds[2 * a] = b.lo[a], n.bounds[2 * a + 1] = b.hi[a];
}
inline Bx join(const Bx &a, const Bx &b) {
    Bx r;
    for (int k = 0; k < 3; k++) r.lo[k] = std::min(a.lo[k], b.lo[k]), r.hi[k] = std::max(a.hi[k], b.hi[k]);
    return r;
}
inline float half_area(const Bx &b) {
    const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    return dx * dy + dy * dz + dz * dx;
}
inline uint32_t sibling(uint32_t i) { return (i & 1u) ? i + 1 : i - 1; }  // pairs are (1,2), (3,4), ...

struct Move {
    uint32_t in, out;
    float gain;
};

struct Tree {
    vt_node *nodes;
    uint32_t count;
    std::vector<uint32_t> parent;

    void link_parents() {
        parent.assign(count, kNone);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < (int64_t)count; i++)
            if (nodes[i].prim_count == 0) parent[nodes[i].first] = parent[nodes[i].first + 1] = (uint32_t)i;
    }

    // Pre-order of the tree as it is linked now (parents before children); false when it is not a tree of `count` nodes.
    bool preorder(std::vector<uint32_t> &order, std::vector<uint8_t> *depth_out, int &max_depth) const {
        order.clear();
        order.reserve(count);
        std::vector<std::pair<uint32_t, int>> stack{{0u, 0}};
        max_depth = 0;
        while (!stack.empty()) {
            const auto [i, d] = stack.back();
            stack.pop_back();
            if (order.size() >= count) return false;
            order.push_back(i);
            max_depth = std::max(max_depth, d);
            if (depth_out) (*depth_out)[i] = (uint8_t)std::min(d, 255);
            if (nodes[i].prim_count == 0) {
                stack.push_back({nodes[i].first + 1, d + 1});
                stack.push_back({nodes[i].first, d + 1});
            }
        }
        return order.size() == count;
    }

    void refit(const std::vector<uint32_t> &order) {
        for (size_t k = order.size(); k-- > 0;) {  // children before parents
            vt_node &n = nodes[order[k]];
            if (n.prim_count != 0) continue;
            store_box(n, join(box_of(nodes[n.first]), box_of(nodes[n.first + 1])));
        }
    }

    double inner_area() const {
        double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
        for (int64_t i = 0; i < (int64_t)count; i++)
            if (nodes[i].prim_count == 0) s += half_area(box_of(nodes[i]));
        return s;
    }

    // Best position for `in` on the current tree.  gain = area(parent) [the parent disappears] + the shrink of the path above it
    // up to the pivot - the growth of the path from the pivot's other child down to `out` - area(out + in) [the new parent].
    Move search(uint32_t in) const {
        Move best{in, kNone, 0.f};
        const Bx in_box = box_of(nodes[in]);
        const float in_area = half_area(in_box);
        const uint32_t p = parent[in];
        const float parent_area = half_area(box_of(nodes[p]));
        struct Item {
            uint32_t node;
            float credit;
        };
        Item stack[128];
        auto descend = [&](uint32_t root, float credit) {
            int sp = 0;
            stack[sp++] = Item{root, credit};
            while (sp) {
                const Item it = stack[--sp];
                if (parent_area + it.credit - in_area <= best.gain) continue;  // even a free slot below cannot beat the best
                const vt_node &n = nodes[it.node];
                const Bx nb = box_of(n);
                const float merged = half_area(join(nb, in_box));
                const float gain = parent_area + it.credit - merged;
                if (gain > best.gain) best.gain = gain, best.out = it.node;
                if (n.prim_count == 0 && sp + 2 <= 128) {
                    const float below = it.credit - (merged - half_area(nb));  // this node grows when `in` goes underneath
                    stack[sp++] = Item{n.first, below};
                    stack[sp++] = Item{n.first + 1, below};
                }
            }
        };
        const uint32_t sib = sibling(in);
        descend(sib, 0.f);  // under the parent itself: the sibling moves up, `in` may go back in below it
        if (best.out == sib) best.out = kNone, best.gain = 0.f;  // the position it already has (rounding aside)
        Bx shrunk = box_of(nodes[sib]);  // the box of the path node once `in` is gone
        float credit = 0.f;
        for (uint32_t cur = p; cur != 0; cur = parent[cur]) {
            const uint32_t other = sibling(cur), up = parent[cur];
            descend(other, credit);
            shrunk = join(shrunk, box_of(nodes[other]));
            credit += half_area(box_of(nodes[up])) - half_area(shrunk);
        }
        return best;
    }

    // Take `in` and its parent out, hang `in` next to `out` under the freed parent slot.  Pair slots never split: the pair
    // {in, sibling} becomes the pair {in, old out}, the old sibling's record moves up into the parent's slot.
    void apply(const Move &m) {
        const uint32_t in = m.in, out = m.out, sib = sibling(in), p = parent[in];
        const vt_node sib_rec = nodes[sib], out_rec = nodes[out];
        nodes[p] = sib_rec;
        if (sib_rec.prim_count == 0) parent[sib_rec.first] = parent[sib_rec.first + 1] = p;
        nodes[sib] = out_rec;
        if (out_rec.prim_count == 0) parent[out_rec.first] = parent[out_rec.first + 1] = sib;
        vt_node fresh;
        std::memset(&fresh, 0, sizeof(fresh));
        store_box(fresh, join(box_of(out_rec), box_of(nodes[in])));
        fresh.prim_count = 0;
        fresh.first = std::min(in, sib);
        nodes[out] = fresh;
        parent[in] = parent[sib] = out;
    }
};

}  // namespace

bool reinsert_optimize(HostBvh &bvh, int iterations, float fraction, double *area_before, double *area_after, uint64_t *moves_out) {
    const size_t n = bvh.nodes.size();
    if (area_before) *area_before = 0.0;
    if (area_after) *area_after = 0.0;
    if (moves_out) *moves_out = 0;
    if (n < 7 || iterations <= 0 || n >= 0xFFFFFFF0ull) return true;  // nothing to move below the root's children
    fraction = std::min(std::max(fraction, 0.001f), 0.5f);
    RawVector<vt_node> work(n);
    std::memcpy(work.data(), bvh.nodes.data(), n * sizeof(vt_node));
    Tree t{work.data(), (uint32_t)n, {}};
    t.link_parents();
    for (uint32_t i = 1; i < n; i++)
        if (t.parent[i] == kNone) return false;  // not the pair-linked tree this pass is written for
    std::vector<uint32_t> order, cand(n - 3);
    int max_depth = 0;
    {   // every node reachable from the root exactly once (a pair referenced twice next to a detached cycle would pass the count)
        if (!t.preorder(order, nullptr, max_depth)) return false;
        std::vector<uint8_t> seen(n, 0);
        for (uint32_t i : order)
            if (seen[i]++) return false;
    }
    const double before = t.inner_area();
    std::vector<float> area(n);
    std::vector<Move> moves;
    std::vector<uint32_t> touched(n, 0);
    uint64_t applied = 0, applied_good = 0;
    RawVector<vt_node> good;  // the tree after the last pass that kept it within the depth bound (empty: the tree as built)
    double area_good = before;
    for (int it = 0; it < iterations; it++) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < (int64_t)n; i++) area[i] = half_area(box_of(t.nodes[i]));
        std::iota(cand.begin(), cand.end(), 3u);
        // the batch shrinks and grows again over the iterations so that small boxes get their turn as well
        const float f = fraction * (1.f + (float)(it % 3));
        const size_t batch = std::min(cand.size(), std::max<size_t>(1, (size_t)((double)n * std::min(f, 0.5f))));
        std::nth_element(cand.begin(), cand.begin() + (batch - 1), cand.end(), [&](uint32_t a, uint32_t b) { return area[a] > area[b] || (area[a] == area[b] && a < b); });
        moves.assign(batch, Move{0, kNone, 0.f});
#pragma omp parallel for schedule(dynamic, 256)
        for (int64_t k = 0; k < (int64_t)batch; k++) moves[k] = t.search(cand[k]);
        moves.erase(std::remove_if(moves.begin(), moves.end(), [](const Move &m) { return m.out == kNone || !(m.gain > 0.f); }), moves.end());
        if (moves.empty()) break;
        std::sort(moves.begin(), moves.end(), [](const Move &a, const Move &b) { return a.gain > b.gain || (a.gain == b.gain && a.in < b.in); });
        const uint32_t stamp = (uint32_t)it + 1;
        uint64_t done = 0;
        for (const Move &m : moves) {
            const uint32_t ids[5] = {m.in, sibling(m.in), t.parent[m.in], m.out, t.parent[m.out]};
            bool clash = false;
            for (uint32_t id : ids) clash = clash || touched[id] == stamp;
            // a move is only valid on the tree it was searched on: `out` must still be outside the sub-tree of `in` and off its
            // root path, which untouched nodes guarantee
            if (clash) continue;
            // earlier moves of this iteration may have carried `out` underneath `in` (its sub-tree now hangs off a node that was
            // moved there): hanging `in` below itself would close a cycle
            bool below_in = false;
            for (uint32_t a = m.out; a != kNone && !below_in; a = t.parent[a]) below_in = a == m.in;
            if (below_in || m.out == ids[1] || m.out == ids[2]) continue;
            t.apply(m);
            for (uint32_t id : ids) touched[id] = stamp;
            touched[t.parent[m.in]] = stamp;
            done++;
        }
        applied += done;
        if (!t.preorder(order, nullptr, max_depth)) return false;
        t.refit(order);
        if (done == 0) break;
        if (max_depth > kMaxDepth) break;  // deeper than the traversal stack allows: keep the result of the pass before
        const double now = t.inner_area();
        if (now < area_good) {
            good.resize(n);
            std::memcpy(good.data(), t.nodes, n * sizeof(vt_node));
            area_good = now, applied_good = applied;
        }
    }
    if (area_before) *area_before = before;
    if (area_after) *area_after = before;
    if (good.empty()) return true;  // keep the tree as it was built
    std::memcpy(t.nodes, good.data(), n * sizeof(vt_node));
    if (!t.preorder(order, nullptr, max_depth)) return false;
    const double after = area_good;
    applied = applied_good;
    // depth-first relayout: root at 0, a node's children adjacent and behind it, left sub-tree before the right one
    std::vector<uint32_t> new_pair(n, 0);  // old index of the first child of a pair -> new index
    {
        uint32_t next = 1;
        for (uint32_t i : order)
            if (t.nodes[i].prim_count == 0) new_pair[t.nodes[i].first] = next, next += 2;
    }
    auto new_index = [&](uint32_t old) { return old == 0 ? 0u : new_pair[(old & 1u) ? old : old - 1] + ((old & 1u) ? 0u : 1u); };
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        vt_node rec = t.nodes[i];
        if (rec.prim_count == 0) rec.first = new_pair[rec.first];
        bvh.nodes[new_index((uint32_t)i)] = rec;
    }
    if (area_after) *area_after = after;
    if (moves_out) *moves_out = applied;
    return true;
}

}  // namespace vt
