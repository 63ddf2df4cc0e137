// vt_bvh_build.cpp — host-side hierarchy construction (the step at
// source/objects/AccelStruct.cpp:762-770) and flattening to the device layout.
//
// Construction stays on the host, as in the reference, and emits the same data structure the
// reference traverses: bvh::Bvh<float> form (libs/bvh/include/bvh/bvh.hpp:17-99) — root at node 0,
// the two children of an inner node adjacent at `first`, `first + 1`, leaves addressing a run of
// `prim_indices`.  The builder itself is our own: a top-down binned-SAH build (OpenMP tasks over
// subtrees) with the reference's cost model (traversal cost 1 vs. one unit per primitive,
// libs/bvh/include/bvh/sah_based_algorithm.hpp:16) and a hard depth bound of 60 so the 64-entry
// traversal stack (single_ray_traverser.hpp:14) can never overflow.  A caller that prefers the
// reference's own PLOC + LeafCollapser tree passes it in through vt_accel_populate_with_bvh.
#include "vt_host.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <omp.h>

namespace vt {

namespace {

struct Box {
    float lo[3], hi[3];
    void reset() {
        for (int k = 0; k < 3; k++) {
            lo[k] = std::numeric_limits<float>::max();
            hi[k] = -std::numeric_limits<float>::max();
        }
    }
    void grow(const float *mn, const float *mx) {
        for (int k = 0; k < 3; k++) {
            lo[k] = std::min(lo[k], mn[k]);
            hi[k] = std::max(hi[k], mx[k]);
        }
    }
    void grow(const Box &b) { grow(b.lo, b.hi); }
    void grow_pt(const float *p) { grow(p, p); }
    float half_area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return dx * dy + dy * dz + dz * dx;
    }
    bool valid() const { return lo[0] <= hi[0]; }
};

constexpr int kBins = 16;
constexpr int kMaxDepth = 60;
constexpr uint32_t kParallelNode = 1u << 17;  // nodes with at least this many primitives are binned / partitioned by the whole team

struct BuildCtx {
    const float *bmin, *bmax, *cen;  // n x 3 each
    uint32_t *idx, *scratch;         // scratch: n entries, the target of the parallel partition
    vt_node *nodes;
    std::atomic<uint32_t> next_node;
    int max_leaf;
    float trav_cost;
    uint32_t sweep_below;  // nodes with at most this many primitives are split by the exact SAH sweep instead of 16 bins (0 = never)
};

void set_bounds(vt_node &nd, const Box &b) {
    nd.bounds[0] = b.lo[0];
    nd.bounds[1] = b.hi[0];
    nd.bounds[2] = b.lo[1];
    nd.bounds[3] = b.hi[1];
    nd.bounds[4] = b.lo[2];
    nd.bounds[5] = b.hi[2];
}

int ceil_log2(uint64_t v) {
    int l = 0;
    while ((1ull << l) < v) l++;
    return l;
}

// One bin of one axis: the primitives whose centroid falls into it — their box, the box of their centroids, their number.
struct Bin {
    Box box, cbox;
    uint32_t count;
    // the boxes of an empty bin are never read (every reader checks `count`) and the first primitive ASSIGNS them, so clearing a
    // bin is one store — 2.6 M inner nodes x 48 bins of 52 bytes would otherwise be 6.5 GB of resets at 5 M triangles
    void add(const float *mn, const float *mx, const float *ce) {
        if (count++ == 0) {
            for (int k = 0; k < 3; k++) box.lo[k] = mn[k], box.hi[k] = mx[k], cbox.lo[k] = cbox.hi[k] = ce[k];
        } else {
            box.grow(mn, mx);
            cbox.grow_pt(ce);
        }
    }
    void merge(const Bin &o) {
        if (!o.count) return;
        if (!count) {
            *this = o;
            return;
        }
        box.grow(o.box);
        cbox.grow(o.cbox);
        count += o.count;
    }
};
struct Bins {
    Bin b[3][kBins];
    void reset() {
        for (int a = 0; a < 3; a++)
            for (int k = 0; k < kBins; k++) b[a][k].count = 0;
    }
};

inline int bin_of(float c, float lo, float scale) {
    int b = (int)((c - lo) * scale);
    return b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
}

// ONE pass over idx[begin, end): all three axes binned at once.  Child boxes and child centroid boxes fall out of the bins
// (min / max are exact, so they equal what a pass over the child's primitives would give), which leaves two passes per node —
// this one and the partition — instead of seven.
void bin_range(const BuildCtx &c, uint32_t begin, uint32_t end, const Box &cb, const float scale[3], Bins &out) {
    out.reset();
    for (uint32_t i = begin; i < end; i++) {
        const size_t p = c.idx[i];
        const float *mn = c.bmin + 3 * p, *mx = c.bmax + 3 * p, *ce = c.cen + 3 * p;
        for (int a = 0; a < 3; a++) {
            if (!(scale[a] > 0.f)) continue;
            out.b[a][bin_of(ce[a], cb.lo[a], scale[a])].add(mn, mx, ce);
        }
    }
}

// Build the subtree over idx[begin, end) into nodes[node]; `box` is its bounds, `cb` the bounds of its centroids.
void build_range(BuildCtx &c, uint32_t node, uint32_t begin, uint32_t end, const Box &box, const Box &cb, int depth) {
    vt_node &nd = c.nodes[node];
    set_bounds(nd, box);
    const uint32_t count = end - begin;
    auto make_leaf = [&]() {
        nd.prim_count = count;
        nd.first = begin;
    };
    if (count <= 1) return make_leaf();

    // binned SAH over the three axes
    float best_cost = std::numeric_limits<float>::max();
    int best_axis = -1, best_bin = -1;
    float scale[3];
    for (int a = 0; a < 3; a++) {
        const float ext = cb.hi[a] - cb.lo[a];
        scale[a] = ext > 0.f ? kBins / ext : 0.f;
    }
    const bool force_balance = depth + ceil_log2((count + c.max_leaf - 1) / c.max_leaf) >= kMaxDepth - 1;
    Bins bins;
    const bool big = count >= kParallelNode;
    if (!force_balance) {
        if (big) {  // the top of the tree: every thread bins a slice, the slices are merged
            const int parts = std::max(1, std::min(omp_get_num_threads() * 4, (int)(count / 32768)));
            std::vector<Bins> part(parts);
#pragma omp taskloop shared(c, part, cb, scale) grainsize(1)
            for (int t = 0; t < parts; t++) {
                const uint32_t b0 = begin + (uint32_t)((uint64_t)count * t / parts), b1 = begin + (uint32_t)((uint64_t)count * (t + 1) / parts);
                bin_range(c, b0, b1, cb, scale, part[t]);
            }
            bins.reset();
            for (int t = 0; t < parts; t++)
                for (int a = 0; a < 3; a++)
                    for (int k = 0; k < kBins; k++) bins.b[a][k].merge(part[t].b[a][k]);
        } else {
            bin_range(c, begin, end, cb, scale, bins);
        }
        for (int axis = 0; axis < 3; axis++) {
            if (!(scale[axis] > 0.f)) continue;
            const Bin *bb = bins.b[axis];
            float right_area[kBins];
            uint32_t right_cnt[kBins];
            Box acc;
            acc.reset();
            uint32_t n = 0;
            for (int b = kBins - 1; b > 0; b--) {
                if (bb[b].count) acc.grow(bb[b].box);
                n += bb[b].count;
                right_area[b] = acc.valid() ? acc.half_area() : 0.f;
                right_cnt[b] = n;
            }
            acc.reset();
            n = 0;
            for (int b = 0; b < kBins - 1; b++) {
                if (bb[b].count) acc.grow(bb[b].box);
                n += bb[b].count;
                if (n == 0 || right_cnt[b + 1] == 0) continue;
                const float cost = acc.half_area() * (float)n + right_area[b + 1] * (float)right_cnt[b + 1];
                if (cost < best_cost) {
                    best_cost = cost;
                    best_axis = axis;
                    best_bin = b;
                }
            }
        }
    }
    // Quality option (what bvh::SweepSahBuilder does for every node, libs/bvh/include/bvh/sweep_sah_builder.hpp:60-120): below
    // `sweep_below` primitives every split position along every axis is evaluated on the sorted centroids instead of 16 bins.
    uint32_t sweep_mid = 0;
    if (!force_balance && c.sweep_below && count <= c.sweep_below && count <= 256) {
        uint32_t order[256], best_order[256];
        float right_area[256];
        float sweep_best = best_cost;
        bool found = false;
        for (int axis = 0; axis < 3; axis++) {
            for (uint32_t i = 0; i < count; i++) order[i] = c.idx[begin + i];
            std::sort(order, order + count, [&](uint32_t a, uint32_t b) {
                const float ca = c.cen[3 * (size_t)a + axis], cb2 = c.cen[3 * (size_t)b + axis];
                return ca < cb2 || (ca == cb2 && a < b);
            });
            Box acc;
            acc.reset();
            for (uint32_t i = count; i-- > 1;) {
                acc.grow(c.bmin + 3 * (size_t)order[i], c.bmax + 3 * (size_t)order[i]);
                right_area[i] = acc.half_area();
            }
            acc.reset();
            for (uint32_t i = 0; i + 1 < count; i++) {
                acc.grow(c.bmin + 3 * (size_t)order[i], c.bmax + 3 * (size_t)order[i]);
                const float cost = acc.half_area() * (float)(i + 1) + right_area[i + 1] * (float)(count - i - 1);
                if (cost < sweep_best) {
                    sweep_best = cost;
                    sweep_mid = i + 1;
                    found = true;
                    std::memcpy(best_order, order, count * sizeof(uint32_t));
                }
            }
        }
        if (found) {  // strictly better than the best binned split
            best_cost = sweep_best;
            best_axis = 3;  // marks "already partitioned"
            std::memcpy(c.idx + begin, best_order, count * sizeof(uint32_t));
        }
    }
    // SAH termination: leaf cost = N * area, split cost = traversal * area + children
    const float area = box.half_area();
    if ((int)count <= c.max_leaf) {
        const float leaf_cost = (float)count * area;
        if (best_axis < 0 || best_cost + c.trav_cost * area >= leaf_cost) return make_leaf();
    }

    uint32_t mid;
    Box lb, rb, lcb, rcb;
    lb.reset(), rb.reset(), lcb.reset(), rcb.reset();
    bool boxes_known = false;
    if (best_axis == 3) {
        mid = begin + sweep_mid;  // the sweep left idx[begin, end) sorted along its axis
    } else if (best_axis >= 0) {
        const float lo = cb.lo[best_axis], sc = scale[best_axis];
        const int axis = best_axis, bin = best_bin;
        uint32_t n_left = 0;
        for (int b = 0; b < kBins; b++) {
            const Bin &bn = bins.b[axis][b];
            if (!bn.count) continue;
            if (b <= bin) lb.grow(bn.box), lcb.grow(bn.cbox), n_left += bn.count;
            else rb.grow(bn.box), rcb.grow(bn.cbox);
        }
        boxes_known = true;
        if (big) {  // stable partition through the scratch array: slices count, offsets follow from the counts, slices scatter
            const int parts = std::max(1, std::min(omp_get_num_threads() * 4, (int)(count / 32768)));
            std::vector<uint32_t> left_in(parts + 1, 0);
#pragma omp taskloop shared(c, left_in) grainsize(1)
            for (int t = 0; t < parts; t++) {
                const uint32_t b0 = begin + (uint32_t)((uint64_t)count * t / parts), b1 = begin + (uint32_t)((uint64_t)count * (t + 1) / parts);
                uint32_t k = 0;
                for (uint32_t i = b0; i < b1; i++) k += bin_of(c.cen[3 * (size_t)c.idx[i] + axis], lo, sc) <= bin ? 1u : 0u;
                left_in[t + 1] = k;
            }
            for (int t = 0; t < parts; t++) left_in[t + 1] += left_in[t];
#pragma omp taskloop shared(c, left_in) grainsize(1)
            for (int t = 0; t < parts; t++) {
                const uint32_t b0 = begin + (uint32_t)((uint64_t)count * t / parts), b1 = begin + (uint32_t)((uint64_t)count * (t + 1) / parts);
                uint32_t l = begin + left_in[t], r = begin + n_left + (b0 - begin - left_in[t]);
                for (uint32_t i = b0; i < b1; i++) {
                    const uint32_t p = c.idx[i];
                    if (bin_of(c.cen[3 * (size_t)p + axis], lo, sc) <= bin) c.scratch[l++] = p;
                    else c.scratch[r++] = p;
                }
            }
#pragma omp taskloop shared(c) grainsize(1)
            for (int t = 0; t < parts; t++) {
                const uint32_t b0 = begin + (uint32_t)((uint64_t)count * t / parts), b1 = begin + (uint32_t)((uint64_t)count * (t + 1) / parts);
                std::memcpy(c.idx + b0, c.scratch + b0, (size_t)(b1 - b0) * sizeof(uint32_t));
            }
            mid = begin + n_left;
        } else {
            uint32_t *m = std::partition(c.idx + begin, c.idx + end, [&](uint32_t p) { return bin_of(c.cen[3 * (size_t)p + axis], lo, sc) <= bin; });
            mid = (uint32_t)(m - c.idx);
        }
    } else {
        // all centroids coincide, or the depth budget forces a balanced split: median along the widest axis
        int axis = 0;
        for (int k = 1; k < 3; k++)
            if (cb.hi[k] - cb.lo[k] > cb.hi[axis] - cb.lo[axis]) axis = k;
        mid = begin + count / 2;
        std::nth_element(c.idx + begin, c.idx + mid, c.idx + end, [&](uint32_t a, uint32_t b) {
            const float ca = c.cen[3 * (size_t)a + axis], cb2 = c.cen[3 * (size_t)b + axis];
            return ca < cb2 || (ca == cb2 && a < b);
        });
    }
    if (mid == begin || mid == end) {
        mid = begin + count / 2;
        boxes_known = false;
    }
    if (!boxes_known) {
        lb.reset(), rb.reset(), lcb.reset(), rcb.reset();
        for (uint32_t i = begin; i < mid; i++) lb.grow(c.bmin + 3 * (size_t)c.idx[i], c.bmax + 3 * (size_t)c.idx[i]), lcb.grow_pt(c.cen + 3 * (size_t)c.idx[i]);
        for (uint32_t i = mid; i < end; i++) rb.grow(c.bmin + 3 * (size_t)c.idx[i], c.bmax + 3 * (size_t)c.idx[i]), rcb.grow_pt(c.cen + 3 * (size_t)c.idx[i]);
    }

    const uint32_t child = c.next_node.fetch_add(2);
    nd.prim_count = 0;
    nd.first = child;
    if (count > 4096) {
#pragma omp task shared(c) firstprivate(child, begin, mid, lb, lcb, depth)
        build_range(c, child, begin, mid, lb, lcb, depth + 1);
#pragma omp task shared(c) firstprivate(child, mid, end, rb, rcb, depth)
        build_range(c, child + 1, mid, end, rb, rcb, depth + 1);
    } else {
        build_range(c, child, begin, mid, lb, lcb, depth + 1);
        build_range(c, child + 1, mid, end, rb, rcb, depth + 1);
    }
}

}  // namespace

// Build a bvh::Bvh<float>-form hierarchy over the triangles.  Deterministic: the tree depends only
// on the input, and the final node order is a depth-first relayout of it.
void build_bvh(const TriangleVec &tris, HostBvh &out, int max_leaf, float trav_cost, uint32_t sweep_below) {
    const size_t n = tris.size();
    out.nodes.clear();
    out.prim_indices.clear();
    if (n == 0) return;
    const bool timing = std::getenv("VT_TIMING") && std::atoi(std::getenv("VT_TIMING")) != 0;
    double t_phase = omp_get_wtime();
    auto lap = [&](const char *what) {
        if (!timing) return;
        const double t = omp_get_wtime();
        std::fprintf(stderr, "[build_bvh] %-24s %.3f s\n", what, t - t_phase);
        t_phase = t;
    };
    RawVector<float> bmin(3 * n), bmax(3 * n), cen(3 * n);
    Box global, global_c;
    global.reset();
    global_c.reset();
#pragma omp parallel
    {
        Box local, local_c;
        local.reset();
        local_c.reset();
#pragma omp for nowait
        for (int64_t i = 0; i < (int64_t)n; i++) {
            const Triangle &t = tris[i];
            // Triangle::bounding_box / center over p0, p1() = p0 - e1, p2() = p0 + e2
            // (source/objects/Primitives.h:104-118): the vertices the intersection test actually sees.
            for (int k = 0; k < 3; k++) {
                const float a = t.p0[k], b = t.p0[k] - t.e1[k], cc = t.p0[k] + t.e2[k];
                bmin[3 * i + k] = std::min(a, std::min(b, cc));
                bmax[3 * i + k] = std::max(a, std::max(b, cc));
                cen[3 * i + k] = (a + b + cc) * (1.0f / 3.0f);
            }
            local.grow(&bmin[3 * i], &bmax[3 * i]);
            local_c.grow_pt(&cen[3 * i]);
        }
#pragma omp critical
        {
            global.grow(local);
            global_c.grow(local_c);
        }
    }
    lap("primitive boxes");
    RawVector<uint32_t> idx(n), scratch(n >= kParallelNode ? n : 0);
#pragma omp parallel for
    for (int64_t i = 0; i < (int64_t)n; i++) idx[i] = (uint32_t)i;
    RawVector<vt_node> tmp(2 * n + 1);
    BuildCtx c{bmin.data(), bmax.data(), cen.data(), idx.data(), scratch.data(), tmp.data(), {1}, max_leaf, trav_cost, sweep_below};
#pragma omp parallel
#pragma omp single
    build_range(c, 0, 0, (uint32_t)n, global, global_c, 0);

    lap("top-down SAH");
    // depth-first relayout: node ids handed out by the task scheduler are not reproducible, this order is — root at 0, the two
    // children of a node adjacent, then everything below the left child, then everything below the right child.  A child's id
    // is always larger than its parent's (ids are handed out when the parent splits), so the descendant counts come from one
    // backward sweep and every subtree knows its slice of the output: the copy itself runs as tasks.
    const uint32_t total = c.next_node.load();
    out.nodes.resize(total);
    out.prim_indices.resize(n);
    std::vector<uint32_t> desc(total, 0);
    for (uint32_t i = total; i-- > 0;)
        if (tmp[i].prim_count == 0) desc[i] = 2 + desc[tmp[i].first] + desc[tmp[i].first + 1];
    out.nodes[0] = tmp[0];
    struct Relayout {
        const vt_node *tmp;
        const uint32_t *desc;
        vt_node *out;
        void place(uint32_t oi, uint32_t ni, uint32_t base) const {  // inner node oi sits at out[ni]; its subtree fills out[base ...)
            const uint32_t oc = tmp[oi].first;
            out[ni].first = base;
            out[base] = tmp[oc];
            out[base + 1] = tmp[oc + 1];
            const uint32_t left_base = base + 2, right_base = base + 2 + desc[oc];
            const bool big = desc[oi] > 8192;
            if (tmp[oc].prim_count == 0) {
                if (big) {
#pragma omp task firstprivate(oc, base, left_base)
                    place(oc, base, left_base);
                } else {
                    place(oc, base, left_base);
                }
            }
            if (tmp[oc + 1].prim_count == 0) {
                if (big) {
#pragma omp task firstprivate(oc, base, right_base)
                    place(oc + 1, base + 1, right_base);
                } else {
                    place(oc + 1, base + 1, right_base);
                }
            }
        }
    } relayout{tmp.data(), desc.data(), out.nodes.data()};
    if (tmp[0].prim_count == 0) {
#pragma omp parallel
#pragma omp single
        relayout.place(0, 0, 1);
    }
#pragma omp parallel for
    for (size_t i = 0; i < n; i++) out.prim_indices[i] = idx[i];
    lap("depth-first relayout");
}

// ------------------------------------------------------------------------------------ refit
// bvh::HierarchyRefitter::refit (libs/bvh/include/bvh/hierarchy_refitter.hpp:20-31) for moved geometry of unchanged
// topology — what `accel:Rebuild` of moving props needs instead of a full build.  Leaf boxes are
// recomputed from Triangle::bounding_box (source/objects/Primitives.h:107-113) of their primitives, the way the library's
// refit test updates leaves (libs/bvh/test/refit_bvh.cpp:79-89), in parallel; inner boxes are the union of their two
// children, children first.  The reference library walks up from every leaf with per-node arrival flags
// (bottom_up_algorithm.hpp:52-80); min/max are exact, so this level-free reverse pre-order pass gives the same boxes.
bool refit_bvh(const TriangleVec &tris, HostBvh &bvh, std::string &err) {
    const size_t node_count = bvh.nodes.size();
    if (node_count == 0) return true;
    if (bvh.prim_indices.size() != tris.size()) {
        err = "refit: triangle count differs from the hierarchy's";
        return false;
    }
    std::vector<uint32_t> order;
    order.reserve(node_count);
    std::vector<uint32_t> stack{0u};
    while (!stack.empty()) {  // pre-order: parents before children
        const uint32_t i = stack.back();
        stack.pop_back();
        order.push_back(i);
        if (order.size() > node_count) {
            err = "refit: hierarchy is not a tree";
            return false;
        }
        const vt_node &nd = bvh.nodes[i];
        if (nd.prim_count == 0) {
            if (nd.first == 0 || (size_t)nd.first + 1 >= node_count) {
                err = "refit: malformed hierarchy";
                return false;
            }
            stack.push_back(nd.first);
            stack.push_back(nd.first + 1);
        } else if ((size_t)nd.first + nd.prim_count > tris.size()) {
            err = "refit: leaf addresses primitives past the end of prim_indices";
            return false;
        }
    }
    bool bad = false;
#pragma omp parallel for schedule(static) reduction(|| : bad)
    for (int64_t k = 0; k < (int64_t)order.size(); k++) {
        vt_node &nd = bvh.nodes[order[k]];
        if (nd.prim_count == 0) continue;
        Box b;
        b.reset();
        for (uint32_t t = 0; t < nd.prim_count; t++) {
            const uint64_t p = bvh.prim_indices[nd.first + t];
            if (p >= tris.size()) {
                bad = true;
                break;
            }
            const Triangle &tr = tris[p];
            float v[3][3];
            for (int a = 0; a < 3; a++) v[0][a] = tr.p0[a], v[1][a] = tr.p0[a] - tr.e1[a], v[2][a] = tr.p0[a] + tr.e2[a];
            b.grow_pt(v[0]);
            b.grow_pt(v[1]);
            b.grow_pt(v[2]);
        }
        set_bounds(nd, b);
    }
    if (bad) {
        err = "refit: primitive index out of range";
        return false;
    }
    for (size_t k = order.size(); k-- > 0;) {  // children before parents
        vt_node &nd = bvh.nodes[order[k]];
        if (nd.prim_count != 0) continue;
        const vt_node &l = bvh.nodes[nd.first], &r = bvh.nodes[nd.first + 1];
        for (int a = 0; a < 3; a++) {
            nd.bounds[2 * a] = std::min(l.bounds[2 * a], r.bounds[2 * a]);
            nd.bounds[2 * a + 1] = std::max(l.bounds[2 * a + 1], r.bounds[2 * a + 1]);
        }
    }
    return true;
}

// ------------------------------------------------------------------------------------ flatten
// bvh::Bvh<float> form -> VtPair array + leaf-order triangle permutation.
// Pair order: breadth-first for the first `bfs_pairs` pairs (the part the traversal kernel stages
// in shared memory), depth-first (left subtree first) below that.  Left/right inside a pair and
// the primitive order inside a leaf are preserved, so traversal order is unchanged.
bool flatten_bvh(const HostBvh &bvh, uint64_t n_tris, uint32_t bfs_pairs, FlatBvh &out, std::string &err) {
    out.pairs.clear();
    out.leaf_order.clear();
    out.root_leaf_count = 0;
    out.max_depth = 0;
    const size_t node_count = bvh.nodes.size();
    if (node_count == 0 || n_tris == 0) return true;
    if (bvh.prim_indices.size() < n_tris) {
        err = "BVH primitive index array is shorter than the triangle count";
        return false;
    }
    const vt_node &root = bvh.nodes[0];
    out.leaf_order.reserve(n_tris);
    auto emit_leaf = [&](const vt_node &nd, uint32_t &first_out) -> bool {
        if ((uint64_t)nd.first + nd.prim_count > n_tris) {
            err = "BVH leaf addresses primitives past the end of prim_indices";
            return false;
        }
        first_out = (uint32_t)out.leaf_order.size();
        for (uint32_t k = 0; k < nd.prim_count; k++) {
            const uint64_t p = bvh.prim_indices[nd.first + k];
            if (p >= n_tris) {
                err = "BVH primitive index out of range";
                return false;
            }
            out.leaf_order.push_back((uint32_t)p);
        }
        return true;
    };
    if (root.prim_count != 0) {  // the root is a leaf (single_ray_traverser.hpp:72-73)
        uint32_t f;
        if (!emit_leaf(root, f)) return false;
        out.root_leaf_count = root.prim_count;
        return true;
    }
    if ((node_count & 1) == 0) {
        err = "BVH node count must be odd (root + sibling pairs)";
        return false;
    }
    const size_t n_pairs = (node_count - 1) / 2;
    out.pairs.resize(n_pairs);
    struct Item {
        uint32_t old_first;  // node index of the left child of the pair
        uint32_t new_pair;
        uint32_t depth;
    };
    // Phase 1: assign new pair ids — BFS for the top, then DFS.  new_id[old pair] = new pair.
    std::vector<uint32_t> new_id(n_pairs, 0xFFFFFFFFu);
    std::vector<Item> order;  // pairs in new order
    order.reserve(n_pairs);
    auto old_pair = [](uint32_t first) { return (first - 1) / 2; };
    auto check_child = [&](uint32_t first) -> bool {
        if (first == 0 || (first & 1) == 0 || (size_t)first + 1 >= node_count) {
            err = "BVH child index invalid (children must be an adjacent pair at an odd index)";
            return false;
        }
        if (new_id[old_pair(first)] != 0xFFFFFFFFu) {
            err = "BVH is not a tree (a node pair is referenced twice)";
            return false;
        }
        return true;
    };
    std::vector<Item> frontier;
    if (!check_child(root.first)) return false;
    frontier.push_back({root.first, 0, 1});
    new_id[old_pair(root.first)] = 0;
    size_t head = 0;
    uint32_t next = 1;
    // BFS
    while (head < frontier.size() && order.size() < bfs_pairs) {
        Item it = frontier[head++];
        order.push_back(it);
        for (int s = 0; s < 2; s++) {
            const vt_node &ch = bvh.nodes[it.old_first + s];
            if (ch.prim_count == 0) {
                if (!check_child(ch.first)) return false;
                new_id[old_pair(ch.first)] = 0xFFFFFFFEu;  // reserved, id assigned when visited
                frontier.push_back({ch.first, 0, it.depth + 1});
            }
        }
    }
    // ids for BFS part are their positions
    for (size_t i = 0; i < order.size(); i++) {
        order[i].new_pair = (uint32_t)i;
        new_id[old_pair(order[i].old_first)] = (uint32_t)i;
    }
    next = (uint32_t)order.size();
    // DFS below each remaining frontier entry, in frontier order
    std::vector<Item> stack;
    for (size_t f = head; f < frontier.size(); f++) {
        stack.push_back(frontier[f]);
        while (!stack.empty()) {
            Item it = stack.back();
            stack.pop_back();
            it.new_pair = next++;
            new_id[old_pair(it.old_first)] = it.new_pair;
            order.push_back(it);
            const vt_node &l = bvh.nodes[it.old_first], &r = bvh.nodes[it.old_first + 1];
            if (r.prim_count == 0) {
                if (!check_child(r.first)) return false;
                new_id[old_pair(r.first)] = 0xFFFFFFFEu;
                stack.push_back({r.first, 0, it.depth + 1});
            }
            if (l.prim_count == 0) {
                if (!check_child(l.first)) return false;
                new_id[old_pair(l.first)] = 0xFFFFFFFEu;
                stack.push_back({l.first, 0, it.depth + 1});
            }
        }
    }
    if (order.size() != n_pairs) {
        err = "BVH has unreachable nodes";
        return false;
    }
    // Phase 2: emit pairs and leaf runs in the new order
    for (const Item &it : order) {
        VtPair &p = out.pairs[it.new_pair];
        out.max_depth = std::max(out.max_depth, it.depth);
        for (int s = 0; s < 2; s++) {
            const vt_node &ch = bvh.nodes[it.old_first + s];
            VtChild &o = s ? p.r : p.l;
            std::memcpy(o.bounds, ch.bounds, sizeof(o.bounds));
            o.count = ch.prim_count;
            if (ch.prim_count == 0) {
                o.first = new_id[old_pair(ch.first)];
            } else if (!emit_leaf(ch, o.first))
                return false;
        }
    }
    if (out.leaf_order.size() != n_tris) {
        err = "BVH leaves do not cover every primitive exactly once";
        return false;
    }
    if (out.max_depth > VT_STACK_SIZE) {
        err = "BVH deeper than the 64-entry traversal stack (single_ray_traverser.hpp:14)";
        return false;
    }
    return true;
}

// ------------------------------------------------------------------------------------ compact
// VtPair[] (depth-first order, bfs_pairs = 0) -> VtCPair[]: see vt_device.h.  Returns false (with a
// reason) when the tree cannot be represented: a leaf of more than 15 triangles, a non-finite
// bound, or a child reference that is not where the depth-first order puts it.
namespace {

// Smallest power-of-two grid 2^E on which [lo, hi] spans <= 255 cells from k = floor(lo / 2^E), with
// |k| small enough that (k - 2^23) * 2^E and (k + q) * 2^E are exact floats.
bool choose_grid(double lo, double hi, int &E, int64_t &k) {
    const double kmax = 8388608.0 - 512.0;
    const double mag = std::max(std::fabs(lo), std::fabs(hi));
    int e = -149;
    if (hi > lo) e = std::max(e, (int)std::ceil(std::log2((hi - lo) / 255.0)) - 1);
    if (mag > 0) e = std::max(e, (int)std::floor(std::log2(mag)) - 23);
    // |E| <= 60: 2^E times a ray's inverse direction (|inv| in [2^-60, 2^24] on the one-fma path of slab_quad) must be
    // a NORMAL float so that the product is exact; it also keeps 2^E normal and (2^23 + q) * 2^E finite.  A coarser
    // grid than the box needs is still conservative (a flat box at 0 gets k = q = 0 on any grid).
    e = std::max(e, -60);
    for (; e <= 60; e++) {
        const double s = std::ldexp(1.0, e);
        const double kl = std::floor(lo / s), kh = std::ceil(hi / s);
        if (kh - kl <= 255.0 && std::fabs(kl) <= kmax && std::fabs(kh) <= kmax) {
            E = e;
            k = (int64_t)kl;
            return true;
        }
    }
    return false;
}

}  // namespace

bool compact_pairs(const std::vector<VtPair> &pairs, std::vector<VtCPair> &out, std::string &err) {
    out.resize(pairs.size());
    std::atomic<int> fail{0};
#pragma omp parallel for
    for (int64_t i = 0; i < (int64_t)pairs.size(); i++) {
        const VtPair &p = pairs[i];
        VtCPair &c = out[i];
        if (p.l.count > 15 || p.r.count > 15) {
            fail = 1;
            continue;
        }
        bool ok = true;
        for (int a = 0; a < 3 && ok; a++) {
            const float llo = p.l.bounds[2 * a], lhi = p.l.bounds[2 * a + 1];
            const float rlo = p.r.bounds[2 * a], rhi = p.r.bounds[2 * a + 1];
            if (!(std::isfinite(llo) && std::isfinite(lhi) && std::isfinite(rlo) && std::isfinite(rhi)) || lhi < llo || rhi < rlo) {
                ok = false;
                break;
            }
            int E;
            int64_t k;
            if (!choose_grid(std::min(llo, rlo), std::max(lhi, rhi), E, k)) {
                ok = false;
                break;
            }
            const double s = std::ldexp(1.0, E);
            c.origin_adj[a] = (float)((double)(k - 8388608) * s);  // exact: |k - 2^23| < 2^24
            c.exp[a] = (uint8_t)(E + 127);
            c.q[a][0] = (uint8_t)((int64_t)std::floor(llo / s) - k);
            c.q[a][1] = (uint8_t)((int64_t)std::ceil(lhi / s) - k);
            c.q[a][2] = (uint8_t)((int64_t)std::floor(rlo / s) - k);
            c.q[a][3] = (uint8_t)((int64_t)std::ceil(rhi / s) - k);
        }
        if (!ok) {
            fail = 2;
            continue;
        }
        c.counts = (uint8_t)(p.l.count | (p.r.count << 4));
        const bool li = p.l.count == 0, ri = p.r.count == 0;
        const uint32_t next = (uint32_t)i + 1;
        if (li) {
            if (p.l.first != next) fail = 3;
            c.ref = p.r.first;
        } else {
            c.ref = p.l.first;
            if (ri ? p.r.first != next : p.r.first != p.l.first + p.l.count) fail = 3;
        }
    }
    switch (fail.load()) {
        case 1: err = "compact layout: a leaf holds more than 15 triangles"; return false;
        case 2: err = "compact layout: non-finite or inverted bounds, or coordinates out of float grid range"; return false;
        case 3: err = "compact layout: pairs are not in depth-first order"; return false;
    }
    return true;
}

// --------------------------------------------------------------------------------------- quads
namespace {

// Two passes over the collapsed tree, both as OpenMP tasks: measure() gives every quad-to-be the number of quads and triangles
// below it (and the stack its subtree needs), so emit() knows the slice of the quad array and of the leaf order each subtree owns
// and the subtrees are written independently.  The order is the sequential one: a quad, its leaf children's triangles, then its
// inner children depth-first.
struct QuadBuilder {
    const HostBvh &bvh;
    uint64_t n_tris;
    QuadBvh &out;
    std::string &err;
    const CollapsePlan &plan;
    std::vector<uint32_t> quads_below, tris_below;  // indexed by the binary node that becomes the quad (itself included)
    std::vector<uint8_t> need_below;
    std::atomic<bool> ok{true};
    std::mutex err_mutex;

    QuadBuilder(const HostBvh &b, uint64_t n, QuadBvh &o, std::string &e, const CollapsePlan &p) : bvh(b), n_tris(n), out(o), err(e), plan(p) {}

    bool fail(const char *why) {
        std::lock_guard<std::mutex> lock(err_mutex);
        if (ok.load()) err = why;
        ok = false;
        return false;
    }

    void measure(uint32_t ni, uint32_t depth) {
        const size_t node_count = bvh.nodes.size();
        quads_below[ni] = 1, tris_below[ni] = 0, need_below[ni] = 0;
        if (!ok.load()) return;
        if (bvh.nodes[ni].first == 0 || (size_t)bvh.nodes[ni].first + 1 >= node_count || depth > 64) {
            fail("quad layout: malformed hierarchy");
            return;
        }
        uint32_t kids[4];
        const int nk = plan.children(bvh, ni, kids);
        const bool spawn = depth < 10;
        for (int i = 0; i < nk; i++) {
            if (bvh.nodes[kids[i]].prim_count != 0) continue;
            if ((size_t)kids[i] >= node_count || quads_below[kids[i]] != 0) {  // not a tree: a node is reachable twice
                fail("BVH is not a tree (a node pair is referenced twice)");
                return;
            }
            const uint32_t k = kids[i];
            if (spawn) {
#pragma omp task firstprivate(k, depth)
                measure(k, depth + 1);
            } else {
                measure(k, depth + 1);
            }
        }
        if (spawn) {
#pragma omp taskwait
        }
        uint32_t q = 1, t = 0, below = 0;
        for (int i = 0; i < nk; i++) {
            const vt_node &c = bvh.nodes[kids[i]];
            if (c.prim_count != 0) {
                t += c.prim_count;
            } else {
                q += quads_below[kids[i]], t += tris_below[kids[i]];
                below = std::max<uint32_t>(below, need_below[kids[i]]);
            }
        }
        // while a child is being walked, up to nk - 1 siblings are pending — and with VT_EMPTY_SENTINEL the kernel does not test
        // "slot in use": an empty slot's inverted box can pass the rounded slab test (vt_traverse.cu, slab_quad) and its sentinel
        // leaf is then pushed like any other child, so every level is budgeted with all 3 non-nearest slots
        const uint32_t need = (uint32_t)(VT_EMPTY_SENTINEL ? 3 : nk - 1) + below;
        quads_below[ni] = q, tris_below[ni] = t, need_below[ni] = (uint8_t)std::min<uint32_t>(need, 255);
    }

    // Write the quad that replaces binary inner node `ni` at out.quads[qi]; its subtree's triangles start at leaf slot `tri_base`.
    void emit(uint32_t ni, uint32_t qi, uint32_t tri_base, uint32_t depth) {
        if (!ok.load()) return;
        uint32_t kids[4];
        const int nk = plan.children(bvh, ni, kids);
        VtQuad q;
        std::memset(&q, 0, sizeof(q));
        // shared grid over the union of the children
        for (int a = 0; a < 3; a++) {
            double lo = 1e300, hi = -1e300;
            for (int i = 0; i < nk; i++) {
                const vt_node &c = bvh.nodes[kids[i]];
                if (!(std::isfinite(c.bounds[2 * a]) && std::isfinite(c.bounds[2 * a + 1])) || c.bounds[2 * a + 1] < c.bounds[2 * a]) {
                    fail("quad layout: non-finite or inverted bounds");
                    return;
                }
                lo = std::min(lo, (double)c.bounds[2 * a]);
                hi = std::max(hi, (double)c.bounds[2 * a + 1]);
            }
            int E = 0;
            int64_t k = 0;
            if (!choose_grid(lo, hi, E, k)) {
                fail("quad layout: coordinates out of float grid range");
                return;
            }
            const double s = std::ldexp(1.0, E);
            q.origin_adj[a] = (float)((double)(k - VT_QUAD_OFFSET) * s);
            q.scale[a] = (float)s;
            for (int i = 0; i < 4; i++) {
                if (i < nk) {
                    const vt_node &c = bvh.nodes[kids[i]];
                    q.q[a][0][i] = (uint8_t)((int64_t)std::floor(c.bounds[2 * a] / s) - k);
                    q.q[a][1][i] = (uint8_t)((int64_t)std::ceil(c.bounds[2 * a + 1] / s) - k);
                } else {
                    q.q[a][0][i] = 255;  // empty slot: inverted box (and valid bit clear)
                    q.q[a][1][i] = 0;
                }
            }
        }
        // leaves first (their triangles sit next to each other), then the inner children, depth-first
        for (int i = 0; i < 4; i++) q.ref[i] = 0xFFFFFFFFu;
        uint32_t slot = tri_base;
        for (int i = 0; i < nk; i++) {
            const vt_node &c = bvh.nodes[kids[i]];
            if (c.prim_count == 0) continue;
            if (c.prim_count > 15) {
                fail("quad layout: a leaf holds more than 15 triangles");
                return;
            }
            if ((uint64_t)c.first + c.prim_count > n_tris) {
                fail("BVH leaf addresses primitives past the end of prim_indices");
                return;
            }
            q.ref[i] = (c.prim_count << 28) | slot;
            for (uint32_t t = 0; t < c.prim_count; t++) {
                const uint64_t p = bvh.prim_indices[c.first + t];
                if (p >= n_tris) {
                    fail("BVH primitive index out of range");
                    return;
                }
                out.leaf_order[slot++] = (uint32_t)p;
            }
        }
        uint32_t next_q = qi + 1;
        const bool spawn = quads_below[ni] > 4096;
        for (int i = 0; i < nk; i++) {
            if (bvh.nodes[kids[i]].prim_count != 0) continue;
            const uint32_t k = kids[i], cq = next_q, ct = slot;
            q.ref[i] = cq;
            if (spawn) {
#pragma omp task firstprivate(k, cq, ct, depth)
                emit(k, cq, ct, depth + 1);
            } else {
                emit(k, cq, ct, depth + 1);
            }
            next_q += quads_below[k];
            slot += tris_below[k];
        }
        out.quads[qi] = q;
    }
};

}  // namespace

bool build_quads(const HostBvh &bvh, uint64_t n_tris, QuadBvh &out, std::string &err) {
    out.quads.clear();
    out.leaf_order.clear();
    out.root_leaf_count = 0;
    out.max_stack = 0;
    if (bvh.nodes.empty() || n_tris == 0) return true;
    if (bvh.prim_indices.size() < n_tris) {
        err = "BVH primitive index array is shorter than the triangle count";
        return false;
    }
    if (n_tris >= (1ull << 28) - 16 || bvh.nodes.size() >= (1ull << 28) - 16) {
        err = "quad layout: more than 2^28 triangles or nodes";
        return false;
    }
    const vt_node &root = bvh.nodes[0];
    out.leaf_order.reserve(n_tris);
    if (root.prim_count != 0) {
        if (root.prim_count > 15 || (uint64_t)root.first + root.prim_count > n_tris) {
            err = "quad layout: the root leaf holds more than 15 triangles";
            return false;
        }
        for (uint32_t t = 0; t < root.prim_count; t++) out.leaf_order.push_back((uint32_t)bvh.prim_indices[root.first + t]);
        out.root_leaf_count = root.prim_count;
        return out.leaf_order.size() == n_tris;
    }
    CollapsePlan plan;
    if (!plan_collapse(bvh, 4, plan, err)) return false;
    out.sibling_overlap = plan.sibling_overlap, out.greedy_collapse = plan.greedy;
    if (std::getenv("VT_TIMING") && std::atoi(std::getenv("VT_TIMING")) != 0)
        std::fprintf(stderr, "[build_quads] sibling overlap %.3f -> %s collapse\n", plan.sibling_overlap, plan.greedy ? "greedy (largest child)" : "SAH-optimal");
    QuadBuilder b(bvh, n_tris, out, err, plan);
    b.quads_below.assign(bvh.nodes.size(), 0);
    b.tris_below.assign(bvh.nodes.size(), 0);
    b.need_below.assign(bvh.nodes.size(), 0);
#pragma omp parallel
#pragma omp single
    b.measure(0, 0);
    if (!b.ok) return false;
    if (b.tris_below[0] != n_tris) {
        err = "BVH leaves do not cover every primitive exactly once";
        return false;
    }
    out.quads.resize(b.quads_below[0]);
    out.leaf_order.resize(n_tris);
#pragma omp parallel
#pragma omp single
    b.emit(0, 0, 0, 0);
    if (!b.ok) return false;
    const uint32_t need = b.need_below[0];
    out.max_stack = need;
    std::vector<uint8_t> seen(n_tris, 0);
    for (uint32_t p : out.leaf_order) {
        if (seen[p]) {
            err = "BVH leaves do not cover every primitive exactly once";
            return false;
        }
        seen[p] = 1;
    }
    if (need > VT_STACK_SIZE) {
        err = "quad layout: worst-case traversal stack deeper than 64 entries";
        return false;
    }
    return true;
}

// Experiment knob (VT_SMEM_QUADS, with a library built -DVT_SMEM_QUADS_BUILD=1): move the top of the quad hierarchy — the first
// `top` quads in BREADTH-first order from the root — to the front of the array, so that the traversal kernel can stage them in shared
// memory as one block (BASELINE.json north_star: "the top BVH levels staged in shared memory"); every other quad keeps its depth-first
// place behind them.  Only inner references change.  Returns the number of quads moved to the front.
uint32_t quads_top_first(QuadBvh &qb, uint32_t top) {
    const uint32_t n = (uint32_t)qb.quads.size();
    top = std::min(top, n);
    if (top <= 1) return top;
    std::vector<uint32_t> order;  // new index -> old index
    order.reserve(n);
    std::vector<uint8_t> in_top(n, 0);
    order.push_back(0);
    in_top[0] = 1;
    for (size_t head = 0; head < order.size() && order.size() < top; head++) {
        const VtQuad &q = qb.quads[order[head]];
        for (int i = 0; i < 4 && order.size() < top; i++) {
            const uint32_t ref = q.ref[i];
            if (ref == 0xFFFFFFFFu || (ref >> 28) != 0) continue;  // empty slot or leaf
            if (!in_top[ref]) in_top[ref] = 1, order.push_back(ref);
        }
    }
    const uint32_t moved = (uint32_t)order.size();
    for (uint32_t i = 0; i < n; i++)
        if (!in_top[i]) order.push_back(i);
    std::vector<uint32_t> new_of(n);
    for (uint32_t i = 0; i < n; i++) new_of[order[i]] = i;
    RawVector<VtQuad> out(n);
    for (uint32_t i = 0; i < n; i++) {
        VtQuad q = qb.quads[order[i]];
        for (int k = 0; k < 4; k++)
            if (q.ref[k] != 0xFFFFFFFFu && (q.ref[k] >> 28) == 0) q.ref[k] = new_of[q.ref[k]];
        out[i] = q;
    }
    qb.quads.swap(out);
    return moved;
}

}  // namespace vt
