// ASan / UBSan driver for the host-only hierarchy code: the product builder, the PLOC + LeafCollapser restatement, flatten, the
// compact and quad layouts, the host refit and the reinsertion optimiser — first on hostile GEOMETRY (degenerate, duplicated, huge,
// non-finite vertices), then on hostile TREES: the arrays a caller may hand to vt_accel_populate_with_bvh / vt_flatten_bvh /
// vt_build_quads / vt_refit_bvh / vt_optimize_bvh (collapse_leaves only ever sees the PLOC builder's own output) with child references, primitive counts, bounds and primitive indices mutated.
// Every call must return (true / false / exception); nothing may read or write out of bounds.
// usage: hierarchy_driver <seed> <cases>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <random>
#include <string>
#include <vector>
#include "vt_host.h"

using namespace vt;

static void make_tris(std::mt19937 &rng, TriangleVec &tris, int kind) {
    std::uniform_real_distribution<float> u(-20.f, 20.f);
    const size_t n = 1 + rng() % 400;
    tris.resize(n);
    for (size_t i = 0; i < n; i++) {
        Triangle &t = tris[i];
        std::memset(&t, 0, sizeof(t));
        float p[3][3];
        for (auto &v : p)
            for (float &c : v) c = u(rng);
        if (kind == 1 && i % 3 == 0) std::memcpy(p[1], p[0], sizeof p[0]);               // zero-area
        if (kind == 2) for (auto &v : p) for (float &c : v) c = std::round(c);           // grid-aligned
        if (kind == 3 && i >= n / 2) { tris[i] = tris[i - n / 2]; continue; }            // duplicated
        if (kind == 4) for (auto &v : p) v[2] = 0.f;                                     // one plane: a zero-extent axis
        if (kind == 5) for (auto &v : p) for (float &c : v) c *= 1e30f;                  // huge: areas overflow
        if (kind == 6 && i % 7 == 0) p[rng() % 3][rng() % 3] = (i % 2) ? std::numeric_limits<float>::quiet_NaN() : std::numeric_limits<float>::infinity();
        for (int k = 0; k < 3; k++) t.p0[k] = p[0][k], t.e1[k] = p[0][k] - p[1][k], t.e2[k] = p[2][k] - p[0][k];
    }
}

static bool pairs_in_range(const HostBvh &b) {  // what vt_optimize_bvh checks before it runs the pass
    const size_t n = b.nodes.size();
    if (n == 0 || n % 2 == 0) return false;
    for (const vt_node &nd : b.nodes)
        if (nd.prim_count == 0 && (nd.first == 0 || nd.first % 2 == 0 || (uint64_t)nd.first + 1 >= n)) return false;
    return true;
}

static void exercise(const TriangleVec &tris, HostBvh bvh, std::mt19937 &rng, unsigned long long *stat) {
    std::string err;
    const uint64_t n = tris.size();
    int step = 0;
    auto guard = [&](auto &&f) {
        if (std::getenv("VT_FUZZ_TRACE")) std::fprintf(stderr, "  call %d\n", step);  // which call of this tree hangs or aborts
        step++;
        try {
            f() ? stat[0]++ : stat[1]++;
        } catch (const std::exception &) {
            stat[2]++;
        }
    };
    FlatBvh flat;
    guard([&] {
        if (!flatten_bvh(bvh, n, rng() % 2 ? 0 : 7, flat, err)) return false;
        std::vector<VtCPair> cp;
        return compact_pairs(flat.pairs, cp, err);
    });
    guard([&] {
        QuadBvh q;
        if (!build_quads(bvh, n, q, err)) return false;
        quads_top_first(q, rng() % 64);
        return true;
    });
    guard([&] { HostBvh c = bvh; return bvh.prim_indices.size() == n && refit_bvh(tris, c, err); });
    guard([&] { HostBvh c = bvh; return pairs_in_range(c) && reinsert_optimize(c, 1 + rng() % 3, 0.3f); });
    guard([&] { CollapsePlan plan; return plan_collapse(bvh, 4, plan, err); });
}

int main(int argc, char **argv) {
    const unsigned seed = argc > 1 ? (unsigned)std::atoi(argv[1]) : 1u;
    const int cases = argc > 2 ? std::atoi(argv[2]) : 200;
    std::mt19937 rng(seed);
    unsigned long long stat[3] = {0, 0, 0};
    for (int c = 0; c < cases; c++) {
        TriangleVec tris;
        make_tris(rng, tris, c % 7);
        HostBvh bvh;
        if (std::getenv("VT_FUZZ_TRACE")) std::fprintf(stderr, "building scene %d (kind %d, %zu triangles)\n", c, c % 7, tris.size());
        try {
            if (c % 3 == 2) {
                build_bvh_ploc(tris, bvh);
                collapse_leaves(bvh);
            } else {
                build_bvh(tris, bvh, 1 + (int)(rng() % 8), 1.0f, rng() % 2 ? 0u : 64u);
            }
        } catch (const std::exception &) {
            stat[2]++;
            continue;
        }
        if (std::getenv("VT_FUZZ_TRACE")) std::fprintf(stderr, "scene %d (kind %d, %zu triangles, %zu nodes)\n", c, c % 7, tris.size(), bvh.nodes.size());
        exercise(tris, bvh, rng, stat);  // the tree as built
        for (int m = 0; m < 12 && !bvh.nodes.empty(); m++) {  // hostile trees
            HostBvh bad = bvh;
            const int edits = 1 + rng() % 3;
            for (int e = 0; e < edits; e++) {
                vt_node &nd = bad.nodes[rng() % bad.nodes.size()];
                switch (rng() % 7) {
                case 0: nd.first = (uint32_t[]){0u, 1u, 2u, (uint32_t)bad.nodes.size(), (uint32_t)bad.nodes.size() - 1, 0x7FFFFFFFu, 0xFFFFFFFFu, (uint32_t)(rng() % (bad.nodes.size() + 2))}[rng() % 8]; break;
                case 1: nd.prim_count = (uint32_t[]){0u, 1u, 15u, 16u, 1000u, 0xFFFFFFFFu, (uint32_t)tris.size()}[rng() % 7]; break;
                case 2: nd.bounds[rng() % 6] = (float[]){std::numeric_limits<float>::quiet_NaN(), std::numeric_limits<float>::infinity(), -std::numeric_limits<float>::infinity(), 1e38f, -1e38f, 0.f}[rng() % 6]; break;
                case 3: std::swap(nd.bounds[0], nd.bounds[1]); break;  // inverted box
                case 4: if (!bad.prim_indices.empty()) bad.prim_indices[rng() % bad.prim_indices.size()] = (uint64_t[]){0ull, tris.size(), ~0ull, 1ull << 40}[rng() % 4]; break;
                case 5: nd = bad.nodes[rng() % bad.nodes.size()]; break;  // a second parent for some pair
                case 6: if (bad.nodes.size() > 2) bad.nodes.resize(bad.nodes.size() - 1 - rng() % 2); break;  // truncated array
                }
            }
            if (rng() % 8 == 0 && !bad.prim_indices.empty()) bad.prim_indices.resize(bad.prim_indices.size() - 1);
            if (std::getenv("VT_FUZZ_TRACE")) std::fprintf(stderr, " hostile tree %d\n", m);
            exercise(tris, bad, rng, stat);
        }
    }
    std::printf("hierarchy: seed %u, %d scenes x 13 trees: %llu ok, %llu refused, %llu exceptions\n", seed, cases, stat[0], stat[1], stat[2]);
    return 0;
}
