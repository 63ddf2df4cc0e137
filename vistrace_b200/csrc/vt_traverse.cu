// vt_traverse.cu — K1: closest-hit / any-hit BVH traversal for sm_100a.
//
// Replaces, per ray, bvh::SingleRayTraverser::intersect
// (libs/bvh/include/bvh/single_ray_traverser.hpp:65-126) + FastNodeIntersector
// (node_intersectors.hpp:15-47,82-103) + ClosestPrimitiveIntersector
// (primitive_intersectors.hpp:48-53) + TriangleBackfaceCull::intersect incl. the alpha test
// (source/objects/Primitives.h:168-215).
//
// Execution model: persistent warps.  The grid is sized to the machine (SMs x resident CTAs),
// every warp pulls rays from a global counter and REFILLS idle lanes whenever too few of its 32
// lanes are still busy (warp-level compaction by replacement: finished lanes never ride along
// for the longest ray of their batch).  Each lane runs the reference's loop verbatim — both
// children of a pair are slab-tested BEFORE any leaf shrinks tmax, left leaf then right leaf,
// near child first with ties going left, far child pushed — so the visit order, the tmax-shrink
// order and therefore exact-tie winners are those of the reference.  The warp schedules that
// per-lane automaton in ROUNDS (all-node or all-triangle, see k_traverse) so the long
// triangle test is never executed for one or two lanes at a time: the first version of this
// kernel ran it inline and spent half of its issued instructions at ~1.5 live lanes
// (profiles/r1_k_traverse_v1_bounce5m.md).
//
// Memory: one traversal step = one 64-byte VtPair = four LDG.128 from a single 128-byte line;
// the first n_smem_pairs pairs (the top of the tree in breadth-first order) are staged in shared
// memory per CTA; a leaf is a contiguous run of 64-byte VtTriRec (four LDG.128 each).  The
// 64-entry traversal stack lives in local memory (lane-interleaved, L1-resident).
#include "vt_kernels.h"
#include "vt_math.cuh"

namespace {

struct RayState {
    V3 o, d;
    float tmin, tmax;
    V3 inv, so;      // safe_inverse(d), -o * inv    (node_intersectors.hpp:89-94)
    float u, v;      // best hit; its t is tmax (single_ray_traverser.hpp:59)
    uint32_t prim;   // original triangle index or VT_MISS
};

// 256-bit read-only load (sm_100+: LDG.E.256): one 32-byte sector per lane per instruction.  A pair
// or a triangle record (64 B, 64-byte aligned) is two of these instead of four LDG.128 — half the
// per-lane requests through the L1 data stage, which is what limits this kernel once its
// instruction count is down (profiles/r1_k_traverse_v3_bounce5m.md: l1tex throughput 84 %).
VT_DEV void ldg256(const void *p, float4 &lo, float4 &hi) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w)
                 : "l"(p));
}

VT_DEV float safe_inverse(float d) {
    // libs/bvh/include/bvh/vector.hpp:69-74
    return 1.0f / (fabsf(d) < FLT_EPSILON ? copysignf(FLT_EPSILON, d) : d);
}

// TriangleBackfaceCull::intersect — source/objects/Primitives.h:168-215.  Updates the ray's best
// hit and tmax when the candidate is accepted (`t <= tmax`: a later equal-t candidate replaces).
template <bool ALPHA>
VT_DEV bool intersect_triangle(const VtSceneView &S, uint32_t slot, RayState &r) {
    float4 q0, q1, q2, q3;
    ldg256(S.tris + slot, q0, q1);
    ldg256(reinterpret_cast<const char *>(S.tris + slot) + 32, q2, q3);
    const V3 p0 = mk3(q0.x, q0.y, q0.z), e1 = mk3(q0.w, q1.x, q1.y), e2 = mk3(q1.z, q1.w, q2.x);
    const V3 n = mk3(q2.y, q2.z, q2.w);  // cross(e1, e2) as stored by the Triangle ctor (Primitives.h:93)
    const uint32_t matflags = __float_as_uint(q3.x);
    const float nDotDir = bvh_dot(n, r.d);
    if ((matflags & VT_TRI_FLAG_CULL) && nDotDir > 0.f) return false;  // :173-174
    const V3 c = p0 - r.o;
    const V3 rr = bvh_cross(r.d, c);
    const float inv_det = 1.0f / nDotDir;
    const float u = bvh_dot(rr, e2) * inv_det;
    const float v = bvh_dot(rr, e1) * inv_det;
    const float w = 1.0f - u - v;
    if (u >= 0.f && v >= 0.f && w >= 0.f) {  // NaN-rejecting compares, tolerance 0 (:184-187)
        const float t = bvh_dot(n, c) * inv_det;
        if (t >= r.tmin && t <= r.tmax) {
            if (ALPHA && (matflags & VT_TRI_FLAG_ALPHATEST)) {  // :195-208
                const VtDevMaterial &m = S.mats[matflags >> 2];
                const float *uv = S.tri_uv + (size_t)slot * 6;
                const float w2 = 1.f - u - v;
                V2 texUV{(w2 * __ldg(uv + 0) + u * __ldg(uv + 2)) + v * __ldg(uv + 4),
                         (w2 * __ldg(uv + 1) + u * __ldg(uv + 3)) + v * __ldg(uv + 5)};
                texUV = transform_texcoord(texUV, m.base_tex_mat, m.tex_scale);
                const uint32_t ti = m.base_texture >= 0 ? (uint32_t)m.base_texture : S.fallback_tex;
                const float alpha = sample_alpha_mip0(S.texs[ti], S.texels, texUV.x, texUV.y);
                if (alpha < m.alphatest_reference) return false;
            }
            r.u = u;
            r.v = v;
            r.prim = __float_as_uint(q3.y);
            r.tmax = t;  // single_ray_traverser.hpp:59; the best t IS the shrunk tmax
            return true;
        }
    }
    return false;
}

// NodeIntersector::intersect for both children of one pair.  robust_max/min chains
// (utilities.hpp:61-71) are evaluated with fmaxf/fminf in the same nesting order: identical
// values whenever the innermost operand (tmin / tmax) is not NaN, which the argument rules
// guarantee; they differ at most in the sign of a zero, which no comparison observes.
VT_DEV void slab_pair(const float4 &a, const float4 &b, const float4 &c, const float4 &d, const RayState &r, float &le,
                      float &lx, float &re, float &rx) {
    // octant[i] = signbit(d[i]) (node_intersectors.hpp:20-26) == signbit(inv[i]): safe_inverse keeps the sign
    // of d, including -0.0 -> -1/eps, and is never zero.
    const bool ox = signbit(r.inv.x), oy = signbit(r.inv.y), oz = signbit(r.inv.z);
    // left child: a = {minx,maxx,miny,maxy}, b = {minz,maxz,count,first}
    float e0 = fmaf(ox ? a.y : a.x, r.inv.x, r.so.x);
    float e1 = fmaf(oy ? a.w : a.z, r.inv.y, r.so.y);
    float e2 = fmaf(oz ? b.y : b.x, r.inv.z, r.so.z);
    float x0 = fmaf(ox ? a.x : a.y, r.inv.x, r.so.x);
    float x1 = fmaf(oy ? a.z : a.w, r.inv.y, r.so.y);
    float x2 = fmaf(oz ? b.x : b.y, r.inv.z, r.so.z);
    le = fmaxf(e0, fmaxf(e1, fmaxf(e2, r.tmin)));
    lx = fminf(x0, fminf(x1, fminf(x2, r.tmax)));
    e0 = fmaf(ox ? c.y : c.x, r.inv.x, r.so.x);
    e1 = fmaf(oy ? c.w : c.z, r.inv.y, r.so.y);
    e2 = fmaf(oz ? d.y : d.x, r.inv.z, r.so.z);
    x0 = fmaf(ox ? c.x : c.y, r.inv.x, r.so.x);
    x1 = fmaf(oy ? c.z : c.w, r.inv.y, r.so.y);
    x2 = fmaf(oz ? d.x : d.y, r.inv.z, r.so.z);
    re = fmaxf(e0, fmaxf(e1, fmaxf(e2, r.tmin)));
    rx = fminf(x0, fminf(x1, fminf(x2, r.tmax)));
}

VT_DEV void init_ray(const vt_ray &in, RayState &r) {
    r.o = mk3(in.ox, in.oy, in.oz);
    r.d = mk3(in.dx, in.dy, in.dz);
    r.tmin = in.tmin;
    r.tmax = in.tmax;
    r.inv = mk3(safe_inverse(r.d.x), safe_inverse(r.d.y), safe_inverse(r.d.z));
    r.so = mk3(-r.o.x * r.inv.x, -r.o.y * r.inv.y, -r.o.z * r.inv.z);
    r.u = r.v = 0.f;
    r.prim = VT_MISS;
}

// Result record: t = tmax of the best hit, or zeros + VT_MISS.
VT_DEV void write_hit(vt_hit *hits, unsigned long long idx, const RayState &r) {
    const bool hit = r.prim != VT_MISS;
    reinterpret_cast<float4 *>(hits)[idx] = make_float4(hit ? r.tmax : 0.f, r.u, r.v, __uint_as_float(r.prim));
}

// Per-lane ray automaton.  A ray alternates between two kinds of work:
//   * a NODE step: slab-test the two children of pair `cur` against the tmax of step entry, queue the
//     leaf children that were hit (left first, then right) and decide the next pair right away —
//     descend / push far / pop do not depend on what the leaves will return;
//   * TRIANGLE tests: drain the queued leaf triangles one by one, in order, shrinking tmax.
// A ray never takes its next node step before its queue is empty, so every ray sees exactly the
// reference's sequence of box tests and triangle tests (single_ray_traverser.hpp:82-123).  What the
// warp is free to choose is WHICH kind of work it executes next: it runs a triangle round when
// enough lanes have a candidate queued (or nobody can walk), a node round otherwise.  Lanes that
// cannot take part in the chosen round wait; this trades a little idling for never running the
// ~100-instruction triangle test with one or two live lanes.
template <bool ANY_HIT, bool ALPHA, bool SMEM>
__global__ void __launch_bounds__(VT_TRAVERSE_BLOCK, VT_TRAVERSE_MIN_BLOCKS)
k_traverse(const VtSceneView S, const vt_ray *__restrict__ rays, vt_hit *__restrict__ hits, unsigned long long n,
           unsigned long long *__restrict__ counters, int persistent, int refill_threshold, int tri_threshold) {
    extern __shared__ float4 s_pairs[];
    if (SMEM) {  // optional: stage the top of the tree (the first n_smem_pairs pairs, breadth-first) per CTA
        for (uint32_t i = threadIdx.x; i < S.n_smem_pairs * 4u; i += blockDim.x)
            s_pairs[i] = __ldg(reinterpret_cast<const float4 *>(S.pairs) + i);
        __syncthreads();
    }

    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    uint32_t stack[VT_STACK_SIZE];
    int sp = 0;
    uint32_t cur = 0;
    bool alive = false;      // lane owns a ray whose result is not written yet
    bool walking = false;    // that ray still has pairs to visit
    uint32_t qa = 0, na = 0, qb = 0, nb = 0;  // queued leaf runs: tris[qa, qa+na) then tris[qb, qb+nb)
    bool exhausted = false;  // warp-uniform: the ray queue has run dry
    unsigned long long ray_idx = 0;
    RayState r;
    unsigned long long n_invalid = 0;

    // invariant: a lane without a ray has walking == false and na == 0
    for (;;) {
        // ---- retire finished rays, refill idle lanes from the global queue
        if (alive && !walking && na == 0) {
            write_hit(hits, ray_idx, r);
            alive = false;
        }
        const unsigned idle = __ballot_sync(0xffffffffu, !alive);
        if (idle == 0xffffffffu && exhausted) break;
        if (idle && !exhausted && __popc(idle) >= 32 - refill_threshold) {
            const int n_idle = __popc(idle);
            const int leader = __ffs(idle) - 1;
            unsigned long long base = 0;
            if (persistent) {
                if ((int)lane == leader) base = atomicAdd(&counters[0], (unsigned long long)n_idle);
                base = __shfl_sync(0xffffffffu, base, leader);
            } else {
                base = ((unsigned long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31u));
                exhausted = true;  // one batch per warp
            }
            if (base + n_idle >= n) exhausted = true;
            if (!alive) {
                ray_idx = base + __popc(idle & lt_mask);
                if (ray_idx < n) {
                    const float4 *rp = reinterpret_cast<const float4 *>(rays + ray_idx);
                    const float4 ra = __ldg(rp), rb = __ldg(rp + 1);
                    vt_ray in{ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
                    init_ray(in, r);
                    alive = true;
                    sp = 0;
                    // argument rules of AccelStruct::Traverse (source/objects/AccelStruct.cpp:805-806):
                    // tMin < 0 or tMax <= tMin is an error there; here the ray becomes a counted miss.
                    if (!(in.tmin >= 0.f) || !(in.tmax > in.tmin)) {
                        if (!(in.tmax < 0.f)) n_invalid++;  // tmax < 0 marks a masked slot of a ray wave: silent miss
                    } else if (S.root_leaf_count) {
                        // root is a leaf: its triangles are tested directly, no slab test (single_ray_traverser.hpp:72-73)
                        qa = 0;
                        na = S.root_leaf_count;
                    } else if (S.n_pairs) {
                        walking = true;
                        cur = 0;  // pair 0 = children of the root (nodes[nodes[0].first], +1)
                    }
                }
            }
        }
        const int keep = exhausted ? 0 : refill_threshold;

        // ---- rounds, until too few lanes have work left
        for (;;) {
            const unsigned want_tri = __ballot_sync(0xffffffffu, na != 0);
            const unsigned want_node = __ballot_sync(0xffffffffu, na == 0 && walking);
            if (__popc(want_tri | want_node) <= keep) break;
            if (want_tri && (want_node == 0 || __popc(want_tri) >= tri_threshold)) {
                // ---- triangle round: one queued candidate per lane (intersect_leaf loop body, :53-61)
                if (na != 0) {
                    const bool hit = intersect_triangle<ALPHA>(S, qa, r);
                    qa++;
                    na--;
                    if (ANY_HIT && hit) {  // any_hit: first accepted candidate ends the ray (:57-58, :91-93)
                        na = nb = 0;
                        walking = false;
                    } else if (na == 0) {
                        qa = qb;
                        na = nb;
                        nb = 0;
                    }
                }
            } else if (na == 0 && walking) {
                // ---- node round (:82-123), written branch-free: everything below is selects and
                // predicated stack accesses, so lanes that take different exits do not serialise.
                float4 a, b, c, d;
                if (SMEM && cur < S.n_smem_pairs) {
                    const float4 *p = s_pairs + cur * 4u;
                    a = p[0], b = p[1], c = p[2], d = p[3];
                } else {
                    ldg256(S.pairs + cur, a, b);
                    ldg256(reinterpret_cast<const char *>(S.pairs + cur) + 32, c, d);
                }
                float le, lx, re, rx;
                slab_pair(a, b, c, d, r, le, lx, re, rx);  // both boxes against the tmax of step entry (:86-87)
                const uint32_t lcount = __float_as_uint(b.z), lfirst = __float_as_uint(b.w);
                const uint32_t rcount = __float_as_uint(d.z), rfirst = __float_as_uint(d.w);
                const bool hit_l = le <= lx, hit_r = re <= rx;
                const bool leaf_l = hit_l && lcount != 0, leaf_r = hit_r && rcount != 0;
                const bool in_l = hit_l && lcount == 0, in_r = hit_r && rcount == 0;
                // leaf children: the left run is tested first (:89-97), then the right run (:99-107)
                qa = leaf_l ? lfirst : rfirst;
                na = leaf_l ? lcount : (leaf_r ? rcount : 0u);
                qb = rfirst;
                nb = (leaf_l && leaf_r) ? rcount : 0u;
                // inner children: near first, far pushed; `le > re` swaps, ties keep the left child first (:109-115)
                const bool both = in_l && in_r;
                const bool take_r = both ? (le > re) : in_r;
                const uint32_t next = take_r ? rfirst : lfirst;
                const uint32_t far_ = take_r ? lfirst : rfirst;
                if (both) {
                    stack[sp & (VT_STACK_SIZE - 1)] = far_;  // reference: unchecked, UB past 64 entries; depth is validated at populate
                    sp++;
                }
                const bool pop = !(in_l || in_r);
                if (pop && sp > 0) {
                    sp--;
                    cur = stack[sp & (VT_STACK_SIZE - 1)];
                } else {
                    cur = next;
                    walking = !pop;
                }
            }
        }
    }
    if (n_invalid) atomicAdd(&counters[1], n_invalid);
}

}  // namespace

cudaError_t vt_launch_traverse(const VtSceneView &S, const vt_ray *rays, vt_hit *hits, uint64_t n, bool any_hit,
                               unsigned long long *counters, const VtLaunchConfig &cfg, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const size_t smem = (size_t)S.n_smem_pairs * sizeof(VtPair);
    int grid;
    if (cfg.persistent) {
        grid = cfg.grid;
    } else {
        grid = (int)((n + VT_TRAVERSE_BLOCK - 1) / VT_TRAVERSE_BLOCK);
    }
    const bool alpha = S.has_alphatest != 0;
    auto launch = [&](auto kernel) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kernel<<<grid, VT_TRAVERSE_BLOCK, smem, stream>>>(S, rays, hits, (unsigned long long)n, counters,
                                                          cfg.persistent ? 1 : 0, cfg.refill_threshold, cfg.tri_threshold);
        return cudaGetLastError();
    };
    if (S.n_smem_pairs) {
        if (any_hit) return alpha ? launch(k_traverse<true, true, true>) : launch(k_traverse<true, false, true>);
        return alpha ? launch(k_traverse<false, true, true>) : launch(k_traverse<false, false, true>);
    }
    if (any_hit) return alpha ? launch(k_traverse<true, true, false>) : launch(k_traverse<true, false, false>);
    return alpha ? launch(k_traverse<false, true, false>) : launch(k_traverse<false, false, false>);
}

cudaError_t vt_traverse_occupancy(int *blocks_per_sm, size_t smem_bytes) {
    cudaError_t e = cudaFuncSetAttribute(k_traverse<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k_traverse<false, true, true>, VT_TRAVERSE_BLOCK, smem_bytes);
}
