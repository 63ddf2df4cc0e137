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
#include <cuda_fp16.h>
#include <algorithm>
#include <cstdlib>
#include <type_traits>

// Build-time knobs of the ALU-pipe diet (A/B builds: make EXTRA=-DVT_...=0); see the notes at slab_quad.
#ifndef VT_SCHED2
#define VT_SCHED2 1
#endif
#ifndef VT_TRI_ADDR_WIDE
#define VT_TRI_ADDR_WIDE 1
#endif
// Ray prefetch ring (quantised kernels): a warp pulls VT_RAY_BATCH rays from the global queue at once — ONE atomic, coalesced
// 32-byte loads — derives safe_inverse / scaled origin for all of them with every lane busy, and parks the prepared ray states in
// shared memory; a lane that finishes its ray later picks the next state up with four LDS.128, no atomic, no global latency and
// no per-lane divisions.  That makes a refill cheap enough to run with only a few idle lanes (VT_REFILL 28 instead of 24), so node
// rounds execute with fuller warps.  0 = the round-1 path (refill straight from global memory).
// Measured (profiles/r2_k1_experiments.md): SLOWER than the round-1 refill at every threshold (3.33 vs 3.51 Grays/s on the bounce
// wave; 64-entry rings 3.21) — the 9 KB of shared memory per CTA come out of the L1 that serves half of the node fetches.  Off.
#ifndef VT_RAY_BATCH
#define VT_RAY_BATCH 0
#endif
// Tail work-sharing (k_traverse_compact): lanes of a warp whose queue is dry help the warp's long rays once at most VT_SHARE_LANES
// lanes still have work; the warp looks for new pairs every VT_SHARE_ROUNDS rounds.  VT_SHARE_GIVE=0 keeps the phase but never
// hands work over (A/B, diagnosis).
#ifndef VT_SHARE_LANES
#define VT_SHARE_LANES 4
#endif
#ifndef VT_SHARE_ROUNDS
#define VT_SHARE_ROUNDS 16
#endif
#ifndef VT_SHARE_REPS
#define VT_SHARE_REPS 1
#endif
#ifndef VT_SHARE_GIVE
#define VT_SHARE_GIVE 1
#endif
#ifndef VT_SHARE_DEBUG
#define VT_SHARE_DEBUG 0
#endif
#ifndef VT_STACK_DIST
#define VT_STACK_DIST 0  // measured: -7 % node visits / -22 % triangle tests on primary rays (+3 %), but -5.6 % on the bounce wave
#endif

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
// Cache-eviction qualifiers (PTX: ld.global.nc{.L1::x}{.L2::y}.v8): the hierarchy is the data every ray
// re-reads, triangles are touched a few times per ray, rays are read once.
#ifndef VT_QUAL_NODE
#define VT_QUAL_NODE ""
#endif
#ifndef VT_QUAL_TRI
#define VT_QUAL_TRI ""
#endif
#ifndef VT_QUAL_RAY
#define VT_QUAL_RAY ""
#endif
VT_DEV void ldg256(const void *p, float4 &lo, float4 &hi) {
    asm volatile("ld.global.nc" VT_QUAL_NODE ".v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w)
                 : "l"(p));
}

VT_DEV void ldg256_tri(const void *p, float4 &lo, float4 &hi) {
    asm volatile("ld.global.nc" VT_QUAL_TRI ".v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w)
                 : "l"(p));
}
VT_DEV void ldg256_ray(const void *p, float4 &lo, float4 &hi) {
    asm volatile("ld.global.nc" VT_QUAL_RAY ".v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w)
                 : "l"(p));
}
VT_DEV float safe_inverse(float d) {
    // libs/bvh/include/bvh/vector.hpp:69-74
    return 1.0f / (fabsf(d) < FLT_EPSILON ? copysignf(FLT_EPSILON, d) : d);
}

// TriangleBackfaceCull::intersect — source/objects/Primitives.h:168-215.  Updates the ray's best
// hit and tmax when the candidate is accepted (`t <= tmax`: a later equal-t candidate replaces).
// CANON (quantised kernels): among candidates with EQUAL t the one with the larger original index wins, whatever the order they
// are met in — the reference keeps the one it meets last (`t <= tmax` replaces), an order the quantised layouts do not promise
// anyway (vt_device.h).  It makes the answer independent of the visit order, which the tail work-sharing below needs: the parts
// of one ray that several lanes traverse can then be merged in any order.
template <bool ALPHA, bool CANON = false>
VT_DEV bool intersect_triangle(const VtSceneView &S, uint32_t slot, RayState &r) {
    float4 q0, q1, q2, q3;
#if VT_TRI_ADDR_WIDE
    // one IMAD.WIDE (FMA pipe) instead of the shift/mask/add-with-carry chain the compiler derives from the tagged reference
    uint64_t rec;
    asm("mad.wide.u32 %0, %1, 64, %2;" : "=l"(rec) : "r"(slot), "l"(S.tris));
    ldg256_tri(reinterpret_cast<const char *>(rec), q0, q1);
    ldg256_tri(reinterpret_cast<const char *>(rec) + 32, q2, q3);
#else
    ldg256_tri(S.tris + slot, q0, q1);
    ldg256_tri(reinterpret_cast<const char *>(S.tris + slot) + 32, q2, q3);
#endif
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
            if (CANON && t == r.tmax && r.prim != VT_MISS && __float_as_uint(q3.y) < r.prim) return false;
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

VT_DEV void ldg256u(const void *p, uint4 &lo, uint4 &hi) {
    asm volatile("ld.global.nc" VT_QUAL_NODE ".v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
                 : "l"(p));
}

// One step over a COMPACT pair (vt_device.h: VtCPair): a single 32-byte load, exact decode of the
// conservative planes ((k + q) * 2^E through one fma on the 2^23 + q float), then the reference's slab
// arithmetic fmaf(plane, inv, so) on them.  fmaf is monotone in `plane`, so entry' <= entry and
// exit' >= exit of the exact box: every node FastNodeIntersector accepts is accepted here.
// Children come back as TAGGED references: bits 28-31 = triangle count (0 = inner), bits 0-27 = pair
// index or first triangle slot.  `magic` is 0x4B000000 read from the kernel parameter block, so PRMT
// takes it as a constant-bank operand and its selector stays an immediate.
// Software prefetch of records whose address is known a round or more before they are read: the first triangle of
// a leaf the lane just selected, and the children it pushed.  0 = off, 1 = prefetch.global.L2, 2 = prefetch.global.L1.
#ifndef VT_PREFETCH
#define VT_PREFETCH 0
#endif
VT_DEV void vt_prefetch(const void *p) {
#if VT_PREFETCH == 1
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#elif VT_PREFETCH == 2
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}
VT_DEV void vt_prefetch_ref(const VtSceneView &S, uint32_t ref, bool quad) {
#if VT_PREFETCH
    const uint32_t idx = ref & 0x0FFFFFFFu;
    if (ref >> 28) vt_prefetch(S.tris + idx);
    else if (quad) vt_prefetch(S.quads + idx);
    else vt_prefetch(S.cpairs + idx);
#else
    (void)S, (void)ref, (void)quad;
#endif
}
#ifndef VT_PREFETCH_LEAF
#define VT_PREFETCH_LEAF 1
#endif
#ifndef VT_PREFETCH_FAR
#define VT_PREFETCH_FAR 1
#endif
#ifndef VT_PREFETCH_ALL
#define VT_PREFETCH_ALL 0  // every pushed child, not only the next one to be popped
#endif
#ifndef VT_LEAF_RUN_PER_ROUND
#define VT_LEAF_RUN_PER_ROUND 1
#endif
#define VT_REF_SHIFT 28
#define VT_REF_MASK 0x0FFFFFFFu
#define VT_REF_DONE 0xFFFFFFFFu
VT_DEV void slab_cpair(const VtCPair *cpairs, uint32_t cur, uint32_t magic, const RayState &r, float &le, float &lx, float &re,
                       float &rx, uint32_t &lref, uint32_t &rref) {
    uint4 w0, w1;
    ldg256u(cpairs + cur, w0, w1);
    const float sx = __uint_as_float((w0.w & 0xffu) << 23);
    const float sy = __uint_as_float((w0.w << 15) & 0x7f800000u);
    const float sz = __uint_as_float((w0.w << 7) & 0x7f800000u);
    const float ax = __uint_as_float(w0.x), ay = __uint_as_float(w0.y), az = __uint_as_float(w0.z);
    // per axis the word holds {l.lo, l.hi, r.lo, r.hi}; with a negative direction the near plane is hi
    // (node_intersectors.hpp:20-26): swap inside the byte pairs so the word reads {l.near, l.far, r.near, r.far}
    const uint32_t qx = __byte_perm(w1.x, 0u, signbit(r.inv.x) ? 0x2301u : 0x3210u);
    const uint32_t qy = __byte_perm(w1.y, 0u, signbit(r.inv.y) ? 0x2301u : 0x3210u);
    const uint32_t qz = __byte_perm(w1.z, 0u, signbit(r.inv.z) ? 0x2301u : 0x3210u);
#define VT_PLANE(q, sel, s, a) fmaf(__uint_as_float(__byte_perm(q, magic, sel)), s, a)
    float e0 = fmaf(VT_PLANE(qx, 0x7650u, sx, ax), r.inv.x, r.so.x);
    float e1 = fmaf(VT_PLANE(qy, 0x7650u, sy, ay), r.inv.y, r.so.y);
    float e2 = fmaf(VT_PLANE(qz, 0x7650u, sz, az), r.inv.z, r.so.z);
    float x0 = fmaf(VT_PLANE(qx, 0x7651u, sx, ax), r.inv.x, r.so.x);
    float x1 = fmaf(VT_PLANE(qy, 0x7651u, sy, ay), r.inv.y, r.so.y);
    float x2 = fmaf(VT_PLANE(qz, 0x7651u, sz, az), r.inv.z, r.so.z);
    le = fmaxf(e0, fmaxf(e1, fmaxf(e2, r.tmin)));
    lx = fminf(x0, fminf(x1, fminf(x2, r.tmax)));
    e0 = fmaf(VT_PLANE(qx, 0x7652u, sx, ax), r.inv.x, r.so.x);
    e1 = fmaf(VT_PLANE(qy, 0x7652u, sy, ay), r.inv.y, r.so.y);
    e2 = fmaf(VT_PLANE(qz, 0x7652u, sz, az), r.inv.z, r.so.z);
    x0 = fmaf(VT_PLANE(qx, 0x7653u, sx, ax), r.inv.x, r.so.x);
    x1 = fmaf(VT_PLANE(qy, 0x7653u, sy, ay), r.inv.y, r.so.y);
    x2 = fmaf(VT_PLANE(qz, 0x7653u, sz, az), r.inv.z, r.so.z);
#undef VT_PLANE
    re = fmaxf(e0, fmaxf(e1, fmaxf(e2, r.tmin)));
    rx = fminf(x0, fminf(x1, fminf(x2, r.tmax)));
    const uint32_t lcount = (w0.w >> 24) & 15u, rcount = w0.w >> 28;
    const uint32_t ref = w1.w, next = cur + 1u;
    const uint32_t lleaf = (lcount << VT_REF_SHIFT) | ref;                // left leaf: `ref` is its slot
    const uint32_t rleaf_after_inner = (rcount << VT_REF_SHIFT) | ref;    // left inner: `ref` belongs to the right child
    const uint32_t rleaf_after_leaf = (rcount << VT_REF_SHIFT) | (ref + lcount);
    lref = lcount == 0 ? next : lleaf;
    rref = lcount == 0 ? rleaf_after_inner : (rcount == 0 ? next : rleaf_after_leaf);
}

// One step over a QUAD (vt_device.h: VtQuad): two 32-byte loads, the four boxes decoded and slab-tested
// exactly like slab_cpair, then the children that were hit are ordered near-to-far by entry distance.
// Out: k[i] ascending sort keys (0x7FFFFFFF = no hit) and the matching tagged references r[i].
// Sort key of a hit child: the bits of its entry distance (en >= tmin >= 0, so the pattern orders like the
// value).  VT_KEY_SLOT=1 additionally breaks ties by slot index inside the key (two more LOP3 per child);
// without it the compare-exchange network leaves equal keys in network order — just as deterministic.
// VT_SLAB_TWO_FMA=1 forces the two-fma plane form for every ray (tuning / A-B builds).
#ifndef VT_SLAB_TWO_FMA
#define VT_SLAB_TWO_FMA 0
#endif
#ifndef VT_KEY_SLOT
#define VT_KEY_SLOT 0
#endif
// ALU-pipe diet (profiles/r1_k1_alu_diet.md: the half-rate ALU pipe — PRMT/SEL/ISETP/FMNMX/VIMNMX — is the busiest unit
// of this kernel, the FMA pipes idle at 20 %):
//   VT_EMPTY_SENTINEL  empty quad slots reference an all-NaN triangle record behind the last real one (vt_accel.cu), so
//                      the step needs no per-child "slot in use" test: a false positive costs one rejected triangle test.
//   VT_SORT_CE         compare-exchanges of the child ordering network: 5 = full sort, 4 = nearest and farthest exact,
//                      the middle two in network order, 3 = nearest exact only.  Any order is correct; order only prunes.
//   VT_DECODE_HALF     planes are decoded two at a time: one PRMT builds the half2 {1024 + q_a, 1024 + q_b} (bytes q, 0x64),
//                      two HADD2.F32 (FMA pipe) widen it — 12 PRMT + 24 conversions instead of 24 PRMT per step.  The host
//                      then stores origin_adj = (k - 1024) * 2^E (vt_device.h).  Measured: IMAD.HI is quarter rate on
//                      B200 (tools/ubench/pipes.cu), so the byte-3 decode through IMAD.HI was slower and is gone.
//   VT_KEY_MID         children are ordered by entry + exit (twice the midpoint of the ray's interval inside the box) instead of
//                      the entry distance.  A box that CONTAINS the ray origin has entry = tmin whatever it holds, so the entry
//                      order sends a camera inside a room's top-level boxes into the largest sub-tree first, however far its
//                      content is; the midpoint prefers the box that ends sooner.  For boxes that do not overlap along the ray the
//                      two orders are the same.  One FADD per child on the FMA pipe.  Foliage in a room, camera rays, SAH-optimal
//                      collapse: 40.7 -> 31.9 quad visits, 20.2 -> 14.6 triangle tests per ray; terrain and props unchanged
//                      (profiles/r2_child_order.md).  Any order is correct (canonical tie rule); order only prunes.
//                      A template parameter of the kernel (KEYMID), chosen per scene by the host from the sibling overlap of
//                      its hierarchy (VtLaunchConfig::key_mid; VT_KEY_ORDER = auto | entry | mid): on the bench terrain, where
//                      the orders coincide, the extra FADDs cost 1 % of the bounce launch (1.610 -> 1.626 ms, session r4e).
#ifndef VT_SORT_CE
#define VT_SORT_CE 5
#endif
#if VT_KEY_SLOT && VT_STACK_DIST
#error "VT_KEY_SLOT alters the low key bits; the stack's distance test (VT_STACK_DIST) needs exact entry distances"
#endif
#if VT_KEY_SLOT
#define VT_QUAD_KEY(en, i) (int)((__float_as_uint(en) & ~3u) | (unsigned)(i))
#else
#define VT_QUAD_KEY(en, i) (int)__float_as_uint(en)
#endif
//
// Plane arithmetic, two forms with the same guarantee (every box FastNodeIntersector accepts is accepted):
//   (OFF = VT_QUAD_OFFSET: 1024 with the half2 decode, 2^23 with the PRMT decode; origin_adj = (k - OFF) * 2^E)
//   TWO_FMA   t = fmaf(fmaf(OFF+q, 2^E, origin_adj), inv, so): exact decode, then the reference's expression.
//   one fma   t = fmaf(OFF+q, 2^E*inv, A) with A = origin_adj*inv + so rounded DOWN for entry planes and UP for exit
//             planes, once per node and axis.  2^E*inv is exact (power of two; the host keeps |E| <= 60 and the lane's
//             |inv| is in [2^-60, 2^24], so the product is a normal float), (OFF+q)*2^E + origin_adj is the decoded
//             plane exactly, hence the real value under the rounding is plane*inv + so minus a non-negative slack
//             (entry) or plus one (exit); rounding is monotone, so entry' <= the reference's entry and exit' >= its exit.
//             A warp holding a ray outside that |inv| range (|d| > 2^60, NaN) takes the TWO_FMA form.
#if VT_DECODE_HALF
// {1024 + byte a, 1024 + byte b} of q as two floats: PRMT interleaves the bytes with 0x64 (fp16 1024 = 0x6400, ulp 1)
VT_DEV float2 decode_half_pair(uint32_t q, uint32_t magic_h, uint32_t sel) {
    const uint32_t h = __byte_perm(q, magic_h, sel);
    return __half22float2(*reinterpret_cast<const __half2 *>(&h));
}
#endif
template <bool TWO_FMA, bool KEYMID>
VT_DEV void slab_quad(const VtQuad *quads, uint32_t cur, uint32_t magic, const RayState &ray, int (&k)[4], uint32_t (&r)[4],
                      const uint4 *s_top = nullptr, uint32_t n_top = 0) {
    uint4 a0, a1, b0, b1;
#if VT_SMEM_QUADS_BUILD
    if (cur < n_top) {  // the top of the hierarchy, staged per CTA (A/B builds)
        const uint4 *p = s_top + cur * 4u;
        a0 = p[0], a1 = p[1], b0 = p[2], b1 = p[3];
    } else
#endif
    {
        ldg256u(quads + cur, a0, a1);
        ldg256u(reinterpret_cast<const char *>(quads + cur) + 32, b0, b1);
    }
    // words: a0 = {origin_adj.xyz, scale.x}, a1 = {scale.y, scale.z, lo_x[4], hi_x[4]}, b0 = {lo_y[4], hi_y[4], lo_z[4], hi_z[4]}, b1 = refs
    const float ax = __uint_as_float(a0.x), ay = __uint_as_float(a0.y), az = __uint_as_float(a0.z);
    const float sx = __uint_as_float(a0.w), sy = __uint_as_float(a1.x), sz = __uint_as_float(a1.y);
    // a negative direction makes hi the near plane (node_intersectors.hpp:20-26)
    const bool ox = signbit(ray.inv.x), oy = signbit(ray.inv.y), oz = signbit(ray.inv.z);
    const uint32_t nx = ox ? a1.w : a1.z, fx = ox ? a1.z : a1.w;
    const uint32_t ny = oy ? b0.y : b0.x, fy = oy ? b0.x : b0.y;
    const uint32_t nz = oz ? b0.w : b0.z, fz = oz ? b0.z : b0.w;
    r[0] = b1.x, r[1] = b1.y, r[2] = b1.z, r[3] = b1.w;
#if VT_DECODE_HALF
    // `magic` is 0x64646464 here; children 0,1 from selector 0x4140, children 2,3 from 0x4342
    const float2 nx_a = decode_half_pair(nx, magic, 0x4140u), nx_b = decode_half_pair(nx, magic, 0x4342u);
    const float2 ny_a = decode_half_pair(ny, magic, 0x4140u), ny_b = decode_half_pair(ny, magic, 0x4342u);
    const float2 nz_a = decode_half_pair(nz, magic, 0x4140u), nz_b = decode_half_pair(nz, magic, 0x4342u);
    const float2 fx_a = decode_half_pair(fx, magic, 0x4140u), fx_b = decode_half_pair(fx, magic, 0x4342u);
    const float2 fy_a = decode_half_pair(fy, magic, 0x4140u), fy_b = decode_half_pair(fy, magic, 0x4342u);
    const float2 fz_a = decode_half_pair(fz, magic, 0x4140u), fz_b = decode_half_pair(fz, magic, 0x4342u);
#define VT_QF(q, sel) ((sel) == 0x7650u ? q##_a.x : (sel) == 0x7651u ? q##_a.y : (sel) == 0x7652u ? q##_b.x : q##_b.y)
#else
#define VT_QF(q, sel) __uint_as_float(__byte_perm(q, magic, sel))
#endif
#define VT_PLANE(q, sel, s, a) fmaf(VT_QF(q, sel), s, a)
    const float six = sx * ray.inv.x, siy = sy * ray.inv.y, siz = sz * ray.inv.z;
    const float alx = __fmaf_rd(ax, ray.inv.x, ray.so.x), aly = __fmaf_rd(ay, ray.inv.y, ray.so.y), alz = __fmaf_rd(az, ray.inv.z, ray.so.z);
    const float ahx = __fmaf_ru(ax, ray.inv.x, ray.so.x), ahy = __fmaf_ru(ay, ray.inv.y, ray.so.y), ahz = __fmaf_ru(az, ray.inv.z, ray.so.z);
#define VT_CHILD(i, sel)                                                                   \
    {                                                                                      \
        float e0, e1, e2, x0, x1, x2;                                                      \
        if (TWO_FMA) {                                                                     \
            e0 = fmaf(VT_PLANE(nx, sel, sx, ax), ray.inv.x, ray.so.x);                     \
            e1 = fmaf(VT_PLANE(ny, sel, sy, ay), ray.inv.y, ray.so.y);                     \
            e2 = fmaf(VT_PLANE(nz, sel, sz, az), ray.inv.z, ray.so.z);                     \
            x0 = fmaf(VT_PLANE(fx, sel, sx, ax), ray.inv.x, ray.so.x);                     \
            x1 = fmaf(VT_PLANE(fy, sel, sy, ay), ray.inv.y, ray.so.y);                     \
            x2 = fmaf(VT_PLANE(fz, sel, sz, az), ray.inv.z, ray.so.z);                     \
        } else {                                                                           \
            e0 = fmaf(VT_QF(nx, sel), six, alx);                                           \
            e1 = fmaf(VT_QF(ny, sel), siy, aly);                                           \
            e2 = fmaf(VT_QF(nz, sel), siz, alz);                                           \
            x0 = fmaf(VT_QF(fx, sel), six, ahx);                                           \
            x1 = fmaf(VT_QF(fy, sel), siy, ahy);                                           \
            x2 = fmaf(VT_QF(fz, sel), siz, ahz);                                           \
        }                                                                                  \
        const float en = fmaxf(e0, fmaxf(e1, fmaxf(e2, ray.tmin)));                        \
        const float ex = fminf(x0, fminf(x1, fminf(x2, ray.tmax)));                        \
        /* an empty slot's inverted box can look hit after rounding when the node is tiny and far: either test the   \
           ref too, or let the slot reference the all-NaN sentinel triangle (VT_EMPTY_SENTINEL) */                     \
        const bool hit = (VT_EMPTY_SENTINEL || r[i] != VT_REF_DONE) && en <= ex;           \
        /* en >= tmin >= 0 (and ex >= en when hit): the bit pattern orders like the value (-0.0 sorts first) */ \
        k[i] = hit ? VT_QUAD_KEY(KEYMID ? en + ex : en, i) : 0x7FFFFFFF;                   \
    }
    VT_CHILD(0, 0x7650u)
    VT_CHILD(1, 0x7651u)
    VT_CHILD(2, 0x7652u)
    VT_CHILD(3, 0x7653u)
#undef VT_CHILD
#undef VT_PLANE
#undef VT_QF
#define VT_CE(a, b)                       \
    if (k[a] > k[b]) {                    \
        const int tk = k[a];              \
        k[a] = k[b];                      \
        k[b] = tk;                        \
        const uint32_t tr = r[a];         \
        r[a] = r[b];                      \
        r[b] = tr;                        \
    }
    VT_CE(0, 1) VT_CE(2, 3) VT_CE(0, 2)
#if VT_SORT_CE >= 4
    VT_CE(1, 3)
#endif
#if VT_SORT_CE >= 5
    VT_CE(1, 2)
#endif
#undef VT_CE
}

// Traversal stack of the quantised kernels: the array lives in local memory and is addressed through a
// 32-bit .local address (st.local / ld.local with a 32-bit register), so push = STL + one VIADD and pop = LDL
// + one VIADD.  A generic `uint32_t *` costs a 64-bit pointer update next to the 32-bit one the compiler
// derives for STL/LDL anyway (profiles/r1_k_traverse_v5_step.md).
VT_DEV uint32_t local_addr(const void *p) {
    uint64_t a;
    asm("cvta.to.local.u64 %0, %1;" : "=l"(a) : "l"(p));
    return (uint32_t)a;
}
VT_DEV void stack_store(uint32_t a, uint32_t v) { asm volatile("st.local.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
VT_DEV uint32_t stack_load(uint32_t a) {
    uint32_t v;
    asm volatile("ld.local.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}

// Stack entries with the entry distance of the pushed child (VT_STACK_DIST, closest hit only): a pop discards every
// entry whose box starts beyond the CURRENT tmax — the ray found something nearer since the push — instead of fetching
// the node (or testing the whole leaf run) just to see every child fail.  The reference's traverser does not carry
// distances (single_ray_traverser.hpp:109-121) and re-tests such nodes; the result is the same: a discarded box cannot
// hold a candidate with t <= tmax, equal t included, because its conservative entry is <= the t of anything inside.
VT_DEV void stack_store2(uint32_t a, uint32_t ref, uint32_t key) {
    asm volatile("st.local.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(ref), "r"(key) : "memory");
}
VT_DEV void stack_load2(uint32_t a, uint32_t &ref, uint32_t &key) {
    asm volatile("ld.local.v2.u32 {%0, %1}, [%2];" : "=r"(ref), "=r"(key) : "r"(a) : "memory");
}
template <bool DIST>
VT_DEV void stack_push(uint32_t &sp, uint32_t ref, uint32_t key) {
    if (DIST) stack_store2(sp, ref, key), sp += 8u;
    else stack_store(sp, ref), sp += 4u;
}
// next reference to visit, or VT_REF_DONE when the stack is empty
template <bool DIST>
VT_DEV uint32_t stack_pop(uint32_t &sp, uint32_t stack, float tmax) {
    if constexpr (!DIST) {
        if (sp == stack) return 0xFFFFFFFFu;
        return stack_load(sp -= 4u);
    } else {
        while (sp != stack) {
            uint32_t ref, key;
            stack_load2(sp -= 8u, ref, key);
            if (__uint_as_float(key) <= tmax) return ref;  // keys are entry distances >= tmin >= 0 (or -0.0)
        }
        return 0xFFFFFFFFu;
    }
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
template <bool ANY_HIT, bool ALPHA, bool SMEM, bool STATS = false>
__global__ void __launch_bounds__(VT_TRAVERSE_BLOCK, VT_TRAVERSE_MIN_BLOCKS)
k_traverse(const VtSceneView S, const vt_ray *__restrict__ rays, vt_hit *__restrict__ hits, unsigned long long n,
           unsigned long long *__restrict__ counters, int persistent, int refill_threshold, int tri_threshold,
           const uint32_t *__restrict__ queue, const unsigned long long *__restrict__ queue_count) {
    extern __shared__ float4 s_pairs[];
    if (queue_count) n = min(n, *queue_count);  // ray queue: only the slots the generator listed are traced
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
    unsigned long long n_steps = 0, n_tests = 0;  // STATS: SingleRayTraverser::Statistics (single_ray_traverser.hpp:132-135)

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
                    if (queue) ray_idx = __ldg(queue + ray_idx);  // queued launch: the slot this entry names
                    float4 ra, rb;
                    ldg256_ray(rays + ray_idx, ra, rb);
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
                    if (STATS) n_tests++;
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
                float le, lx, re, rx;
                uint32_t lcount, lfirst, rcount, rfirst;
                if (STATS) n_steps++;
                {
                    float4 a, b, c, d;
                    if (SMEM && cur < S.n_smem_pairs) {
                        const float4 *p = s_pairs + cur * 4u;
                        a = p[0], b = p[1], c = p[2], d = p[3];
                    } else {
                        ldg256(S.pairs + cur, a, b);
                        ldg256(reinterpret_cast<const char *>(S.pairs + cur) + 32, c, d);
                    }
                    slab_pair(a, b, c, d, r, le, lx, re, rx);  // both boxes against the tmax of step entry (:86-87)
                    lcount = __float_as_uint(b.z), lfirst = __float_as_uint(b.w);
                    rcount = __float_as_uint(d.z), rfirst = __float_as_uint(d.w);
                }
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
    if (STATS) {
        atomicAdd(&counters[2], n_steps);
        atomicAdd(&counters[3], n_tests);
    }
}

// K1 on the COMPACT layout.  Same persistent-warp / refill / warp-round machinery; the per-lane automaton
// is simpler because this layout does not promise the reference's visit order (only its result, see
// vt_device.h): a lane holds ONE tagged reference `cur` — an inner pair to step through, or a leaf run
// whose remaining count and next slot are both encoded in it — leaves are ordered against inner siblings
// by entry distance like any other child (ties: left first), and the far child of either kind is pushed.
//
// TAIL WORK-SHARING (SHARE; closest hit only).  Once the ray queue is dry a launch is as long as its longest rays: measured on the
// bench scene, 0.1 % of the primary rays (chains of 100 - 440 dependent rounds against a median of 24) hold the last 0.1 ms of
// every launch — 40 % of a 262 k-ray launch, what one rank of an 8-GPU frame traces (profiles/r2_tail_sharing.md).  A ray's pending
// sub-trees are independent, so in that phase an idle lane TAKES the nearest pending sub-tree of a lane that has one (with a copy of
// the ray and its current tmax), walks it on its own stack, and hands its best candidate back; the owner keeps the closest (ties: the
// canonical rule above) and writes the record when all its helpers have reported.  A helper may itself give work away; the count of
// outstanding helpers is kept by the ray's owner.  Nothing changes while the queue still has rays: refills keep the lanes busy.
template <bool ANY_HIT, bool ALPHA, bool STATS, bool QUAD, bool SHARE = false, bool KEYMID = false>
__global__ void __launch_bounds__(VT_TRAVERSE_BLOCK, VT_COMPACT_MIN_BLOCKS)
k_traverse_compact(const VtSceneView S, const vt_ray *__restrict__ rays, vt_hit *__restrict__ hits, unsigned long long n,
                   unsigned long long *__restrict__ counters, int persistent, int refill_threshold, int tri_threshold,
                   const uint32_t *__restrict__ queue, const unsigned long long *__restrict__ queue_count) {
    if (queue_count) n = min(n, *queue_count);  // ray queue: only the slots the generator listed are traced
#if VT_SMEM_QUADS_BUILD
    extern __shared__ uint4 s_top_quads[];
    const uint32_t n_top = QUAD ? S.n_smem_pairs : 0u;
    if (QUAD && n_top) {
        for (uint32_t i = threadIdx.x; i < n_top * 4u; i += blockDim.x) s_top_quads[i] = __ldg(reinterpret_cast<const uint4 *>(S.quads) + i);
        __syncthreads();
    }
#else
    const uint4 *const s_top_quads = nullptr;
    const uint32_t n_top = 0;
#endif
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    // the layouts served by this kernel validate the worst-case stack depth on the host (<= VT_STACK_SIZE), so
    // the stack is a bare pointer: push = store + increment, pop = decrement + load
    constexpr bool DIST = VT_STACK_DIST && !ANY_HIT;  // any-hit rays never shrink tmax before they end
    uint64_t stack_mem[DIST ? VT_STACK_SIZE : VT_STACK_SIZE / 2];
    const uint32_t stack = local_addr(stack_mem);
    uint32_t sp = stack;         // byte address of the next free entry
    uint32_t cur = VT_REF_DONE;  // VT_REF_DONE: nothing left to visit
    bool alive = false;          // lane owns a ray whose result is not written yet
    bool exhausted = false;      // warp-uniform: the ray queue has run dry
    unsigned long long ray_idx = 0;
    RayState r;
    unsigned long long n_invalid = 0, n_steps = 0, n_tests = 0;
    // STATS: optional per-ray record steps | tests << 16 (counters[4] holds the device address of a uint32 array, or 0)
    uint32_t *per_ray = STATS ? reinterpret_cast<uint32_t *>(counters[4]) : nullptr;
    uint32_t ray_steps = 0, ray_tests = 0;
    const uint32_t magic = (QUAD && VT_DECODE_HALF) ? S.magic_h : S.magic;
    bool warp_wild = false;  // warp-uniform: some lane holds a ray the one-fma plane form is not proven for (slab_quad)
    constexpr bool TEAM = SHARE && !ANY_HIT && !STATS && !VT_RAY_BATCH;
    int owner = -1;    // TEAM: >= 0 on a helper lane: the lane that owns the ray this lane walks a sub-tree of
    int pending = 0;   // TEAM: on an owner lane: helpers that have not reported yet
    bool last_tri = false;  // TEAM, warp-uniform: the kind of the previous round (tail phase: kinds alternate when both are wanted)

#if VT_RAY_BATCH
    // prepared ray states of this warp: {o.xyz, tmin} {d.xyz, tmax} {inv.xyz, flags} {so.xyz, -} + the slot each belongs to
    __shared__ float4 s_state[VT_TRAVERSE_BLOCK / 32][VT_RAY_BATCH][4];
    __shared__ unsigned long long s_slot[VT_TRAVERSE_BLOCK / 32][VT_RAY_BATCH];
    const unsigned warp_in_cta = threadIdx.x >> 5;
    int buf_pos = 0, buf_cnt = 0;  // warp-uniform: entries [buf_pos, buf_cnt) of the ring are still to be handed out
#endif

    for (;;) {
        if (TEAM && exhausted) {
            // helpers that have finished report to their owners (warp-uniform loop over the finished helpers)
            unsigned fin = __ballot_sync(0xffffffffu, alive && owner >= 0 && cur == VT_REF_DONE);
            while (fin) {
                const int h = __ffs(fin) - 1;
                fin &= fin - 1;
                const int o = __shfl_sync(0xffffffffu, owner, h);
                const float ht = __shfl_sync(0xffffffffu, r.tmax, h), hu = __shfl_sync(0xffffffffu, r.u, h), hv = __shfl_sync(0xffffffffu, r.v, h);
                const uint32_t hp = __shfl_sync(0xffffffffu, r.prim, h);
                if ((int)lane == o) {
                    pending--;
                    if (hp != VT_MISS && (ht < r.tmax || (ht == r.tmax && (r.prim == VT_MISS || hp > r.prim)))) r.tmax = ht, r.u = hu, r.v = hv, r.prim = hp;
                }
                if ((int)lane == h) alive = false, owner = -1;
            }
        }
        if (alive && cur == VT_REF_DONE && (!TEAM || pending == 0)) {
            if (VT_SHARE_DEBUG && TEAM && owner >= 0) atomicAdd(&counters[1], 1ull << 32);
            if (VT_SHARE_DEBUG && TEAM && pending < 0) atomicAdd(&counters[1], 1ull << 40);
            write_hit(hits, ray_idx, r);
            if (STATS && per_ray) per_ray[ray_idx] = min(ray_steps, 0xFFFFu) | (min(ray_tests, 0xFFFFu) << 16);
            ray_steps = ray_tests = 0;
            alive = false;
        }
        const unsigned idle = __ballot_sync(0xffffffffu, !alive);
        int round_budget = 0x7FFFFFFF;  // TEAM, tail phase: rounds before the warp looks for idle lanes again
#if VT_RAY_BATCH
        if (idle == 0xffffffffu && exhausted && buf_pos == buf_cnt) break;
        if (idle && !(exhausted && buf_pos == buf_cnt) && __popc(idle) >= 32 - refill_threshold) {
            const int n_idle = __popc(idle);
            if (buf_pos == buf_cnt && !exhausted) {
                // ---- fill the ring: one atomic for the batch, coalesced loads, ray set-up with all lanes busy
                unsigned long long base = 0;
                if (persistent) {
                    if (lane == 0) base = atomicAdd(&counters[0], (unsigned long long)VT_RAY_BATCH);
                    base = __shfl_sync(0xffffffffu, base, 0);
                } else {
                    base = ((unsigned long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31u));
                    exhausted = true;
                }
                if (base + VT_RAY_BATCH >= n) exhausted = true;
                bool fresh_wild = false;
                for (unsigned e = lane; e < (unsigned)VT_RAY_BATCH; e += 32u) {
                    const unsigned long long q = base + e;
                    if (q < n) {
                        const unsigned long long slot = queue ? (unsigned long long)__ldg(queue + q) : q;  // queued launch: the slot this entry names
                        float4 ra, rb;
                        ldg256_ray(rays + slot, ra, rb);
                        const float ix = safe_inverse(rb.x), iy = safe_inverse(rb.y), iz = safe_inverse(rb.z);
                        // argument rules of AccelStruct::Traverse (source/objects/AccelStruct.cpp:805-806) -> counted miss; tmax < 0: masked slot
                        uint32_t flags = 0;
                        if (!(ra.w >= 0.f) || !(rb.w > ra.w)) {
                            flags = 1u;
                            if (!(rb.w < 0.f)) n_invalid++;
                        }
                        fresh_wild |= VT_SLAB_TWO_FMA || !(fminf(fabsf(ix), fminf(fabsf(iy), fabsf(iz))) >= 0x1p-60f);
                        s_state[warp_in_cta][e][0] = ra;
                        s_state[warp_in_cta][e][1] = rb;
                        s_state[warp_in_cta][e][2] = make_float4(ix, iy, iz, __uint_as_float(flags));
                        s_state[warp_in_cta][e][3] = make_float4(-ra.x * ix, -ra.y * iy, -ra.z * iz, 0.f);
                        s_slot[warp_in_cta][e] = slot;
                    }
                }
                // sticky for this warp's share of the launch: such rays are pathological input, the flag only has to be right, not tight
                if (__any_sync(0xffffffffu, fresh_wild)) warp_wild = true;
                buf_pos = 0;
                buf_cnt = base < n ? (int)min((unsigned long long)VT_RAY_BATCH, n - base) : 0;
                __syncwarp();
            }
            const int avail = buf_cnt - buf_pos;
            if (!alive) {
                const int rank_in_idle = __popc(idle & lt_mask);
                if (rank_in_idle < avail) {
                    const int e = buf_pos + rank_in_idle;
                    const float4 a = s_state[warp_in_cta][e][0], b = s_state[warp_in_cta][e][1], c = s_state[warp_in_cta][e][2], d = s_state[warp_in_cta][e][3];
                    ray_idx = s_slot[warp_in_cta][e];
                    r.o = mk3(a.x, a.y, a.z), r.tmin = a.w;
                    r.d = mk3(b.x, b.y, b.z), r.tmax = b.w;
                    r.inv = mk3(c.x, c.y, c.z), r.so = mk3(d.x, d.y, d.z);
                    r.u = r.v = 0.f;
                    r.prim = VT_MISS;
                    alive = true;
                    sp = stack;
                    if (__float_as_uint(c.w)) cur = VT_REF_DONE;                                             // invalid or masked: a miss
                    else if (S.root_leaf_count) cur = S.root_leaf_count << VT_REF_SHIFT;                     // the root is a leaf over tris[0, count)
                    else cur = S.n_pairs ? 0u : VT_REF_DONE;                                                 // pair 0 / quad 0: the children of the root
                }
            }
            buf_pos += min(n_idle, avail);
            __syncwarp();  // every pick-up has read its entry before a later refill may overwrite the ring
        }
        const int keep = (exhausted && buf_pos == buf_cnt) ? 0 : refill_threshold;
#else
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
                exhausted = true;
            }
            if (base + n_idle >= n) exhausted = true;
            bool fresh_wild = false;
            if (!alive) {
                ray_idx = base + __popc(idle & lt_mask);
                if (ray_idx < n) {
                    if (queue) ray_idx = __ldg(queue + ray_idx);  // queued launch: the slot this entry names
                    float4 ra, rb;
                    ldg256_ray(rays + ray_idx, ra, rb);
                    vt_ray in{ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
                    init_ray(in, r);
                    alive = true;
                    sp = stack;
                    fresh_wild = VT_SLAB_TWO_FMA || !(fminf(fabsf(r.inv.x), fminf(fabsf(r.inv.y), fabsf(r.inv.z))) >= 0x1p-60f);
                    if (!(in.tmin >= 0.f) || !(in.tmax > in.tmin)) {  // AccelStruct.cpp:805-806 -> counted miss; tmax < 0: masked slot
                        if (!(in.tmax < 0.f)) n_invalid++;
                    } else if (S.root_leaf_count) {
                        cur = S.root_leaf_count << VT_REF_SHIFT;  // the root is a leaf over tris[0, count)
                    } else if (S.n_pairs) {
                        cur = 0;  // pair 0 / quad 0: the children of the root
                    }
                }
            }
            // sticky for this warp's share of the launch: such rays are pathological input, the flag only has to be right, not tight
            if (__any_sync(0xffffffffu, fresh_wild)) warp_wild = true;
        }
        int keep = exhausted ? 0 : refill_threshold;
        bool tail = false;  // warp-uniform: the queue is dry and few lanes still have work — the phase in which lanes share
        if (TEAM && exhausted) {
            tail = __popc(__ballot_sync(0xffffffffu, alive && owner < 0)) <= VT_SHARE_LANES;  // rays still in flight (helpers not counted)
            if (!tail) keep = VT_SHARE_LANES;  // ordinary rounds until only the long rays are left
        }
        if (TEAM && tail) {
            // idle lanes take the nearest pending sub-tree of lanes that have one: the k-th idle lane from the k-th such lane; repeated
            // (VT_SHARE_REPS) so that a lone long ray with several pending sub-trees gets several helpers at once
            for (int rep = 0; rep < VT_SHARE_REPS; rep++) {
            const bool can_give = VT_SHARE_GIVE && alive && cur != VT_REF_DONE && sp != stack;
            const unsigned givers = __ballot_sync(0xffffffffu, can_give);
            // NOT `idle` from above: in the iteration in which the queue runs dry some of those lanes have just been refilled, and a
            // giver whose partner does not take would drop the sub-tree it popped
            const unsigned takers = __ballot_sync(0xffffffffu, !alive);
            const int n_pairs = min(__popc(givers), __popc(takers));
            if (n_pairs == 0) break;
            {
                const int my_rank = __popc((alive ? givers : takers) & lt_mask);
                const bool gives = can_give && my_rank < n_pairs, takes = !alive && my_rank < n_pairs;
                uint32_t given = VT_REF_DONE;
                if (gives) given = stack_pop<DIST>(sp, stack, r.tmax);
                const int src = takes ? (int)__fns(givers, 0, my_rank + 1) : (int)lane;
                const uint32_t ref_in = __shfl_sync(0xffffffffu, given, src);
                const int root_in = __shfl_sync(0xffffffffu, owner >= 0 ? owner : (int)lane, src);
                const float ox = __shfl_sync(0xffffffffu, r.o.x, src), oy = __shfl_sync(0xffffffffu, r.o.y, src), oz = __shfl_sync(0xffffffffu, r.o.z, src);
                const float dx = __shfl_sync(0xffffffffu, r.d.x, src), dy = __shfl_sync(0xffffffffu, r.d.y, src), dz = __shfl_sync(0xffffffffu, r.d.z, src);
                const float ix = __shfl_sync(0xffffffffu, r.inv.x, src), iy = __shfl_sync(0xffffffffu, r.inv.y, src), iz = __shfl_sync(0xffffffffu, r.inv.z, src);
                const float sx = __shfl_sync(0xffffffffu, r.so.x, src), sy = __shfl_sync(0xffffffffu, r.so.y, src), sz = __shfl_sync(0xffffffffu, r.so.z, src);
                const float tmn = __shfl_sync(0xffffffffu, r.tmin, src), tmx = __shfl_sync(0xffffffffu, r.tmax, src);
                if (takes && ref_in != VT_REF_DONE) {  // DONE: the giver's entry was discarded by the distance test of its stack
                    r.o = mk3(ox, oy, oz), r.d = mk3(dx, dy, dz), r.inv = mk3(ix, iy, iz), r.so = mk3(sx, sy, sz);
                    r.tmin = tmn, r.tmax = tmx;
                    r.u = r.v = 0.f;
                    r.prim = VT_MISS;
                    cur = ref_in;
                    sp = stack;
                    alive = true;
                    owner = root_in;
                }
                // the owners count their new helpers
                unsigned joined = __ballot_sync(0xffffffffu, takes && ref_in != VT_REF_DONE);
                while (joined) {
                    const int t = __ffs(joined) - 1;
                    joined &= joined - 1;
                    if (__shfl_sync(0xffffffffu, owner, t) == (int)lane) pending++;
                }
            }
            }
            // back here as soon as a lane runs out of work, and after a few rounds while lanes are idle (new sub-trees appear on the stacks)
            const unsigned working = __ballot_sync(0xffffffffu, alive && cur != VT_REF_DONE);
            keep = max(0, __popc(working) - 1);
            if (working != 0xffffffffu) round_budget = VT_SHARE_ROUNDS;
        }
#endif

        // the rounds, until too few lanes have work left; instantiated twice so that the bookkeeping of the tail phase (round budget,
        // alternating round kinds) costs nothing while the queue still has rays
        auto run_rounds = [&](auto tail_tag) {
        constexpr bool TAIL = decltype(tail_tag)::value;
        for (;;) {
            if (TAIL && round_budget-- <= 0) break;
#if VT_SCHED2
            // one compare per class: inner references are < 2^28, VT_REF_DONE is the only value that is neither
            const bool has = cur != VT_REF_DONE;
            const bool is_node = cur < (1u << VT_REF_SHIFT);
            const bool is_leaf = has && !is_node;
            const unsigned want_node = __ballot_sync(0xffffffffu, is_node);
            const unsigned want_any = __ballot_sync(0xffffffffu, has);
            const unsigned want_tri = want_any & ~want_node;
            if (__popc(want_any) <= keep) break;
#else
            const bool has = cur != VT_REF_DONE;
            const bool is_leaf = has && (cur >> VT_REF_SHIFT) != 0;
            const unsigned want_tri = __ballot_sync(0xffffffffu, is_leaf);
            const unsigned want_node = __ballot_sync(0xffffffffu, has && !is_leaf);
            if (__popc(want_tri | want_node) <= keep) break;
#endif
            bool tri_round = want_tri && (want_node == 0 || __popc(want_tri) >= tri_threshold);
            if (TAIL) {  // few lanes, idle issue slots: nobody waits longer than one round for its kind
                tri_round = want_tri && (want_node == 0 || !last_tri);
                last_tri = tri_round;
            }
            if (tri_round) {
                if (is_leaf) {
                    // the whole leaf run in one round, in order (tmax shrinks between candidates exactly as in
                    // intersect_leaf, single_ray_traverser.hpp:41-63); the run is contiguous, so after the first
                    // record the next ones mostly come from the same 128-byte line
#if VT_LEAF_RUN_PER_ROUND
                    bool any = false;
                    for (;;) {
                        if (STATS) n_tests++, ray_tests++;
                        any |= intersect_triangle<ALPHA, !ANY_HIT>(S, cur & VT_REF_MASK, r);
                        if ((ANY_HIT && any) || (cur >> VT_REF_SHIFT) == 1u) break;
                        cur -= VT_REF_MASK;  // count - 1, slot + 1
                    }
                    if (ANY_HIT && any) {
                        cur = VT_REF_DONE;
                        sp = stack;
                    } else {
                        cur = stack_pop<DIST>(sp, stack, r.tmax);
                    }
#else
                    if (STATS) n_tests++;
                    const bool hit = intersect_triangle<ALPHA, !ANY_HIT>(S, cur & VT_REF_MASK, r);
                    if (ANY_HIT && hit) {
                        cur = VT_REF_DONE;
                        sp = stack;
                    } else if ((cur >> VT_REF_SHIFT) == 1u) {  // run finished: pop
                        cur = stack_pop<DIST>(sp, stack, r.tmax);
                    } else {
                        cur -= VT_REF_MASK;  // count - 1, slot + 1
                    }
#endif
                }
            } else if (has && !is_leaf) {
                if (STATS) n_steps++, ray_steps++;
                if (QUAD) {
                    int k[4];
                    uint32_t cr[4];
                    if (warp_wild) slab_quad<true, KEYMID>(S.quads, cur, magic, r, k, cr, s_top_quads, n_top);
                    else slab_quad<false, KEYMID>(S.quads, cur, magic, r, k, cr, s_top_quads, n_top);
                    // farthest first, so the nearest pending child is popped first
                    if (k[3] != 0x7FFFFFFF) {
                        stack_push<DIST>(sp, cr[3], (uint32_t)k[3]);
                        if (VT_PREFETCH_ALL) vt_prefetch_ref(S, cr[3], true);
                    }
                    if (k[2] != 0x7FFFFFFF) {
                        stack_push<DIST>(sp, cr[2], (uint32_t)k[2]);
                        if (VT_PREFETCH_ALL) vt_prefetch_ref(S, cr[2], true);
                    }
                    if (k[1] != 0x7FFFFFFF) {
                        stack_push<DIST>(sp, cr[1], (uint32_t)k[1]);
                        if (VT_PREFETCH_FAR) vt_prefetch_ref(S, cr[1], true);  // the next one to be popped
                    }
                    if (k[0] != 0x7FFFFFFF) {
                        cur = cr[0];
                        if (VT_PREFETCH_LEAF && (cur >> VT_REF_SHIFT)) vt_prefetch(S.tris + (cur & VT_REF_MASK));
                    } else {
                        cur = stack_pop<DIST>(sp, stack, r.tmax);
                    }
                    continue;
                }
                float le, lx, re, rx;
                uint32_t lref, rref;
                slab_cpair(S.cpairs, cur, magic, r, le, lx, re, rx, lref, rref);
                const bool hit_l = le <= lx, hit_r = re <= rx;
                const bool take_r = hit_r && (!hit_l || le > re);  // near child first; ties keep the left child first
                const uint32_t next = take_r ? rref : lref;
                const uint32_t far_ = take_r ? lref : rref;
                if (hit_l && hit_r) stack_push<DIST>(sp, far_, __float_as_uint(take_r ? le : re));
                if (hit_l || hit_r) {
                    cur = next;
                } else {
                    cur = stack_pop<DIST>(sp, stack, r.tmax);
                }
            }
        }
        };
        if (TEAM && tail) run_rounds(std::true_type{});
        else run_rounds(std::false_type{});
    }
    if (n_invalid) atomicAdd(&counters[1], n_invalid);
    if (STATS) {
        atomicAdd(&counters[2], n_steps);
        atomicAdd(&counters[3], n_tests);
    }
}

}  // namespace

cudaError_t vt_launch_traverse(const VtSceneView &S, const vt_ray *rays, vt_hit *hits, uint64_t n, bool any_hit,
                               unsigned long long *counters, const VtLaunchConfig &cfg, cudaStream_t stream, bool stats,
                               const uint32_t *queue, const unsigned long long *queue_count) {
    if (n == 0) return cudaSuccess;
    if ((queue == nullptr) != (queue_count == nullptr)) return cudaErrorInvalidValue;
    // VT_K1_DYN_SMEM (bytes, tuning only): unused dynamic shared memory for the quantised kernels — shrinks the L1 carve-out so the
    // L1-capacity sensitivity of the kernel can be measured in isolation
    const char *dyn_smem_env = std::getenv("VT_K1_DYN_SMEM");
    const int dyn_smem_probe = (dyn_smem_env && *dyn_smem_env) ? std::atoi(dyn_smem_env) : 0;
    const size_t smem = (S.cpairs || S.quads) ? (size_t)dyn_smem_probe + (VT_SMEM_QUADS_BUILD && S.quads ? (size_t)S.n_smem_pairs * sizeof(VtQuad) : 0)
                                              : (size_t)S.n_smem_pairs * sizeof(VtPair);
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
                                                          cfg.persistent ? 1 : 0, cfg.refill_threshold, cfg.tri_threshold, queue, queue_count);
        return cudaGetLastError();
    };
    if (S.quads && cfg.key_mid && !VT_STACK_DIST) {  // children ordered by entry + exit (scenes whose sibling boxes overlap: VT_KEY_MID above)
        if (stats) {
            if (any_hit) return cudaErrorInvalidValue;
            return alpha ? launch(k_traverse_compact<false, true, true, true, false, true>) : launch(k_traverse_compact<false, false, true, true, false, true>);
        }
        if (any_hit) return alpha ? launch(k_traverse_compact<true, true, false, true, false, true>) : launch(k_traverse_compact<true, false, false, true, false, true>);
        if (cfg.tail_share) return alpha ? launch(k_traverse_compact<false, true, false, true, true, true>) : launch(k_traverse_compact<false, false, false, true, true, true>);
        return alpha ? launch(k_traverse_compact<false, true, false, true, false, true>) : launch(k_traverse_compact<false, false, false, true, false, true>);
    }
    if (S.quads) {
        if (stats) {  // closest hit only; counters[2] += traversal steps, counters[3] += triangle tests
            if (any_hit) return cudaErrorInvalidValue;
            return alpha ? launch(k_traverse_compact<false, true, true, true>) : launch(k_traverse_compact<false, false, true, true>);
        }
        if (any_hit) return alpha ? launch(k_traverse_compact<true, true, false, true>) : launch(k_traverse_compact<true, false, false, true>);
        if (cfg.tail_share) return alpha ? launch(k_traverse_compact<false, true, false, true, true>) : launch(k_traverse_compact<false, false, false, true, true>);
        return alpha ? launch(k_traverse_compact<false, true, false, true>) : launch(k_traverse_compact<false, false, false, true>);
    }
    if (S.cpairs) {
        if (stats) {
            if (any_hit) return cudaErrorInvalidValue;
            return alpha ? launch(k_traverse_compact<false, true, true, false>) : launch(k_traverse_compact<false, false, true, false>);
        }
        if (any_hit) return alpha ? launch(k_traverse_compact<true, true, false, false>) : launch(k_traverse_compact<true, false, false, false>);
        if (cfg.tail_share) return alpha ? launch(k_traverse_compact<false, true, false, false, true>) : launch(k_traverse_compact<false, false, false, false, true>);
        return alpha ? launch(k_traverse_compact<false, true, false, false>) : launch(k_traverse_compact<false, false, false, false>);
    }
    if (stats) {
        if (any_hit || S.n_smem_pairs) return cudaErrorInvalidValue;
        return alpha ? launch(k_traverse<false, true, false, true>) : launch(k_traverse<false, false, false, true>);
    }
    if (S.n_smem_pairs) {
        if (any_hit) return alpha ? launch(k_traverse<true, true, true>) : launch(k_traverse<true, false, true>);
        return alpha ? launch(k_traverse<false, true, true>) : launch(k_traverse<false, false, true>);
    }
    if (any_hit) return alpha ? launch(k_traverse<true, true, false>) : launch(k_traverse<true, false, false>);
    return alpha ? launch(k_traverse<false, true, false>) : launch(k_traverse<false, false, false>);
}

cudaError_t vt_traverse_occupancy(int *blocks_per_sm, size_t smem_bytes, int layout) {
    if (layout == VT_LAYOUT_QUAD)
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k_traverse_compact<false, true, false, true>, VT_TRAVERSE_BLOCK, 0);
    if (layout == VT_LAYOUT_COMPACT)
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k_traverse_compact<false, true, false, false>, VT_TRAVERSE_BLOCK, 0);
    cudaError_t e = cudaFuncSetAttribute(k_traverse<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k_traverse<false, true, true>, VT_TRAVERSE_BLOCK, smem_bytes);
}
