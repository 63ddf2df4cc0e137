// vt_refit.cu — K5: device-side refit of the resident QUAD hierarchy for moved geometry of unchanged topology.
//
// What `accel:Rebuild` (source/VisTrace.cpp:798-818) needs when props merely moved: the reference rebuilds from
// scratch (PopulateAccel); the bvh library it builds with offers bvh::HierarchyRefitter
// (libs/bvh/include/bvh/hierarchy_refitter.hpp:20-31: leaf boxes from their primitives, inner boxes = union of the
// children, bottom-up with per-node arrival flags, bottom_up_algorithm.hpp:52-80).  This file is that algorithm on
// the device, over the 4-wide quantised nodes the traversal kernel reads:
//
//   k_refit_prepare   once per resident hierarchy: parent[] and inner-child counts from the quad references,
//                     slot_of[] (original triangle -> leaf-order slot) from the triangle records.
//   k_refit_tris      per ORIGINAL triangle: the Triangle constructor + ComputeNormalAndLoD
//                     (source/objects/Primitives.h:75-102) on the caller's new vertices, written straight into the
//                     resident leaf-order VtTriRec / tri_uv and original-order VtTriAttr records.
//   k_refit_quads     bottom-up: a thread starts at every quad whose children are all leaves, computes the exact child
//                     boxes (Triangle::bounding_box, Primitives.h:107-113, over p0, p0 - e1, p0 + e2), re-derives the
//                     quad's power-of-two grid and conservative plane bytes exactly as the host's build_quads does
//                     (vt_bvh_build.cpp: choose_grid), stores the quad's exact union box, and walks up: the LAST
//                     child to arrive at a parent (atomic arrival counter) processes it.
//
// HBM-bound streaming: 152 B in + 264 B out per triangle, 64 B in/out + 24 B per quad.  Arithmetic is IEEE, uncontracted
// (--fmad=false), so p0/e1/e2/n/nNorm equal the host constructor's bit for bit; `lod` goes through log2f, whose device
// and glibc implementations may differ in the last place (it only selects a texture LOD; parity bar 1e-5 relative).
#include <cfloat>

#include "vt_kernels.h"

namespace {

#define VT_REF_SHIFT 28
#define VT_REF_MASK 0x0FFFFFFFu

__global__ void k_refit_prepare(const VtQuad *__restrict__ quads, uint32_t n_quads, const VtTriRec *__restrict__ tris, uint32_t n_tris,
                                uint32_t *__restrict__ parent, uint32_t *__restrict__ n_inner, uint32_t *__restrict__ slot_of,
                                uint32_t *__restrict__ leaf_quad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_quads) {
        uint32_t inner = 0;
        for (int c = 0; c < 4; c++) {
            const uint32_t ref = quads[i].ref[c];
            if (ref == 0xFFFFFFFFu) continue;
            const uint32_t count = ref >> VT_REF_SHIFT, idx = ref & VT_REF_MASK;
            if (count == 0) {
                parent[ref] = i;
                inner++;
            } else if (idx < n_tris) {  // a real leaf (not the sentinel of an empty slot): its slots hang off this quad
                for (uint32_t t = 0; t < count; t++) leaf_quad[idx + t] = i;
            }
        }
        n_inner[i] = inner;
        if (i == 0) parent[0] = 0xFFFFFFFFu;
    }
    if (i < n_tris) slot_of[tris[i].orig] = i;
}

// in[j] holds the new vertices of ORIGINAL triangle first + j, j < count.  A block of 128 threads handles 128 consecutive
// triangles: their 152-byte input records and their 176-byte attribute records are CONTIGUOUS in global memory, so both go through
// shared memory with 16-byte coalesced accesses (a thread reading / writing its own record walks a 152- / 176-byte stride: 2.6 TB/s,
// profiles/r1). Only the 64-byte geometry record and the 24 bytes of UVs go to the triangle's leaf slot, which is scattered by nature.
constexpr int kRefitBlock = 128;
__global__ void __launch_bounds__(kRefitBlock)
k_refit_tris(const vt_tri_in *__restrict__ in, uint32_t first, uint32_t count, const uint32_t *__restrict__ slot_of,
             const VtDevMaterial *__restrict__ mats, VtTriRec *__restrict__ recs, float *__restrict__ tri_uv,
             VtTriAttr *__restrict__ attrs) {
    static_assert(sizeof(vt_tri_in) % 8 == 0 && sizeof(VtTriAttr) % 16 == 0, "staging moves 8- / 16-byte words");
    __shared__ __align__(16) unsigned char s_in[kRefitBlock * sizeof(vt_tri_in)];
    __shared__ __align__(16) unsigned char s_attr[kRefitBlock * sizeof(VtTriAttr)];
    const uint32_t j0 = blockIdx.x * kRefitBlock;
    const uint32_t m = min((uint32_t)kRefitBlock, count - j0);
    {   // vt_tri_in is 152 bytes (8-byte multiples; the array base is at least 8-byte aligned): coalesced 8-byte loads
        const uint2 *src = reinterpret_cast<const uint2 *>(in + j0);
        uint2 *dst = reinterpret_cast<uint2 *>(s_in);
        for (uint32_t w = threadIdx.x; w < m * (uint32_t)(sizeof(vt_tri_in) / 8); w += kRefitBlock) dst[w] = __ldg(src + w);
    }
    __syncthreads();
    const uint32_t j = j0 + threadIdx.x;
    if (threadIdx.x < m) {
        const uint32_t i = first + j;
        const vt_tri_in &t = *reinterpret_cast<const vt_tri_in *>(s_in + threadIdx.x * sizeof(vt_tri_in));
        float p0[3], e1[3], e2[3], n[3], nn[3];
        for (int k = 0; k < 3; k++) {
            p0[k] = t.p[0][k];
            e1[k] = t.p[0][k] - t.p[1][k];  // e1 = p0 - p1, e2 = p2 - p0 (Primitives.h:82)
            e2[k] = t.p[2][k] - t.p[0][k];
        }
        n[0] = e1[1] * e2[2] - e1[2] * e2[1];  // ComputeNormalAndLoD, Primitives.h:91-102
        n[1] = e1[2] * e2[0] - e1[0] * e2[2];
        n[2] = e1[0] * e2[1] - e1[1] * e2[0];
        const float uv10x = t.uvs[1][0] - t.uvs[0][0], uv10y = t.uvs[1][1] - t.uvs[0][1];
        const float uv20x = t.uvs[2][0] - t.uvs[0][0], uv20y = t.uvs[2][1] - t.uvs[0][1];
        const float area = fabsf(uv10x * uv20y - uv20x * uv10y);
        float d = n[0] * n[0];
        d += n[1] * n[1];
        d += n[2] * n[2];
        const float len = sqrtf(d);
        const float lod = 0.5f * log2f(area / len);
        for (int k = 0; k < 3; k++) nn[k] = n[k] / len;

        const uint32_t s = slot_of[i];
        VtTriRec r;
        for (int k = 0; k < 3; k++) r.p0[k] = p0[k], r.e1[k] = e1[k], r.e2[k] = e2[k], r.n[k] = n[k];
        const uint32_t mflags = mats[t.material].flags;
        uint32_t fl = 0;
        if (t.one_sided && (mflags & VT_MATFLAG_NOCULL) == 0) fl |= VT_TRI_FLAG_CULL;  // Primitives.h:174
        if (mflags & VT_MATFLAG_ALPHATEST) fl |= VT_TRI_FLAG_ALPHATEST;               // Primitives.h:195
        r.matflags = (t.material << 2) | fl;
        r.orig = i;
        r.pad[0] = r.pad[1] = 0;
        recs[s] = r;
        for (int k = 0; k < 6; k++) tri_uv[(size_t)s * 6 + k] = (&t.uvs[0][0])[k];
        VtTriAttr &a = *reinterpret_cast<VtTriAttr *>(s_attr + threadIdx.x * sizeof(VtTriAttr));
        for (int k = 0; k < 3; k++) a.p0[k] = p0[k], a.e1[k] = e1[k], a.e2[k] = e2[k], a.nNorm[k] = nn[k], a.alphas[k] = t.alphas[k];
        for (int k = 0; k < 9; k++) (&a.normals[0][0])[k] = (&t.normals[0][0])[k], (&a.tangents[0][0])[k] = (&t.tangents[0][0])[k];
        for (int k = 0; k < 6; k++) (&a.uvs[0][0])[k] = (&t.uvs[0][0])[k];
        a.lod = lod;
        a.material = t.material;
        a.ent_idx = t.ent_idx;
    }
    __syncthreads();
    {   // VtTriAttr is 176 bytes, 16-byte aligned: coalesced 16-byte stores of the block's contiguous output range
        const uint4 *src = reinterpret_cast<const uint4 *>(s_attr);
        uint4 *dst = reinterpret_cast<uint4 *>(attrs + first + j0);
        for (uint32_t w = threadIdx.x; w < m * (uint32_t)(sizeof(VtTriAttr) / 16); w += kRefitBlock) dst[w] = src[w];
    }
}

// choose_grid of vt_bvh_build.cpp: the smallest power-of-two cell 2^E on which [lo, hi] spans <= 255 cells from
// k = floor(lo / 2^E) with |k| small enough that (k - OFFSET) * 2^E and (k + q) * 2^E are exact floats; |E| <= 60.
__device__ bool choose_grid_dev(double lo, double hi, int &E, long long &k) {
    const double kmax = 8388608.0 - 512.0;
    const double mag = fmax(fabs(lo), fabs(hi));
    int e = -149;
    if (hi > lo) e = max(e, (int)ceil(log2((hi - lo) / 255.0)) - 1);
    if (mag > 0) e = max(e, (int)floor(log2(mag)) - 23);
    e = max(e, -60);
    for (; e <= 60; e++) {
        const double s = ldexp(1.0, e);
        const double kl = floor(lo / s), kh = ceil(hi / s);
        if (kh - kl <= 255.0 && fabs(kl) <= kmax && fabs(kh) <= kmax) {
            E = e;
            k = (long long)kl;
            return true;
        }
    }
    return false;
}

struct Box6 {
    float lo[3], hi[3];
};

// One quad: exact child boxes (leaf children from their triangles, inner children from qbox), grid + plane bytes re-derived, the
// quad and its exact union box stored.  false: the box cannot be put on a float grid (the caller raises the error flag).
__device__ bool refit_one_quad(VtQuad *quads, uint32_t q, const VtTriRec *__restrict__ tris, uint32_t n_tris, Box6 *qbox) {
    VtQuad node = quads[q];
    Box6 cb[4], un;
    bool used[4];
    for (int a = 0; a < 3; a++) un.lo[a] = FLT_MAX, un.hi[a] = -FLT_MAX;
    for (int c = 0; c < 4; c++) {
        const uint32_t ref = node.ref[c];
        const uint32_t count = ref >> VT_REF_SHIFT, idx = ref & VT_REF_MASK;
        used[c] = ref != 0xFFFFFFFFu && !(count != 0 && idx >= n_tris);  // 0xFFFFFFFF or the sentinel leaf: empty slot
        if (!used[c]) continue;
        Box6 b;
        if (count == 0) {
            // written by the thread that finished that child before it signalled arrival (threadfence + atomic in the callers),
            // or — ranged walk, untouched child — by an earlier pass
            const volatile Box6 *src = qbox + idx;
            for (int a = 0; a < 3; a++) b.lo[a] = src->lo[a], b.hi[a] = src->hi[a];
        } else {
            for (int a = 0; a < 3; a++) b.lo[a] = FLT_MAX, b.hi[a] = -FLT_MAX;
            for (uint32_t t = 0; t < count; t++) {
                const VtTriRec &tr = tris[idx + t];
                for (int a = 0; a < 3; a++) {
                    const float v0 = tr.p0[a], v1 = tr.p0[a] - tr.e1[a], v2 = tr.p0[a] + tr.e2[a];
                    b.lo[a] = fminf(b.lo[a], fminf(v0, fminf(v1, v2)));
                    b.hi[a] = fmaxf(b.hi[a], fmaxf(v0, fmaxf(v1, v2)));
                }
            }
        }
        cb[c] = b;
        for (int a = 0; a < 3; a++) un.lo[a] = fminf(un.lo[a], b.lo[a]), un.hi[a] = fmaxf(un.hi[a], b.hi[a]);
    }
    for (int a = 0; a < 3; a++) {
        int E = 0;
        long long k = 0;
        if (!(isfinite(un.lo[a]) && isfinite(un.hi[a])) || un.hi[a] < un.lo[a] || !choose_grid_dev(un.lo[a], un.hi[a], E, k)) return false;
        const double s = ldexp(1.0, E);
        node.origin_adj[a] = (float)((double)(k - VT_QUAD_OFFSET) * s);
        node.scale[a] = (float)s;
        for (int c = 0; c < 4; c++) {
            if (used[c]) {
                node.q[a][0][c] = (uint8_t)((long long)floor((double)cb[c].lo[a] / s) - k);
                node.q[a][1][c] = (uint8_t)((long long)ceil((double)cb[c].hi[a] / s) - k);
            } else {
                node.q[a][0][c] = 255;  // empty slot: inverted box
                node.q[a][1][c] = 0;
            }
        }
    }
    quads[q] = node;
    qbox[q] = un;
    return true;
}

__global__ void k_refit_quads(VtQuad *quads, uint32_t n_quads, const VtTriRec *__restrict__ tris, uint32_t n_tris,
                              const uint32_t *__restrict__ parent, const uint32_t *__restrict__ n_inner, uint32_t *arrive,
                              Box6 *qbox, unsigned int *error) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_quads || n_inner[q] != 0) return;  // start at the quads all of whose children are leaves
    for (;;) {
        if (!refit_one_quad(quads, q, tris, n_tris, qbox)) {
            atomicExch(error, 1u);  // non-finite geometry or coordinates out of float grid range: the host re-derives the layout
            return;
        }
        const uint32_t p = parent[q];
        if (p == 0xFFFFFFFFu) return;  // the root
        __threadfence();               // box and node visible before the arrival is counted
        if (atomicAdd(&arrive[p], 1u) + 1u < n_inner[p]) return;  // a sibling subtree is still being refitted
        __threadfence();
        q = p;
    }
}

// Ranged refit (vt_accel_refit_range: one entity moved): only the quads that hold a touched triangle and their ANCESTORS change;
// every other quad keeps its bytes and its qbox entry from the previous pass.  Three kernels, one thread per touched triangle, over
// per-quad state that is all-zero between calls (stamp[] compares against a per-call epoch E, E += 3 per call):
//   k_refit_mark_range   walks from the triangle's quad to the root stamping E; the first thread to stamp a quad counts it as a dirty
//                        child of its parent and goes on, later ones stop there (the first carries on upwards).
//   k_refit_quads_range  claims the triangle's quad (E -> E + 1: one winner per quad); a quad without dirty inner children is a
//                        starting point, the others are processed by their last dirty child to arrive — the bottom-up protocol of
//                        k_refit_quads with kids[] in the place of n_inner.  kids[] is read-only here: a quad that holds touched
//                        triangles AND dirty children must see the same count whenever its claimant looks.
//   k_refit_clear_range  walks the same chains stamping E + 2 and zeroes kids[] / arrive[] (also after an aborted pass).
// A one-entity range at 1.07 M quads touches a few thousand quads instead of all of them.
__global__ void k_refit_mark_range(uint32_t first, uint32_t count, const uint32_t *__restrict__ slot_of, const uint32_t *__restrict__ leaf_quad,
                                   const uint32_t *__restrict__ parent, uint32_t *stamp, uint32_t *kids, uint32_t epoch) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    uint32_t q = leaf_quad[slot_of[first + j]];
    for (;;) {
        if (atomicExch(&stamp[q], epoch) == epoch) return;
        const uint32_t p = parent[q];
        if (p == 0xFFFFFFFFu) return;
        atomicAdd(&kids[p], 1u);
        q = p;
    }
}

__global__ void k_refit_quads_range(VtQuad *quads, const VtTriRec *__restrict__ tris, uint32_t n_tris, uint32_t first, uint32_t count,
                                    const uint32_t *__restrict__ slot_of, const uint32_t *__restrict__ leaf_quad,
                                    const uint32_t *__restrict__ parent, uint32_t *stamp, const uint32_t *kids, uint32_t *arrive,
                                    Box6 *qbox, uint32_t epoch, unsigned int *error) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    uint32_t q = leaf_quad[slot_of[first + j]];
    if (atomicCAS(&stamp[q], epoch, epoch + 1u) != epoch) return;  // another touched triangle of the same quad got here first
    if (kids[q] != 0) return;                                      // its last dirty child processes it
    for (;;) {
        if (!refit_one_quad(quads, q, tris, n_tris, qbox)) {
            atomicExch(error, 1u);
            return;
        }
        const uint32_t p = parent[q];
        if (p == 0xFFFFFFFFu) return;
        __threadfence();
        if (atomicAdd(&arrive[p], 1u) + 1u < kids[p]) return;
        __threadfence();
        q = p;
    }
}

__global__ void k_refit_clear_range(uint32_t first, uint32_t count, const uint32_t *__restrict__ slot_of, const uint32_t *__restrict__ leaf_quad,
                                    const uint32_t *__restrict__ parent, uint32_t *stamp, uint32_t *kids, uint32_t *arrive, uint32_t epoch) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    uint32_t q = leaf_quad[slot_of[first + j]];
    while (q != 0xFFFFFFFFu && atomicExch(&stamp[q], epoch + 2u) != epoch + 2u) {
        kids[q] = 0, arrive[q] = 0;
        q = parent[q];
    }
}

// Sum of the half surface areas of all quad boxes: the SAH's inner-node term of the resident hierarchy (the expected number of
// node visits of a random ray is proportional to it, libs/bvh/include/bvh/sah_based_algorithm.hpp:16-41).  Compared with the value
// right after the build it tells how much a sequence of refits has loosened the tree — the rebuild trigger of vt_accel_refit.
__global__ void k_refit_cost(const Box6 *__restrict__ qbox, uint32_t n_quads, double *__restrict__ sum) {
    double local = 0.0;
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n_quads; q += gridDim.x * blockDim.x) {
        const Box6 b = qbox[q];
        const double dx = (double)b.hi[0] - b.lo[0], dy = (double)b.hi[1] - b.lo[1], dz = (double)b.hi[2] - b.lo[2];
        local += dx * dy + dy * dz + dz * dx;
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local != 0.0) atomicAdd(sum, local);
}

}  // namespace

cudaError_t vt_launch_refit_cost(const void *qbox, uint32_t n_quads, double *sum, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(sum, 0, sizeof(double), stream);
    if (e != cudaSuccess || n_quads == 0) return e;
    k_refit_cost<<<296, 256, 0, stream>>>(static_cast<const Box6 *>(qbox), n_quads, sum);
    return cudaGetLastError();
}

cudaError_t vt_launch_refit_prepare(const VtSceneView &S, uint32_t *parent, uint32_t *n_inner, uint32_t *slot_of, uint32_t *leaf_quad,
                                    cudaStream_t stream) {
    const uint32_t n = S.n_pairs > S.n_tris ? S.n_pairs : S.n_tris;
    if (n == 0) return cudaSuccess;
    k_refit_prepare<<<(n + 255) / 256, 256, 0, stream>>>(S.quads, S.n_pairs, S.tris, S.n_tris, parent, n_inner, slot_of, leaf_quad);
    return cudaGetLastError();
}

cudaError_t vt_launch_refit_tris(const VtSceneView &S, const vt_tri_in *in, uint32_t first, uint32_t count, const uint32_t *slot_of,
                                 cudaStream_t stream) {
    if (count == 0) return cudaSuccess;
    if ((uint64_t)first + count > S.n_tris) return cudaErrorInvalidValue;
    if (((uintptr_t)in & 7) != 0) return cudaErrorInvalidValue;  // cudaMalloc'd staging: always true
    k_refit_tris<<<(count + kRefitBlock - 1) / kRefitBlock, kRefitBlock, 0, stream>>>(in, first, count, slot_of, S.mats, const_cast<VtTriRec *>(S.tris),
                                                          const_cast<float *>(S.tri_uv), const_cast<VtTriAttr *>(S.attrs));
    return cudaGetLastError();
}

cudaError_t vt_launch_refit_quads(const VtSceneView &S, const uint32_t *parent, const uint32_t *n_inner, uint32_t *arrive, void *qbox,
                                  unsigned int *error, cudaStream_t stream) {
    if (S.n_pairs == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(arrive, 0, (size_t)S.n_pairs * sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    k_refit_quads<<<(S.n_pairs + 127) / 128, 128, 0, stream>>>(const_cast<VtQuad *>(S.quads), S.n_pairs, S.tris, S.n_tris, parent, n_inner,
                                                               arrive, static_cast<Box6 *>(qbox), error);
    return cudaGetLastError();
}

cudaError_t vt_launch_refit_quads_range(const VtSceneView &S, uint32_t first, uint32_t count, const uint32_t *slot_of, const uint32_t *leaf_quad,
                                        const uint32_t *parent, uint32_t *stamp, uint32_t *kids, uint32_t *arrive, void *qbox, uint32_t epoch,
                                        unsigned int *error, cudaStream_t stream) {
    if (S.n_pairs == 0 || count == 0) return cudaSuccess;
    if ((uint64_t)first + count > S.n_tris || epoch == 0 || epoch > 0xFFFFFFF0u) return cudaErrorInvalidValue;
    const uint32_t blocks = (count + 127) / 128;
    k_refit_mark_range<<<blocks, 128, 0, stream>>>(first, count, slot_of, leaf_quad, parent, stamp, kids, epoch);
    k_refit_quads_range<<<blocks, 128, 0, stream>>>(const_cast<VtQuad *>(S.quads), S.tris, S.n_tris, first, count, slot_of, leaf_quad, parent, stamp,
                                                    kids, arrive, static_cast<Box6 *>(qbox), epoch, error);
    k_refit_clear_range<<<blocks, 128, 0, stream>>>(first, count, slot_of, leaf_quad, parent, stamp, kids, arrive, epoch);
    return cudaGetLastError();
}
