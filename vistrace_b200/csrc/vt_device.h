// vt_device.h — HBM-resident layouts shared by the host flattening code and the kernels.
//
// Everything a ray touches lives in four flat arrays:
//
//   pairs   64 B / sibling pair   the two children of an inner node, i.e. the two adjacent
//                                 32-byte bvh::Bvh<float>::Node records the reference reads per
//                                 traversal step (libs/bvh/include/bvh/bvh.hpp:25-31,
//                                 single_ray_traverser.hpp:85-87), 64-byte aligned so one step is
//                                 four 128-bit loads from one or two 32-byte sectors of ONE line.
//   tris    64 B / triangle       p0, e1, e2, n + packed {material, alphatest, cull} + original index,
//                                 stored in LEAF ORDER (the permutation primitive_indices encodes,
//                                 primitive_intersectors.hpp:17-20) so a leaf is one contiguous run.
//   tri_uv  24 B / triangle       leaf-order UVs, read only by the alpha test (Primitives.h:198).
//   attrs  176 B / triangle       ORIGINAL order; everything the TraceResult stage interpolates.
#pragma once
#include <stdint.h>

#define VT_STACK_SIZE 64          // libs/bvh/include/bvh/single_ray_traverser.hpp:14
#define VT_TRI_FLAG_CULL 1u       // oneSided && !(mat.flags & nocull)   (Primitives.h:174)
#define VT_TRI_FLAG_ALPHATEST 2u  // mat.flags & alphatest               (Primitives.h:195)
#define VT_LEAF_BIT 0x80000000u
// Quad layout: empty child slots reference a one-triangle leaf whose record (slot n_tris, behind the last real
// triangle) is all NaN and can never be hit, instead of carrying a "slot in use" test through every node step.
#ifndef VT_EMPTY_SENTINEL
#define VT_EMPTY_SENTINEL 1
#endif
// Quad layout: the kernel forms the float OFFSET + q from a quantised byte q and decodes a plane as
// fmaf(OFFSET + q, 2^E, origin_adj) with origin_adj = (k - OFFSET) * 2^E.  OFFSET = 2^23 when the float is built by
// OR-ing q into the mantissa of 0x4B000000, 1024 when it comes from the fp16 0x6400 | q (VT_DECODE_HALF, vt_traverse.cu).
#ifndef VT_DECODE_HALF
#define VT_DECODE_HALF 1
#endif
#define VT_QUAD_OFFSET (VT_DECODE_HALF ? 1024 : 8388608)

// One child inside a pair: bounds in bvh order {minx,maxx,miny,maxy,minz,maxz}; count != 0 marks a
// leaf.  `first` = pair index of the child's own children (inner) or first slot in `tris` (leaf).
struct VtChild {
    float bounds[6];
    uint32_t count;
    uint32_t first;
};
struct alignas(64) VtPair {
    VtChild l, r;
};
static_assert(sizeof(VtPair) == 64, "pair layout");

// Compact sibling pair: the same two children in ONE 32-byte sector (one LDG.E.256 per traversal
// step instead of two).  The boxes are quantised to 8 bits per plane on a per-pair power-of-two grid
// and are CONSERVATIVE: plane' = (k + q) * 2^E with lo' <= lo and hi' >= hi, decoded exactly (no
// rounding) by ONE fma:  fmaf(as_float(0x4B000000 | q), 2^E, origin_adj),  origin_adj = (k - 2^23) * 2^E.
// Because fmaf(plane, inv_dir, scaled_origin) is monotone in `plane`, every node the reference's
// FastNodeIntersector accepts (node_intersectors.hpp:35-47) is accepted here too; only the ORDER of
// equal-distance visits can differ, i.e. which of two exactly tied candidates wins.
// Needs the depth-first pair order of flatten_bvh (bfs_pairs = 0) and leaves of <= 15 triangles:
//   left child inner  -> its pair is cur + 1, `ref` belongs to the right child (pair index or triangle slot)
//   left child leaf   -> `ref` is its triangle slot; right child = pair cur + 1 (inner) or slot ref + lcount (leaf)
struct alignas(32) VtCPair {
    float origin_adj[3];  // (k - 2^23) * 2^E per axis
    uint8_t exp[3];       // biased exponent of 2^E per axis
    uint8_t counts;       // lcount | rcount << 4; 0 = inner child
    uint8_t q[3][4];      // per axis: {l.lo, l.hi, r.lo, r.hi}
    uint32_t ref;
};
static_assert(sizeof(VtCPair) == 32, "compact pair layout");

// Quad node: a 4-wide hierarchy obtained by collapsing the binary tree (each node adopts its
// grandchildren, largest box first), so a ray needs about half as many dependent node fetches.  Same
// conservative power-of-two grid as VtCPair, shared by the four children; 64 bytes = two LDG.E.256.
// Children are TAGGED references: bits 28-31 = triangle count (0 = inner quad), bits 0-27 = quad index
// or first triangle slot; empty slots have ref 0xFFFFFFFF.
struct alignas(64) VtQuad {
    float origin_adj[3];  // (k - VT_QUAD_OFFSET) * 2^E per axis
    float scale[3];       // 2^E per axis, as a float
    uint8_t q[3][2][4];   // [axis][lo, hi][child]
    uint32_t ref[4];      // tagged child references; 0xFFFFFFFF = empty slot (its box is inverted: q_lo = 255, q_hi = 0)
};
static_assert(sizeof(VtQuad) == 64, "quad layout");

struct alignas(64) VtTriRec {
    float p0[3];
    float e1[3];
    float e2[3];
    float n[3];         // cross(e1, e2) exactly as the Triangle constructor rounded it (Primitives.h:93)
    uint32_t matflags;  // (material << 2) | VT_TRI_FLAG_*
    uint32_t orig;      // index into the caller's triangle array
    uint32_t pad[2];
};
static_assert(sizeof(VtTriRec) == 64, "triangle layout");

// Per ORIGINAL triangle: inputs of TraceResult::TraceResult (source/objects/TraceResult.cpp:45-86).
struct alignas(16) VtTriAttr {
    float p0[3], e1[3], e2[3];  // v0 = p0, v1 = p0 - e1, v2 = p0 + e2
    float nNorm[3];             // geometricNormal
    float normals[3][3];
    float tangents[3][3];
    float uvs[3][2];
    float alphas[3];
    float lod;
    uint32_t material;
    uint32_t ent_idx;
};
static_assert(sizeof(VtTriAttr) == 176, "attr layout");

// Texture header; texels of all textures live in one byte buffer, each chain in VTF order
// (smallest mip first).  mip_offset[m] = byte offset of mip m inside the chain, precomputed once —
// the loop the reference runs per sample (libs/VTFParser/VTFParser.cpp:219-229, "TODO: Cache these").
struct VtDevTexture {
    uint32_t width, height, mips, flags;
    uint64_t base;  // byte offset of the chain inside the texel buffer
    uint32_t mip_offset[16];  // TEXEL index of mip m inside the chain
    uint32_t layout;  // vt_texture.texel_layout: 0 = RGBA8888 (4 bytes per texel), else wide (8 bytes: four uint16 numerators + divisor codes)
    uint32_t pad;
};

struct VtDevMaterial {
    uint32_t flags, surf_flags;
    float alphatest_reference, tex_scale;
    float colour[4];
    float base_tex_mat[8], base_tex_mat2[8], normal_map_mat[8], normal_map_mat2[8], blend_tex_mat[8], detail_mat[8];
    float detail_scale, detail_blend_factor;
    int32_t base_texture, base_texture2, normal_map, normal_map2, mrao, mrao2, blend_texture, detail;
    uint32_t detail_blend_mode, masked_blending, water, pad;
};

struct VtDevEntity {
    uint32_t id;
    float colour[4];
};

// Kernel argument block (passed by value).
struct VtSceneView {
    const VtPair *pairs;     // exact layout (nullptr when the compact layout is resident)
    const VtCPair *cpairs;   // compact layout (nullptr when the exact layout is resident)
    const VtQuad *quads;     // quad layout (nullptr unless resident)
    const VtTriRec *tris;
    const float *tri_uv;  // 6 floats per leaf-order triangle
    const VtTriAttr *attrs;
    const VtDevMaterial *mats;
    const VtDevEntity *ents;
    const VtDevTexture *texs;
    const uint8_t *texels;
    uint32_t n_pairs;
    uint32_t n_tris;
    uint32_t root_leaf_count;  // != 0: the root itself is a leaf over tris[0, count)  (single_ray_traverser.hpp:72-73)
    uint32_t n_smem_pairs;     // leading pairs staged in shared memory by the traversal kernel
    uint32_t has_alphatest;    // any triangle carries VT_TRI_FLAG_ALPHATEST
    uint32_t fallback_tex;     // index of the 1x1 white stand-in for a null baseTexture
    uint32_t magic;            // 0x4B000000 (2^23 as float bits), a constant-bank operand of the compact decode
    uint32_t magic_h;          // 0x64646464 (fp16 1024 = 0x6400): the half2 decode of the quad planes (VT_DECODE_HALF)
};
