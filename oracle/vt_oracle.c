/*
 * oracle/vt_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's accel:Traverse hot path
 * (Derpius/VisTrace @06ba9ee).  Every function cites the reference file:line it
 * follows; paths are relative to the reference tree.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load the library built from this file (oracle/libvt_oracle.so); the product
 * (vistrace_b200/csrc) never links, loads or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement bit-for-bit
 * against (a) the UNMODIFIED reference compiled from /root/reference into
 * oracle/_ref/libvt_ref.so, (b) the golden vectors under tests/golden/ that were
 * generated from that reference build (tests/golden/make_golden.py), and (c) the
 * known-answer tests the vendored bvh library ships
 * (libs/bvh/test/node_intersectors.cpp:18-36, libs/bvh/test/simple_example.cpp:63-83).
 *
 * Build: -mfma -ffp-contract=off (oracle/Makefile): fmaf() below is a real fused
 * multiply-add exactly where the reference uses fast_multiply_add with FP_FAST_FMAF
 * defined, and nothing else is contracted.
 *
 * Not restated: hierarchy CONSTRUCTION (PLOC + LeafCollapser,
 * source/objects/AccelStruct.cpp:762-770) — construction is host-side and outside
 * the traversal path; the port traverses whatever bvh::Bvh<float>-form hierarchy it
 * is handed through vto_set_bvh (the product builder's or the reference's).
 */
#include <float.h>
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../include/vistrace_b200.h"

/* ------------------------------------------------------------------ containers */

/* TriangleBackfaceCull<float> (source/objects/Primitives.h:43-73) */
typedef struct {
    float p0[3], e1[3], e2[3], n[3], nNorm[3];
    float lod;
    float normals[3][3], tangents[3][3], uvs[3][2], alphas[3];
    uint32_t material;
    uint16_t entIdx;
    uint8_t oneSided;
} OTri;

/* VTFTexture after load: RGBA8888 chain, smallest mip first (libs/VTFParser/VTFParser.cpp:13-93) */
typedef struct {
    uint16_t width, height, mips;
    uint32_t flags;
    uint32_t layout; /* vt_texture.texel_layout: 0 = RGBA8888, else wide texels */
    uint8_t *data;
} OTex;

typedef struct {
    OTri *tris;
    uint64_t n_tris;
    vt_material *mats;
    uint32_t n_mats;
    vt_entity *ents;
    uint32_t n_ents;
    OTex *texs;
    uint32_t n_texs; /* caller textures + 1 fallback (see vto_create) */
    vt_node *nodes;
    uint64_t node_count;
    uint64_t *prim_indices;
    int built;
} OScene;

/* ------------------------------------------------------------ bvh vector math */

/* bvh::dot (libs/bvh/include/bvh/vector.hpp:134-141): sum = a0*b0; sum += a1*b1; sum += a2*b2 */
static inline float bdot(const float a[3], const float b[3]) {
    float s = a[0] * b[0];
    s += a[1] * b[1];
    s += a[2] * b[2];
    return s;
}
/* bvh::cross (vector.hpp:159-167): r[i] = a[j]*b[k] - a[k]*b[j], j=(i+1)%3, k=(i+2)%3 */
static inline void bcross(const float a[3], const float b[3], float r[3]) {
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
}

/* ------------------------------------------------------------------- glm math */
typedef struct { float x, y, z; } v3;
typedef struct { float x, y; } v2;
typedef struct { float r, g, b, a; } px4;

static inline v3 V3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 v3add(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3sub(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3mul(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 v3s(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
/* glm::dot(vec3) (libs/glm/glm/detail/func_geometric.inl:48-54): tmp = a*b; tmp.x + tmp.y + tmp.z */
static inline float gdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/* glm::cross (func_geometric.inl:68-78) */
static inline v3 gcross(v3 x, v3 y) { return V3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
/* glm::normalize = v * inversesqrt(dot(v,v)), inversesqrt = 1/sqrt (func_geometric.inl:88, func_exponential.inl:138) */
static inline v3 gnormalize(v3 v) { return v3s(v, 1.0f / sqrtf(gdot(v, v))); }
/* glm::mix / gtx lerp (gtx/compatibility.hpp:41-48, detail/func_common.inl:104-111): x*(1-a) + y*a */
static inline float lerp1(float x, float y, float a) { return x * (1.0f - a) + y * a; }
static inline v3 lerp3(v3 x, v3 y, float a) { return v3add(v3s(x, 1.0f - a), v3s(y, a)); }
/* glm::min/max/clamp (func_common.inl): min(a,b) = b<a ? b : a; max(a,b) = a<b ? b : a; clamp = min(max(x,lo),hi) */
static inline float gmin(float a, float b) { return (b < a) ? b : a; }
static inline float gmax(float a, float b) { return (a < b) ? b : a; }
static inline float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
static inline float saturate(float x) { return gclamp(x, 0.0f, 1.0f); } /* gtx/compatibility.hpp:50-53 */
/* glm::smoothstep (func_common.inl:564-570) */
static inline float gsmoothstep(float e0, float e1, float x) {
    float t = gclamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}

/* --------------------------------------------------------------- texture path */

/* inline intmod (libs/VTFParser/VTFParser.cpp:9-11) */
static inline int intmod(int a, int b) { return (a % b + b) % b; }
static inline float fclampf(float v, float lo, float hi) { return (v < lo) ? lo : (hi < v) ? hi : v; } /* std::clamp */
static inline int iclamp(int v, int lo, int hi) { return (v < lo) ? lo : (hi < v) ? hi : v; }

/* VTFParser::ParsePixel, RGBA8888 case (libs/VTFParser/FileFormat/Parser.cpp:161-168) */
static inline px4 parse_pixel(const uint8_t *p) {
    px4 r = {p[0] / 255.f, p[1] / 255.f, p[2] / 255.f, p[3] / 255.f};
    return r;
}
/* The 16-bit formats (Parser.cpp:190-196,238-262,281-294) as vt_texture's wide texels hold them: four uint16 numerators over the
 * divisor each channel's 2-bit code names (include/vistrace_b200.h: VT_TEXEL_WIDE) */
static inline float wide_divisor(uint32_t layout, int c) {
    const uint32_t code = (layout >> (2 * c)) & 3u;
    return code == VT_TEXEL_DIV_255 ? 255.f : (code == VT_TEXEL_DIV_65535 ? 65535.f : 1.f);
}
static inline px4 parse_pixel_wide(const uint8_t *p, uint32_t layout) {
    uint16_t n[4];
    memcpy(n, p, 8);
    px4 r = {(float)n[0] / wide_divisor(layout, 0), (float)n[1] / wide_divisor(layout, 1), (float)n[2] / wide_divisor(layout, 2),
             (float)n[3] / wide_divisor(layout, 3)};
    return r;
}

/* VTFTexture::SampleBilinear (libs/VTFParser/VTFParser.cpp:207-309), z = frame = face = 0, depth 1 */
static px4 sample_bilinear(const OTex *t, float u, float v, uint8_t mipLevel) {
    uint32_t offset = 0;
    uint16_t width = t->width >> mipLevel, height = t->height >> mipLevel;
    for (uint8_t i = mipLevel + 1; i < t->mips; i++) { /* :219-229 mip offset = sizes of all smaller mips */
        width >>= 1;
        height >>= 1;
        if (width < 1) width = 1;
        if (height < 1) height = 1;
        offset += (uint32_t)width * height * (t->layout ? 8u : 4u);
    }
    width = t->width >> mipLevel;
    height = t->height >> mipLevel;
    if (width < 1) width = 1;
    if (height < 1) height = 1;
    const uint32_t pixelSize = t->layout ? 8 : 4;
    int clampX = (t->flags & VT_TEXFLAG_CLAMPS) != 0, clampY = (t->flags & VT_TEXFLAG_CLAMPT) != 0;
    if (clampX) u = fclampf(u, 0.f, 0.9999f); else u -= floorf(u); /* :250-258 */
    if (clampY) v = fclampf(v, 0.f, 0.9999f); else v -= floorf(v);
    u = u * width - 0.5f; /* :261-262 pixel centres */
    v = v * height - 0.5f;
    int x = (int)floorf(u), y = (int)floorf(v);
    float uFract = u - x, vFract = v - y;
    float uFractInv = 1.f - uFract, vFractInv = 1.f - vFract;
    px4 c[2][2];
    for (int xOff = 0; xOff < 2; xOff++)
        for (int yOff = 0; yOff < 2; yOff++) {
            int xc = x + xOff, yc = y + yOff;
            xc = clampX ? iclamp(xc, 0, (int)width - 1) : intmod(xc, width);
            yc = clampY ? iclamp(yc, 0, (int)height - 1) : intmod(yc, height);
            const uint8_t *px = t->data + offset + (uint32_t)yc * width * pixelSize + (uint32_t)xc * pixelSize;
            c[xOff][yOff] = t->layout ? parse_pixel_wide(px, t->layout) : parse_pixel(px);
        }
    px4 r; /* :295-308 */
    r.r = (c[0][0].r * uFractInv + c[1][0].r * uFract) * vFractInv + (c[0][1].r * uFractInv + c[1][1].r * uFract) * vFract;
    r.g = (c[0][0].g * uFractInv + c[1][0].g * uFract) * vFractInv + (c[0][1].g * uFractInv + c[1][1].g * uFract) * vFract;
    r.b = (c[0][0].b * uFractInv + c[1][0].b * uFract) * vFractInv + (c[0][1].b * uFractInv + c[1][1].b * uFract) * vFract;
    r.a = (c[0][0].a * uFractInv + c[1][0].a * uFract) * vFractInv + (c[0][1].a * uFractInv + c[1][1].a * uFract) * vFract;
    return r;
}

/* VTFTexture::Sample (libs/VTFParser/VTFParser.cpp:311-330) through IVTFTexture::Sample(u,v,mip)
 * (include/vistrace/IVTFTexture.h:139-142) */
static px4 tex_sample(const OTex *t, float u, float v, float mipLevel) {
    mipLevel = fclampf(mipLevel, 0.f, (float)(t->mips - 1));
    float mipHigh = floorf(mipLevel), mipLow = ceilf(mipLevel);
    px4 high = sample_bilinear(t, u, v, (uint8_t)mipHigh);
    if (mipLow == mipHigh) return high;
    px4 low = sample_bilinear(t, u, v, (uint8_t)mipLow);
    float fract = mipLevel - mipHigh, fractInv = 1.f - fract;
    px4 r = {low.r * fract + high.r * fractInv, low.g * fract + high.g * fractInv, low.b * fract + high.b * fractInv,
             low.a * fract + high.a * fractInv};
    return r;
}

/* TransformTexcoord (source/Utils.h:65-72); glm::dot(vec4) = (x+y)+(z+w) of the products (func_geometric.inl:57-64) */
static inline v2 transform_texcoord(v2 tc, const float m[8], float scale) {
    float x = (tc.x * m[0] + tc.y * m[1]) + (1.f * m[2] + 1.f * m[3]);
    float y = (tc.x * m[4] + tc.y * m[5]) + (1.f * m[6] + 1.f * m[7]);
    v2 r = {x * scale, y * scale};
    return r;
}

/* TriUVInfoToTexLOD (source/Utils.h:75-78) */
static inline float tri_uv_info_to_tex_lod(const OTex *t, v2 uvInfo) {
    return uvInfo.x + 0.5f * log2f((float)((int)t->width * (int)t->height) * uvInfo.y);
}

/* ------------------------------------------------------------ triangle set-up */

/* Triangle ctor + ComputeNormalAndLoD (source/objects/Primitives.h:75-102) */
static void derive_triangle(const vt_tri_in *in, OTri *t) {
    for (int k = 0; k < 3; k++) {
        t->p0[k] = in->p[0][k];
        t->e1[k] = in->p[0][k] - in->p[1][k]; /* e1 = p0 - p1 (:82) */
        t->e2[k] = in->p[2][k] - in->p[0][k]; /* e2 = p2 - p0 */
    }
    memcpy(t->uvs, in->uvs, sizeof(t->uvs));
    bcross(t->e1, t->e2, t->n); /* LeftHandedNormal = true (:93) */
    float uv10x = t->uvs[1][0] - t->uvs[0][0], uv10y = t->uvs[1][1] - t->uvs[0][1];
    float uv20x = t->uvs[2][0] - t->uvs[0][0], uv20y = t->uvs[2][1] - t->uvs[0][1];
    float triUVArea = fabsf(uv10x * uv20y - uv20x * uv10y); /* :97 */
    float len = sqrtf(bdot(t->n, t->n));                  /* bvh::length, vector.hpp:143-147 */
    t->lod = 0.5f * log2f(triUVArea / len);               /* :100 */
    for (int k = 0; k < 3; k++) t->nNorm[k] = t->n[k] / len;
    memcpy(t->normals, in->normals, sizeof(t->normals));
    memcpy(t->tangents, in->tangents, sizeof(t->tangents));
    memcpy(t->alphas, in->alphas, sizeof(t->alphas));
    t->material = in->material;
    t->entIdx = in->ent_idx;
    t->oneSided = in->one_sided;
}

/* ------------------------------------------------------------------ hot path */

typedef struct { float o[3], d[3], tmin, tmax; } ORay;

/* TriangleBackfaceCull::intersect (source/objects/Primitives.h:168-215).  Returns 1 on hit. */
static int tri_intersect(const OScene *s, const OTri *t, const ORay *ray, float *ot, float *ou, float *ov) {
    const vt_material *mat = &s->mats[t->material]; /* :170 */
    float nDotDir = bdot(t->n, ray->d);
    if (t->oneSided && (mat->flags & VT_MATFLAG_NOCULL) == 0 && nDotDir > 0) return 0; /* :173-174 */
    float c[3] = {t->p0[0] - ray->o[0], t->p0[1] - ray->o[1], t->p0[2] - ray->o[2]};
    float r[3];
    bcross(ray->d, c, r);
    float inv_det = 1.0f / nDotDir; /* negate_when_right_handed(1.0) is the identity for LeftHandedNormal */
    float u = bdot(r, t->e2) * inv_det;
    float v = bdot(r, t->e1) * inv_det;
    float w = 1.0f - u - v;
    if (u >= 0.0f && v >= 0.0f && w >= 0.0f) { /* tolerance = 0 (NonZeroTolerance = false) */
        float tt = bdot(t->n, c) * inv_det;
        if (tt >= ray->tmin && tt <= ray->tmax) {
            if ((mat->flags & VT_MATFLAG_ALPHATEST) != 0) { /* :195-208 */
                float w2 = 1.f - u - v;
                v2 texUV = {(w2 * t->uvs[0][0] + u * t->uvs[1][0]) + v * t->uvs[2][0],
                            (w2 * t->uvs[0][1] + u * t->uvs[1][1]) + v * t->uvs[2][1]};
                texUV = transform_texcoord(texUV, mat->base_tex_mat, mat->tex_scale);
                int ti = mat->base_texture >= 0 ? mat->base_texture : (int)s->n_texs - 1;
                float alpha = tex_sample(&s->texs[ti], texUV.x, texUV.y, 0.f).a;
                if (alpha < mat->alphatest_reference) return 0;
            }
            *ot = tt;
            *ou = u;
            *ov = v;
            return 1;
        }
    }
    return 0;
}

/* FastNodeIntersector (libs/bvh/include/bvh/node_intersectors.hpp:82-103) + base class (:15-47) */
typedef struct { int octant[3]; float inv[3], so[3]; } ONodeIsect;

static void node_isect_init(ONodeIsect *ni, const ORay *ray) {
    for (int i = 0; i < 3; i++) {
        ni->octant[i] = signbit(ray->d[i]) ? 1 : 0; /* :20-26 */
        float d = ray->d[i];                        /* safe_inverse, vector.hpp:69-74 */
        ni->inv[i] = 1.0f / (fabsf(d) < FLT_EPSILON ? copysignf(FLT_EPSILON, d) : d);
        ni->so[i] = -ray->o[i] * ni->inv[i]; /* scaled_origin = -origin * inverse_direction (:93) */
    }
}
static inline float rmax(float x, float y) { return x > y ? x : y; } /* robust_max, utilities.hpp:68-71 */
static inline float rmin(float x, float y) { return x < y ? x : y; } /* robust_min, utilities.hpp:61-64 */

static inline void node_isect(const ONodeIsect *ni, const vt_node *nd, const ORay *ray, float *entry, float *exit_) {
    /* fast_multiply_add = fmaf with FP_FAST_FMAF (utilities.hpp:44-54); :33-46 */
    float e0 = fmaf(nd->bounds[0 + ni->octant[0]], ni->inv[0], ni->so[0]);
    float e1 = fmaf(nd->bounds[2 + ni->octant[1]], ni->inv[1], ni->so[1]);
    float e2 = fmaf(nd->bounds[4 + ni->octant[2]], ni->inv[2], ni->so[2]);
    float x0 = fmaf(nd->bounds[0 + 1 - ni->octant[0]], ni->inv[0], ni->so[0]);
    float x1 = fmaf(nd->bounds[2 + 1 - ni->octant[1]], ni->inv[1], ni->so[1]);
    float x2 = fmaf(nd->bounds[4 + 1 - ni->octant[2]], ni->inv[2], ni->so[2]);
    *entry = rmax(e0, rmax(e1, rmax(e2, ray->tmin)));
    *exit_ = rmin(x0, rmin(x1, rmin(x2, ray->tmax)));
}

typedef struct { int hit; uint64_t prim; float t, u, v; } OBest;

/* SingleRayTraverser::intersect_leaf (libs/bvh/include/bvh/single_ray_traverser.hpp:41-63) with
 * ClosestPrimitiveIntersector::intersect (primitive_intersectors.hpp:48-53, Permuted = false) */
static inline void intersect_leaf(const OScene *s, const vt_node *nd, ORay *ray, OBest *best, uint64_t *isects) {
    uint64_t begin = nd->first, end = begin + nd->prim_count;
    *isects += end - begin;
    for (uint64_t i = begin; i < end; i++) {
        uint64_t idx = s->prim_indices[i]; /* primitive_at, primitive_intersectors.hpp:17-20 */
        float t, u, v;
        if (tri_intersect(s, &s->tris[idx], ray, &t, &u, &v)) {
            best->hit = 1;
            best->prim = idx;
            best->t = t;
            best->u = u;
            best->v = v;
            ray->tmax = t; /* :59 — later candidates with t == tmax replace this one */
        }
    }
}

#define STACK_SIZE 64 /* single_ray_traverser.hpp:14 */

/* SingleRayTraverser::intersect (single_ray_traverser.hpp:65-126).  Returns -1 if the 64-entry stack
 * would overflow (the reference has no check in release builds: undefined behaviour there). */
static int traverse_one(const OScene *s, ORay ray, OBest *best, uint64_t *steps, uint64_t *isects) {
    best->hit = 0;
    const vt_node *nodes = s->nodes;
    if (nodes[0].prim_count != 0) { /* :72-73 root is a leaf */
        intersect_leaf(s, &nodes[0], &ray, best, isects);
        return 0;
    }
    ONodeIsect ni;
    node_isect_init(&ni, &ray);
    uint32_t stack[STACK_SIZE];
    int sp = 0;
    const vt_node *left = &nodes[nodes[0].first];
    for (;;) {
        (*steps)++;
        const vt_node *right = left + 1;
        float le, lx, re, rx;
        node_isect(&ni, left, &ray, &le, &lx); /* both boxes are tested before any leaf shrinks tmax (:86-87) */
        node_isect(&ni, right, &ray, &re, &rx);
        if (le <= lx) {
            if (left->prim_count != 0) {
                intersect_leaf(s, left, &ray, best, isects);
                left = NULL;
            }
        } else
            left = NULL;
        if (re <= rx) {
            if (right->prim_count != 0) {
                intersect_leaf(s, right, &ray, best, isects);
                right = NULL;
            }
        } else
            right = NULL;
        if (left) {
            if (right) {
                if (le > re) { /* :111 far child pushed; ties keep left first */
                    const vt_node *tmp = left;
                    left = right;
                    right = tmp;
                }
                if (sp >= STACK_SIZE) return -1;
                stack[sp++] = right->first;
            }
            left = &nodes[left->first];
        } else if (right) {
            left = &nodes[right->first];
        } else {
            if (sp == 0) break;
            left = &nodes[stack[--sp]];
        }
    }
    return 0;
}

/* --------------------------------------------------------------- TraceResult */

/* TextureCombine (source/objects/TraceResult.cpp:11-43) */
static px4 texture_combine(px4 base, px4 det, uint8_t mode, float bf) {
    px4 r = base;
    switch (mode) {
    case 0: { /* DecalModulate */
        r.r = base.r * lerp1(1.f, 2.f * det.r, bf);
        r.g = base.g * lerp1(1.f, 2.f * det.g, bf);
        r.b = base.b * lerp1(1.f, 2.f * det.b, bf);
        r.a = base.a * 1.f;
        return r;
    }
    case 5: case 6: case 1: { /* UnlitAdditive, UnlitAdditiveThresholdFade, Additive */
        r.r = base.r + bf * det.r;
        r.g = base.g + bf * det.g;
        r.b = base.b + bf * det.b;
        r.a = base.a + 0.f;
        return r;
    }
    case 2: { /* TranslucentDetail */
        float blend = bf * det.a;
        r.r = lerp1(base.r, det.r, blend);
        r.g = lerp1(base.g, det.g, blend);
        r.b = lerp1(base.b, det.b, blend);
        r.a = base.a;
        return r;
    }
    case 3: { /* BlendFactorFade */
        r.r = lerp1(base.r, det.r, bf);
        r.g = lerp1(base.g, det.g, bf);
        r.b = lerp1(base.b, det.b, bf);
        r.a = lerp1(base.a, det.a, bf);
        return r;
    }
    case 4: { /* TranslucentBase */
        float blend = bf * (1.f - base.a);
        r.r = lerp1(base.r, det.r, blend);
        r.g = lerp1(base.g, det.g, blend);
        r.b = lerp1(base.b, det.b, blend);
        r.a = det.a;
        return r;
    }
    case 7: { /* TwoPatternDecalModulate */
        float dc = lerp1(det.r, det.a, base.a);
        float m = lerp1(1.f, 2.f * dc, bf);
        r.r = base.r * m;
        r.g = base.g * m;
        r.b = base.b * m;
        r.a = base.a * 1.f;
        return r;
    }
    case 8: { /* Multiply */
        r.r = lerp1(base.r, base.r * det.r, bf);
        r.g = lerp1(base.g, base.g * det.g, bf);
        r.b = lerp1(base.b, base.b * det.b, bf);
        r.a = lerp1(base.a, base.a * det.a, bf);
        return r;
    }
    case 9: { /* BaseMaskDetailAlpha */
        r.a = lerp1(base.a, base.a * det.a, bf);
        return r;
    }
    default: /* SSBump, SSBumpAlbedo */
        return base;
    }
}

static inline const OTex *mat_tex(const OScene *s, int32_t idx) { return (idx >= 0 && (uint32_t)idx < s->n_texs) ? &s->texs[idx] : NULL; }

/* TraceResult ctor + every getter (source/objects/TraceResult.cpp:45-262), evaluated eagerly.
 * dir_n is glm::normalize(direction) as passed by AccelStruct::Traverse (AccelStruct.cpp:826). */
static void trace_result(const OScene *s, const vt_ray *ray, uint64_t prim, float dist, float u, float v, float coneWidth,
                         float coneAngle, vt_attr *o) {
    const OTri *tri = &s->tris[prim];
    const vt_entity *ent = &s->ents[tri->entIdx];
    const vt_material *mat = &s->mats[tri->material];
    const OTex *baseTexture = mat_tex(s, mat->base_texture);
    if (!baseTexture) baseTexture = &s->texs[s->n_texs - 1]; /* ingestion fallback, see vto_create */

    v3 dn = gnormalize(V3(ray->dx, ray->dy, ray->dz)); /* AccelStruct.cpp:826 */
    v3 wo = V3(-dn.x, -dn.y, -dn.z);                    /* wo = -direction (:56) */
    int mipOverride = (coneWidth < 0.f || coneAngle <= 0.f); /* :53 */
    v3 vN[3], vT[3], vB[3], vv[3];
    v2 vUV[3];
    for (int i = 0; i < 3; i++) { /* :58-63 */
        vN[i] = V3(tri->normals[i][0], tri->normals[i][1], tri->normals[i][2]);
        vT[i] = V3(tri->tangents[i][0], tri->tangents[i][1], tri->tangents[i][2]);
        vB[i] = gcross(vT[i], vN[i]);
        vUV[i].x = tri->uvs[i][0];
        vUV[i].y = tri->uvs[i][1];
    }
    vv[0] = V3(tri->p0[0], tri->p0[1], tri->p0[2]); /* :65-68; p1 = p0 - e1, p2 = p0 + e2 (Primitives.h:104-105) */
    vv[1] = V3(tri->p0[0] - tri->e1[0], tri->p0[1] - tri->e1[1], tri->p0[2] - tri->e1[2]);
    vv[2] = V3(tri->p0[0] + tri->e2[0], tri->p0[1] + tri->e2[1], tri->p0[2] + tri->e2[2]);
    v3 uvw = V3(u, v, 1.f - u - v); /* :70 */
    v3 gN = V3(tri->nNorm[0], tri->nNorm[1], tri->nNorm[2]);
    float blendFactor = uvw.z * tri->alphas[0] + uvw.x * tri->alphas[1] + uvw.y * tri->alphas[2]; /* :73 */
    v2 texUV = {uvw.z * vUV[0].x + uvw.x * vUV[1].x + uvw.y * vUV[2].x, uvw.z * vUV[0].y + uvw.x * vUV[1].y + uvw.y * vUV[2].y};
    v3 albedo = V3(ent->colour[0] * mat->colour[0], ent->colour[1] * mat->colour[1], ent->colour[2] * mat->colour[2]); /* :80 */
    float alpha = ent->colour[3] * mat->colour[3];
    int hitSky = (mat->surf_flags & VT_SURF_SKY) != 0; /* :83 */
    int frontFacing = gdot(wo, gN) >= 0.f;             /* :85 */

    /* CalcFootprint (:89-104) */
    v2 lodInfo = {0.f, 0.f};
    if (!mipOverride) {
        coneWidth = coneAngle * dist + coneWidth;
        float normalTerm = gdot(wo, gN);
        lodInfo.x = tri->lod;
        lodInfo.y = (coneWidth * coneWidth) / (normalTerm * normalTerm);
    }
#define LOD_OF(tex) (mipOverride ? 0.f : tri_uv_info_to_tex_lod((tex), lodInfo))

    /* CalcBlendFactor (:106-130) */
    const OTex *blendTexture = mat_tex(s, mat->blend_texture);
    if (mat->masked_blending) blendFactor = 0.5f;
    if (blendTexture) {
        v2 sc = transform_texcoord(texUV, mat->blend_tex_mat, mat->tex_scale);
        px4 pb = tex_sample(blendTexture, sc.x, sc.y, LOD_OF(blendTexture));
        if (mat->masked_blending) {
            blendFactor = pb.g;
        } else {
            float minb = saturate(pb.g - pb.r);
            float maxb = saturate(pb.g + pb.r);
            blendFactor = gsmoothstep(minb, maxb, blendFactor);
        }
    }

    /* GetPos (:255-262) */
    v3 pos = v3add(v3add(v3s(vv[0], uvw.z), v3s(vv[1], uvw.x)), v3s(vv[2], uvw.y));

    /* CalcTBN (:132-187) */
    v3 normal = gnormalize(v3add(v3add(v3s(vN[0], uvw.z), v3s(vN[1], uvw.x)), v3s(vN[2], uvw.y)));
    v3 tangent = gnormalize(v3add(v3add(v3s(vT[0], uvw.z), v3s(vT[1], uvw.x)), v3s(vT[2], uvw.y)));
    v3 binormal = gnormalize(v3add(v3add(v3s(vB[0], uvw.z), v3s(vB[1], uvw.x)), v3s(vB[2], uvw.y)));
    const OTex *normalMap = mat_tex(s, mat->normal_map);
    if (normalMap) { /* :140-174 */
        v2 sc = transform_texcoord(texUV, mat->normal_map_mat, mat->tex_scale);
        px4 pn = tex_sample(normalMap, sc.x, sc.y, LOD_OF(normalMap));
        v3 mapped = V3(pn.r * 2.f - 1.f, pn.g * 2.f - 1.f, pn.b * 2.f - 1.f);
        const OTex *normalMap2 = mat_tex(s, mat->normal_map2);
        if (normalMap2) {
            sc = transform_texcoord(texUV, mat->normal_map_mat2, mat->tex_scale);
            pn = tex_sample(normalMap2, sc.x, sc.y, LOD_OF(normalMap2));
            v3 mapped2 = V3(pn.r * 2.f - 1.f, pn.g * 2.f - 1.f, pn.b * 2.f - 1.f);
            mapped = gnormalize(lerp3(mapped, mapped2, blendFactor));
        }
        /* mat3 columns = tangent, binormal, normal; m*v (libs/glm/glm/detail/type_mat3x3.inl:468-474) */
        v3 wn = V3(tangent.x * mapped.x + binormal.x * mapped.y + normal.x * mapped.z,
                   tangent.y * mapped.x + binormal.y * mapped.y + normal.y * mapped.z,
                   tangent.z * mapped.x + binormal.z * mapped.y + normal.z * mapped.z);
        wn = gnormalize(wn);
        if (isfinite(wn.x) && isfinite(wn.y) && isfinite(wn.z)) {
            normal = wn;
            tangent = gnormalize(v3sub(tangent, v3s(normal, gdot(tangent, normal))));
            binormal = gcross(tangent, normal);
        }
    }
    const float kCosThetaThreshold = 0.1f; /* :176-184 */
    float cosTheta = fabsf(gdot(wo, normal));
    if (cosTheta <= kCosThetaThreshold) {
        float t = saturate(cosTheta * (1.f / kCosThetaThreshold));
        normal = gnormalize(lerp3(gN, normal, t));
        tangent = gnormalize(v3sub(tangent, v3s(normal, gdot(tangent, normal))));
        binormal = gcross(tangent, normal);
    }

    /* CalcShadingData (:189-253) */
    v2 scaled = transform_texcoord(texUV, mat->base_tex_mat, mat->tex_scale);
    v2 scaled2 = transform_texcoord(texUV, mat->base_tex_mat2, mat->tex_scale);
    px4 colour = tex_sample(baseTexture, scaled.x, scaled.y, LOD_OF(baseTexture));
    const OTex *baseTexture2 = mat_tex(s, mat->base_texture2);
    if (baseTexture2) {
        px4 c2 = tex_sample(baseTexture2, scaled2.x, scaled2.y, LOD_OF(baseTexture2));
        colour.r = lerp1(colour.r, c2.r, blendFactor);
        colour.g = lerp1(colour.g, c2.g, blendFactor);
        colour.b = lerp1(colour.b, c2.b, blendFactor);
        colour.a = lerp1(colour.a, c2.a, blendFactor);
    }
    const OTex *detail = mat_tex(s, mat->detail);
    if (detail) {
        v2 duv = transform_texcoord(texUV, mat->detail_mat, mat->detail_scale);
        px4 dc = tex_sample(detail, duv.x, duv.y, LOD_OF(detail));
        colour = texture_combine(colour, dc, mat->detail_blend_mode, mat->detail_blend_factor);
        colour.r = gclamp(colour.r, 0.f, 1.f);
        colour.g = gclamp(colour.g, 0.f, 1.f);
        colour.b = gclamp(colour.b, 0.f, 1.f);
        colour.a = gclamp(colour.a, 0.f, 1.f);
    }
    albedo = v3mul(albedo, V3(colour.r, colour.g, colour.b));
    alpha *= colour.a;
    float metalness = 0.f, roughness = 1.f; /* TraceResult.h:46-47 */
    const OTex *mrao = mat_tex(s, mat->mrao);
    if (mrao) {
        px4 pm = tex_sample(mrao, scaled.x, scaled.y, LOD_OF(mrao));
        float mr = pm.r, mg = pm.g;
        const OTex *mrao2 = mat_tex(s, mat->mrao2);
        if (mrao2) {
            px4 pm2 = tex_sample(mrao2, scaled2.x, scaled2.y, LOD_OF(mrao2));
            mr = lerp1(mr, pm2.r, blendFactor);
            mg = lerp1(mg, pm2.g, blendFactor);
        }
        metalness = mr;
        roughness = mg;
    }

    o->pos[0] = pos.x; o->pos[1] = pos.y; o->pos[2] = pos.z;
    o->distance = dist;
    o->normal[0] = normal.x; o->normal[1] = normal.y; o->normal[2] = normal.z;
    o->alpha = alpha;
    o->tangent[0] = tangent.x; o->tangent[1] = tangent.y; o->tangent[2] = tangent.z;
    o->metalness = metalness;
    o->binormal[0] = binormal.x; o->binormal[1] = binormal.y; o->binormal[2] = binormal.z;
    o->roughness = roughness;
    o->geometric_normal[0] = gN.x; o->geometric_normal[1] = gN.y; o->geometric_normal[2] = gN.z;
    o->base_mip = LOD_OF(baseTexture); /* GetBaseMIPLevel (:305-309) */
    o->albedo[0] = albedo.x; o->albedo[1] = albedo.y; o->albedo[2] = albedo.z;
    o->ent_id = ent->id;
    o->uvw[0] = uvw.x; o->uvw[1] = uvw.y; o->uvw[2] = uvw.z;
    o->submat_idx = tri->material;
    o->tex_uv[0] = texUV.x; o->tex_uv[1] = texUV.y;
    o->flags = (frontFacing ? VT_ATTR_FRONT_FACING : 0u) | (hitSky ? VT_ATTR_HIT_SKY : 0u) | (mat->water ? VT_ATTR_HIT_WATER : 0u);
    o->prim = (uint32_t)prim;
#undef LOD_OF
}

/* ---------------------------------------------------------------- C interface
 * Same shape as the vtref_* functions of oracle/ref_harness.cpp so tests drive both alike. */

int vto_max_threads(void) { return omp_get_max_threads(); }

void vto_destroy(void *h) {
    OScene *s = (OScene *)h;
    if (!s) return;
    for (uint32_t i = 0; i < s->n_texs; i++) free(s->texs[i].data);
    free(s->texs);
    free(s->tris);
    free(s->mats);
    free(s->ents);
    free(s->nodes);
    free(s->prim_indices);
    free(s);
}

/* `build` must be 0: hierarchy construction is not restated (see header); use vto_set_bvh. */
void *vto_create(const vt_scene *sc, int build) {
    if (build) return NULL;
    OScene *s = (OScene *)calloc(1, sizeof(OScene));
    s->n_tris = sc->n_tris;
    s->tris = (OTri *)malloc(sizeof(OTri) * (sc->n_tris ? sc->n_tris : 1));
#pragma omp parallel for
    for (int64_t i = 0; i < (int64_t)sc->n_tris; i++) derive_triangle(&sc->tris[i], &s->tris[i]);
    s->n_mats = sc->n_materials;
    s->mats = (vt_material *)malloc(sizeof(vt_material) * (s->n_mats ? s->n_mats : 1));
    memcpy(s->mats, sc->materials, sizeof(vt_material) * s->n_mats);
    s->n_ents = sc->n_entities;
    s->ents = (vt_entity *)malloc(sizeof(vt_entity) * (s->n_ents ? s->n_ents : 1));
    memcpy(s->ents, sc->entities, sizeof(vt_entity) * s->n_ents);
    /* Ingestion never leaves baseTexture null (fallback MISSING_TEXTURE, source/objects/AccelStruct.cpp:120,286).
     * Headless stand-in, identical in ref_harness.cpp: a 1x1 opaque white texture appended after the caller's. */
    s->n_texs = sc->n_textures + 1;
    s->texs = (OTex *)calloc(s->n_texs, sizeof(OTex));
    for (uint32_t i = 0; i < sc->n_textures; i++) {
        const vt_texture *t = &sc->textures[i];
        s->texs[i].width = t->width;
        s->texs[i].height = t->height;
        s->texs[i].mips = t->mip_count;
        s->texs[i].flags = t->flags;
        s->texs[i].layout = t->texel_layout;
        s->texs[i].data = (uint8_t *)malloc(t->nbytes);
        memcpy(s->texs[i].data, t->rgba, t->nbytes);
    }
    OTex *fb = &s->texs[s->n_texs - 1];
    fb->width = fb->height = fb->mips = 1;
    fb->flags = 0;
    fb->data = (uint8_t *)malloc(4);
    memset(fb->data, 255, 4);
    return s;
}

void vto_get_tri_derived(void *h, float *out16) {
    OScene *s = (OScene *)h;
    for (uint64_t i = 0; i < s->n_tris; i++) {
        const OTri *t = &s->tris[i];
        float *o = out16 + i * 16;
        for (int k = 0; k < 3; k++) {
            o[k] = t->p0[k];
            o[3 + k] = t->e1[k];
            o[6 + k] = t->e2[k];
            o[9 + k] = t->n[k];
            o[12 + k] = t->nNorm[k];
        }
        o[15] = t->lod;
    }
}

void vto_get_bvh(void *h, vt_node *nodes, uint64_t *node_count, uint64_t *prim_indices, uint64_t *n_tris) {
    OScene *s = (OScene *)h;
    if (node_count) *node_count = s->node_count;
    if (n_tris) *n_tris = s->n_tris;
    if (nodes) memcpy(nodes, s->nodes, s->node_count * sizeof(vt_node));
    if (prim_indices) memcpy(prim_indices, s->prim_indices, s->n_tris * sizeof(uint64_t));
}

void vto_set_bvh(void *h, const vt_node *nodes, uint64_t node_count, const uint64_t *prim_indices) {
    OScene *s = (OScene *)h;
    free(s->nodes);
    free(s->prim_indices);
    s->nodes = (vt_node *)malloc(sizeof(vt_node) * node_count);
    memcpy(s->nodes, nodes, sizeof(vt_node) * node_count);
    s->node_count = node_count;
    s->prim_indices = (uint64_t *)malloc(sizeof(uint64_t) * (s->n_tris ? s->n_tris : 1));
    memcpy(s->prim_indices, prim_indices, sizeof(uint64_t) * s->n_tris);
    s->built = 1;
}

/* bvh::HierarchyRefitter::refit (libs/bvh/include/bvh/hierarchy_refitter.hpp:20-31) over moved geometry of the same
 * topology, leaf update as the library's own test drives it (libs/bvh/test/refit_bvh.cpp:79-89): a leaf's box is the
 * union of Triangle::bounding_box() (source/objects/Primitives.h:107-113: p0, p1() = p0 - e1, p2() = p0 + e2) of its
 * primitives, an inner node's box the union of its two children.  BoundingBox::extend is a per-component min / max
 * (bounding_box.hpp), so any bottom-up order gives the same bits; the parallel schedule of bottom_up_algorithm.hpp is
 * not restated, a reverse pre-order walk is used instead. */
void vto_refit(void *h, const vt_scene *sc) {
    OScene *s = (OScene *)h;
    if (!s->built || sc->n_tris != s->n_tris) return;
#pragma omp parallel for
    for (int64_t i = 0; i < (int64_t)sc->n_tris; i++) derive_triangle(&sc->tris[i], &s->tris[i]);
    uint64_t *order = (uint64_t *)malloc(sizeof(uint64_t) * s->node_count), n_order = 0;
    uint64_t *stack = (uint64_t *)malloc(sizeof(uint64_t) * s->node_count), sp = 0;
    stack[sp++] = 0;
    while (sp) { /* pre-order: parents before children */
        uint64_t i = stack[--sp];
        order[n_order++] = i;
        if (s->nodes[i].prim_count == 0) {
            stack[sp++] = s->nodes[i].first;
            stack[sp++] = (uint64_t)s->nodes[i].first + 1;
        }
    }
    for (uint64_t k = n_order; k-- > 0;) { /* children before parents */
        vt_node *nd = &s->nodes[order[k]];
        float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX}; /* BoundingBox::empty() */
        if (nd->prim_count) {
            for (uint32_t t = 0; t < nd->prim_count; t++) {
                const OTri *tr = &s->tris[s->prim_indices[nd->first + t]];
                for (int a = 0; a < 3; a++) {
                    const float v[3] = {tr->p0[a], tr->p0[a] - tr->e1[a], tr->p0[a] + tr->e2[a]};
                    for (int j = 0; j < 3; j++) {
                        lo[a] = fminf(lo[a], v[j]);
                        hi[a] = fmaxf(hi[a], v[j]);
                    }
                }
            }
        } else {
            for (int c = 0; c < 2; c++) {
                const vt_node *ch = &s->nodes[nd->first + c];
                for (int a = 0; a < 3; a++) {
                    lo[a] = fminf(lo[a], ch->bounds[2 * a]);
                    hi[a] = fmaxf(hi[a], ch->bounds[2 * a + 1]);
                }
            }
        }
        for (int a = 0; a < 3; a++) {
            nd->bounds[2 * a] = lo[a];
            nd->bounds[2 * a + 1] = hi[a];
        }
    }
    free(order);
    free(stack);
}

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* Batched loop over the single-ray body of AccelStruct::Traverse (source/objects/AccelStruct.cpp:810-837). */
double vto_traverse(void *h, const vt_ray *rays, uint64_t n, vt_hit *hits, vt_attr *attrs, int threads, uint64_t *stats) {
    OScene *s = (OScene *)h;
    if (!s->built) return -1.0;
    if (threads <= 0) threads = omp_get_max_threads();
    uint64_t steps = 0, isects = 0;
    double t0 = now_s();
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads) reduction(+ : steps, isects)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        const vt_ray *r = &rays[i];
        ORay ray = {{r->ox, r->oy, r->oz}, {r->dx, r->dy, r->dz}, r->tmin, r->tmax};
        OBest best;
        uint64_t st = 0, is = 0;
        int rc = traverse_one(s, ray, &best, &st, &is);
        steps += st;
        isects += is;
        if (rc == 0 && best.hit) {
            if (hits) {
                hits[i].t = best.t;
                hits[i].u = best.u;
                hits[i].v = best.v;
                hits[i].prim = (uint32_t)best.prim;
            }
            if (attrs) trace_result(s, r, best.prim, best.t, best.u, best.v, -1.f, -1.f, &attrs[i]);
        } else {
            if (hits) {
                hits[i].t = hits[i].u = hits[i].v = 0.f;
                hits[i].prim = VT_MISS;
            }
            if (attrs) {
                memset(&attrs[i], 0, sizeof(vt_attr));
                attrs[i].prim = VT_MISS;
            }
        }
    }
    double t1 = now_s();
    if (stats) {
        stats[0] = steps;
        stats[1] = isects;
    }
    return t1 - t0;
}

void vto_trace_result(void *h, const vt_ray *rays, const vt_hit *hits, uint64_t n, vt_attr *attrs, int threads) {
    OScene *s = (OScene *)h;
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        if (hits[i].prim == VT_MISS) {
            memset(&attrs[i], 0, sizeof(vt_attr));
            attrs[i].prim = VT_MISS;
        } else {
            trace_result(s, &rays[i], hits[i].prim, hits[i].t, hits[i].u, hits[i].v, -1.f, -1.f, &attrs[i]);
        }
    }
}

/* Same with per-ray texture-LOD cones {coneWidth, coneAngle} (the 5th/6th arguments of accel:Traverse,
 * source/objects/AccelStruct.cpp:795-803); cones == NULL -> (-1, -1). */
void vto_trace_result_cones(void *h, const vt_ray *rays, const vt_hit *hits, const float *cones, uint64_t n, vt_attr *attrs, int threads) {
    OScene *s = (OScene *)h;
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        if (hits[i].prim == VT_MISS) {
            memset(&attrs[i], 0, sizeof(vt_attr));
            attrs[i].prim = VT_MISS;
        } else {
            trace_result(s, &rays[i], hits[i].prim, hits[i].t, hits[i].u, hits[i].v, cones ? cones[2 * i] : -1.f,
                         cones ? cones[2 * i + 1] : -1.f, &attrs[i]);
        }
    }
}

void vto_sample(void *h, int tex, const float *uvm, uint64_t n, float *rgba) {
    OScene *s = (OScene *)h;
    for (uint64_t i = 0; i < n; i++) {
        px4 p = tex_sample(&s->texs[tex], uvm[i * 3], uvm[i * 3 + 1], uvm[i * 3 + 2]);
        rgba[i * 4] = p.r;
        rgba[i * 4 + 1] = p.g;
        rgba[i * 4 + 2] = p.b;
        rgba[i * 4 + 3] = p.a;
    }
}

void vto_node_intersect(const vt_node *node, const vt_ray *r, float *out2) {
    ORay ray = {{r->ox, r->oy, r->oz}, {r->dx, r->dy, r->dz}, r->tmin, r->tmax};
    ONodeIsect ni;
    node_isect_init(&ni, &ray);
    node_isect(&ni, node, &ray, &out2[0], &out2[1]);
}

int vto_tri_intersect(void *h, uint64_t prim, const vt_ray *r, float *tuv) {
    OScene *s = (OScene *)h;
    ORay ray = {{r->ox, r->oy, r->oz}, {r->dx, r->dy, r->dz}, r->tmin, r->tmax};
    return tri_intersect(s, &s->tris[prim], &ray, &tuv[0], &tuv[1], &tuv[2]);
}

/* ---- SkinTriangle: source/objects/AccelStruct.cpp:33-47 (TransformToBone) and :66-101 (SkinTriangle), with
 * glm::mat4 * mat4 (libs/glm/glm/detail/type_mat4x4.inl:630-648) and mat4 * vec4 (:561-572) in glm's operation order.
 * out27[i] = {p0, e1, e2, normals[3], tangents[3]} as SkinTriangle leaves them. */
static void m4_mul(const float *a, const float *b, float *r) { /* column-major: m[col*4 + row] */
    for (int j = 0; j < 4; j++)
        for (int k = 0; k < 4; k++)
            r[j * 4 + k] = ((a[0 * 4 + k] * b[j * 4 + 0] + a[1 * 4 + k] * b[j * 4 + 1]) + a[2 * 4 + k] * b[j * 4 + 2]) + a[3 * 4 + k] * b[j * 4 + 3];
}
static void m4_mul_v(const float *m, const float *v, float *o) {
    for (int k = 0; k < 4; k++) o[k] = (m[0 * 4 + k] * v[0] + m[1 * 4 + k] * v[1]) + (m[2 * 4 + k] * v[2] + m[3 * 4 + k] * v[3]);
}
static void transform_to_bone(const float *vec, const float *bones, const float *binds, unsigned num, const float *w, const int8_t *ids,
                              int angle_only, float *out) {
    float fin[4] = {0.f, 0.f, 0.f, 0.f};
    const float vertex[4] = {vec[0], vec[1], vec[2], angle_only ? 0.f : 1.f};
    for (unsigned i = 0; i < num; i++) {
        float bb[16], t[4];
        m4_mul(bones + 16 * ids[i], binds + 16 * ids[i], bb);
        m4_mul_v(bb, vertex, t);
        for (int k = 0; k < 4; k++) fin[k] += t[k] * w[i];
    }
    out[0] = fin[0], out[1] = fin[1], out[2] = fin[2];
}
void vto_skin_triangles(const vt_tri_in *tris, const vt_tri_skin *skin, uint64_t n, const float *bones, const float *binds,
                        uint32_t n_bones, float *out27) {
    (void)n_bones;
    static const float one_w[3] = {1.f, 0.f, 0.f};
    static const int8_t one_id[3] = {0, 0, 0};
    for (uint64_t i = 0; i < n; i++) {
        const vt_tri_in *in = &tris[i];
        float pos[3][3], *o = out27 + 27 * i;
        for (int k = 0; k < 3; k++) { /* p0, p0 - e1, p0 + e2 with the constructor's rounded edges (Primitives.h:82) */
            const float e1 = in->p[0][k] - in->p[1][k], e2 = in->p[2][k] - in->p[0][k];
            pos[0][k] = in->p[0][k];
            pos[1][k] = in->p[0][k] - e1;
            pos[2][k] = in->p[0][k] + e2;
        }
        float v[3][3];
        for (int vi = 0; vi < 3; vi++) {
            const unsigned num = skin ? skin[i].num_bones[vi] : 1u;
            const float *w = skin ? skin[i].weights[vi] : one_w;
            const int8_t *ids = skin ? skin[i].bone_ids[vi] : one_id;
            transform_to_bone(pos[vi], bones, binds, num, w, ids, 0, v[vi]);
            transform_to_bone(in->normals[vi], bones, binds, num, w, ids, 1, o + 9 + 3 * vi);
            transform_to_bone(in->tangents[vi], bones, binds, num, w, ids, 1, o + 18 + 3 * vi);
        }
        for (int k = 0; k < 3; k++) {
            o[k] = v[0][k];
            o[3 + k] = v[0][k] - v[1][k]; /* e1 = v0 - v1, e2 = v2 - v0 (AccelStruct.cpp:96-98) */
            o[6 + k] = v[2][k] - v[0][k];
        }
    }
}

/* SampleBSDF restricted to the diffuse lobe — source/libraries/BSDF.cpp:770-825 with a BSDFMaterial whose activeLobes is
 * LobeType::DiffuseReflection (BSDF.h:13), prepared per hit by BSDFMaterial::PrepShadingData (BSDF.cpp:11-21) with the defaults of
 * BSDF.h:58-92 (dielectricInput = 1, specularTransmission = 0, anisotropicRotation = 0).  rnd3 = the three numbers the reference
 * draws from its ISampler, in call order: lobeSelect (:780), then r1, r2 of hemisphere_cos (:69-77).
 * glm::rotate(v, 0, n) (to_local / from_local, :760-767) is the identity for finite vectors: cos 0 = 1, sin 0 = 0. */
typedef struct { float scattered[3]; float pdf; float weight[3]; uint32_t lobe; } vto_bsdf_sample;
static inline float schlick_dielectric(float f0, float cosTheta, float f90) { return f0 + (f90 - f0) * powf(1.f - cosTheta, 5.f); } /* :157-160 */
void vto_sample_bsdf_diffuse(const vt_attr *attrs, const float *wo3, const float *rnd3, uint64_t n, vto_bsdf_sample *out, int32_t *returned) {
    const float pi = 3.14159265358979323846264338327950288f; /* glm::pi<float>() */
#pragma omp parallel for
    for (int64_t i = 0; i < (int64_t)n; i++) {
        const vt_attr *a = &attrs[i];
        /* PrepShadingData, :11-21 */
        const float metallic = gclamp(a->metalness, 0.f, 1.f);
        const float linearRoughness = gclamp(a->roughness, 0.f, 1.f);
        const v3 dielectric = V3(gclamp(1.f * a->albedo[0], 0.f, 1.f), gclamp(1.f * a->albedo[1], 0.f, 1.f), gclamp(1.f * a->albedo[2], 0.f, 1.f));
        /* CalculateLobePDFs, :23-56: only the diffuse lobe is active */
        float pDiffuse = (1.f - metallic) * (1.f - 0.f);
        float normFactor = pDiffuse + 0.f + 0.f + 0.f;
        if (normFactor > 0.f) {
            normFactor = 1.f / normFactor;
            pDiffuse *= normFactor;
        }
        const float lobeSelect = rnd3[3 * i];
        const v3 N = V3(a->normal[0], a->normal[1], a->normal[2]), T = V3(a->tangent[0], a->tangent[1], a->tangent[2]),
                 B = V3(a->binormal[0], a->binormal[1], a->binormal[2]), wo = V3(wo3[3 * i], wo3[3 * i + 1], wo3[3 * i + 2]);
        const int entering = gdot(wo, N) >= 0.f; /* :782 */
        const v3 Ns = entering ? N : V3(-N.x, -N.y, -N.z);
        const v3 incident = V3(gdot(wo, T), gdot(wo, B), gdot(wo, Ns)); /* to_local, :783 */
        v3 scattered = V3(0.f, 0.f, 0.f), weight = V3(0.f, 0.f, 0.f);
        float pdf = 0.f;
        uint32_t lobe = 0;
        if (lobeSelect < pDiffuse) { /* :785-793, SampleDiffuse :252-278 */
            lobe = 1;
            const float r1 = rnd3[3 * i + 1];
            const float z = sqrtf(r1), sinTheta = sqrtf(1.f - r1), phi = 2.f * pi * rnd3[3 * i + 2];
            scattered = V3(sinTheta * cosf(phi), sinTheta * sinf(phi), z);
            pdf = (scattered.z > 0.f) ? (scattered.z / pi) : 0.f;
            const v3 halfway = gnormalize(v3add(incident, scattered));
            const float iDotN = incident.z, sDotH = gdot(scattered, halfway), sDotN = scattered.z;
            const float energyBias = lerp1(0.f, 0.5f, linearRoughness);
            const float energyFactor = lerp1(1.f, 1.f / 1.51f, linearRoughness);
            const float fd90 = energyBias + 2.f * sDotH * sDotH * linearRoughness;
            const float lightScatter = schlick_dielectric(1.f, sDotN, fd90), viewScatter = schlick_dielectric(1.f, iDotN, fd90);
            weight = v3s(v3s(v3s(dielectric, lightScatter), viewScatter), energyFactor);
            weight = v3s(weight, (1.f - metallic) * (1.f - 0.f) / pDiffuse); /* :788 */
            pdf *= pDiffuse;                                                 /* :790; the other lobes' probabilities are 0 */
        }
        /* from_local, :824: T * x + B * y + N * z */
        const v3 world = v3add(v3add(v3s(T, scattered.x), v3s(B, scattered.y)), v3s(Ns, scattered.z));
        out[i].scattered[0] = world.x, out[i].scattered[1] = world.y, out[i].scattered[2] = world.z;
        out[i].pdf = pdf;
        out[i].weight[0] = weight.x, out[i].weight[1] = weight.y, out[i].weight[2] = weight.z;
        out[i].lobe = lobe;
        returned[i] = 1;
    }
}
