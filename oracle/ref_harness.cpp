// oracle/ref_harness.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Headless driver around the UNMODIFIED Derpius/VisTrace reference.  It is
// compiled by oracle/Makefile against the sources where they lie under
// /root/reference (nothing is copied) into oracle/_ref/libvt_ref.so.  Only
// tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference arm
// load it.  The product library (vistrace_b200/csrc) never does.
//
// What runs here is the reference's own code:
//   * Triangle ctor + ComputeNormalAndLoD           source/objects/Primitives.h:75-102
//   * the build sequence                            source/objects/AccelStruct.cpp:762-775
//   * per ray, the statements of AccelStruct::Traverse  source/objects/AccelStruct.cpp:810-831
//     (bvh::SingleRayTraverser + ClosestPrimitiveIntersector + Triangle::intersect)
//   * TraceResult ctor and getters                  source/objects/TraceResult.cpp:45-262
//   * VTFTexture(const uint8_t*, size_t)::Sample    libs/VTFParser/VTFParser.cpp:311-330
//   * SampleBSDF + BSDFMaterial::PrepShadingData    source/libraries/BSDF.cpp:11-21,770-825 (diffuse lobe)
//   * MDL / VVD / VTX parsers + BodyGroup / Mesh    libs/MDLParser/source/*.cpp, source/objects/Model.cpp:11-176
//   * BSPMap (parse + Triangulate + displacements)  libs/BSPParser/BSPParser.cpp, Displacements/*.cpp
// The ingestion paths (Lua, engine filesystem) are bypassed by filling the
// private containers directly, which is why `private` is opened up below.
#include <algorithm>
#include <array>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>
#include <omp.h>

#define private public
#include "AccelStruct.h"
#include "Model.h"
#include "TraceResult.h"
#undef private

#include "BSDF.h"
#include "VTFParser.h"
#include "BSPParser.h"
#include "bvh/hierarchy_refitter.hpp"
#include "bvh/leaf_collapser.hpp"
#include "bvh/locally_ordered_clustering_builder.hpp"
#include "bvh/node_intersectors.hpp"
#include "bvh/parallel_reinsertion_optimizer.hpp"

#include "../include/vistrace_b200.h"

// defined in source/objects/AccelStruct.cpp:66 (no header declares it)
void SkinTriangle(Triangle &tri, const std::vector<glm::mat4> &bones, const std::vector<glm::mat4> &binds);
void SkinTriangle(Triangle &tri, const glm::mat4 bone, glm::mat4 bind);

namespace {

// In-memory IVTFTexture over the reference's VTFTexture parser/sampler.
class MemVTF final : public VisTrace::IVTFTexture {
    std::unique_ptr<VTFTexture> tex;

public:
    explicit MemVTF(const vt_texture &t) {
        VTFHeader hdr;
        std::memset(&hdr, 0, sizeof(hdr));
        std::memcpy(hdr.signature, "VTF\0", 4);
        hdr.version[0] = 7;
        hdr.version[1] = 2;
        hdr.headerSize = 80;
        hdr.width = t.width;
        hdr.height = t.height;
        hdr.flags = t.flags;
        hdr.frames = 1;
        hdr.firstFrame = 0;
        hdr.bumpmapScale = 1.f;
        hdr.highResImageFormat = IMAGE_FORMAT::RGBA8888;
        if (t.texel_layout != 0) {
            // wide texels: numerators over 65535 in every channel are the raw bytes of an RGBA16161616 file; the other wide layouts
            // (the packed 16-bit formats after ParsePixel) have no file the reference could be handed — the C port covers them
            if (t.texel_layout != (VT_TEXEL_WIDE | (VT_TEXEL_DIV_65535 * 0x55u))) throw std::runtime_error("reference harness: wide texture that is not RGBA16161616");
            hdr.highResImageFormat = IMAGE_FORMAT::RGBA16161616;
        }
        hdr.mipmapCount = static_cast<uint8_t>(t.mip_count);
        hdr.lowResImageFormat = IMAGE_FORMAT::NONE;
        hdr.depth = 1;
        std::vector<uint8_t> file(80 + t.nbytes);
        std::memcpy(file.data(), &hdr, 80);
        std::memcpy(file.data() + 80, t.rgba, t.nbytes);
        tex = std::make_unique<VTFTexture>(file.data(), file.size());
    }
    bool IsValid() const override { return tex && tex->IsValid(); }
    VisTrace::VTFTextureFormatInfo GetFormat() const override {
        ImageFormatInfo f = tex->GetFormat();
        VisTrace::VTFTextureFormatInfo o;
        std::memcpy(&o, &f, sizeof(o));
        return o;
    }
    uint32_t GetVersionMajor() const override { return tex->GetVersionMajor(); }
    uint32_t GetVersionMinor() const override { return tex->GetVersionMinor(); }
    uint16_t GetWidth(uint8_t m = 0) const override { return tex->GetWidth(m); }
    uint16_t GetHeight(uint8_t m = 0) const override { return tex->GetHeight(m); }
    uint16_t GetDepth(uint8_t m = 0) const override { return tex->GetDepth(m); }
    uint8_t GetFaces() const override { return tex->GetFaces(); }
    uint16_t GetMIPLevels() const override { return tex->GetMIPLevels(); }
    uint16_t GetFrames() const override { return tex->GetFrames(); }
    uint16_t GetFirstFrame() const override { return tex->GetFirstFrame(); }
    VisTrace::Pixel GetPixel(uint16_t x, uint16_t y, uint16_t z, uint8_t m, uint16_t fr, uint8_t fa) const override {
        VTFPixel p = tex->GetPixel(x, y, z, m, fr, fa);
        return VisTrace::Pixel{p.r, p.g, p.b, p.a};
    }
    VisTrace::Pixel Sample(float u, float v, uint16_t z, float m, uint16_t fr, uint8_t fa) const override {
        VTFPixel p = tex->Sample(u, v, z, m, fr, fa);
        return VisTrace::Pixel{p.r, p.g, p.b, p.a};
    }
};

struct RefScene {
    AccelStruct accel;
    std::vector<std::unique_ptr<MemVTF>> textures;
};

glm::mat2x4 to_mat(const float m[8]) {
    glm::mat2x4 r;
    r[0] = glm::vec4(m[0], m[1], m[2], m[3]);
    r[1] = glm::vec4(m[4], m[5], m[6], m[7]);
    return r;
}

// mTriangles from the caller's vertices through the reference constructor (source/objects/Primitives.h:75-89)
void fill_triangles(AccelStruct &a, const vt_scene *s) {
    a.mTriangles.resize(s->n_tris);
#pragma omp parallel for
    for (int64_t i = 0; i < static_cast<int64_t>(s->n_tris); i++) {
        const vt_tri_in &t = s->tris[i];
        glm::vec2 uvs[3] = {glm::vec2(t.uvs[0][0], t.uvs[0][1]), glm::vec2(t.uvs[1][0], t.uvs[1][1]),
                            glm::vec2(t.uvs[2][0], t.uvs[2][1])};
        // Reference constructor: source/objects/Primitives.h:75-89
        Triangle tri(Vector3(t.p[0][0], t.p[0][1], t.p[0][2]), Vector3(t.p[1][0], t.p[1][1], t.p[1][2]),
                     Vector3(t.p[2][0], t.p[2][1], t.p[2][2]), 0, uvs, t.one_sided != 0);
        tri.material = t.material; // ctor takes int16_t; ingestion overwrites it with the global index anyway
        tri.entIdx = t.ent_idx;
        for (int k = 0; k < 3; k++) {
            tri.normals[k] = glm::vec3(t.normals[k][0], t.normals[k][1], t.normals[k][2]);
            tri.tangents[k] = glm::vec3(t.tangents[k][0], t.tangents[k][1], t.tangents[k][2]);
            tri.alphas[k] = t.alphas[k];
            tri.numBones[k] = 0;
        }
        a.mTriangles[i] = tri;
    }
}

void build_reference_sequence(AccelStruct &a) {
    // Verbatim statement sequence of source/objects/AccelStruct.cpp:762-775
    if (a.mAccelBuilt) {
        delete a.mpIntersector;
        delete a.mpTraverser;
        a.mAccelBuilt = false;
    }
    a.mAccel = BVH();
    bvh::LocallyOrderedClusteringBuilder<BVH, uint32_t> builder(a.mAccel);
    auto [bboxes, centers] = bvh::compute_bounding_boxes_and_centers(a.mTriangles.data(), a.mTriangles.size());
    auto global_bbox = bvh::compute_bounding_boxes_union(bboxes.get(), a.mTriangles.size());
    builder.build(global_bbox, bboxes.get(), centers.get(), a.mTriangles.size());

    bvh::LeafCollapser collapser(a.mAccel);
    collapser.collapse();

    a.mpIntersector = new Intersector(a.mAccel, a.mTriangles.data());
    a.mpTraverser = new Traverser(a.mAccel);
    a.mAccelBuilt = true;
}

} // namespace

extern "C" {

int vtref_max_threads() { return omp_get_max_threads(); }

// Fill the containers the way ingestion would and (optionally) build.
void *vtref_create(const vt_scene *s, int build) {
    auto *rs = new RefScene();
    AccelStruct &a = rs->accel;
    a.mpWorld = nullptr;

    for (uint32_t i = 0; i < s->n_textures; i++) rs->textures.push_back(std::make_unique<MemVTF>(s->textures[i]));
    // Ingestion never leaves baseTexture null (fallback MISSING_TEXTURE, source/objects/AccelStruct.cpp:120,286;
    // TraceResult::CalcShadingData dereferences it unconditionally, TraceResult.cpp:199).  Headless stand-in:
    // a 1x1 opaque white RGBA8888 texture appended after the caller's textures.
    static const uint8_t white[4] = {255, 255, 255, 255};
    vt_texture fb{1, 1, 1, 0, 0, 0, white, 4};
    rs->textures.push_back(std::make_unique<MemVTF>(fb));
    const VisTrace::IVTFTexture *fallback = rs->textures.back().get();
    auto tex = [&](int32_t idx) -> const VisTrace::IVTFTexture * {
        return (idx >= 0 && static_cast<uint32_t>(idx) < s->n_textures) ? rs->textures[idx].get() : nullptr;
    };

    a.mMaterials.resize(s->n_materials);
    for (uint32_t i = 0; i < s->n_materials; i++) {
        const vt_material &m = s->materials[i];
        Material &o = a.mMaterials[i];
        o.path = "synthetic/" + std::to_string(i);
        o.colour = glm::vec4(m.colour[0], m.colour[1], m.colour[2], m.colour[3]);
        o.baseTexture = tex(m.base_texture);
        if (!o.baseTexture) o.baseTexture = fallback;
        o.baseTexMat = to_mat(m.base_tex_mat);
        o.normalMap = tex(m.normal_map);
        o.normalMapMat = to_mat(m.normal_map_mat);
        o.mrao = tex(m.mrao);
        o.baseTexture2 = tex(m.base_texture2);
        o.baseTexMat2 = to_mat(m.base_tex_mat2);
        o.normalMap2 = tex(m.normal_map2);
        o.normalMapMat2 = to_mat(m.normal_map_mat2);
        o.mrao2 = tex(m.mrao2);
        o.blendTexture = tex(m.blend_texture);
        o.blendTexMat = to_mat(m.blend_tex_mat);
        o.maskedBlending = m.masked_blending != 0;
        o.detail = tex(m.detail);
        o.detailMat = to_mat(m.detail_mat);
        o.detailScale = m.detail_scale;
        o.detailBlendFactor = m.detail_blend_factor;
        o.detailBlendMode = static_cast<DetailBlendMode>(m.detail_blend_mode);
        o.detailTint = glm::vec3(m.detail_tint[0], m.detail_tint[1], m.detail_tint[2]);
        o.detailAlphaMaskBaseTexture = m.detail_alpha_mask_base_texture != 0;
        o.texScale = m.tex_scale;
        o.flags = static_cast<MaterialFlags>(m.flags);
        o.surfFlags = static_cast<BSPEnums::SURF>(m.surf_flags);
        o.alphatestreference = m.alphatest_reference;
        o.water = m.water != 0;
    }

    a.mEntities.resize(s->n_entities);
    for (uint32_t i = 0; i < s->n_entities; i++) {
        Entity &e = a.mEntities[i];
        e.rawEntity = nullptr;
        e.id = s->entities[i].id;
        e.colour = glm::vec4(s->entities[i].colour[0], s->entities[i].colour[1], s->entities[i].colour[2],
                             s->entities[i].colour[3]);
    }

    fill_triangles(a, s);

    if (build) build_reference_sequence(a);
    return rs;
}

void vtref_destroy(void *h) { delete static_cast<RefScene *>(h); }

// p0,e1,e2,n,nNorm (15 floats) + lod as derived by the reference constructor.
void vtref_get_tri_derived(void *h, float *out16) {
    AccelStruct &a = static_cast<RefScene *>(h)->accel;
    for (size_t i = 0; i < a.mTriangles.size(); i++) {
        const Triangle &t = a.mTriangles[i];
        float *o = out16 + i * 16;
        for (int k = 0; k < 3; k++) {
            o[k] = t.p0[k];
            o[3 + k] = t.e1[k];
            o[6 + k] = t.e2[k];
            o[9 + k] = t.n[k];
            o[12 + k] = t.nNorm[k];
        }
        o[15] = t.lod;
    }
}

void vtref_get_bvh(void *h, vt_node *nodes, uint64_t *node_count, uint64_t *prim_indices, uint64_t *n_tris) {
    AccelStruct &a = static_cast<RefScene *>(h)->accel;
    static_assert(sizeof(vt_node) == sizeof(BVH::Node), "node layout");
    if (node_count) *node_count = a.mAccel.node_count;
    if (n_tris) *n_tris = a.mTriangles.size();
    if (nodes) std::memcpy(nodes, a.mAccel.nodes.get(), a.mAccel.node_count * sizeof(vt_node));
    if (prim_indices)
        for (size_t i = 0; i < a.mTriangles.size(); i++) prim_indices[i] = a.mAccel.primitive_indices[i];
}

// Every texel of one frame / face (z slice 0) of a VTF file as the reference's own VTFTexture reports it
// (libs/VTFParser/VTFParser.cpp: constructor :13-96 incl. DXT decompression, GetPixel :166-205), in VTF storage order
// (smallest mip first, rows top to bottom).  Returns the number of texels written, -1 when the parser rejects the file,
// -2 when `capacity` (in texels) is too small.
int64_t vtref_vtf_pixels(const uint8_t *file, uint64_t size, uint32_t frame, uint32_t face, float *rgba, uint64_t capacity) {
    VTFTexture tex(file, size);
    if (!tex.IsValid()) return -1;
    uint64_t n = 0;
    for (int m = (int)tex.GetMIPLevels() - 1; m >= 0; m--) {
        const uint32_t w = tex.GetWidth((uint8_t)m), h = tex.GetHeight((uint8_t)m);
        for (uint32_t y = 0; y < h; y++)
            for (uint32_t x = 0; x < w; x++) {
                if (n >= capacity) return -2;
                const VTFPixel p = tex.GetPixel((uint16_t)x, (uint16_t)y, 0, (uint8_t)m, (uint16_t)frame, (uint8_t)face);
                rgba[4 * n + 0] = p.r, rgba[4 * n + 1] = p.g, rgba[4 * n + 2] = p.b, rgba[4 * n + 3] = p.a;
                n++;
            }
    }
    return (int64_t)n;
}

// Experiment hook (tools/reinsertion_probe.py): the library's bvh::ParallelReinsertionOptimizer over the CURRENT hierarchy — how
// much traversal work would a reinsertion pass in the product's builder save?  Returns the SAH cost before and after.
void vtref_reinsertion_optimize(void *h, double *cost_before_after) {
    AccelStruct &a = static_cast<RefScene *>(h)->accel;
    if (!a.mAccelBuilt) throw std::runtime_error("vtref_reinsertion_optimize: nothing built");
    auto sah_cost = [&]() {  // SahBasedAlgorithm::compute_cost (sah_based_algorithm.hpp:20-35; protected there), traversal cost 1
        double cost = 0;
        for (size_t i = 0; i < a.mAccel.node_count; i++) {
            const auto &n = a.mAccel.nodes[i];
            cost += (double)n.bounding_box_proxy().half_area() * (n.is_leaf() ? (double)n.primitive_count : 1.0);
        }
        return cost / (double)a.mAccel.nodes[0].bounding_box_proxy().half_area();
    };
    bvh::ParallelReinsertionOptimizer<BVH> optimizer(a.mAccel);
    if (cost_before_after) cost_before_after[0] = sah_cost();
    optimizer.optimize();
    if (cost_before_after) cost_before_after[1] = sah_cost();
    delete a.mpIntersector;
    delete a.mpTraverser;
    a.mpIntersector = new Intersector(a.mAccel, a.mTriangles.data());
    a.mpTraverser = new Traverser(a.mAccel);
}

// Moved geometry, same topology: rebuild mTriangles from `s` and refit the CURRENT hierarchy with the reference
// library's own bvh::HierarchyRefitter, leaf update exactly as its test drives it (libs/bvh/test/refit_bvh.cpp:79-89).
void vtref_refit(void *h, const vt_scene *s) {
    AccelStruct &a = static_cast<RefScene *>(h)->accel;
    if (s->n_tris != a.mTriangles.size() || !a.mAccelBuilt) throw std::runtime_error("vtref_refit: topology changed / nothing built");
    fill_triangles(a, s);
    bvh::HierarchyRefitter<BVH> refitter(a.mAccel);
    refitter.refit([&](BVH::Node &leaf) {
        auto bbox = bvh::BoundingBox<float>::empty();
        for (size_t i = 0; i < leaf.primitive_count; ++i) {
            auto &triangle = a.mTriangles[a.mAccel.primitive_indices[leaf.first_child_or_primitive + i]];
            bbox.extend(triangle.bounding_box());
        }
        leaf.bounding_box_proxy() = bbox;
    });
    delete a.mpIntersector;
    delete a.mpTraverser;
    a.mpIntersector = new Intersector(a.mAccel, a.mTriangles.data());
    a.mpTraverser = new Traverser(a.mAccel);
}

// Replace the hierarchy with one built elsewhere (same bvh::Bvh<float> form) so the
// reference traverser can be run over the product builder's tree.
void vtref_set_bvh(void *h, const vt_node *nodes, uint64_t node_count, const uint64_t *prim_indices) {
    AccelStruct &a = static_cast<RefScene *>(h)->accel;
    if (a.mAccelBuilt) {
        delete a.mpIntersector;
        delete a.mpTraverser;
        a.mAccelBuilt = false;
    }
    a.mAccel = BVH();
    a.mAccel.nodes = std::make_unique<BVH::Node[]>(node_count);
    std::memcpy(a.mAccel.nodes.get(), nodes, node_count * sizeof(vt_node));
    a.mAccel.node_count = node_count;
    a.mAccel.primitive_indices = std::make_unique<size_t[]>(a.mTriangles.size());
    for (size_t i = 0; i < a.mTriangles.size(); i++) a.mAccel.primitive_indices[i] = prim_indices[i];
    a.mpIntersector = new Intersector(a.mAccel, a.mTriangles.data());
    a.mpTraverser = new Traverser(a.mAccel);
    a.mAccelBuilt = true;
}

static inline void fill_attr(const AccelStruct &a, const vt_ray &r, size_t prim, float t, float u, float v,
                             float coneWidth, float coneAngle, vt_attr &o) {
    // source/objects/AccelStruct.cpp:819-831
    const Triangle &tri = a.mTriangles[prim];
    const Entity &ent = a.mEntities[tri.entIdx];
    const Material &mat = a.mMaterials[tri.material];
    TraceResult res(glm::normalize(glm::vec3(r.dx, r.dy, r.dz)), t, coneWidth, coneAngle, tri, glm::vec2(u, v), ent, mat);
    const glm::vec3 &pos = res.GetPos();
    const glm::vec3 &n = res.GetNormal();
    const glm::vec3 &tg = res.GetTangent();
    const glm::vec3 &bn = res.GetBinormal();
    const glm::vec3 &alb = res.GetAlbedo();
    for (int k = 0; k < 3; k++) {
        o.pos[k] = pos[k];
        o.normal[k] = n[k];
        o.tangent[k] = tg[k];
        o.binormal[k] = bn[k];
        o.geometric_normal[k] = res.geometricNormal[k];
        o.albedo[k] = alb[k];
        o.uvw[k] = res.uvw[k];
    }
    o.distance = res.distance;
    o.alpha = res.GetAlpha();
    o.metalness = res.GetMetalness();
    o.roughness = res.GetRoughness();
    o.base_mip = res.GetBaseMIPLevel();
    o.ent_id = res.entIdx;
    o.submat_idx = res.submatIdx;
    o.tex_uv[0] = res.texUV[0];
    o.tex_uv[1] = res.texUV[1];
    o.flags = (res.frontFacing ? VT_ATTR_FRONT_FACING : 0u) | (res.hitSky ? VT_ATTR_HIT_SKY : 0u) |
              (res.HitWater() ? VT_ATTR_HIT_WATER : 0u);
    o.prim = static_cast<uint32_t>(prim);
}

// Batched loop over the reference's single-ray Traverse body.  stats (nullable) =
// {traversal_steps, intersections} from SingleRayTraverser::Statistics
// (libs/bvh/include/bvh/single_ray_traverser.hpp:132-135).  Returns seconds spent in the loop.
double vtref_traverse(void *h, const vt_ray *rays, uint64_t n, vt_hit *hits, vt_attr *attrs, int threads,
                      uint64_t *stats) {
    AccelStruct &a = static_cast<RefScene *>(h)->accel;
    if (!a.mAccelBuilt) return -1.0;
    if (threads <= 0) threads = omp_get_max_threads();
    uint64_t steps = 0, isects = 0;
    const bool want_stats = stats != nullptr;
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads) reduction(+ : steps, isects)
    for (int64_t i = 0; i < static_cast<int64_t>(n); i++) {
        const vt_ray &r = rays[i];
        // source/objects/AccelStruct.cpp:810-818
        Ray ray(Vector3(r.ox, r.oy, r.oz), Vector3(r.dx, r.dy, r.dz), &a, r.tmin, r.tmax);
        std::optional<Intersector::Result> hit;
        if (want_stats) {
            Traverser::Statistics st;
            hit = a.mpTraverser->traverse(ray, *a.mpIntersector, st);
            steps += st.traversal_steps;
            isects += st.intersections;
        } else {
            hit = a.mpTraverser->traverse(ray, *a.mpIntersector);
        }
        if (hit) {
            if (hits) hits[i] = vt_hit{hit->intersection.t, hit->intersection.u, hit->intersection.v,
                                       static_cast<uint32_t>(hit->primitive_index)};
            if (attrs) fill_attr(a, r, hit->primitive_index, hit->distance(), hit->intersection.u, hit->intersection.v, -1.f, -1.f, attrs[i]);
        } else {
            if (hits) hits[i] = vt_hit{0.f, 0.f, 0.f, VT_MISS};
            if (attrs) {
                std::memset(&attrs[i], 0, sizeof(vt_attr));
                attrs[i].prim = VT_MISS;
            }
        }
    }
    auto t1 = std::chrono::steady_clock::now();
    if (stats) {
        stats[0] = steps;
        stats[1] = isects;
    }
    return std::chrono::duration<double>(t1 - t0).count();
}

// TraceResult for given hits (so the attribute stage can be checked on its own).
void vtref_trace_result(void *h, const vt_ray *rays, const vt_hit *hits, uint64_t n, vt_attr *attrs, int threads) {
    AccelStruct &a = static_cast<RefScene *>(h)->accel;
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads)
    for (int64_t i = 0; i < static_cast<int64_t>(n); i++) {
        if (hits[i].prim == VT_MISS) {
            std::memset(&attrs[i], 0, sizeof(vt_attr));
            attrs[i].prim = VT_MISS;
        } else {
            fill_attr(a, rays[i], hits[i].prim, hits[i].t, hits[i].u, hits[i].v, -1.f, -1.f, attrs[i]);
        }
    }
}

// Same with per-ray texture-LOD cones {coneWidth, coneAngle} (source/objects/AccelStruct.cpp:795-803, :827).
void vtref_trace_result_cones(void *h, const vt_ray *rays, const vt_hit *hits, const float *cones, uint64_t n, vt_attr *attrs,
                              int threads) {
    AccelStruct &a = static_cast<RefScene *>(h)->accel;
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads)
    for (int64_t i = 0; i < static_cast<int64_t>(n); i++) {
        if (hits[i].prim == VT_MISS) {
            std::memset(&attrs[i], 0, sizeof(vt_attr));
            attrs[i].prim = VT_MISS;
        } else {
            fill_attr(a, rays[i], hits[i].prim, hits[i].t, hits[i].u, hits[i].v, cones ? cones[2 * i] : -1.f,
                      cones ? cones[2 * i + 1] : -1.f, attrs[i]);
        }
    }
}

// IVTFTexture::Sample(u, v, mip) through the reference sampler: in = n × (u, v, mip), out = n × rgba.
void vtref_sample(void *h, int tex, const float *uvm, uint64_t n, float *rgba) {
    auto *rs = static_cast<RefScene *>(h);
    const VisTrace::IVTFTexture &t = *rs->textures[tex]; // 3-arg inline overload, include/vistrace/IVTFTexture.h:139-142
    for (uint64_t i = 0; i < n; i++) {
        VisTrace::Pixel p = t.Sample(uvm[i * 3], uvm[i * 3 + 1], uvm[i * 3 + 2]);
        rgba[i * 4] = p.r;
        rgba[i * 4 + 1] = p.g;
        rgba[i * 4 + 2] = p.b;
        rgba[i * 4 + 3] = p.a;
    }
}

// FastNodeIntersector on one node (libs/bvh/include/bvh/node_intersectors.hpp:82-103): out = {entry, exit}.
void vtref_node_intersect(const vt_node *node, const vt_ray *r, float *out2) {
    BVH::Node nd;
    std::memcpy(&nd, node, sizeof(nd));
    Ray ray(Vector3(r->ox, r->oy, r->oz), Vector3(r->dx, r->dy, r->dz), nullptr, r->tmin, r->tmax);
    bvh::FastNodeIntersector<BVH> ni(ray);
    auto d = ni.intersect(nd, ray);
    out2[0] = d.first;
    out2[1] = d.second;
}

// Triangle::intersect on one input triangle without a hierarchy (source/objects/Primitives.h:168-215).
// Returns 1 and fills tuv on a hit.
int vtref_tri_intersect(void *h, uint64_t prim, const vt_ray *r, float *tuv) {
    AccelStruct &a = static_cast<RefScene *>(h)->accel;
    Ray ray(Vector3(r->ox, r->oy, r->oz), Vector3(r->dx, r->dy, r->dz), &a, r->tmin, r->tmax);
    auto hit = a.mTriangles[prim].intersect(ray);
    if (!hit) return 0;
    tuv[0] = hit->t;
    tuv[1] = hit->u;
    tuv[2] = hit->v;
    return 1;
}

// The reference's own SkinTriangle (source/objects/AccelStruct.cpp:66-101; external linkage, no header) over
// vt_tri_in records: construct the Triangle as Model.cpp does, attach the skinning fields, skin.
// out27[i] = {p0, e1, e2, normals[3], tangents[3]} as SkinTriangle left them (the fields the build reads next).
void vtref_skin_triangles(const vt_tri_in *tris, const vt_tri_skin *skin, uint64_t n, const float *bones, const float *binds,
                          uint32_t n_bones, float *out27) {
    std::vector<glm::mat4> vb(n_bones), vbind(n_bones);
    std::memcpy((void *)vb.data(), bones, n_bones * sizeof(glm::mat4));
    std::memcpy((void *)vbind.data(), binds, n_bones * sizeof(glm::mat4));
    for (uint64_t i = 0; i < n; i++) {
        const vt_tri_in &in = tris[i];
        glm::vec2 uvs[3] = {{in.uvs[0][0], in.uvs[0][1]}, {in.uvs[1][0], in.uvs[1][1]}, {in.uvs[2][0], in.uvs[2][1]}};
        Triangle tri(Vector3(in.p[0][0], in.p[0][1], in.p[0][2]), Vector3(in.p[1][0], in.p[1][1], in.p[1][2]),
                     Vector3(in.p[2][0], in.p[2][1], in.p[2][2]), in.material, uvs, in.one_sided != 0);
        for (int v = 0; v < 3; v++) {
            tri.normals[v] = glm::vec3(in.normals[v][0], in.normals[v][1], in.normals[v][2]);
            tri.tangents[v] = glm::vec3(in.tangents[v][0], in.tangents[v][1], in.tangents[v][2]);
            tri.numBones[v] = skin ? skin[i].num_bones[v] : 1;
            for (int k = 0; k < 3; k++) {
                tri.weights[v][k] = skin ? skin[i].weights[v][k] : (k == 0 ? 1.f : 0.f);
                tri.boneIds[v][k] = skin ? skin[i].bone_ids[v][k] : 0;
            }
        }
        if (skin) SkinTriangle(tri, vb, vbind);
        else SkinTriangle(tri, vb[0], vbind[0]);  // the one-bone overload, AccelStruct.cpp:103-108
        float *o = out27 + 27 * i;
        for (int k = 0; k < 3; k++) {
            o[k] = tri.p0[k];
            o[3 + k] = tri.e1[k];
            o[6 + k] = tri.e2[k];
            for (int v = 0; v < 3; v++) {
                o[9 + 3 * v + k] = tri.normals[v][k];
                o[18 + 3 * v + k] = tri.tangents[v][k];
            }
        }
    }
}


// SampleBSDF (source/libraries/BSDF.cpp:770-825) for a BSDFMaterial restricted to the diffuse lobe, prepared per hit by
// PrepShadingData (BSDF.cpp:11-21) from the TraceResult values; the ISampler hands out the caller's numbers in call order.
namespace {
class ScriptedSampler final : public VisTrace::ISampler {
    const float *v;
    int i = 0;

public:
    explicit ScriptedSampler(const float *values) : v(values) {}
    float GetFloat() override { return v[i++]; }
    void GetFloat2D(float &r1, float &r2) override { r1 = v[i++], r2 = v[i++]; }
};
}  // namespace
void vtref_sample_bsdf_diffuse(const vt_attr *attrs, const float *wo3, const float *rnd3, uint64_t n, vt_bsdf_sample *out, int32_t *returned) {
#pragma omp parallel for
    for (int64_t i = 0; i < (int64_t)n; i++) {
        const vt_attr &a = attrs[i];
        BSDFMaterial mat;
        mat.activeLobes = LobeType::DiffuseReflection;
        mat.PrepShadingData(glm::vec3(a.albedo[0], a.albedo[1], a.albedo[2]), a.metalness, a.roughness);
        ScriptedSampler sg(rnd3 + 3 * i);
        BSDFSample res;
        const bool ok = SampleBSDF(mat, &sg, glm::vec3(a.normal[0], a.normal[1], a.normal[2]), glm::vec3(a.tangent[0], a.tangent[1], a.tangent[2]),
                                   glm::vec3(a.binormal[0], a.binormal[1], a.binormal[2]), glm::vec3(wo3[3 * i], wo3[3 * i + 1], wo3[3 * i + 2]), res);
        returned[i] = ok ? 1 : 0;
        out[i].scattered[0] = res.scattered.x, out[i].scattered[1] = res.scattered.y, out[i].scattered[2] = res.scattered.z;
        out[i].pdf = res.pdf;
        out[i].weight[0] = res.weight.x, out[i].weight[1] = res.weight.y, out[i].weight[2] = res.weight.z;
        out[i].lobe = (uint32_t)res.lobe;
    }
}

// ---- model ingestion: the reference's own MDL / VVD / VTX parsers and its BodyGroup / Mesh constructors (source/objects/Model.cpp:11-176).
// Model::Model(path) reads through the game's file system, so the Model object is assembled by hand: raw storage, the MDL member
// constructed from the caller's bytes, and the bind matrices by the statements of Model.cpp:242-254.
namespace {
struct RefModel {
    alignas(Model) unsigned char storage[sizeof(Model)];
    Model *m = nullptr;
    std::vector<BodyGroup *> groups;
    std::vector<glm::mat4> binds;
    bool valid = false;
    ~RefModel() {
        for (BodyGroup *g : groups) delete g;
        if (m) m->mMDL.~MDL();
    }
};
}  // namespace
void *vtref_mdl_open(const uint8_t *mdl, uint64_t mdl_size, const uint8_t *vvd, uint64_t vvd_size, const uint8_t *vtx, uint64_t vtx_size) {
    auto *r = new RefModel();
    std::memset(r->storage, 0, sizeof(r->storage));
    r->m = reinterpret_cast<Model *>(r->storage);
    new (&r->m->mMDL) MDL(mdl, mdl_size, vvd, vvd_size, vtx, vtx_size);
    r->valid = r->m->mMDL.IsValid();
    if (!r->valid) return r;
    const MDL &M = r->m->mMDL;
    for (int i = 0; i < M.GetNumBodyParts(); i++) {  // Model.cpp:230-238
        const MDLStructs::BodyPart *bodypart;
        const VTXStructs::BodyPart *vtxBodypart;
        M.GetBodyPart(i, &bodypart, &vtxBodypart);
        r->groups.push_back(new BodyGroup(r->m, bodypart, vtxBodypart));
        if (!r->groups.back()->IsValid()) r->valid = false;
    }
    for (int i = 0; i < M.GetNumBones(); i++) {  // Model.cpp:242-254
        const MDLStructs::Matrix3x4 &m = M.GetBone(i)->poseToBone;
        glm::mat4x4 bind(m[0][0], m[1][0], m[2][0], 0, m[0][1], m[1][1], m[2][1], 0, m[0][2], m[1][2], m[2][2], 0, m[0][3], m[1][3], m[2][3], 1);
        r->binds.push_back(bind);
    }
    return r;
}
void vtref_mdl_close(void *h) { delete static_cast<RefModel *>(h); }
// out: {valid, body groups, bones, materials, skin families, vertices}
void vtref_mdl_info(void *h, int32_t *out6) {
    auto *r = static_cast<RefModel *>(h);
    std::memset(out6, 0, 6 * sizeof(int32_t));
    out6[0] = r->valid ? 1 : 0;
    if (!r->valid) return;
    const MDL &M = r->m->mMDL;
    out6[1] = M.GetNumBodyParts(), out6[2] = M.GetNumBones(), out6[3] = M.GetNumMaterials(), out6[4] = M.GetNumSkinFamilies(), out6[5] = M.GetNumVertices();
}
int32_t vtref_mdl_bodygroup_values(void *h, uint32_t bodygroup) {
    auto *r = static_cast<RefModel *>(h);
    return (r->valid && bodygroup < r->groups.size()) ? r->groups[bodygroup]->GetNumMeshes() : -1;
}
// Triangles of BodyGroup::GetMesh(value) as the reference's Triangle holds them: tris[i].p = {p0, e1, e2} (NOT three vertices),
// normals / tangents / uvs / alphas / material as stored, skin = numBones / weights / boneIds.  Returns the count (-1: no such mesh).
int64_t vtref_mdl_mesh(void *h, uint32_t bodygroup, uint32_t value, vt_tri_in *tris, vt_tri_skin *skin, uint64_t capacity) {
    auto *r = static_cast<RefModel *>(h);
    if (!r->valid || bodygroup >= r->groups.size() || (int32_t)value >= r->groups[bodygroup]->GetNumMeshes()) return -1;
    const Mesh *mesh = r->groups[bodygroup]->GetMesh((int)value);
    const int32_t n = mesh->GetNumTriangles();
    if (!tris) return n;
    const Triangle *T = mesh->GetTriangles();
    for (int32_t i = 0; i < n && (uint64_t)i < capacity; i++) {
        const Triangle &t = T[i];
        vt_tri_in &o = tris[i];
        vt_tri_skin &sk = skin[i];
        std::memset(&o, 0, sizeof(o));
        std::memset(&sk, 0, sizeof(sk));
        for (int k = 0; k < 3; k++) o.p[0][k] = t.p0[k], o.p[1][k] = t.e1[k], o.p[2][k] = t.e2[k];
        std::memcpy(o.normals, t.normals, sizeof o.normals);
        std::memcpy(o.tangents, t.tangents, sizeof o.tangents);
        std::memcpy(o.uvs, t.uvs, sizeof o.uvs);
        std::memcpy(o.alphas, t.alphas, sizeof o.alphas);
        o.material = (uint32_t)t.material, o.one_sided = t.oneSided;
        for (int j = 0; j < 3; j++) {
            sk.num_bones[j] = t.numBones[j];
            // the reference fills all three slots when the VTX vertex has bones and only slot 0 otherwise (Model.cpp:105-116)
            const int filled = t.numBones[j] == 1 && t.weights[j][0] == 1.f && t.boneIds[j][0] == 0 ? 1 : 3;
            for (int b = 0; b < filled; b++) sk.weights[j][b] = t.weights[j][b], sk.bone_ids[j][b] = t.boneIds[j][b];
        }
    }
    return n;
}
void vtref_mdl_bind_matrices(void *h, float *out16) {
    auto *r = static_cast<RefModel *>(h);
    for (size_t i = 0; i < r->binds.size(); i++) std::memcpy(out16 + 16 * i, &r->binds[i][0][0], 64);
}
int32_t vtref_mdl_material_index(void *h, int32_t skin, int32_t material_id) {  // Model::GetMaterialIdx, Model.cpp:349-357
    auto *r = static_cast<RefModel *>(h);
    return r->m->GetMaterialIdx(skin, material_id);
}
int32_t vtref_mdl_material_path(void *h, int32_t material_id, int32_t dir, char *out, uint64_t cap) {
    auto *r = static_cast<RefModel *>(h);
    const std::string s = std::string(r->m->mMDL.GetMaterialDirectory(dir)) + r->m->mMDL.GetMaterialName(material_id);
    std::snprintf(out, cap, "%s", s.c_str());
    return (int32_t)s.size();
}
// ---------------------------------------------------------------- BSPMap (libs/BSPParser): the checker of vt_bsp_*
// The world half of World::World (source/objects/AccelStruct.cpp:236-414) needs the engine (Material(), IMaterial); what it does with
// the map is restated here statement by statement: texture of every triangle through BSPMap::GetTexture, one material per distinct
// path in order of first use.
namespace {
struct RefBsp {
    BSPMap *map = nullptr;
    std::vector<uint32_t> tri_material;
    std::vector<int16_t> material_texinfo;
    bool textures_ok = true;
    ~RefBsp() { delete map; }
};
}  // namespace
void *vtref_bsp_open(const uint8_t *file, uint64_t size) {
    auto *r = new RefBsp();
    r->map = new BSPMap(file, size);
    if (!r->map->IsValid()) return r;
    std::unordered_map<std::string, size_t> ids;
    const int16_t *textures = r->map->GetTriTextures();
    for (size_t i = 0; i < r->map->GetNumTris(); i++) {
        BSPTexture tex;
        try {
            tex = r->map->GetTexture(textures[i]);  // AccelStruct.cpp:243-249
        } catch (const std::out_of_range &) {
            r->textures_ok = false;
            break;
        }
        const std::string path = tex.path;
        if (ids.find(path) == ids.end()) {
            ids.emplace(path, r->material_texinfo.size());
            r->material_texinfo.push_back(textures[i]);
        }
        r->tri_material.push_back((uint32_t)ids[path]);
    }
    return r;
}
void vtref_bsp_close(void *h) { delete static_cast<RefBsp *>(h); }
// out: {valid, textures resolvable, triangles, materials, static props}
void vtref_bsp_info(void *h, int64_t *out5) {
    auto *r = static_cast<RefBsp *>(h);
    std::memset(out5, 0, 5 * sizeof(int64_t));
    out5[0] = r->map->IsValid() ? 1 : 0;
    if (!out5[0]) return;
    out5[1] = r->textures_ok ? 1 : 0;
    out5[2] = (int64_t)r->map->GetNumTris();
    out5[3] = (int64_t)r->material_texinfo.size();
    out5[4] = r->map->GetNumStaticProps();
}
// BSPMap's arrays packed the way World::World packs them into Triangles (AccelStruct.cpp:390-412): three vertices, normals, tangents,
// uvs, alphas, material, one-sided; binormals (9 floats per triangle) and texinfo indices beside them
void vtref_bsp_triangles(void *h, vt_tri_in *tris, float *binormals, int16_t *texinfo) {
    auto *r = static_cast<RefBsp *>(h);
    const BSPMap &m = *r->map;
    const float *pos = reinterpret_cast<const float *>(m.GetVertices()), *nrm = reinterpret_cast<const float *>(m.GetNormals());
    const float *tan = reinterpret_cast<const float *>(m.GetTangents()), *bin = reinterpret_cast<const float *>(m.GetBinormals());
    const size_t n = r->textures_ok ? m.GetNumTris() : 0;
    for (size_t i = 0; i < n; i++) {
        vt_tri_in &o = tris[i];
        std::memset(&o, 0, sizeof(o));
        std::memcpy(o.p, pos + 9 * i, 36);
        std::memcpy(o.normals, nrm + 9 * i, 36);
        std::memcpy(o.tangents, tan + 9 * i, 36);
        std::memcpy(o.uvs, m.GetUVs() + 6 * i, 24);
        std::memcpy(o.alphas, m.GetAlphas() + 3 * i, 12);
        o.material = r->tri_material[i], o.ent_idx = 0, o.one_sided = 1;
        std::memcpy(binormals + 9 * i, bin + 9 * i, 36);
        texinfo[i] = m.GetTriTextures()[i];
    }
}
int32_t vtref_bsp_material(void *h, uint32_t material, vt_bsp_material *out) {
    auto *r = static_cast<RefBsp *>(h);
    if (material >= r->material_texinfo.size()) return -1;
    const BSPTexture t = r->map->GetTexture(r->material_texinfo[material]);
    std::memset(out, 0, sizeof(*out));
    out->surf_flags = (uint32_t)t.flags, out->texinfo = r->material_texinfo[material], out->width = t.width, out->height = t.height;
    out->reflectivity[0] = t.reflectivity.x, out->reflectivity[1] = t.reflectivity.y, out->reflectivity[2] = t.reflectivity.z;
    std::snprintf(out->path, sizeof(out->path), "%s", t.path);
    return 0;
}
int32_t vtref_bsp_static_prop(void *h, int32_t index, vt_bsp_static_prop *out) {
    auto *r = static_cast<RefBsp *>(h);
    try {
        const BSPStaticProp p = r->map->GetStaticProp(index);
        std::memset(out, 0, sizeof(*out));
        out->pos[0] = p.pos.x, out->pos[1] = p.pos.y, out->pos[2] = p.pos.z;
        out->ang[0] = p.ang.x, out->ang[1] = p.ang.y, out->ang[2] = p.ang.z;
        out->skin = p.skin;
        std::snprintf(out->model, sizeof(out->model), "%s", p.model);
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}
} // extern "C"
