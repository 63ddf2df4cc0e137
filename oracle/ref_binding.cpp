// oracle/ref_binding.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// INTEGRATION.md sections 2-4 as COMPILED code: the binding a VisTrace maintainer would add to the real AccelStruct
// (source/objects/AccelStruct.h:61-86), built against the unmodified reference headers and linked against the product's C ABI
// (libvistrace_b200.so).  It proves that the boundary is a drop-in for the reference's own containers:
//
//   * GpuAccelBinding::Upload           what goes at the end of PopulateAccel (source/objects/AccelStruct.cpp:762-775): the
//                                       reference's mTriangles / mMaterials / mEntities and its freshly built, collapsed bvh::Bvh
//                                       (mAccel.nodes, mAccel.primitive_indices) handed to vt_accel_populate_with_bvh, untouched;
//   * GpuAccelBinding::DecodeTexture    IVTFTexture* -> vt_texture through the PUBLIC interface only (GetWidth / GetHeight /
//                                       GetMIPLevels / GetPixel, include/vistrace/IVTFTexture.h:44-97), so any extension's texture
//                                       class works, not just VTFTexture;
//   * GpuAccelBinding::TraverseBatch    the batched entry next to AccelStruct::Traverse;
//   * vtbind_optimize_check            vt_optimize_bvh on the real mAccel.nodes, the reference's own traverser before and after (host only)
//   * vtbind_selfcheck                  runs the reference's own per-ray statements (AccelStruct.cpp:810-831) and the batched GPU
//                                       call over the same rays and reports every difference.
//
// The scene-filling code (what ingestion does in the game) and the verbatim build sequence come from ref_harness.cpp, included
// here as source so that both shared objects are built from one definition.
#include "ref_harness.cpp"

#include "../include/vistrace_b200_ivtf.hpp"

#include <map>

namespace {

struct GpuAccelBinding {
    vt_accel *mpGpu = nullptr;  // owned; freed in the destructor and replaced on rebuild (AccelStruct.cpp:525-542 pattern)
    std::vector<std::vector<uint8_t>> mTexelStorage;
    std::string error;

    ~GpuAccelBinding() {
        if (mpGpu) vt_accel_destroy(mpGpu);
    }

    // IVTFTexture -> RGBA8888 mip chain through the public interface only: the product's header-only helper
    // (include/vistrace_b200_ivtf.hpp), instantiated here with the reference's own VisTrace::IVTFTexture
    vt_texture DecodeTexture(const VisTrace::IVTFTexture *tex) {
        mTexelStorage.emplace_back();
        return vt::decode_ivtf_texture(tex, mTexelStorage.back());
    }

    static void mat_out(const glm::mat2x4 &m, float out[8]) {
        for (int c = 0; c < 2; c++)
            for (int r = 0; r < 4; r++) out[4 * c + r] = m[c][r];
    }

    // after LeafCollapser::collapse
    bool Upload(AccelStruct &a, const std::map<const VisTrace::IVTFTexture *, uint32_t> &tex_flags) {
        std::vector<vt_tri_in> tris(a.mTriangles.size());
        for (size_t i = 0; i < a.mTriangles.size(); i++) {
            const Triangle &t = a.mTriangles[i];
            vt_tri_in &o = tris[i];
            std::memset(&o, 0, sizeof(o));
            const Vector3 p1 = t.p1(), p2 = t.p2();  // Primitives.h:104-105
            for (int k = 0; k < 3; k++) o.p[0][k] = t.p0[k], o.p[1][k] = p1[k], o.p[2][k] = p2[k];
            std::memcpy(o.normals, t.normals, sizeof o.normals);  // glm::vec3[3] == float[3][3]
            std::memcpy(o.tangents, t.tangents, sizeof o.tangents);
            std::memcpy(o.uvs, t.uvs, sizeof o.uvs);
            std::memcpy(o.alphas, t.alphas, sizeof o.alphas);
            o.material = (uint32_t)t.material, o.ent_idx = t.entIdx, o.one_sided = t.oneSided;
        }
        // textures: one vt_texture per distinct IVTFTexture* the materials reference
        mTexelStorage.clear();
        std::vector<vt_texture> texs;
        std::map<const VisTrace::IVTFTexture *, int32_t> index;
        auto slot = [&](const VisTrace::IVTFTexture *p) -> int32_t {
            if (!p) return -1;
            auto it = index.find(p);
            if (it != index.end()) return it->second;
            vt_texture t = DecodeTexture(p);
            auto fl = tex_flags.find(p);
            if (fl != tex_flags.end()) t.flags = fl->second;
            texs.push_back(t);
            return index[p] = (int32_t)texs.size() - 1;
        };
        std::vector<vt_material> mats(a.mMaterials.size());
        for (size_t i = 0; i < mats.size(); i++) {
            const Material &m = a.mMaterials[i];
            vt_material &o = mats[i];
            std::memset(&o, 0, sizeof(o));
            o.flags = (uint32_t)m.flags, o.surf_flags = (uint32_t)m.surfFlags;
            o.alphatest_reference = m.alphatestreference, o.tex_scale = m.texScale;
            for (int k = 0; k < 4; k++) o.colour[k] = m.colour[k];
            mat_out(m.baseTexMat, o.base_tex_mat), mat_out(m.baseTexMat2, o.base_tex_mat2);
            mat_out(m.normalMapMat, o.normal_map_mat), mat_out(m.normalMapMat2, o.normal_map_mat2);
            mat_out(m.blendTexMat, o.blend_tex_mat), mat_out(m.detailMat, o.detail_mat);
            o.detail_scale = m.detailScale, o.detail_blend_factor = m.detailBlendFactor;
            for (int k = 0; k < 3; k++) o.detail_tint[k] = m.detailTint[k];
            o.base_texture = slot(m.baseTexture), o.base_texture2 = slot(m.baseTexture2);
            o.normal_map = slot(m.normalMap), o.normal_map2 = slot(m.normalMap2);
            o.mrao = slot(m.mrao), o.mrao2 = slot(m.mrao2);
            o.blend_texture = slot(m.blendTexture), o.detail = slot(m.detail);
            o.detail_blend_mode = (uint8_t)m.detailBlendMode, o.masked_blending = m.maskedBlending;
            o.detail_alpha_mask_base_texture = m.detailAlphaMaskBaseTexture, o.water = m.water;
        }
        for (vt_texture &t : texs) t.rgba = nullptr;  // storage may have moved while the list grew: re-point
        for (size_t i = 0; i < texs.size(); i++) texs[i].rgba = mTexelStorage[i].data();
        std::vector<vt_entity> ents(a.mEntities.size());
        for (size_t i = 0; i < ents.size(); i++) {
            ents[i].id = a.mEntities[i].id;
            for (int k = 0; k < 4; k++) ents[i].colour[k] = a.mEntities[i].colour[k];
        }
        vt_scene scene{tris.data(), tris.size(), mats.data(), (uint32_t)mats.size(), ents.data(), (uint32_t)ents.size(), texs.data(), (uint32_t)texs.size()};
        if (!mpGpu) mpGpu = vt_accel_create(/*device*/ 0);
        if (!mpGpu) return fail();
        std::vector<uint64_t> prim(a.mTriangles.size());
        for (size_t i = 0; i < prim.size(); i++) prim[i] = a.mAccel.primitive_indices[i];
        static_assert(sizeof(vt_node) == sizeof(BVH::Node), "vt_node is bit-compatible with bvh::Bvh<float>::Node");
        if (vt_accel_populate_with_bvh(mpGpu, &scene, reinterpret_cast<const vt_node *>(a.mAccel.nodes.get()), a.mAccel.node_count, prim.data()) != 0)
            return fail();
        return true;
    }

    // INTEGRATION.md section 2
    int TraverseBatch(const vt_ray *rays, size_t n, vt_hit *hits, vt_attr *attrs, uint32_t flags) {
        return vt_accel_traverse(mpGpu, rays, n, hits, attrs, flags, nullptr);
    }
    bool fail() {
        error = vt_last_error();
        return false;
    }
};

}  // namespace

extern "C" {

// layout: 0 exact (the reference's visit order: whole hit buffers must be byte-identical), 2 quad (default; up to counted ties).
// report: [0] rays, [1] hit-record bytes differing, [2] hit/miss mismatches, [3] t/u/v bit mismatches on the same primitive,
//         [4] primitive mismatches, [5] attr records compared, [6] textures decoded through IVTFTexture, [7] triangles uploaded.
// max_attr_err: largest relative error over pos / normal / tangent / binormal / albedo / alpha / uvw / tex_uv of the hit records.
int vtbind_selfcheck(const vt_scene *s, const vt_ray *rays, uint64_t n, int layout, uint64_t *report, double *max_attr_err, char *err, uint64_t err_cap) {
    auto say = [&](const std::string &m) {
        if (err && err_cap) std::snprintf(err, err_cap, "%s", m.c_str());
        return 1;
    };
    std::unique_ptr<RefScene> rs(static_cast<RefScene *>(vtref_create(s, 1)));  // ingestion stand-in + the reference's own build sequence
    AccelStruct &a = rs->accel;
    std::map<const VisTrace::IVTFTexture *, uint32_t> tex_flags;  // TEXTURE_FLAGS are not part of IVTFTexture: carried on the side
    for (uint32_t i = 0; i < s->n_textures; i++) tex_flags[rs->textures[i].get()] = s->textures[i].flags;
    GpuAccelBinding gpu;
    gpu.mpGpu = vt_accel_create(0);
    if (!gpu.mpGpu) return say(std::string("vt_accel_create: ") + vt_last_error());
    if (vt_accel_set_layout(gpu.mpGpu, layout) != 0) return say(vt_last_error());
    if (!gpu.Upload(a, tex_flags)) return say("upload: " + gpu.error);
    if (vt_accel_get_layout(gpu.mpGpu) != layout) return say("the engine fell back to another node layout");
    std::vector<vt_hit> got(n);
    std::vector<vt_attr> got_attr(n);
    if (gpu.TraverseBatch(rays, n, got.data(), got_attr.data(), 0) != 0) return say(std::string("TraverseBatch: ") + vt_last_error());
    std::vector<vt_hit> want(n);
    std::vector<vt_attr> want_attr(n);
    vtref_traverse(rs.get(), rays, n, want.data(), want_attr.data(), 0, nullptr);  // the statements of AccelStruct.cpp:810-831 per ray
    std::memset(report, 0, 8 * sizeof(uint64_t));
    report[0] = n, report[6] = gpu.mTexelStorage.size(), report[7] = a.mTriangles.size();
    double worst = 0.0;
    for (uint64_t i = 0; i < n; i++) {
        const vt_hit &g = got[i], &w = want[i];
        if (std::memcmp(&g, &w, sizeof(g)) != 0) report[1]++;
        const bool gm = g.prim == VT_MISS, wm = w.prim == VT_MISS;
        if (gm != wm) {
            report[2]++;
            continue;
        }
        if (gm) continue;
        if (g.prim != w.prim) {
            report[4]++;
            continue;
        }
        if (std::memcmp(&g.t, &w.t, 12) != 0) report[3]++;
        report[5]++;
        const vt_attr &x = got_attr[i], &y = want_attr[i];
        auto rel = [&](const float *p, const float *q, int k) {
            double num = 0, den = 0;
            for (int c = 0; c < k; c++) num = std::max(num, (double)std::fabs(p[c] - q[c])), den += (double)q[c] * q[c];
            worst = std::max(worst, num / std::max(std::sqrt(den), 1e-6));
        };
        rel(x.pos, y.pos, 3), rel(x.normal, y.normal, 3), rel(x.tangent, y.tangent, 3), rel(x.binormal, y.binormal, 3);
        rel(x.albedo, y.albedo, 3), rel(&x.alpha, &y.alpha, 1), rel(x.uvw, y.uvw, 3), rel(x.tex_uv, y.tex_uv, 2);
        if (x.ent_id != y.ent_id || x.submat_idx != y.submat_idx || x.flags != y.flags) worst = 1e30;
    }
    if (max_attr_err) *max_attr_err = worst;
    return 0;
}

// INTEGRATION.md section 5, "Builder options": vt_optimize_bvh in place on the REAL bvh::Bvh<float> of the reference's AccelStruct
// (its own PLOC + LeafCollapser build), then the reference's own traverser over its own — now optimised — containers.  Host only.
// report: [0] rays, [1] hit/miss mismatches, [2] t/u/v bit mismatches, [3] primitive mismatches (exact ties: same t), [4] reinsertions
//         applied, [5] traversal steps before, [6] after, [7] node count.  areas: inner-node area before / after.
int vtbind_optimize_check(const vt_scene *s, const vt_ray *rays, uint64_t n, int iterations, double fraction, uint64_t *report, double *areas,
                          char *err, uint64_t err_cap) {
    auto say = [&](const std::string &m) {
        if (err && err_cap) std::snprintf(err, err_cap, "%s", m.c_str());
        return 1;
    };
    std::unique_ptr<RefScene> rs(static_cast<RefScene *>(vtref_create(s, 1)));
    AccelStruct &a = rs->accel;
    std::vector<vt_hit> before(n), after(n);
    uint64_t st0[2] = {0, 0}, st1[2] = {0, 0};
    vtref_traverse(rs.get(), rays, n, before.data(), nullptr, 0, st0);
    static_assert(sizeof(vt_node) == sizeof(BVH::Node), "vt_node is bit-compatible with bvh::Bvh<float>::Node");
    uint64_t moves = 0;
    if (vt_optimize_bvh(reinterpret_cast<vt_node *>(a.mAccel.nodes.get()), a.mAccel.node_count, iterations, fraction, &areas[0], &areas[1], &moves) != 0)
        return say(std::string("vt_optimize_bvh: ") + vt_last_error());
    vtref_traverse(rs.get(), rays, n, after.data(), nullptr, 0, st1);
    std::memset(report, 0, 8 * sizeof(uint64_t));
    report[0] = n, report[4] = moves, report[5] = st0[0], report[6] = st1[0], report[7] = a.mAccel.node_count;
    for (uint64_t i = 0; i < n; i++) {
        const vt_hit &g = after[i], &w = before[i];
        if ((g.prim == VT_MISS) != (w.prim == VT_MISS)) report[1]++;
        else if (g.prim != VT_MISS && std::memcmp(&g.t, &w.t, 4) != 0) report[2]++;
        else if (g.prim != w.prim) report[3]++;
    }
    return 0;
}

}  // extern "C"
