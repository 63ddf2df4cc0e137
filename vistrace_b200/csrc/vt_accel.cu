// vt_accel.cu — host objects + the extern "C" boundary of include/vistrace_b200.h.
//
// Host code stays C++ (as in the reference); the kernels are reached only from here.  There is
// NO CPU fallback: without a CUDA device vt_accel_create fails and every later call errors out.
#include <limits>
#include <cuda_runtime.h>

#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <omp.h>

#include "vt_accel_internal.h"
#include "vt_math.cuh"

namespace vt {

thread_local std::string g_last_error;

// ------------------------------------------------------------------------------ Triangle
Triangle::Triangle(const float p0_[3], const float p1[3], const float p2[3], uint32_t material_, const float uvs_[3][2],
                   bool oneSided_)
    : oneSided(oneSided_), material(material_), entIdx(0), lod(0.f) {
    for (int k = 0; k < 3; k++) {
        p0[k] = p0_[k];
        e1[k] = p0_[k] - p1[k];  // e1 = p0 - p1, e2 = p2 - p0  (Primitives.h:82)
        e2[k] = p2[k] - p0_[k];
    }
    std::memcpy(uvs, uvs_, sizeof(uvs));
    std::memset(normals, 0, sizeof(normals));
    std::memset(tangents, 0, sizeof(tangents));
    alphas[0] = alphas[1] = alphas[2] = 0.f;
    ComputeNormalAndLoD();
}

// Primitives.h:91-102.  Compiled with -ffp-contract=off / --fmad=false: every product and sum
// below rounds separately, as in the reference build.
void Triangle::ComputeNormalAndLoD() {
    n[0] = e1[1] * e2[2] - e1[2] * e2[1];  // cross(e1, e2), LeftHandedNormal (vector.hpp:159-167)
    n[1] = e1[2] * e2[0] - e1[0] * e2[2];
    n[2] = e1[0] * e2[1] - e1[1] * e2[0];
    const float uv10x = uvs[1][0] - uvs[0][0], uv10y = uvs[1][1] - uvs[0][1];
    const float uv20x = uvs[2][0] - uvs[0][0], uv20y = uvs[2][1] - uvs[0][1];
    const float triUVArea = std::fabs(uv10x * uv20y - uv20x * uv10y);
    float d = n[0] * n[0];  // bvh::dot / bvh::length (vector.hpp:134-147)
    d += n[1] * n[1];
    d += n[2] * n[2];
    const float len = std::sqrt(d);
    lod = 0.5f * std::log2(triUVArea / len);
    for (int k = 0; k < 3; k++) nNorm[k] = n[k] / len;
}

// ------------------------------------------------------------------------------ SkinTriangle
// glm::mat4 arithmetic in glm's operation order (libs/glm/glm/detail/type_mat4x4.inl), host side,
// compiled with -ffp-contract=off so every product and sum rounds separately as in the reference build.
namespace {
struct M4 {
    float c[4][4];  // c[column][row]
};
// operator*(mat4, mat4), type_mat4x4.inl:630-648: Result[j] = ((A0*B[j][0] + A1*B[j][1]) + A2*B[j][2]) + A3*B[j][3]
M4 mul(const M4 &a, const M4 &b) {
    M4 r;
    for (int j = 0; j < 4; j++)
        for (int k = 0; k < 4; k++) r.c[j][k] = ((a.c[0][k] * b.c[j][0] + a.c[1][k] * b.c[j][1]) + a.c[2][k] * b.c[j][2]) + a.c[3][k] * b.c[j][3];
    return r;
}
// operator*(mat4, vec4), type_mat4x4.inl:561-572: (m[0]*v0 + m[1]*v1) + (m[2]*v2 + m[3]*v3)
void mul(const M4 &m, const float v[4], float out[4]) {
    for (int k = 0; k < 4; k++) out[k] = (m.c[0][k] * v[0] + m.c[1][k] * v[1]) + (m.c[2][k] * v[2] + m.c[3][k] * v[3]);
}
// TransformToBone, source/objects/AccelStruct.cpp:33-47
void transform_to_bone(const float vec[3], const M4 *bones, const M4 *binds, uint32_t n_bones, uint8_t num, const float *weights,
                       const int8_t *ids, bool angle_only, float out[3]) {
    float fin[4] = {0.f, 0.f, 0.f, 0.f};
    const float vertex[4] = {vec[0], vec[1], vec[2], angle_only ? 0.f : 1.f};
    for (uint8_t i = 0; i < num; i++) {
        const int8_t b = ids[i];
        if (b < 0 || (uint32_t)b >= n_bones) throw std::runtime_error("skin_triangles: bone id out of range");
        float t[4];
        mul(mul(bones[b], binds[b]), vertex, t);  // (bones * binds) * vertex ...
        for (int k = 0; k < 4; k++) fin[k] += t[k] * weights[i];  // ... * weight, accumulated
    }
    out[0] = fin[0], out[1] = fin[1], out[2] = fin[2];
}
}  // namespace

void SkinTriangles(vt_tri_in *tris, const vt_tri_skin *skin, uint64_t n, const float *bones_, const float *binds_, uint32_t n_bones) {
    if (n == 0) return;
    if (!tris || !bones_ || !binds_ || n_bones == 0) throw std::runtime_error("skin_triangles: null argument");
    const M4 *bones = reinterpret_cast<const M4 *>(bones_), *binds = reinterpret_cast<const M4 *>(binds_);
    vt_tri_skin one;
    std::memset(&one, 0, sizeof(one));
    for (int v = 0; v < 3; v++) {
        one.num_bones[v] = 1;
        one.weights[v][0] = 1.f;
    }
    std::string error;
#pragma omp parallel for
    for (int64_t i = 0; i < (int64_t)n; i++) {
        vt_tri_in &t = tris[i];
        const vt_tri_skin &sk = skin ? skin[i] : one;
        // the vertices SkinTriangle sees: p0, p0 - e1, p0 + e2 with the constructor's rounded edges (AccelStruct.cpp:68-72)
        float pos[3][3];
        for (int k = 0; k < 3; k++) {
            const float e1 = t.p[0][k] - t.p[1][k], e2 = t.p[2][k] - t.p[0][k];
            pos[0][k] = t.p[0][k];
            pos[1][k] = t.p[0][k] - e1;
            pos[2][k] = t.p[0][k] + e2;
        }
        try {
            for (int v = 0; v < 3; v++) {
                float o[3];
                transform_to_bone(pos[v], bones, binds, n_bones, sk.num_bones[v], sk.weights[v], sk.bone_ids[v], false, o);
                std::memcpy(t.p[v], o, 12);
                transform_to_bone(t.normals[v], bones, binds, n_bones, sk.num_bones[v], sk.weights[v], sk.bone_ids[v], true, o);
                std::memcpy(t.normals[v], o, 12);
                transform_to_bone(t.tangents[v], bones, binds, n_bones, sk.num_bones[v], sk.weights[v], sk.bone_ids[v], true, o);
                std::memcpy(t.tangents[v], o, 12);
            }
        } catch (const std::exception &e) {
#pragma omp critical
            error = e.what();
        }
    }
    if (!error.empty()) throw std::runtime_error(error);
}

// ----------------------------------------------------------------------------- AccelStruct
namespace {
struct PhaseTimer {  // VT_TIMING=1: one stderr line per host phase of a populate (rebuild latency)
    bool on = env_int("VT_TIMING", 0) != 0;
    double t0 = omp_get_wtime();
    void lap(const char *what) {
        if (!on) return;
        const double t = omp_get_wtime();
        std::fprintf(stderr, "[populate] %-28s %.3f s\n", what, t - t0);
        t0 = t;
    }
};
}  // namespace

double AccelStruct::RefitQuality() const {
    const DeviceScene &D = *mpDevice;
    if (!D.refit_ready || !(D.refit_cost_built > 0.0) || !(D.refit_cost_now > 0.0)) return 1.0;  // never refitted on the device
    return D.refit_cost_now / D.refit_cost_built;
}

static void check_built(bool built) {
    // source/objects/AccelStruct.cpp:780
    if (!built) throw std::runtime_error("Unable to perform traversal, acceleration structure invalid (use AccelStruct:Rebuild to rebuild it)");
}

AccelStruct::AccelStruct(int device) : mDevice(device) {
    if (const char *v = std::getenv("VT_LAYOUT")) {  // default for handles that do not call SetLayout
        const std::string l(v);
        if (l == "exact") mWantLayout = VT_LAYOUT_EXACT;
        else if (l == "compact") mWantLayout = VT_LAYOUT_COMPACT;
        else if (l == "quad") mWantLayout = VT_LAYOUT_QUAD;
        else if (!l.empty()) throw std::runtime_error("VT_LAYOUT must be 'exact', 'compact' or 'quad'");
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        throw std::runtime_error(std::string("no CUDA device (the engine has no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= count) throw std::runtime_error("device index out of range");
    VT_CUDA(cudaSetDevice(device));
    mpDevice = new DeviceScene();
    cudaDeviceProp prop;
    VT_CUDA(cudaGetDeviceProperties(&prop, device));
    mpDevice->sm_count = prop.multiProcessorCount;
    mpDevice->counters.ensure(kCounterSlots * 2);
    VT_CUDA(cudaMemset(mpDevice->counters.p, 0, kCounterSlots * 2 * sizeof(unsigned long long)));
    VT_CUDA(cudaStreamCreateWithFlags(&mpDevice->own_stream, cudaStreamNonBlocking));
}

void AccelStruct::JoinAttrUpload() {
    if (mAttrUpload.joinable()) mAttrUpload.join();
    RawVector<VtTriAttr>().swap(mAttrStage);
    if (!mAttrUploadError.empty()) {
        const std::string e = mAttrUploadError;
        mAttrUploadError.clear();
        throw std::runtime_error(e);
    }
}

AccelStruct::~AccelStruct() {
    DrainWaveFrames();
    if (mAttrUpload.joinable()) mAttrUpload.join();
    if (mpDevice) {
        cudaSetDevice(mDevice);
        delete mpDevice;
    }
}

uint64_t AccelStruct::DeviceBytes() const { return mpDevice ? mpDevice->scene_bytes() : 0; }

void AccelStruct::Ingest(const vt_scene &scene, bool stage_attrs) {
    // PopulateAccel prologue (source/objects/AccelStruct.cpp:537-556): drop the old structure, refill containers
    mAccelBuilt = false;
    mTriangles.clear();
    mEntities.clear();
    mMaterials.clear();
    if (scene.n_tris > 0xFFFFFFF0ull) throw std::runtime_error("too many triangles (Bvh::IndexType is uint32_t, bvh.hpp:20)");
    if (scene.n_tris && !scene.tris) throw std::runtime_error("scene.tris is null");
    if (scene.n_materials == 0 || scene.n_entities == 0) throw std::runtime_error("scene needs at least one material and one entity");
    if (scene.n_materials >= (1u << 30)) throw std::runtime_error("too many materials");
    mMaterials.assign(scene.materials, scene.materials + scene.n_materials);
    mEntities.assign(scene.entities, scene.entities + scene.n_entities);
    for (const Material &m : mMaterials) {
        const int32_t slots[8] = {m.base_texture, m.base_texture2, m.normal_map, m.normal_map2, m.mrao, m.mrao2, m.blend_texture, m.detail};
        for (int32_t s : slots)
            if (s >= (int32_t)scene.n_textures) throw std::runtime_error("material references a texture index past n_textures");
    }
    for (uint32_t i = 0; i < scene.n_textures; i++) {
        const vt_texture &t = scene.textures[i];
        if (t.width == 0 || t.height == 0 || t.mip_count == 0 || t.mip_count > 16 || !t.rgba)
            throw std::runtime_error("texture " + std::to_string(i) + ": bad header");
        if (t.texel_layout != 0 && ((t.texel_layout & ~0xFFu) != VT_TEXEL_WIDE))
            throw std::runtime_error("texture " + std::to_string(i) + ": unknown texel layout");
        for (int c = 0; c < 4; c++)
            if (t.texel_layout && ((t.texel_layout >> (2 * c)) & 3u) == 3u) throw std::runtime_error("texture " + std::to_string(i) + ": unknown divisor code");
        uint64_t need = 0;
        for (uint32_t m = 0; m < t.mip_count; m++) need += (uint64_t)std::max(1, t.width >> m) * std::max(1, t.height >> m) * (t.texel_layout ? 8 : 4);
        if (need != t.nbytes) throw std::runtime_error("texture " + std::to_string(i) + ": nbytes does not match the mip chain of its texel layout");
    }
    if (mAttrUpload.joinable()) mAttrUpload.join();
    mTriangles.resize(scene.n_tris);
    mAttrStage.resize(stage_attrs ? scene.n_tris : 0);
    bool bad = false;
#pragma omp parallel for
    for (int64_t i = 0; i < (int64_t)scene.n_tris; i++) {
        const vt_tri_in &in = scene.tris[i];
        Triangle t(in.p[0], in.p[1], in.p[2], in.material, in.uvs, in.one_sided != 0);
        std::memcpy(t.normals, in.normals, sizeof(t.normals));
        std::memcpy(t.tangents, in.tangents, sizeof(t.tangents));
        std::memcpy(t.alphas, in.alphas, sizeof(t.alphas));
        t.entIdx = in.ent_idx;
        if (in.material >= scene.n_materials || in.ent_idx >= scene.n_entities) bad = true;
        mTriangles[i] = t;
        if (!stage_attrs) continue;
        VtTriAttr &a = mAttrStage[i];  // everything TraceResult::TraceResult copies (TraceResult.cpp:58-78), original order
        std::memset(&a, 0, sizeof(a));
        std::memcpy(a.p0, t.p0, 12);
        std::memcpy(a.e1, t.e1, 12);
        std::memcpy(a.e2, t.e2, 12);
        std::memcpy(a.nNorm, t.nNorm, 12);
        std::memcpy(a.normals, t.normals, 36);
        std::memcpy(a.tangents, t.tangents, 36);
        std::memcpy(a.uvs, t.uvs, 24);
        std::memcpy(a.alphas, t.alphas, 12);
        a.lod = t.lod;
        a.material = t.material;
        a.ent_idx = t.entIdx;
    }
    if (bad) {
        RawVector<VtTriAttr>().swap(mAttrStage);
        throw std::runtime_error("triangle references a material or entity out of range");
    }
    if (!stage_attrs) return;
    // async upload: the attribute records go up while the caller builds / flattens the hierarchy; Upload() joins
    mAttrUploadError.clear();
    mAttrUpload = std::thread([this] {
        try {
            VT_CUDA(cudaSetDevice(mDevice));
            mpDevice->attrs.upload(mAttrStage.data(), mAttrStage.size());
        } catch (const std::exception &e) {
            mAttrUploadError = e.what();
        }
    });
}

void AccelStruct::Upload(const vt_scene &scene) {
    VT_CUDA(cudaSetDevice(mDevice));
    DeviceScene &D = *mpDevice;
    D.cfg.persistent = env_int("VT_PERSISTENT", 1);
    D.cfg.refill_threshold = env_int("VT_REFILL", D.cfg.refill_threshold);
    D.cfg.tri_threshold = env_int("VT_TRI_ROUND", D.cfg.tri_threshold);
    D.cfg.tail_share = env_int("VT_TAIL_SHARE", D.cfg.tail_share);
    const size_t n = mTriangles.size();

    // Node layout (include/vistrace_b200.h: VT_LAYOUT_*).  A tree the quantised layouts cannot hold (a leaf of
    // more than 15 triangles, non-finite bounds, too deep) falls back to the exact layout.
    PhaseTimer timer;
    int layout = mWantLayout;
    uint32_t smem_pairs = layout == VT_LAYOUT_EXACT ? (uint32_t)env_int("VT_SMEM_PAIRS", 0) : 0u;  // exact only: breadth-first prefix staged in smem
    FlatBvh flat;
    QuadBvh quad;
    std::vector<VtCPair> cpairs;
    std::string err;
    if (layout == VT_LAYOUT_QUAD && !build_quads(mAccel, n, quad, err)) layout = VT_LAYOUT_EXACT;
    {   // child order of the quad kernel (vt_traverse.cu: VT_KEY_MID): by entry + exit where sibling boxes overlap, by entry elsewhere
        const char *order = std::getenv("VT_KEY_ORDER");
        const std::string o = order ? order : "auto";
        D.cfg.key_mid = o == "mid" ? 1 : o == "entry" ? 0 : (quad.sibling_overlap > (double)env_float("VT_COLLAPSE_OVERLAP", 0.225f) ? 1 : 0);
    }
#if VT_SMEM_QUADS_BUILD
    if (layout == VT_LAYOUT_QUAD) smem_pairs = quads_top_first(quad, (uint32_t)std::max(0, env_int("VT_SMEM_QUADS", 0)));  // A/B builds only
#endif
    if (layout != VT_LAYOUT_QUAD) {
        if (!flatten_bvh(mAccel, n, smem_pairs, flat, err)) throw std::runtime_error(err);
        smem_pairs = (uint32_t)std::min<size_t>(smem_pairs, flat.pairs.size());
        // tagged references (vt_traverse.cu): 4 bits of leaf count + 28 bits of pair index / triangle slot
        if (n >= (1u << 28) - 16 || flat.pairs.size() >= (1u << 28) - 16 || flat.root_leaf_count > 15) layout = VT_LAYOUT_EXACT;
        if (layout == VT_LAYOUT_COMPACT && !compact_pairs(flat.pairs, cpairs, err)) {
            layout = VT_LAYOUT_EXACT;
            cpairs.clear();
        }
    }
    timer.lap("  node layout (collapse/quantise)");
    mLayout = layout;
    const uint32_t *leaf_order = layout == VT_LAYOUT_QUAD ? quad.leaf_order.data() : flat.leaf_order.data();
    const uint32_t root_leaf_count = layout == VT_LAYOUT_QUAD ? quad.root_leaf_count : flat.root_leaf_count;
    const uint32_t n_inner = layout == VT_LAYOUT_QUAD ? (uint32_t)quad.quads.size() : (uint32_t)flat.pairs.size();

    // leaf-order geometry records + UVs; original-order attribute records
    // quad layout: one extra all-NaN record behind the last triangle, the target of empty child slots (vt_device.h)
    const bool sentinel = VT_EMPTY_SENTINEL && layout == VT_LAYOUT_QUAD;
    RawVector<VtTriRec> recs(n + (sentinel ? 1 : 0));  // every byte is written by the loops below: no zero-fill (vt_host.h)
    RawVector<float> uv((n + (sentinel ? 1 : 0)) * 6);
    if (sentinel) {
        VtTriRec &r = recs[n];
        const float nan = std::numeric_limits<float>::quiet_NaN();
        for (int k = 0; k < 3; k++) r.p0[k] = r.e1[k] = r.e2[k] = r.n[k] = nan;  // every accept test is NaN-rejecting (Primitives.h:184-189)
        r.matflags = 0;
        r.orig = VT_MISS;
        r.pad[0] = r.pad[1] = 0;
        const uint32_t ref = (1u << 28) | (uint32_t)n;
        for (VtQuad &q : quad.quads)
            for (int i = 0; i < 4; i++)
                if (q.ref[i] == 0xFFFFFFFFu) q.ref[i] = ref;
    }
    uint32_t any_alpha = 0;
#pragma omp parallel for reduction(| : any_alpha)
    for (int64_t s = 0; s < (int64_t)n; s++) {
        const uint32_t orig = leaf_order[s];
        const Triangle &t = mTriangles[orig];
        const Material &m = mMaterials[t.material];
        VtTriRec &r = recs[s];
        for (int k = 0; k < 3; k++) {
            r.p0[k] = t.p0[k];
            r.e1[k] = t.e1[k];
            r.e2[k] = t.e2[k];
            r.n[k] = t.n[k];
        }
        uint32_t fl = 0;
        if (t.oneSided && (m.flags & VT_MATFLAG_NOCULL) == 0) fl |= VT_TRI_FLAG_CULL;  // Primitives.h:174
        if (m.flags & VT_MATFLAG_ALPHATEST) fl |= VT_TRI_FLAG_ALPHATEST;               // Primitives.h:195
        any_alpha |= (fl & VT_TRI_FLAG_ALPHATEST);
        r.matflags = (t.material << 2) | fl;
        r.orig = orig;
        r.pad[0] = r.pad[1] = 0;
        std::memcpy(&uv[6 * s], t.uvs, 6 * sizeof(float));
    }
    // materials / entities / textures (+ the 1x1 white stand-in for a null baseTexture: ingestion never
    // leaves it null — fallback MISSING_TEXTURE, source/objects/AccelStruct.cpp:120,286)
    std::vector<VtDevMaterial> dm(mMaterials.size());
    for (size_t i = 0; i < mMaterials.size(); i++) {
        const Material &m = mMaterials[i];
        VtDevMaterial &o = dm[i];
        std::memset(&o, 0, sizeof(o));
        o.flags = m.flags;
        o.surf_flags = m.surf_flags;
        o.alphatest_reference = m.alphatest_reference;
        o.tex_scale = m.tex_scale;
        std::memcpy(o.colour, m.colour, 16);
        std::memcpy(o.base_tex_mat, m.base_tex_mat, 32);
        std::memcpy(o.base_tex_mat2, m.base_tex_mat2, 32);
        std::memcpy(o.normal_map_mat, m.normal_map_mat, 32);
        std::memcpy(o.normal_map_mat2, m.normal_map_mat2, 32);
        std::memcpy(o.blend_tex_mat, m.blend_tex_mat, 32);
        std::memcpy(o.detail_mat, m.detail_mat, 32);
        o.detail_scale = m.detail_scale;
        o.detail_blend_factor = m.detail_blend_factor;
        o.base_texture = m.base_texture;
        o.base_texture2 = m.base_texture2;
        o.normal_map = m.normal_map;
        o.normal_map2 = m.normal_map2;
        o.mrao = m.mrao;
        o.mrao2 = m.mrao2;
        o.blend_texture = m.blend_texture;
        o.detail = m.detail;
        o.detail_blend_mode = m.detail_blend_mode;
        o.masked_blending = m.masked_blending;
        o.water = m.water;
    }
    std::vector<VtDevEntity> de(mEntities.size());
    for (size_t i = 0; i < mEntities.size(); i++) {
        de[i].id = mEntities[i].id;
        std::memcpy(de[i].colour, mEntities[i].colour, 16);
    }
    std::vector<VtDevTexture> dt(scene.n_textures + 1);
    std::vector<uint8_t> texels;
    auto add_texture = [&](VtDevTexture &o, uint32_t w, uint32_t h, uint32_t mips, uint32_t flags, uint32_t layout, const uint8_t *px, uint64_t nbytes) {
        std::memset(&o, 0, sizeof(o));
        o.width = w;
        o.height = h;
        o.mips = mips;
        o.flags = flags;
        o.layout = layout;
        o.base = texels.size();
        // chain is smallest mip first: offset of mip m = sizes of mips m+1 .. last (VTFParser.cpp:219-229), in texels
        for (uint32_t m = 0; m < mips; m++) {
            uint32_t off = 0;
            for (uint32_t i = m + 1; i < mips; i++) off += std::max(1u, w >> i) * std::max(1u, h >> i);
            o.mip_offset[m] = off;
        }
        texels.insert(texels.end(), px, px + nbytes);
        while (texels.size() & 15) texels.push_back(0);
    };
    for (uint32_t i = 0; i < scene.n_textures; i++) {
        const vt_texture &t = scene.textures[i];
        add_texture(dt[i], t.width, t.height, t.mip_count, t.flags, t.texel_layout, t.rgba, t.nbytes);
    }
    const uint8_t white[4] = {255, 255, 255, 255};
    add_texture(dt[scene.n_textures], 1, 1, 1, 0, 0, white, 4);

    timer.lap("  triangle / attr records");
    D.refit_ready = false;
    D.refit_cost_built = D.refit_cost_now = 0.0;
    mBvhStale = false;
    mReplica = false;
    D.pairs.release();
    D.cpairs.release();
    D.quads.release();
    if (layout == VT_LAYOUT_QUAD) D.quads.upload(quad.quads.data(), quad.quads.size());
    else if (layout == VT_LAYOUT_COMPACT) D.cpairs.upload(cpairs.data(), cpairs.size());
    else D.pairs.upload(flat.pairs.data(), flat.pairs.size());
    D.tris.upload(recs.data(), recs.size());
    D.tri_uv.upload(uv.data(), uv.size());
    JoinAttrUpload();  // D.attrs: uploaded by Ingest's helper thread while the hierarchy was being built
    D.mats.upload(dm.data(), dm.size());
    D.ents.upload(de.data(), de.size());
    D.texs.upload(dt.data(), dt.size());
    D.texels.upload(texels.data(), texels.size());

    timer.lap("  cudaMalloc + H2D");
    VtSceneView &V = D.view;
    V.pairs = layout == VT_LAYOUT_EXACT ? D.pairs.p : nullptr;
    V.cpairs = layout == VT_LAYOUT_COMPACT ? D.cpairs.p : nullptr;
    V.quads = layout == VT_LAYOUT_QUAD ? D.quads.p : nullptr;
    V.tris = D.tris.p;
    V.tri_uv = D.tri_uv.p;
    V.attrs = D.attrs.p;
    V.mats = D.mats.p;
    V.ents = D.ents.p;
    V.texs = D.texs.p;
    V.texels = D.texels.p;
    V.n_pairs = n_inner;
    V.n_tris = (uint32_t)n;
    V.root_leaf_count = root_leaf_count;
    V.n_smem_pairs = smem_pairs;
    V.has_alphatest = any_alpha ? 1u : 0u;
    V.fallback_tex = scene.n_textures;
    V.magic = 0x4B000000u;
    V.magic_h = 0x64646464u;

    int blocks = 0;
    VT_CUDA(vt_traverse_occupancy(&blocks, (size_t)smem_pairs * sizeof(VtPair), layout));
    if (blocks < 1) blocks = 1;
    const int mult = env_int("VT_GRID_BLOCKS_PER_SM", blocks);
    D.cfg.grid = D.sm_count * std::max(1, std::min(mult, blocks));
    mAccelBuilt = true;
}

void AccelStruct::Populate(const vt_scene &scene) {
    DrainWaveFrames();
    PhaseTimer timer;
    Ingest(scene);
    timer.lap("ingest (Triangle ctor)");
    // the build step of source/objects/AccelStruct.cpp:762-770, host side: the product's binned-SAH builder, or (VT_BUILDER=ploc)
    // the reference's own PLOC + LeafCollapser tree, node for node (vt_bvh_ploc.cpp)
    const char *builder = std::getenv("VT_BUILDER");
    if (builder && std::string(builder) == "ploc") {
        build_bvh_ploc(mTriangles, mAccel);
        collapse_leaves(mAccel);
    } else {
        build_bvh(mTriangles, mAccel, env_int("VT_MAX_LEAF", 4), env_float("VT_TRAV_COST", 1.0f), (uint32_t)std::max(0, env_int("VT_SAH_SWEEP", 0)));
    }
    timer.lap("hierarchy build");
    if (const int passes = env_int("VT_REINSERT", 0); passes > 0) {  // builder-quality option, off by default (vt_bvh_reinsert.cpp)
        if (!reinsert_optimize(mAccel, passes, env_float("VT_REINSERT_FRACTION", 0.05f))) throw std::runtime_error("reinsertion: malformed hierarchy");
        timer.lap("reinsertion optimisation");
    }
    Upload(scene);
    timer.lap("flatten + records + upload");
}

void AccelStruct::PopulateWithBvh(const vt_scene &scene, const vt_node *nodes, uint64_t node_count, const uint64_t *prim_indices) {
    DrainWaveFrames();
    Ingest(scene);
    if (!nodes || !prim_indices || node_count == 0) throw std::runtime_error("populate_with_bvh: null hierarchy");
    mAccel.nodes.assign(nodes, nodes + node_count);
    mAccel.prim_indices.assign(prim_indices, prim_indices + scene.n_tris);
    Upload(scene);
}

// ---- replicas (multi-GPU): the device image of a populated handle, copied GPU to GPU by vt_group.cu
void AccelStruct::ExportReplica(ReplicaImage &img, const void *bufs[10]) const {
    check_built(mAccelBuilt);
    const DeviceScene &D = *mpDevice;
    const VtSceneView &V = D.view;
    std::memset(&img, 0, sizeof(img));
    const size_t n_inner = V.n_pairs, n_tri_recs = (size_t)V.n_tris + ((VT_EMPTY_SENTINEL && mLayout == VT_LAYOUT_QUAD) ? 1 : 0);
    const uint64_t bytes[10] = {V.pairs ? n_inner * sizeof(VtPair) : 0,
                                V.cpairs ? n_inner * sizeof(VtCPair) : 0,
                                V.quads ? n_inner * sizeof(VtQuad) : 0,
                                n_tri_recs * sizeof(VtTriRec),
                                n_tri_recs * 6 * sizeof(float),
                                (size_t)V.n_tris * sizeof(VtTriAttr),
                                D.mats.cap * sizeof(VtDevMaterial),
                                D.ents.cap * sizeof(VtDevEntity),
                                D.texs.cap * sizeof(VtDevTexture),
                                D.texels.cap};
    const void *ptrs[10] = {D.pairs.p, D.cpairs.p, D.quads.p, D.tris.p, D.tri_uv.p, D.attrs.p, D.mats.p, D.ents.p, D.texs.p, D.texels.p};
    for (int i = 0; i < 10; i++) img.bytes[i] = bytes[i], bufs[i] = bytes[i] ? ptrs[i] : nullptr;
    img.n_pairs = V.n_pairs, img.n_tris = V.n_tris, img.root_leaf_count = V.root_leaf_count, img.n_smem_pairs = V.n_smem_pairs;
    img.has_alphatest = V.has_alphatest, img.fallback_tex = V.fallback_tex;
    img.key_mid = (uint32_t)D.cfg.key_mid;
    img.layout = mLayout;
    img.n_materials = (uint32_t)D.mats.cap;
}

void AccelStruct::AllocReplica(const ReplicaImage &img, void *bufs[10]) {
    VT_CUDA(cudaSetDevice(mDevice));
    DeviceScene &D = *mpDevice;
    mAccelBuilt = false;
    mTriangles.clear();
    mEntities.clear();
    mMaterials.clear();
    mAccel = HostBvh();
    D.refit_ready = false;
    D.pairs.release();
    D.cpairs.release();
    D.quads.release();
    if (img.bytes[0]) D.pairs.ensure(img.bytes[0] / sizeof(VtPair));
    if (img.bytes[1]) D.cpairs.ensure(img.bytes[1] / sizeof(VtCPair));
    if (img.bytes[2]) D.quads.ensure(img.bytes[2] / sizeof(VtQuad));
    D.tris.ensure(std::max<size_t>(1, img.bytes[3] / sizeof(VtTriRec)));
    D.tri_uv.ensure(std::max<size_t>(1, img.bytes[4] / sizeof(float)));
    D.attrs.ensure(std::max<size_t>(1, img.bytes[5] / sizeof(VtTriAttr)));
    D.mats.ensure(std::max<size_t>(1, img.bytes[6] / sizeof(VtDevMaterial)));
    D.ents.ensure(std::max<size_t>(1, img.bytes[7] / sizeof(VtDevEntity)));
    D.texs.ensure(std::max<size_t>(1, img.bytes[8] / sizeof(VtDevTexture)));
    D.texels.ensure(std::max<size_t>(1, img.bytes[9]));
    void *ptrs[10] = {D.pairs.p, D.cpairs.p, D.quads.p, D.tris.p, D.tri_uv.p, D.attrs.p, D.mats.p, D.ents.p, D.texs.p, D.texels.p};
    for (int i = 0; i < 10; i++) bufs[i] = img.bytes[i] ? ptrs[i] : nullptr;
    mLayout = mWantLayout = img.layout;
    VtSceneView &V = D.view;
    V = VtSceneView{};
    V.pairs = img.layout == VT_LAYOUT_EXACT ? D.pairs.p : nullptr;
    V.cpairs = img.layout == VT_LAYOUT_COMPACT ? D.cpairs.p : nullptr;
    V.quads = img.layout == VT_LAYOUT_QUAD ? D.quads.p : nullptr;
    V.tris = D.tris.p, V.tri_uv = D.tri_uv.p, V.attrs = D.attrs.p, V.mats = D.mats.p, V.ents = D.ents.p, V.texs = D.texs.p, V.texels = D.texels.p;
    V.n_pairs = img.n_pairs, V.n_tris = img.n_tris, V.root_leaf_count = img.root_leaf_count, V.n_smem_pairs = img.n_smem_pairs;
    V.has_alphatest = img.has_alphatest, V.fallback_tex = img.fallback_tex;
    D.cfg.key_mid = (int)img.key_mid;  // the child order chosen for the scene travels with its image
    V.magic = 0x4B000000u;
    V.magic_h = 0x64646464u;
    D.cfg.persistent = env_int("VT_PERSISTENT", 1);
    D.cfg.refill_threshold = env_int("VT_REFILL", D.cfg.refill_threshold);
    D.cfg.tri_threshold = env_int("VT_TRI_ROUND", D.cfg.tri_threshold);
    D.cfg.tail_share = env_int("VT_TAIL_SHARE", D.cfg.tail_share);
    int blocks = 0;
    VT_CUDA(vt_traverse_occupancy(&blocks, (size_t)V.n_smem_pairs * sizeof(VtPair), img.layout));
    if (blocks < 1) blocks = 1;
    const int mult = env_int("VT_GRID_BLOCKS_PER_SM", blocks);
    D.cfg.grid = D.sm_count * std::max(1, std::min(mult, blocks));
    mReplica = true;
    mBvhStale = false;
    mAccelBuilt = true;  // valid once the caller has filled the buffers (it does so before anything is enqueued)
}

const HostBvh &AccelStruct::Bvh() const {
    if (mReplica) throw std::runtime_error("this handle is a replica (vt_group): the host-side hierarchy lives in the group's first member");
    std::lock_guard<std::mutex> lock(mBvhMutex);
    if (mBvhStale) {
        std::string err;
        if (!refit_bvh(mTriangles, mAccel, err)) throw std::runtime_error(err);
        mBvhStale = false;
    }
    return mAccel;
}

// K5 state of a resident quad hierarchy, built on its first refit: parent / inner-child count per quad, original triangle -> leaf
// slot, leaf slot -> quad, the zeroed per-quad counters of the ranged walk — and the tree AS BUILT: one bottom-up pass over the
// unchanged geometry fills the box table and gives the node-area sum that later refits are measured against (the pass re-derives the
// very bytes it reads: idempotent).  Returns the number of kernels launched.
static unsigned prepare_refit_state(DeviceScene &D, cudaStream_t stream) {
    const VtSceneView &V = D.view;
    D.refit_parent.ensure(V.n_pairs);
    D.refit_n_inner.ensure(V.n_pairs);
    D.refit_arrive.ensure(V.n_pairs);
    D.refit_qbox.ensure((size_t)V.n_pairs * 6);
    D.refit_slot_of.ensure(V.n_tris);
    D.refit_leaf_quad.ensure(V.n_tris);
    D.refit_range_state.ensure((size_t)V.n_pairs * 3);  // stamp | dirty children | arrivals (vt_refit.cu: k_refit_*_range)
    D.refit_error.ensure(1);
    D.refit_cost.ensure(1);
    VT_CUDA(cudaMemsetAsync(D.refit_range_state.p, 0, (size_t)V.n_pairs * 3 * sizeof(uint32_t), stream));
    D.refit_epoch = 1;
    VT_CUDA(vt_launch_refit_prepare(V, D.refit_parent.p, D.refit_n_inner.p, D.refit_slot_of.p, D.refit_leaf_quad.p, stream));
    VT_CUDA(cudaMemsetAsync(D.refit_error.p, 0, sizeof(uint32_t), stream));
    VT_CUDA(vt_launch_refit_quads(V, D.refit_parent.p, D.refit_n_inner.p, D.refit_arrive.p, D.refit_qbox.p, D.refit_error.p, stream));
    VT_CUDA(vt_launch_refit_cost(D.refit_qbox.p, V.n_pairs, D.refit_cost.p, stream));
    VT_CUDA(cudaMemcpyAsync(&D.refit_cost_built, D.refit_cost.p, sizeof(double), cudaMemcpyDeviceToHost, stream));
    D.refit_ready = true;
    return 3;
}

void AccelStruct::Refit(const vt_scene &scene) {
    DrainWaveFrames();
    if (mReplica) throw std::runtime_error("refit: this handle is a replica (vt_group): refit the group");
    if (!mAccelBuilt) throw std::runtime_error("refit: nothing built yet (use Populate)");
    if (scene.n_tris != mAccel.prim_indices.size()) throw std::runtime_error("refit: triangle count changed (use Populate to rebuild)");
    if (scene.n_materials != mMaterials.size() || scene.n_entities != mEntities.size())
        throw std::runtime_error("refit: material or entity count changed (use Populate to rebuild)");
    const bool on_device = mLayout == VT_LAYOUT_QUAD && mpDevice->view.n_pairs != 0 && env_int("VT_REFIT_DEVICE", 1) != 0;
    if (on_device) {
        // K5: new vertices up, Triangle constructor + bottom-up quad refit on the device (vt_refit.cu); the host copies
        // (mTriangles now, the bvh::Bvh-form boxes on demand) are brought up to date while the GPU works.
        // the kernel indexes the resident material table with the caller's indices: check them before anything is launched
        bool bad = false, any_alpha = false;
#pragma omp parallel for reduction(|| : bad, any_alpha)
        for (int64_t i = 0; i < (int64_t)scene.n_tris; i++) {
            const bool oob = scene.tris[i].material >= scene.n_materials || scene.tris[i].ent_idx >= scene.n_entities;
            bad = bad || oob;
            // k_refit_tris rebuilds the CULL / ALPHATEST bits from the RESIDENT material table (materials are unchanged by a refit)
            any_alpha = any_alpha || (!oob && (mMaterials[scene.tris[i].material].flags & VT_MATFLAG_ALPHATEST) != 0);
        }
        if (bad) throw std::runtime_error("triangle references a material or entity out of range");
        VT_CUDA(cudaSetDevice(mDevice));
        DeviceScene &D = *mpDevice;
        // K1's ALPHA template is chosen from this flag: a refit may move a triangle onto (or off) an alpha-tested material
        D.view.has_alphatest = any_alpha ? 1u : 0u;
        const VtSceneView &V = D.view;
        cudaStream_t stream = D.own_stream;
        D.refit_in.ensure(scene.n_tris);
        VT_CUDA(cudaMemcpyAsync(D.refit_in.p, scene.tris, scene.n_tris * sizeof(vt_tri_in), cudaMemcpyHostToDevice, stream));
        if (!D.refit_ready) mLaunches += prepare_refit_state(D, stream);
        VT_CUDA(cudaMemsetAsync(D.refit_error.p, 0, sizeof(uint32_t), stream));
        VT_CUDA(vt_launch_refit_tris(V, D.refit_in.p, 0, (uint32_t)scene.n_tris, D.refit_slot_of.p, stream));
        VT_CUDA(vt_launch_refit_quads(V, D.refit_parent.p, D.refit_n_inner.p, D.refit_arrive.p, D.refit_qbox.p, D.refit_error.p, stream));
        VT_CUDA(vt_launch_refit_cost(D.refit_qbox.p, V.n_pairs, D.refit_cost.p, stream));
        VT_CUDA(cudaMemcpyAsync(&D.refit_cost_now, D.refit_cost.p, sizeof(double), cudaMemcpyDeviceToHost, stream));
        mLaunches += 3;
        Ingest(scene, false);  // host copies of the containers only (K5 writes the device records; clears mAccelBuilt, mAccel keeps the structure)
        uint32_t failed = 0;
        VT_CUDA(cudaMemcpyAsync(&failed, D.refit_error.p, sizeof(failed), cudaMemcpyDeviceToHost, stream));
        VT_CUDA(cudaStreamSynchronize(stream));
        D.refit_in.release();  // 152 B per triangle of staging: not part of the resident scene
        if (!failed) {
            mBvhStale = true;
            mAccelBuilt = true;
            // rebuild trigger: refits keep the topology of the ORIGINAL geometry; when the moved geometry has loosened the boxes
            // beyond the caller's tolerance the structure is rebuilt from scratch (what accel:Rebuild always does in the reference)
            const double limit = mRefitRebuildRatio > 0.0 ? mRefitRebuildRatio : (double)env_float("VT_REFIT_REBUILD_RATIO", 0.f);
            if (limit > 0.0 && RefitQuality() > limit) {
                Populate(scene);
                mRebuilds++;
            }
            return;
        }
        // a box left the float grid the quantised layout can hold: re-derive the layout on the host (may fall back to exact)
    } else {
        Ingest(scene);
    }
    std::string err;
    if (!refit_bvh(mTriangles, mAccel, err)) throw std::runtime_error(err);
    Upload(scene);
}

void AccelStruct::RefitRange(const vt_tri_in *tris, uint64_t first, uint64_t count) {
    DrainWaveFrames();
    if (mReplica) throw std::runtime_error("refit: this handle is a replica (vt_group): refit the group");
    if (!mAccelBuilt) throw std::runtime_error("refit: nothing built yet (use Populate)");
    if (count == 0) return;
    if (!tris) throw std::runtime_error("refit_range: null triangles");
    if (first + count > mTriangles.size()) throw std::runtime_error("refit_range: range past the end of the triangle array");
    if (mLayout != VT_LAYOUT_QUAD || mpDevice->view.n_pairs == 0)
        throw std::runtime_error("refit_range: needs the resident quad layout (use vt_accel_refit with the whole scene)");
    bool bad = false, any_alpha = false;
#pragma omp parallel for reduction(|| : bad, any_alpha)
    for (int64_t i = 0; i < (int64_t)count; i++) {
        const bool oob = tris[i].material >= mMaterials.size() || tris[i].ent_idx >= mEntities.size();
        bad = bad || oob;
        any_alpha = any_alpha || (!oob && (mMaterials[tris[i].material].flags & VT_MATFLAG_ALPHATEST) != 0);
    }
    if (bad) throw std::runtime_error("triangle references a material or entity out of range");
    VT_CUDA(cudaSetDevice(mDevice));
    DeviceScene &D = *mpDevice;
    // a range can only ADD alpha-tested triangles as far as this call can tell; the ALPHA template is correct for scenes without any
    if (any_alpha) D.view.has_alphatest = 1u;
    const VtSceneView &V = D.view;
    cudaStream_t stream = D.own_stream;
    D.refit_in.ensure(count);
    VT_CUDA(cudaMemcpyAsync(D.refit_in.p, tris, count * sizeof(vt_tri_in), cudaMemcpyHostToDevice, stream));
    if (!D.refit_ready) mLaunches += prepare_refit_state(D, stream);
    VT_CUDA(cudaMemsetAsync(D.refit_error.p, 0, sizeof(uint32_t), stream));
    VT_CUDA(vt_launch_refit_tris(V, D.refit_in.p, (uint32_t)first, (uint32_t)count, D.refit_slot_of.p, stream));
    const bool timing = env_int("VT_TIMING", 0) != 0;  // device time of the bottom-up walk, one stderr line
    cudaEvent_t walk_t0 = nullptr, walk_t1 = nullptr;
    if (timing) {
        VT_CUDA(cudaEventCreate(&walk_t0));
        VT_CUDA(cudaEventCreate(&walk_t1));
        VT_CUDA(cudaEventRecord(walk_t0, stream));
    }
    const bool ranged_walk = env_int("VT_REFIT_RANGE_WALK", 1) != 0;
    if (ranged_walk) {
        // only the quads holding a touched triangle and their ancestors (every other quad keeps its bytes and its box-table entry)
        if (D.refit_epoch > 0xFFFFFF00u) {
            VT_CUDA(cudaMemsetAsync(D.refit_range_state.p, 0, (size_t)V.n_pairs * 3 * sizeof(uint32_t), stream));
            D.refit_epoch = 1;
        }
        uint32_t *st = D.refit_range_state.p;
        VT_CUDA(vt_launch_refit_quads_range(V, (uint32_t)first, (uint32_t)count, D.refit_slot_of.p, D.refit_leaf_quad.p, D.refit_parent.p, st,
                                            st + V.n_pairs, st + 2 * (size_t)V.n_pairs, D.refit_qbox.p, D.refit_epoch, D.refit_error.p, stream));
        D.refit_epoch += 3;
        mLaunches += 2;
    } else {
        VT_CUDA(vt_launch_refit_quads(V, D.refit_parent.p, D.refit_n_inner.p, D.refit_arrive.p, D.refit_qbox.p, D.refit_error.p, stream));
    }
    if (timing) VT_CUDA(cudaEventRecord(walk_t1, stream));
    VT_CUDA(vt_launch_refit_cost(D.refit_qbox.p, V.n_pairs, D.refit_cost.p, stream));
    VT_CUDA(cudaMemcpyAsync(&D.refit_cost_now, D.refit_cost.p, sizeof(double), cudaMemcpyDeviceToHost, stream));
    mLaunches += 3;
    mAccelBuilt = false;  // until the device reports success
#pragma omp parallel for
    for (int64_t i = 0; i < (int64_t)count; i++) {  // the host copy of the containers, as Ingest fills them
        const vt_tri_in &in = tris[i];
        Triangle t(in.p[0], in.p[1], in.p[2], in.material, in.uvs, in.one_sided != 0);
        std::memcpy(t.normals, in.normals, sizeof(t.normals));
        std::memcpy(t.tangents, in.tangents, sizeof(t.tangents));
        std::memcpy(t.alphas, in.alphas, sizeof(t.alphas));
        t.entIdx = in.ent_idx;
        mTriangles[first + i] = t;
    }
    uint32_t failed = 0;
    VT_CUDA(cudaMemcpyAsync(&failed, D.refit_error.p, sizeof(failed), cudaMemcpyDeviceToHost, stream));
    VT_CUDA(cudaStreamSynchronize(stream));
    if (timing) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, walk_t0, walk_t1);
        std::fprintf(stderr, "[refit_range] %llu triangles, %s walk %.3f ms on the device\n", (unsigned long long)count, ranged_walk ? "ranged" : "whole-tree", ms);
        cudaEventDestroy(walk_t0);
        cudaEventDestroy(walk_t1);
    }
    D.refit_in.release();
    mBvhStale = true;
    if (failed)  // the resident records are half updated: the handle stays invalid until the caller rebuilds
        throw std::runtime_error("refit_range: a box left the float grid of the quad layout; rebuild with vt_accel_refit or vt_accel_populate");
    mAccelBuilt = true;
}


void AccelStruct::TraverseBatch(const vt_ray *rays, uint64_t n, vt_hit *hits, vt_attr *attrs, const float *cones,
                                uint32_t flags, void *stream_) {
    check_built(mAccelBuilt);
    if (n == 0) return;
    if (!rays || !hits) throw std::runtime_error("traverse: rays and hits must not be null");
    VT_CUDA(cudaSetDevice(mDevice));
    DeviceScene &D = *mpDevice;
    const bool dev_ptrs = (flags & VT_TRAVERSE_DEVICE_PTRS) != 0;
    const bool any_hit = (flags & VT_TRAVERSE_ANY_HIT) != 0;
    if (dev_ptrs && (((uintptr_t)rays & 31) || ((uintptr_t)hits & 15)))
        throw std::runtime_error("traverse: device ray buffers must be 32-byte aligned and hit buffers 16-byte aligned");
    cudaStream_t stream = stream_ ? (cudaStream_t)stream_ : (dev_ptrs ? (cudaStream_t) nullptr : D.own_stream);
    const uint32_t slot = D.next_slot.fetch_add(1) % kCounterSlots;
    unsigned long long *ctr = D.counters.p + 2 * slot;
    VT_CUDA(cudaMemsetAsync(ctr, 0, 2 * sizeof(unsigned long long), stream));

    const vt_ray *d_rays = rays;
    vt_hit *d_hits = hits;
    vt_attr *d_attrs = attrs;
    const float *d_cones = cones;
    if (!dev_ptrs) {
        D.s_rays.ensure(n);
        D.s_hits.ensure(n);
        VT_CUDA(cudaMemcpyAsync(D.s_rays.p, rays, n * sizeof(vt_ray), cudaMemcpyHostToDevice, stream));
        d_rays = D.s_rays.p;
        d_hits = D.s_hits.p;
        if (attrs) {
            D.s_attrs.ensure(n);
            d_attrs = D.s_attrs.p;
            if (cones) {
                D.s_cones.ensure(2 * n);
                VT_CUDA(cudaMemcpyAsync(D.s_cones.p, cones, 2 * n * sizeof(float), cudaMemcpyHostToDevice, stream));
                d_cones = D.s_cones.p;
            }
        }
    }
    VT_CUDA(vt_launch_traverse(D.view, d_rays, d_hits, n, any_hit, ctr, D.cfg, stream));
    mLaunches++;
    if (attrs) {
        VT_CUDA(vt_launch_trace_result(D.view, d_rays, d_hits, d_cones, d_attrs, n, stream));
        mLaunches++;
    }
    if (!dev_ptrs) {
        unsigned long long invalid[2] = {0, 0};
        VT_CUDA(cudaMemcpyAsync(hits, d_hits, n * sizeof(vt_hit), cudaMemcpyDeviceToHost, stream));
        if (attrs) VT_CUDA(cudaMemcpyAsync(attrs, d_attrs, n * sizeof(vt_attr), cudaMemcpyDeviceToHost, stream));
        VT_CUDA(cudaMemcpyAsync(invalid, ctr, sizeof(invalid), cudaMemcpyDeviceToHost, stream));
        VT_CUDA(cudaStreamSynchronize(stream));
        mInvalidRays = invalid[1];
    }
}

void AccelStruct::TraverseStats(const vt_ray *rays, uint64_t n, uint32_t flags, uint64_t *steps, uint64_t *tests, uint32_t *per_ray) {
    check_built(mAccelBuilt);
    if (steps) *steps = 0;
    if (tests) *tests = 0;
    if (n == 0) return;
    if (!rays) throw std::runtime_error("traverse_stats: rays must not be null");
    VT_CUDA(cudaSetDevice(mDevice));
    DeviceScene &D = *mpDevice;
    if (D.view.n_smem_pairs && D.view.pairs) throw std::runtime_error("traverse_stats: not available with VT_SMEM_PAIRS");
    cudaStream_t stream = D.own_stream;
    D.stat_counters.ensure(5);
    VT_CUDA(cudaMemsetAsync(D.stat_counters.p, 0, 5 * sizeof(unsigned long long), stream));
    if (per_ray) {  // per-ray records (quantised layouts): counters[4] carries the device address of the array
        if (!D.view.quads && !D.view.cpairs) throw std::runtime_error("traverse_ray_stats: quad or compact layout only");
        D.s_ray_stats.ensure(n);
        VT_CUDA(cudaMemsetAsync(D.s_ray_stats.p, 0, n * sizeof(uint32_t), stream));
        const unsigned long long addr = (unsigned long long)(uintptr_t)D.s_ray_stats.p;
        VT_CUDA(cudaMemcpyAsync(D.stat_counters.p + 4, &addr, sizeof(addr), cudaMemcpyHostToDevice, stream));
    }
    const vt_ray *d_rays = rays;
    D.s_hits.ensure(n);
    if (!(flags & VT_TRAVERSE_DEVICE_PTRS)) {
        D.s_rays.ensure(n);
        VT_CUDA(cudaMemcpyAsync(D.s_rays.p, rays, n * sizeof(vt_ray), cudaMemcpyHostToDevice, stream));
        d_rays = D.s_rays.p;
    } else {
        VT_CUDA(cudaDeviceSynchronize());  // the caller's stream may still be producing the rays
    }
    VT_CUDA(vt_launch_traverse(D.view, d_rays, D.s_hits.p, n, false, D.stat_counters.p, D.cfg, stream, true));
    mLaunches++;
    unsigned long long c[4];
    VT_CUDA(cudaMemcpyAsync(c, D.stat_counters.p, sizeof(c), cudaMemcpyDeviceToHost, stream));
    if (per_ray) VT_CUDA(cudaMemcpyAsync(per_ray, D.s_ray_stats.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    VT_CUDA(cudaStreamSynchronize(stream));
    if (steps) *steps = c[2];
    if (tests) *tests = c[3];
}

void AccelStruct::TraceResultBatch(const vt_ray *rays, const vt_hit *hits, uint64_t n, vt_attr *attrs, const float *cones,
                                   uint32_t flags, void *stream_) {
    check_built(mAccelBuilt);
    if (n == 0) return;
    if (!rays || !hits || !attrs) throw std::runtime_error("trace_result: rays, hits and attrs must not be null");
    VT_CUDA(cudaSetDevice(mDevice));
    DeviceScene &D = *mpDevice;
    const bool dev_ptrs = (flags & VT_TRAVERSE_DEVICE_PTRS) != 0;
    cudaStream_t stream = stream_ ? (cudaStream_t)stream_ : (dev_ptrs ? (cudaStream_t) nullptr : D.own_stream);
    if (dev_ptrs) {
        VT_CUDA(vt_launch_trace_result(D.view, rays, hits, cones, attrs, n, stream));
        mLaunches++;
        return;
    }
    D.s_rays.ensure(n);
    D.s_hits.ensure(n);
    D.s_attrs.ensure(n);
    VT_CUDA(cudaMemcpyAsync(D.s_rays.p, rays, n * sizeof(vt_ray), cudaMemcpyHostToDevice, stream));
    VT_CUDA(cudaMemcpyAsync(D.s_hits.p, hits, n * sizeof(vt_hit), cudaMemcpyHostToDevice, stream));
    const float *d_cones = nullptr;
    if (cones) {
        D.s_cones.ensure(2 * n);
        VT_CUDA(cudaMemcpyAsync(D.s_cones.p, cones, 2 * n * sizeof(float), cudaMemcpyHostToDevice, stream));
        d_cones = D.s_cones.p;
    }
    VT_CUDA(vt_launch_trace_result(D.view, D.s_rays.p, D.s_hits.p, d_cones, D.s_attrs.p, n, stream));
    mLaunches++;
    VT_CUDA(cudaMemcpyAsync(attrs, D.s_attrs.p, n * sizeof(vt_attr), cudaMemcpyDeviceToHost, stream));
    VT_CUDA(cudaStreamSynchronize(stream));
}

void AccelStruct::BounceRays(const vt_attr *attrs, uint64_t n, uint32_t spp, uint64_t seed, vt_ray *out_rays,
                             uint64_t *live_out, uint32_t flags, void *stream_) {
    if (n == 0 || spp == 0) {
        if (live_out) *live_out = 0;
        return;
    }
    if (!attrs || !out_rays) throw std::runtime_error("bounce_rays: attrs and out_rays must not be null");
    VT_CUDA(cudaSetDevice(mDevice));
    DeviceScene &D = *mpDevice;
    const bool dev_ptrs = (flags & VT_TRAVERSE_DEVICE_PTRS) != 0;
    cudaStream_t stream = stream_ ? (cudaStream_t)stream_ : (dev_ptrs ? (cudaStream_t) nullptr : D.own_stream);
    D.live.ensure(1);
    if (live_out) VT_CUDA(cudaMemsetAsync(D.live.p, 0, sizeof(unsigned long long), stream));
    const vt_attr *d_attrs = attrs;
    vt_ray *d_out = out_rays;
    if (!dev_ptrs) {
        D.s_attrs.ensure(n);
        D.s_rays.ensure(n * spp);
        VT_CUDA(cudaMemcpyAsync(D.s_attrs.p, attrs, n * sizeof(vt_attr), cudaMemcpyHostToDevice, stream));
        d_attrs = D.s_attrs.p;
        d_out = D.s_rays.p;
    }
    VT_CUDA(vt_launch_bounce_rays(d_attrs, n, spp, seed, 0, d_out, live_out ? D.live.p : nullptr, stream));
    mLaunches++;
    if (!dev_ptrs) VT_CUDA(cudaMemcpyAsync(out_rays, d_out, n * spp * sizeof(vt_ray), cudaMemcpyDeviceToHost, stream));
    if (live_out) {
        unsigned long long v = 0;
        VT_CUDA(cudaMemcpyAsync(&v, D.live.p, sizeof(v), cudaMemcpyDeviceToHost, stream));
        VT_CUDA(cudaStreamSynchronize(stream));
        *live_out = v;
    } else if (!dev_ptrs) {
        VT_CUDA(cudaStreamSynchronize(stream));
    }
}

void AccelStruct::SampleBsdfRays(const vt_ray *rays, const vt_attr *attrs, uint64_t n, uint32_t spp, uint64_t seed, vt_ray *out_rays,
                                 vt_bsdf_sample *samples, uint64_t *live_out, uint32_t *queue, uint64_t *queue_count, vt_hit *miss_hits,
                                 uint32_t flags, void *stream_) {
    if (live_out) *live_out = 0;
    if (n == 0 || spp == 0) return;
    if (!rays || !attrs || !out_rays || !samples) throw std::runtime_error("sample_bsdf_rays: null argument");
    const bool dev_ptrs = (flags & VT_TRAVERSE_DEVICE_PTRS) != 0;
    if ((queue || queue_count || miss_hits) && !(queue && queue_count && miss_hits && dev_ptrs))
        throw std::runtime_error("sample_bsdf_rays: the queue needs queue, queue_count and miss_hits, as device pointers");
    VT_CUDA(cudaSetDevice(mDevice));
    DeviceScene &D = *mpDevice;
    cudaStream_t stream = stream_ ? (cudaStream_t)stream_ : (dev_ptrs ? (cudaStream_t) nullptr : D.own_stream);
    D.live.ensure(1);
    if (live_out) VT_CUDA(cudaMemsetAsync(D.live.p, 0, sizeof(unsigned long long), stream));
    if (queue_count) VT_CUDA(cudaMemsetAsync(queue_count, 0, sizeof(uint64_t), stream));
    const vt_ray *d_rays = rays;
    const vt_attr *d_attrs = attrs;
    vt_ray *d_out = out_rays;
    vt_bsdf_sample *d_samples = samples;
    if (!dev_ptrs) {
        D.s_rays2.ensure(n);
        D.s_attrs.ensure(n);
        D.s_rays.ensure(n * spp);
        D.s_samples.ensure(n * spp);
        VT_CUDA(cudaMemcpyAsync(D.s_rays2.p, rays, n * sizeof(vt_ray), cudaMemcpyHostToDevice, stream));
        VT_CUDA(cudaMemcpyAsync(D.s_attrs.p, attrs, n * sizeof(vt_attr), cudaMemcpyHostToDevice, stream));
        d_rays = D.s_rays2.p, d_attrs = D.s_attrs.p, d_out = D.s_rays.p, d_samples = D.s_samples.p;
    }
    VT_CUDA(vt_launch_bsdf_diffuse_rays(d_rays, d_attrs, n, spp, seed, d_out, d_samples, live_out ? D.live.p : nullptr, stream, queue,
                                        (unsigned long long *)queue_count, miss_hits));
    mLaunches++;
    if (!dev_ptrs) {
        VT_CUDA(cudaMemcpyAsync(out_rays, d_out, n * spp * sizeof(vt_ray), cudaMemcpyDeviceToHost, stream));
        VT_CUDA(cudaMemcpyAsync(samples, d_samples, n * spp * sizeof(vt_bsdf_sample), cudaMemcpyDeviceToHost, stream));
    }
    if (live_out) {
        unsigned long long v = 0;
        VT_CUDA(cudaMemcpyAsync(&v, D.live.p, sizeof(v), cudaMemcpyDeviceToHost, stream));
        VT_CUDA(cudaStreamSynchronize(stream));
        *live_out = v;
    } else if (!dev_ptrs) {
        VT_CUDA(cudaStreamSynchronize(stream));
    }
}

void AccelStruct::ShadowRays(const vt_attr *attrs, uint64_t n, const float light[3], bool point_light, float tmax, vt_ray *out_rays,
                             uint64_t *live_out, uint32_t flags, void *stream_) {
    if (live_out) *live_out = 0;
    if (n == 0) return;
    if (!attrs || !out_rays || !light) throw std::runtime_error("shadow_rays: null argument");
    VT_CUDA(cudaSetDevice(mDevice));
    DeviceScene &D = *mpDevice;
    const bool dev_ptrs = (flags & VT_TRAVERSE_DEVICE_PTRS) != 0;
    cudaStream_t stream = stream_ ? (cudaStream_t)stream_ : (dev_ptrs ? (cudaStream_t) nullptr : D.own_stream);
    D.live.ensure(1);
    if (live_out) VT_CUDA(cudaMemsetAsync(D.live.p, 0, sizeof(unsigned long long), stream));
    const vt_attr *d_attrs = attrs;
    vt_ray *d_out = out_rays;
    if (!dev_ptrs) {
        D.s_attrs.ensure(n);
        D.s_rays.ensure(n);
        VT_CUDA(cudaMemcpyAsync(D.s_attrs.p, attrs, n * sizeof(vt_attr), cudaMemcpyHostToDevice, stream));
        d_attrs = D.s_attrs.p;
        d_out = D.s_rays.p;
    }
    VT_CUDA(vt_launch_shadow_rays(d_attrs, n, light, point_light, tmax, d_out, live_out ? D.live.p : nullptr, stream));
    mLaunches++;
    if (!dev_ptrs) VT_CUDA(cudaMemcpyAsync(out_rays, d_out, n * sizeof(vt_ray), cudaMemcpyDeviceToHost, stream));
    if (live_out) {
        unsigned long long v = 0;
        VT_CUDA(cudaMemcpyAsync(&v, D.live.p, sizeof(v), cudaMemcpyDeviceToHost, stream));
        VT_CUDA(cudaStreamSynchronize(stream));
        *live_out = v;
    } else if (!dev_ptrs) {
        VT_CUDA(cudaStreamSynchronize(stream));
    }
}

void AccelStruct::BounceRaysQueued(const vt_attr *attrs, uint64_t n, uint32_t spp, uint64_t seed, vt_ray *out_rays, uint32_t *queue,
                                   uint64_t *queue_count, vt_hit *miss_hits, void *stream_, const uint32_t *in_queue, const uint64_t *in_count) {
    if ((in_queue == nullptr) != (in_count == nullptr)) throw std::runtime_error("bounce_rays_queued: in_queue and in_count go together");
    if (in_queue && (in_queue == queue || in_count == queue_count)) throw std::runtime_error("bounce_rays_queued: the input queue must not be the output queue");
    if (!attrs || !out_rays || !queue || !queue_count || !miss_hits) throw std::runtime_error("bounce_rays_queued: null argument");
    if (spp == 0) throw std::runtime_error("bounce_rays_queued: spp must be positive");
    if (n * spp > 0xFFFFFFFFull) throw std::runtime_error("bounce_rays_queued: more than 2^32 slots");
    VT_CUDA(cudaSetDevice(mDevice));
    cudaStream_t stream = (cudaStream_t)stream_;
    VT_CUDA(cudaMemsetAsync(queue_count, 0, sizeof(uint64_t), stream));
    if (n == 0) return;
    VT_CUDA(vt_launch_bounce_rays(attrs, n, spp, seed, 0, out_rays, nullptr, stream, queue, (unsigned long long *)queue_count, miss_hits, nullptr,
                                  in_queue, (const unsigned long long *)in_count));
    mLaunches++;
}

void AccelStruct::ShadowRaysQueued(const vt_attr *attrs, uint64_t n, const float light[3], bool point_light, float tmax,
                                   vt_ray *out_rays, uint32_t *queue, uint64_t *queue_count, vt_hit *miss_hits, void *stream_,
                                   const uint32_t *in_queue, const uint64_t *in_count) {
    if ((in_queue == nullptr) != (in_count == nullptr)) throw std::runtime_error("shadow_rays_queued: in_queue and in_count go together");
    if (in_queue && (in_queue == queue || in_count == queue_count)) throw std::runtime_error("shadow_rays_queued: the input queue must not be the output queue");
    if (!attrs || !out_rays || !light || !queue || !queue_count || !miss_hits) throw std::runtime_error("shadow_rays_queued: null argument");
    if (n > 0xFFFFFFFFull) throw std::runtime_error("shadow_rays_queued: more than 2^32 slots");
    VT_CUDA(cudaSetDevice(mDevice));
    cudaStream_t stream = (cudaStream_t)stream_;
    VT_CUDA(cudaMemsetAsync(queue_count, 0, sizeof(uint64_t), stream));
    if (n == 0) return;
    VT_CUDA(vt_launch_shadow_rays(attrs, n, light, point_light, tmax, out_rays, nullptr, stream, queue, (unsigned long long *)queue_count,
                                  miss_hits, in_queue, (const unsigned long long *)in_count));
    mLaunches++;
}

void AccelStruct::TraverseQueued(const vt_ray *rays, const uint32_t *queue, const uint64_t *queue_count, uint64_t capacity,
                                 vt_hit *hits, vt_attr *attrs, uint32_t flags, void *stream_) {
    check_built(mAccelBuilt);
    if (capacity == 0) return;
    if (!rays || !hits || !queue || !queue_count) throw std::runtime_error("traverse_queued: null argument");
    if (((uintptr_t)rays & 31) || ((uintptr_t)hits & 15))
        throw std::runtime_error("traverse_queued: device ray buffers must be 32-byte aligned and hit buffers 16-byte aligned");
    VT_CUDA(cudaSetDevice(mDevice));
    DeviceScene &D = *mpDevice;
    cudaStream_t stream = (cudaStream_t)stream_;
    unsigned long long *ctr = D.counters.p + 2 * (D.next_slot.fetch_add(1) % kCounterSlots);
    VT_CUDA(cudaMemsetAsync(ctr, 0, 2 * sizeof(unsigned long long), stream));
    VT_CUDA(vt_launch_traverse(D.view, rays, hits, capacity, (flags & VT_TRAVERSE_ANY_HIT) != 0, ctr, D.cfg, stream, false, queue,
                               (const unsigned long long *)queue_count));
    mLaunches++;
    if (attrs) {
        // eager TraceResult of every slot (hits[] is complete once the generator has written its miss records), or — wave
        // compaction, VT_TRAVERSE_QUEUE_ATTRS — only of the slots the queue lists
        if (flags & VT_TRAVERSE_QUEUE_ATTRS)
            VT_CUDA(vt_launch_trace_result(D.view, rays, hits, nullptr, attrs, capacity, stream, queue, (const unsigned long long *)queue_count));
        else
            VT_CUDA(vt_launch_trace_result(D.view, rays, hits, nullptr, attrs, capacity, stream));
        mLaunches++;
    }
}

void AccelStruct::TraceDiffuseWave(const vt_ray *rays, uint64_t n, uint32_t spp, uint64_t seed, vt_hit *hits, vt_attr *attrs,
                                   vt_ray *bounce_rays, vt_hit *bounce_hits, uint64_t *live_out, uint32_t flags,
                                   void *stream_) {
    check_built(mAccelBuilt);
    if (live_out) *live_out = 0;
    if (n == 0) return;
    if (!rays || !hits || !bounce_hits) throw std::runtime_error("diffuse_wave: rays, hits and bounce_hits must not be null");
    VT_CUDA(cudaSetDevice(mDevice));
    DeviceScene &D = *mpDevice;
    const bool dev_ptrs = (flags & VT_TRAVERSE_DEVICE_PTRS) != 0;
    D.live.ensure(1);
    auto next_counter = [&]() { return D.counters.p + 2 * (D.next_slot.fetch_add(1) % kCounterSlots); };
    if (dev_ptrs) {
        // everything resident: five launches on the caller's stream, no synchronisation
        if (!attrs || !bounce_rays) throw std::runtime_error("diffuse_wave (device pointers): attrs and bounce_rays scratch required");
        cudaStream_t stream = (cudaStream_t)stream_;
        unsigned long long *c0 = next_counter(), *c1 = next_counter();
        DeviceScene::WaveScratch *ws;
        {
            std::lock_guard<std::mutex> lock(D.wave_mutex);
            ws = &D.wave_scratch[stream];
        }
        ws->queue.ensure(n * spp);
        ws->queue_count.ensure(1);
        VT_CUDA(cudaMemsetAsync(c0, 0, 16, stream));
        VT_CUDA(cudaMemsetAsync(c1, 0, 16, stream));
        VT_CUDA(cudaMemsetAsync(ws->queue_count.p, 0, sizeof(unsigned long long), stream));
        VT_CUDA(vt_launch_traverse(D.view, rays, hits, n, false, c0, D.cfg, stream));
        VT_CUDA(vt_launch_trace_result(D.view, rays, hits, nullptr, attrs, n, stream));
        VT_CUDA(vt_launch_bounce_rays(attrs, n, spp, seed, 0, bounce_rays, nullptr, stream, ws->queue.p, ws->queue_count.p, bounce_hits));
        VT_CUDA(vt_launch_traverse(D.view, bounce_rays, bounce_hits, n * spp, false, c1, D.cfg, stream, false, ws->queue.p,
                                   ws->queue_count.p));
        mLaunches += 4;
        return;
    }
    // host pointers: tiles round-robin over three streams — H2D(rays) | K1 K2 K3 K1 | D2H(results) overlap across tiles
    const uint64_t tile = (uint64_t)std::max(1, env_int("VT_WAVE_TILE", 1 << 19));
    VT_CUDA(cudaMemsetAsync(D.live.p, 0, sizeof(unsigned long long), D.own_stream));
    VT_CUDA(cudaStreamSynchronize(D.own_stream));
    const int n_lanes = std::max(1, std::min(8, env_int("VT_WAVE_LANES", 4)));
    for (int i = 0; i < n_lanes; i++)
        if (!D.lanes[i].stream) VT_CUDA(cudaStreamCreateWithFlags(&D.lanes[i].stream, cudaStreamNonBlocking));
    // VT_WAVE_CTAS_PER_SM: persistent-grid size of the tiles' K1 launches.  Several tiles are in flight, so a tile's K1 takes
    // 6 of the 9 CTA slots per SM and leaves room for another tile's kernels to start (3.08 vs 3.15 ms per e2e step; 3: 3.38)
    VtLaunchConfig tile_cfg = D.cfg;
    {
        const int per_sm = env_int("VT_WAVE_CTAS_PER_SM", 6);
        if (per_sm > 0) tile_cfg.grid = std::min(D.cfg.grid, D.sm_count * per_sm);
    }
    // uploads go through one copy stream into a frame-sized staging buffer (see DeviceScene::wave_rays)
    D.wave_rays.ensure(n);
    if (!D.copy_stream) VT_CUDA(cudaStreamCreateWithFlags(&D.copy_stream, cudaStreamNonBlocking));
    size_t n_uploads = 0;
    auto upload_tile = [&](uint64_t base, uint64_t m, cudaStream_t consumer) -> const vt_ray * {
        if (n_uploads == D.upload_done.size()) {
            cudaEvent_t e;
            VT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            D.upload_done.push_back(e);
        }
        cudaEvent_t done = D.upload_done[n_uploads++];
        VT_CUDA(cudaMemcpyAsync(D.wave_rays.p + base, rays + base, m * sizeof(vt_ray), cudaMemcpyHostToDevice, D.copy_stream));
        VT_CUDA(cudaEventRecord(done, D.copy_stream));
        VT_CUDA(cudaStreamWaitEvent(consumer, done, 0));
        return D.wave_rays.p + base;
    };
    int li = 0, tiles_since_fence = 0;
    uint64_t cur_tile = std::max<uint64_t>(1, std::min<uint64_t>(tile, (uint64_t)std::max(1, env_int("VT_WAVE_FIRST", (int)(tile / 8)))));  // small first tiles, doubling up to `tile`
    for (uint64_t base = 0, m = 0; base < n; base += m, li = (li + 1) % n_lanes, cur_tile = std::min(tile, cur_tile * 2)) {
        m = std::min(cur_tile, n - base);
        if (n - base - m < cur_tile / 2) m = n - base;  // no small tail tile: its launch-latency chain would run alone at the end
        DeviceScene::WaveLane &l = D.lanes[li];
        const uint64_t cap = tile + tile / 2;
        l.hits.ensure(cap);
        l.attrs.ensure(cap);
        l.brays.ensure(cap * spp);
        l.bhits.ensure(cap * spp);
        l.queue.ensure(cap * spp);
        l.queue_count.ensure(1);
        // the counter ring has kCounterSlots entries and a tile takes two: drain the lanes before a live slot could be reused
        if (++tiles_since_fence >= kCounterSlots / 2 - 16) {
            for (int i = 0; i < n_lanes; i++) VT_CUDA(cudaStreamSynchronize(D.lanes[i].stream));
            tiles_since_fence = 0;
        }
        unsigned long long *c0 = next_counter(), *c1 = next_counter();
        VT_CUDA(cudaMemsetAsync(c0, 0, 16, l.stream));
        VT_CUDA(cudaMemsetAsync(c1, 0, 16, l.stream));
        VT_CUDA(cudaMemsetAsync(l.queue_count.p, 0, sizeof(unsigned long long), l.stream));
        const vt_ray *d_tile = upload_tile(base, m, l.stream);
        VT_CUDA(vt_launch_traverse(D.view, d_tile, l.hits.p, m, false, c0, tile_cfg, l.stream));
        VT_CUDA(vt_launch_trace_result(D.view, d_tile, l.hits.p, nullptr, l.attrs.p, m, l.stream));
        VT_CUDA(vt_launch_bounce_rays(l.attrs.p, m, spp, seed, base * spp, l.brays.p, D.live.p, l.stream, l.queue.p, l.queue_count.p,
                                      l.bhits.p));
        VT_CUDA(vt_launch_traverse(D.view, l.brays.p, l.bhits.p, m * spp, false, c1, tile_cfg, l.stream, false, l.queue.p,
                                   l.queue_count.p));
        mLaunches += 4;
        VT_CUDA(cudaMemcpyAsync(hits + base, l.hits.p, m * sizeof(vt_hit), cudaMemcpyDeviceToHost, l.stream));
        VT_CUDA(cudaMemcpyAsync(bounce_hits + base * spp, l.bhits.p, m * spp * sizeof(vt_hit), cudaMemcpyDeviceToHost, l.stream));
        if (attrs) VT_CUDA(cudaMemcpyAsync(attrs + base, l.attrs.p, m * sizeof(vt_attr), cudaMemcpyDeviceToHost, l.stream));
        if (bounce_rays)
            VT_CUDA(cudaMemcpyAsync(bounce_rays + base * spp, l.brays.p, m * spp * sizeof(vt_ray), cudaMemcpyDeviceToHost, l.stream));
    }
    for (int i = 0; i < n_lanes; i++) VT_CUDA(cudaStreamSynchronize(D.lanes[i].stream));
    if (live_out) {
        unsigned long long v = 0;
        VT_CUDA(cudaMemcpy(&v, D.live.p, sizeof(v), cudaMemcpyDeviceToHost));
        *live_out = v;
    }
}

// The frame is ENQUEUED here (tiles round-robin over the wave lanes, uploads on the copy stream) and awaited in RenderDiffuseWaveWait:
// the synchronous call is the two back to back; a caller that begins frame k + 1 before it waits for frame k keeps two frames in
// flight — frame k's last tiles, launch tails and download run under frame k + 1's first uploads and kernels.  Tiles of consecutive
// frames follow each other on the same lane streams, so a lane's scratch needs no extra fencing; the ray staging buffer alternates.
void AccelStruct::RenderDiffuseWaveBegin(const vt_ray *rays, uint64_t n, uint32_t spp, uint64_t seed, float weight, float *fb, bool count_live,
                                         bool pipelined) {
    check_built(mAccelBuilt);
    if (!rays || !fb) throw std::runtime_error("render_diffuse_wave: rays and framebuffer must not be null");
    if (mWaveFrames.size() >= 2) throw std::runtime_error("render_diffuse_wave: two frames are already in flight (wait for one first)");
    VT_CUDA(cudaSetDevice(mDevice));
    DeviceScene &D = *mpDevice;
    const uint64_t *live_out = count_live ? &n : nullptr;  // only its non-nullness is used below
    D.live.ensure(1);
    auto next_counter = [&]() { return D.counters.p + 2 * (D.next_slot.fetch_add(1) % kCounterSlots); };
    // tiles round-robin over three streams: H2D(rays) | K1 K2 K3 K1 K4 | D2H(framebuffer tile) overlap across tiles
    const uint64_t tile = (uint64_t)std::max(1, env_int("VT_WAVE_TILE", 1 << 19));
    if (count_live) {
        if (!mWaveFrames.empty()) throw std::runtime_error("render_diffuse_wave: the live-ray count needs the handle to itself");
        VT_CUDA(cudaMemsetAsync(D.live.p, 0, sizeof(unsigned long long), D.own_stream));
        VT_CUDA(cudaStreamSynchronize(D.own_stream));
    }
    const int n_lanes = std::max(1, std::min(8, env_int("VT_WAVE_LANES", 4)));
    for (int i = 0; i < n_lanes; i++)
        if (!D.lanes[i].stream) VT_CUDA(cudaStreamCreateWithFlags(&D.lanes[i].stream, cudaStreamNonBlocking));
    // VT_WAVE_TRACE=1: one line per tile with the stream-time (ms since the first submission) at which each stage finished
    const bool trace = env_int("VT_WAVE_TRACE", 0) != 0;
    using TileTrace = WaveFrame::TileTrace;
    std::vector<TileTrace> tiles;
    cudaEvent_t ev_begin = nullptr;
    if (trace) {
        VT_CUDA(cudaEventCreate(&ev_begin));
        VT_CUDA(cudaEventRecord(ev_begin, D.lanes[0].stream));
    }
    auto mark = [&](TileTrace &t, int k, cudaStream_t st) {
        if (!trace) return;
        VT_CUDA(cudaEventCreate(&t.ev[k]));
        VT_CUDA(cudaEventRecord(t.ev[k], st));
    };
    // VT_WAVE_CTAS_PER_SM: persistent-grid size of the tiles' K1 launches.  Several tiles are in flight, so a tile's K1 takes
    // 6 of the 9 CTA slots per SM and leaves room for another tile's kernels to start (3.08 vs 3.15 ms per e2e step; 3: 3.38)
    VtLaunchConfig tile_cfg = D.cfg;
    {
        const int per_sm = env_int("VT_WAVE_CTAS_PER_SM", 6);
        if (per_sm > 0) tile_cfg.grid = std::min(D.cfg.grid, D.sm_count * per_sm);
    }
    // uploads go through one copy stream into a frame-sized staging buffer (see DeviceScene::wave_rays); consecutive frames alternate
    // between two of them: frame k + 1's uploads run while frame k's kernels still read theirs
    DevBuf<vt_ray> &staging = (mWaveFrameCount++ & 1) ? D.wave_rays_b : D.wave_rays;
    staging.ensure(n);
    if (!D.copy_stream) VT_CUDA(cudaStreamCreateWithFlags(&D.copy_stream, cudaStreamNonBlocking));
    size_t n_uploads = 0;
    auto upload_tile = [&](uint64_t base, uint64_t m, cudaStream_t consumer) -> const vt_ray * {
        if (n_uploads == D.upload_done.size()) {
            cudaEvent_t e;
            VT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            D.upload_done.push_back(e);
        }
        cudaEvent_t done = D.upload_done[n_uploads++];
        VT_CUDA(cudaMemcpyAsync(staging.p + base, rays + base, m * sizeof(vt_ray), cudaMemcpyHostToDevice, D.copy_stream));
        VT_CUDA(cudaEventRecord(done, D.copy_stream));
        VT_CUDA(cudaStreamWaitEvent(consumer, done, 0));
        return staging.p + base;
    };
    int li = 0, tiles_since_fence = 0;
    // the first tiles are small so the first kernel starts after a short upload; sizes double up to `tile`
    // With another frame in flight the fill of the pipeline is hidden by that frame, so fewer, larger tiles win (measured,
    // tools/e2e_chunk_probe.py, two frames in flight: 2.07 M rays 2.78 / 2.68 / 2.56 ms with a first tile of 64 k / 128 k / 256 k rays;
    // 262 k rays 0.63 / 0.57 ms with 64 k / 128 k); a frame on its own starts small so that its first kernel follows a short upload
    // (2.07 M rays, synchronous: 3.05 ms with 64 k).
    uint64_t first_default = tile / 8;
    if (pipelined) first_default = n <= tile ? std::max<uint64_t>(1, n / 2) : (n >= 4 * tile ? tile / 2 : tile / 8);  // 1.04 M rays: 64 k is best (1.44 vs 1.48 / 1.51 ms)
    uint64_t cur_tile = std::max<uint64_t>(1, std::min<uint64_t>(tile, (uint64_t)std::max(1, env_int("VT_WAVE_FIRST", (int)first_default))));
    for (uint64_t base = 0, m = 0; base < n; base += m, li = (li + 1) % n_lanes, cur_tile = std::min(tile, cur_tile * 2)) {
        m = std::min(cur_tile, n - base);
        if (n - base - m < cur_tile / 2) m = n - base;  // no small tail tile: its launch-latency chain would run alone at the end
        DeviceScene::WaveLane &l = D.lanes[li];
        const uint64_t cap = std::max<uint64_t>(tile + tile / 2, m);
        l.hits.ensure(cap);
        l.attrs.ensure(cap);
        l.brays.ensure(cap * spp);
        l.bhits.ensure(cap * spp);
        l.fb.ensure(cap * 3);
        l.queue.ensure(cap * spp);
        l.queue_count.ensure(1);
        if (++tiles_since_fence >= kCounterSlots / 4 - 8) {  // see TraceDiffuseWave: the counter ring must not wrap onto a live slot (two frames may be in flight)
            for (int i = 0; i < n_lanes; i++) VT_CUDA(cudaStreamSynchronize(D.lanes[i].stream));
            tiles_since_fence = 0;
        }
        unsigned long long *c0 = next_counter(), *c1 = next_counter();
        VT_CUDA(cudaMemsetAsync(c0, 0, 16, l.stream));
        VT_CUDA(cudaMemsetAsync(c1, 0, 16, l.stream));
        VT_CUDA(cudaMemsetAsync(l.queue_count.p, 0, sizeof(unsigned long long), l.stream));
        VT_CUDA(cudaMemsetAsync(l.fb.p, 0, m * 3 * sizeof(float), l.stream));
        TileTrace tt{base, m, li, {}};
        const vt_ray *d_tile = upload_tile(base, m, l.stream);
        mark(tt, 0, l.stream);
        VT_CUDA(vt_launch_traverse(D.view, d_tile, l.hits.p, m, false, c0, tile_cfg, l.stream));
        mark(tt, 1, l.stream);
        VT_CUDA(vt_launch_trace_result(D.view, d_tile, l.hits.p, nullptr, l.attrs.p, m, l.stream));
        VT_CUDA(vt_launch_bounce_rays(l.attrs.p, m, spp, seed, base * spp, l.brays.p, live_out ? D.live.p : nullptr, l.stream, l.queue.p,
                                      l.queue_count.p, l.bhits.p));
        mark(tt, 2, l.stream);
        VT_CUDA(vt_launch_traverse(D.view, l.brays.p, l.bhits.p, m * spp, false, c1, tile_cfg, l.stream, false, l.queue.p,
                                   l.queue_count.p));
        mark(tt, 3, l.stream);
        VT_CUDA(vt_launch_accumulate_sky(D.view, l.attrs.p, l.bhits.p, m, spp, weight, l.fb.p, l.stream));
        mark(tt, 4, l.stream);
        mLaunches += 5;
        VT_CUDA(cudaMemcpyAsync(fb + base * 3, l.fb.p, m * 3 * sizeof(float), cudaMemcpyDeviceToHost, l.stream));
        mark(tt, 5, l.stream);
        if (trace) tiles.push_back(tt);
    }
    WaveFrame f;
    f.n_lanes = n_lanes;
    for (int i = 0; i < n_lanes; i++) {
        VT_CUDA(cudaEventCreateWithFlags(&f.done[i], cudaEventDisableTiming | cudaEventBlockingSync));
        VT_CUDA(cudaEventRecord(f.done[i], D.lanes[i].stream));
    }
    f.tiles = std::move(tiles);
    f.ev_begin = ev_begin;
    mWaveFrames.push_back(std::move(f));
}

// populate / refit / destruction: nothing may still be reading the scene or the staging buffers
void AccelStruct::DrainWaveFrames() {
    while (!mWaveFrames.empty()) {
        try {
            RenderDiffuseWaveWait();
        } catch (const std::exception &) {  // a failed frame is dropped with its events; the caller's own operation reports what it finds
        }
    }
}

// Waits for the OLDEST frame in flight (frames complete in the order they were begun).
void AccelStruct::RenderDiffuseWaveWait() {
    if (mWaveFrames.empty()) throw std::runtime_error("render_diffuse_wave: no frame in flight");
    VT_CUDA(cudaSetDevice(mDevice));
    WaveFrame f = std::move(mWaveFrames.front());
    mWaveFrames.pop_front();
    cudaError_t err = cudaSuccess;
    for (int i = 0; i < f.n_lanes; i++) {
        const cudaError_t e = cudaEventSynchronize(f.done[i]);
        if (e != cudaSuccess) err = e;
        cudaEventDestroy(f.done[i]);
    }
    VT_CUDA(err);
    if (f.ev_begin) {
        for (WaveFrame::TileTrace &t : f.tiles) {
            float ms[6];
            for (int k = 0; k < 6; k++) {
                VT_CUDA(cudaEventElapsedTime(&ms[k], f.ev_begin, t.ev[k]));
                cudaEventDestroy(t.ev[k]);
            }
            std::fprintf(stderr, "[wave] tile base %8llu rays %7llu lane %d: H2D %.3f  K1p %.3f  K2K3 %.3f  K1b %.3f  K4 %.3f  D2H %.3f ms\n",
                         (unsigned long long)t.base, (unsigned long long)t.m, t.lane, ms[0], ms[1], ms[2], ms[3], ms[4], ms[5]);
        }
        cudaEventDestroy(f.ev_begin);
    }
}

void AccelStruct::RenderDiffuseWave(const vt_ray *rays, uint64_t n, uint32_t spp, uint64_t seed, float weight, float *fb,
                                    uint64_t *live_out) {
    if (live_out) *live_out = 0;
    if (n == 0) return;
    if (!mWaveFrames.empty()) throw std::runtime_error("render_diffuse_wave: frames begun asynchronously are still in flight");
    RenderDiffuseWaveBegin(rays, n, spp, seed, weight, fb, live_out != nullptr, false);
    RenderDiffuseWaveWait();
    if (live_out) {
        unsigned long long v = 0;
        VT_CUDA(cudaMemcpy(&v, mpDevice->live.p, sizeof(v), cudaMemcpyDeviceToHost));
        *live_out = v;
    }
}

void AccelStruct::TracePaths(const vt_ray *rays, uint64_t n, uint32_t bounces, const float sun_dir[3], const float sun_rgb[3], uint64_t seed,
                             float weight, float *fb, uint64_t *ray_counts, bool compact, void *stream_, int slot) {
    check_built(mAccelBuilt);
    if (n == 0) return;
    if (!rays || !fb || !sun_dir || !sun_rgb) throw std::runtime_error("trace_paths: null argument");
    if (bounces > 8) throw std::runtime_error("trace_paths: at most 8 bounces");
    if (n > 0xFFFFFFFFull) throw std::runtime_error("trace_paths: more than 2^32 paths per call");
    if (((uintptr_t)rays & 31)) throw std::runtime_error("trace_paths: device ray buffers must be 32-byte aligned");
    VT_CUDA(cudaSetDevice(mDevice));
    DeviceScene &D = *mpDevice;
    DeviceScene::PathScratch &P = D.path[slot & 1];
    cudaStream_t stream = (cudaStream_t)stream_;
    for (int i = 0; i < 2; i++) P.hits[i].ensure(n), P.attrs[i].ensure(n);
    for (int i = 0; i < 3; i++) P.queue[i].ensure(n);
    P.shits.ensure(n), P.brays.ensure(n), P.srays.ensure(n), P.throughput.ensure(3 * n);
    P.counts.ensure(32);
    VT_CUDA(cudaMemsetAsync(P.counts.p, 0, 32 * sizeof(unsigned long long), stream));
    auto next_counter = [&]() {
        unsigned long long *c = D.counters.p + 2 * (D.next_slot.fetch_add(1) % kCounterSlots);
        VT_CUDA(cudaMemsetAsync(c, 0, 16, stream));
        return c;
    };
    auto count = [&](int i) { return P.counts.p + i; };
    auto record = [&](int slot, const unsigned long long *src) {  // rays of a wave = what its queue counted
        VT_CUDA(cudaMemcpyAsync(P.counts.p + slot, src, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, stream));
    };
    // ---- primary wave: K1 + K2 over every ray, then one shadow ray per surface hit (any hit), then shading
    VT_CUDA(vt_launch_traverse(D.view, rays, P.hits[0].p, n, false, next_counter(), D.cfg, stream));
    VT_CUDA(vt_launch_trace_result(D.view, rays, P.hits[0].p, nullptr, P.attrs[0].p, n, stream));
    uint32_t *q_live = P.queue[0].p, *q_next = P.queue[1].p, *q_shadow = P.queue[2].p;  // live vertices of this wave / of the next / shadow rays
    unsigned long long *c_live = count(0), *c_next = count(1), *c_shadow = count(2);
    VT_CUDA(vt_launch_shadow_rays(P.attrs[0].p, n, sun_dir, false, FLT_MAX, P.srays.p, nullptr, stream, q_shadow, c_shadow, P.shits.p));
    VT_CUDA(vt_launch_traverse(D.view, P.srays.p, P.shits.p, n, true, next_counter(), D.cfg, stream, false, q_shadow, c_shadow));
    VT_CUDA(vt_launch_path_shade(P.attrs[0].p, P.shits.p, nullptr, nullptr, n, true, weight, sun_rgb, P.throughput.p, fb, stream));
    record(9, c_shadow);
    mLaunches += 5;
    int cur = 0;
    bool have_live_queue = false;  // the primary wave has no queue: every slot is a vertex
    for (uint32_t k = 0; k < bounces; k++) {
        const int nxt = cur ^ 1;
        // bounce rays of the live vertices of wave k -> slots of wave k + 1 (slot = pixel), listed in q_next
        VT_CUDA(cudaMemsetAsync(c_next, 0, sizeof(unsigned long long), stream));
        const bool from_queue = compact && have_live_queue;
        VT_CUDA(vt_launch_bounce_rays(P.attrs[cur].p, n, 1, seed + 0x9E3779B97F4A7C15ull * (k + 1), 0, P.brays.p, nullptr, stream, q_next, c_next, P.hits[nxt].p, nullptr,
                                      from_queue ? q_live : nullptr, from_queue ? c_live : nullptr));
        VT_CUDA(vt_launch_traverse(D.view, P.brays.p, P.hits[nxt].p, n, false, next_counter(), D.cfg, stream, false, q_next, c_next));
        if (compact) VT_CUDA(vt_launch_trace_result(D.view, P.brays.p, P.hits[nxt].p, nullptr, P.attrs[nxt].p, n, stream, q_next, c_next));
        else VT_CUDA(vt_launch_trace_result(D.view, P.brays.p, P.hits[nxt].p, nullptr, P.attrs[nxt].p, n, stream));
        record(10 + 2 * k, c_next);
        // shadow rays of wave k + 1's vertices
        VT_CUDA(cudaMemsetAsync(c_shadow, 0, sizeof(unsigned long long), stream));
        VT_CUDA(vt_launch_shadow_rays(P.attrs[nxt].p, n, sun_dir, false, FLT_MAX, P.srays.p, nullptr, stream, q_shadow, c_shadow, P.shits.p,
                                      compact ? q_next : nullptr, compact ? c_next : nullptr));
        VT_CUDA(vt_launch_traverse(D.view, P.srays.p, P.shits.p, n, true, next_counter(), D.cfg, stream, false, q_shadow, c_shadow));
        record(11 + 2 * k, c_shadow);
        VT_CUDA(vt_launch_path_shade(P.attrs[nxt].p, P.shits.p, q_next, c_next, n, false, weight, sun_rgb, P.throughput.p, fb, stream));
        mLaunches += 6;
        std::swap(q_live, q_next);
        std::swap(c_live, c_next);
        have_live_queue = true;
        cur = nxt;
    }
    if (ray_counts) {  // [0] = primary rays, [1] = their shadow rays, [2 + 2k], [3 + 2k] = bounce k + 1 and its shadow rays
        unsigned long long h[32];
        VT_CUDA(cudaMemcpyAsync(h, P.counts.p, sizeof(h), cudaMemcpyDeviceToHost, stream));
        VT_CUDA(cudaStreamSynchronize(stream));
        ray_counts[0] = n;
        ray_counts[1] = h[9];
        for (uint32_t k = 0; k < bounces; k++) ray_counts[2 + 2 * k] = h[10 + 2 * k], ray_counts[3 + 2 * k] = h[11 + 2 * k];
    }
}

void AccelStruct::AccumulateSky(const vt_attr *attrs, const vt_hit *bounce_hits, uint64_t n, uint32_t spp, float weight,
                                float *fb, void *stream_) {
    check_built(mAccelBuilt);
    if (n == 0) return;
    if (!attrs || !bounce_hits || !fb) throw std::runtime_error("accumulate_sky: null argument");
    VT_CUDA(cudaSetDevice(mDevice));
    VT_CUDA(vt_launch_accumulate_sky(mpDevice->view, attrs, bounce_hits, n, spp, weight, fb, (cudaStream_t)stream_));
    mLaunches++;
}

TraceResult *AccelStruct::Traverse(const float origin[3], const float direction[3], float tMin, float tMax, float coneWidth,
                                   float coneAngle) {
    check_built(mAccelBuilt);
    // argument rules and messages of source/objects/AccelStruct.cpp:802-806
    if (coneWidth >= 0 && coneAngle <= 0.f) throw std::invalid_argument("Valid cone width but invalid cone angle passed");
    if (coneWidth < 0 && coneAngle > 0.f) throw std::invalid_argument("Valid cone angle but invalid cone width passed");
    if (tMin < 0.f) throw std::invalid_argument("tMin cannot be less than 0");
    if (tMax <= tMin) throw std::invalid_argument("tMax must be greater than tMin");
    vt_ray ray{origin[0], origin[1], origin[2], tMin, direction[0], direction[1], direction[2], tMax};
    vt_hit hit;
    vt_attr attr;
    const float cone[2] = {coneWidth, coneAngle};
    TraverseBatch(&ray, 1, &hit, &attr, cone, 0, nullptr);
    if (hit.prim == VT_MISS) return nullptr;  // Lua nil (AccelStruct.cpp:837)
    return new TraceResult(attr);
}

}  // namespace vt

// ================================================================================ C ABI

extern "C" {

const char *vt_last_error(void) { return vt::g_last_error.c_str(); }

int vt_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        vt::g_last_error = cudaGetErrorString(e);
        return -1;
    }
    return n;
}

vt_accel *vt_accel_create(int device) {
    VT_TRY
    return new vt_accel(device);
    VT_CATCH(nullptr)
}

void vt_accel_destroy(vt_accel *a) { delete a; }

int vt_accel_populate(vt_accel *a, const vt_scene *scene) {
    VT_TRY
    if (!a || !scene) throw std::runtime_error("null argument");
    a->impl.Populate(*scene);
    return 0;
    VT_CATCH(1)
}

int vt_accel_populate_with_bvh(vt_accel *a, const vt_scene *scene, const vt_node *nodes, uint64_t node_count,
                               const uint64_t *prim_indices) {
    VT_TRY
    if (!a || !scene) throw std::runtime_error("null argument");
    a->impl.PopulateWithBvh(*scene, nodes, node_count, prim_indices);
    return 0;
    VT_CATCH(1)
}

int vt_accel_refit(vt_accel *a, const vt_scene *scene) {
    VT_TRY
    if (!a || !scene) throw std::runtime_error("null argument");
    a->impl.Refit(*scene);
    return 0;
    VT_CATCH(1)
}

int vt_accel_refit_range(vt_accel *a, const vt_tri_in *tris, uint64_t first, uint64_t count) {
    VT_TRY
    if (!a) throw std::runtime_error("null argument");
    a->impl.RefitRange(tris, first, count);
    return 0;
    VT_CATCH(1)
}

int vt_refit_bvh(const vt_scene *scene, vt_node *nodes, uint64_t node_count, const uint64_t *prim_indices) {
    VT_TRY
    if (!scene || !nodes || !prim_indices) throw std::runtime_error("null argument");
    vt::TriangleVec tris(scene->n_tris);
#pragma omp parallel for
    for (int64_t i = 0; i < (int64_t)scene->n_tris; i++) {
        const vt_tri_in &in = scene->tris[i];
        tris[i] = vt::Triangle(in.p[0], in.p[1], in.p[2], in.material, in.uvs, in.one_sided != 0);
    }
    vt::HostBvh bvh;
    bvh.nodes.assign(nodes, nodes + node_count);
    bvh.prim_indices.assign(prim_indices, prim_indices + scene->n_tris);
    std::string err;
    if (!vt::refit_bvh(tris, bvh, err)) throw std::runtime_error(err);
    std::memcpy(nodes, bvh.nodes.data(), node_count * sizeof(vt_node));
    return 0;
    VT_CATCH(1)
}

int vt_accel_get_bvh(const vt_accel *a, vt_node *nodes, uint64_t *node_count, uint64_t *prim_indices, uint64_t *n_tris) {
    VT_TRY
    if (!a) throw std::runtime_error("null argument");
    const vt::HostBvh &b = a->impl.Bvh();
    if (node_count) *node_count = b.nodes.size();
    if (n_tris) *n_tris = b.prim_indices.size();
    if (nodes) std::memcpy(nodes, b.nodes.data(), b.nodes.size() * sizeof(vt_node));
    if (prim_indices) std::memcpy(prim_indices, b.prim_indices.data(), b.prim_indices.size() * sizeof(uint64_t));
    return 0;
    VT_CATCH(1)
}

int vt_accel_traverse(vt_accel *a, const vt_ray *rays, uint64_t n, vt_hit *hits, vt_attr *attrs, uint32_t flags, void *stream) {
    VT_TRY
    if (!a) throw std::runtime_error("null argument");
    a->impl.TraverseBatch(rays, n, hits, attrs, nullptr, flags, stream);
    return 0;
    VT_CATCH(1)
}

int vt_accel_traverse_cones(vt_accel *a, const vt_ray *rays, const float *cones, uint64_t n, vt_hit *hits, vt_attr *attrs,
                            uint32_t flags, void *stream) {
    VT_TRY
    if (!a) throw std::runtime_error("null argument");
    a->impl.TraverseBatch(rays, n, hits, attrs, cones, flags, stream);
    return 0;
    VT_CATCH(1)
}

int vt_accel_trace_result(vt_accel *a, const vt_ray *rays, const vt_hit *hits, uint64_t n, vt_attr *attrs, uint32_t flags,
                          void *stream) {
    VT_TRY
    if (!a) throw std::runtime_error("null argument");
    a->impl.TraceResultBatch(rays, hits, n, attrs, nullptr, flags, stream);
    return 0;
    VT_CATCH(1)
}

int vt_accel_bounce_rays(vt_accel *a, const vt_attr *attrs, uint64_t n, uint32_t spp, uint64_t seed, vt_ray *out_rays,
                         uint64_t *live_out, uint32_t flags, void *stream) {
    VT_TRY
    if (!a) throw std::runtime_error("null argument");
    a->impl.BounceRays(attrs, n, spp, seed, out_rays, live_out, flags, stream);
    return 0;
    VT_CATCH(1)
}

int vt_accel_sample_bsdf_rays(vt_accel *a, const vt_ray *rays, const vt_attr *attrs, uint64_t n, uint32_t spp, uint64_t seed, vt_ray *out_rays,
                              vt_bsdf_sample *out_samples, uint64_t *live_out, uint32_t *queue, uint64_t *queue_count, vt_hit *miss_hits,
                              uint32_t flags, void *stream) {
    VT_TRY
    if (!a) throw std::runtime_error("null argument");
    a->impl.SampleBsdfRays(rays, attrs, n, spp, seed, out_rays, out_samples, live_out, queue, queue_count, miss_hits, flags, stream);
    return 0;
    VT_CATCH(1)
}

float vt_sample_uniform01(uint64_t slot, uint32_t dim, uint64_t seed) { return vt_uniform01(slot, dim, seed); }

int vt_accel_shadow_rays(vt_accel *a, const vt_attr *attrs, uint64_t n, const float light[3], int point_light, float tmax,
                         vt_ray *out_rays, uint64_t *live_out, uint32_t flags, void *stream) {
    VT_TRY
    if (!a) throw std::runtime_error("null argument");
    a->impl.ShadowRays(attrs, n, light, point_light != 0, tmax, out_rays, live_out, flags, stream);
    return 0;
    VT_CATCH(1)
}

int vt_accel_bounce_rays_queued(vt_accel *a, const vt_attr *attrs, uint64_t n, uint32_t spp, uint64_t seed, vt_ray *out_rays,
                                uint32_t *queue, uint64_t *queue_count, vt_hit *miss_hits, void *stream) {
    VT_TRY
    if (!a) throw std::runtime_error("null argument");
    a->impl.BounceRaysQueued(attrs, n, spp, seed, out_rays, queue, queue_count, miss_hits, stream);
    return 0;
    VT_CATCH(1)
}

int vt_accel_shadow_rays_queued(vt_accel *a, const vt_attr *attrs, uint64_t n, const float light[3], int point_light, float tmax,
                                vt_ray *out_rays, uint32_t *queue, uint64_t *queue_count, vt_hit *miss_hits, void *stream) {
    VT_TRY
    if (!a) throw std::runtime_error("null argument");
    a->impl.ShadowRaysQueued(attrs, n, light, point_light != 0, tmax, out_rays, queue, queue_count, miss_hits, stream);
    return 0;
    VT_CATCH(1)
}

int vt_accel_bounce_rays_requeued(vt_accel *a, const vt_attr *attrs, const uint32_t *in_queue, const uint64_t *in_count, uint64_t n, uint32_t spp,
                                  uint64_t seed, vt_ray *out_rays, uint32_t *queue, uint64_t *queue_count, vt_hit *miss_hits, void *stream) {
    VT_TRY
    if (!a || !in_queue || !in_count) throw std::runtime_error("null argument");
    a->impl.BounceRaysQueued(attrs, n, spp, seed, out_rays, queue, queue_count, miss_hits, stream, in_queue, in_count);
    return 0;
    VT_CATCH(1)
}

int vt_accel_shadow_rays_requeued(vt_accel *a, const vt_attr *attrs, const uint32_t *in_queue, const uint64_t *in_count, uint64_t n,
                                  const float light[3], int point_light, float tmax, vt_ray *out_rays, uint32_t *queue, uint64_t *queue_count,
                                  vt_hit *miss_hits, void *stream) {
    VT_TRY
    if (!a || !in_queue || !in_count) throw std::runtime_error("null argument");
    a->impl.ShadowRaysQueued(attrs, n, light, point_light != 0, tmax, out_rays, queue, queue_count, miss_hits, stream, in_queue, in_count);
    return 0;
    VT_CATCH(1)
}

int vt_accel_traverse_queued(vt_accel *a, const vt_ray *rays, const uint32_t *queue, const uint64_t *queue_count, uint64_t capacity,
                             vt_hit *hits, vt_attr *attrs, uint32_t flags, void *stream) {
    VT_TRY
    if (!a) throw std::runtime_error("null argument");
    a->impl.TraverseQueued(rays, queue, queue_count, capacity, hits, attrs, flags, stream);
    return 0;
    VT_CATCH(1)
}

int vt_accel_trace_diffuse_wave(vt_accel *a, const vt_ray *rays, uint64_t n, uint32_t spp, uint64_t seed, vt_hit *hits,
                                vt_attr *attrs, vt_ray *bounce_rays, vt_hit *bounce_hits, uint64_t *live_out, uint32_t flags,
                                void *stream) {
    VT_TRY
    if (!a) throw std::runtime_error("null argument");
    a->impl.TraceDiffuseWave(rays, n, spp, seed, hits, attrs, bounce_rays, bounce_hits, live_out, flags, stream);
    return 0;
    VT_CATCH(1)
}

int vt_accel_render_diffuse_wave(vt_accel *a, const vt_ray *rays, uint64_t n, uint32_t spp, uint64_t seed, float weight,
                                 float *framebuffer_rgb, uint64_t *live_out) {
    VT_TRY
    if (!a) throw std::runtime_error("null argument");
    a->impl.RenderDiffuseWave(rays, n, spp, seed, weight, framebuffer_rgb, live_out);
    return 0;
    VT_CATCH(1)
}

int vt_accel_trace_paths(vt_accel *a, const vt_ray *rays, uint64_t n, uint32_t bounces, const float sun_dir[3], const float sun_rgb[3],
                         uint64_t seed, float weight, float *framebuffer_rgb, uint64_t *ray_counts, uint32_t flags, void *stream) {
    VT_TRY
    if (!a) throw std::runtime_error("null argument");
    if (!(flags & VT_TRAVERSE_DEVICE_PTRS)) throw std::runtime_error("vt_accel_trace_paths: device pointers only (VT_TRAVERSE_DEVICE_PTRS)");
    a->impl.TracePaths(rays, n, bounces, sun_dir, sun_rgb, seed, weight, framebuffer_rgb, ray_counts, !(flags & VT_PATHS_NO_COMPACTION), stream,
                       (flags & VT_PATHS_SLOT1) ? 1 : 0);
    return 0;
    VT_CATCH(1)
}

int vt_accel_accumulate_sky(vt_accel *a, const vt_attr *attrs, const vt_hit *bounce_hits, uint64_t n, uint32_t spp,
                            float weight, float *framebuffer_rgb, void *stream) {
    VT_TRY
    if (!a) throw std::runtime_error("null argument");
    a->impl.AccumulateSky(attrs, bounce_hits, n, spp, weight, framebuffer_rgb, stream);
    return 0;
    VT_CATCH(1)
}

int vt_accel_set_layout(vt_accel *a, int layout) {
    VT_TRY
    if (!a) throw std::runtime_error("null argument");
    if (layout != VT_LAYOUT_EXACT && layout != VT_LAYOUT_COMPACT && layout != VT_LAYOUT_QUAD) throw std::runtime_error("unknown layout");
    a->impl.SetLayout(layout);
    return 0;
    VT_CATCH(1)
}

int vt_accel_get_layout(const vt_accel *a) { return a ? a->impl.Layout() : VT_LAYOUT_EXACT; }

int vt_accel_traverse_stats(vt_accel *a, const vt_ray *rays, uint64_t n, uint32_t flags, uint64_t *steps, uint64_t *tests) {
    VT_TRY
    if (!a) throw std::runtime_error("null argument");
    a->impl.TraverseStats(rays, n, flags, steps, tests);
    return 0;
    VT_CATCH(1)
}

int vt_accel_traverse_ray_stats(vt_accel *a, const vt_ray *rays, uint64_t n, uint32_t flags, uint32_t *per_ray) {
    VT_TRY
    if (!a || !per_ray) throw std::runtime_error("null argument");
    a->impl.TraverseStats(rays, n, flags, nullptr, nullptr, per_ray);
    return 0;
    VT_CATCH(1)
}

int vt_accel_refit_quality(const vt_accel *a, double *area_ratio, uint64_t *rebuilds) {
    VT_TRY
    if (!a) throw std::runtime_error("null argument");
    if (area_ratio) *area_ratio = a->impl.RefitQuality();
    if (rebuilds) *rebuilds = a->impl.Rebuilds();
    return 0;
    VT_CATCH(1)
}

int vt_accel_set_refit_rebuild_ratio(vt_accel *a, double ratio) {
    VT_TRY
    if (!a) throw std::runtime_error("null argument");
    a->impl.SetRefitRebuildRatio(ratio);
    return 0;
    VT_CATCH(1)
}

uint64_t vt_accel_invalid_rays(const vt_accel *a) { return a ? a->impl.InvalidRays() : 0; }
uint64_t vt_accel_launch_count(const vt_accel *a) { return a ? a->impl.Launches() : 0; }

int vt_accel_stats(const vt_accel *a, uint64_t *n_tris, uint64_t *node_count, uint64_t *device_bytes) {
    VT_TRY
    if (!a) throw std::runtime_error("null argument");
    if (n_tris) *n_tris = a->impl.Triangles().size();
    if (node_count) *node_count = a->impl.Bvh().nodes.size();
    if (device_bytes) *device_bytes = a->impl.DeviceBytes();
    return 0;
    VT_CATCH(1)
}

// Triangle constructor output for parity checks of the derived fields: n x 16 floats
// {p0, e1, e2, n, nNorm, lod} (source/objects/Primitives.h:75-102).
int vt_accel_get_tri_derived(const vt_accel *a, float *out16) {
    VT_TRY
    if (!a || !out16) throw std::runtime_error("null argument");
    const auto &tris = a->impl.Triangles();
    for (size_t i = 0; i < tris.size(); i++) {
        const vt::Triangle &t = tris[i];
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
    return 0;
    VT_CATCH(1)
}

// ---- host-only entry points (no GPU touched): the build and flatten steps on their own

int vt_build_bvh(const vt_scene *scene, vt_node *nodes, uint64_t *node_count, uint64_t *prim_indices) {
    VT_TRY
    if (!scene || !node_count) throw std::runtime_error("null argument");
    vt::TriangleVec tris(scene->n_tris);
#pragma omp parallel for
    for (int64_t i = 0; i < (int64_t)scene->n_tris; i++) {
        const vt_tri_in &in = scene->tris[i];
        tris[i] = vt::Triangle(in.p[0], in.p[1], in.p[2], in.material, in.uvs, in.one_sided != 0);
    }
    vt::HostBvh bvh;
    vt::build_bvh(tris, bvh, vt::env_int("VT_MAX_LEAF", 4), vt::env_float("VT_TRAV_COST", 1.0f), (uint32_t)std::max(0, vt::env_int("VT_SAH_SWEEP", 0)));
    if (const int passes = vt::env_int("VT_REINSERT", 0); passes > 0 && !vt::reinsert_optimize(bvh, passes, vt::env_float("VT_REINSERT_FRACTION", 0.05f)))
        throw std::runtime_error("reinsertion: malformed hierarchy");
    if (nodes) {
        if (*node_count < bvh.nodes.size()) throw std::runtime_error("node buffer too small");
        std::memcpy(nodes, bvh.nodes.data(), bvh.nodes.size() * sizeof(vt_node));
        if (prim_indices) std::memcpy(prim_indices, bvh.prim_indices.data(), bvh.prim_indices.size() * sizeof(uint64_t));
    }
    *node_count = bvh.nodes.size();
    return 0;
    VT_CATCH(1)
}

int vt_optimize_bvh(vt_node *nodes, uint64_t node_count, int iterations, double fraction, double *area_before, double *area_after, uint64_t *moves) {
    VT_TRY
    if (!nodes || node_count == 0) throw std::runtime_error("null hierarchy");
    if (node_count % 2 == 0) throw std::runtime_error("optimize_bvh: a bvh::Bvh-form hierarchy has an odd node count (root + sibling pairs)");
    for (uint64_t i = 0; i < node_count; i++)  // untrusted array: every inner node must name a pair inside it
        if (nodes[i].prim_count == 0 && (nodes[i].first == 0 || nodes[i].first % 2 == 0 || (uint64_t)nodes[i].first + 1 >= node_count))
            throw std::runtime_error("optimize_bvh: inner node " + std::to_string(i) + " does not reference a sibling pair of the array");
    vt::HostBvh bvh;
    bvh.nodes.assign(nodes, nodes + node_count);
    if (!vt::reinsert_optimize(bvh, iterations, (float)fraction, area_before, area_after, moves)) throw std::runtime_error("optimize_bvh: the node array is not one tree");
    std::memcpy(nodes, bvh.nodes.data(), node_count * sizeof(vt_node));
    return 0;
    VT_CATCH(1)
}

int vt_build_bvh_ploc(const vt_scene *scene, int collapse, vt_node *nodes, uint64_t *node_count, uint64_t *prim_indices) {
    VT_TRY
    if (!scene || !node_count) throw std::runtime_error("null argument");
    vt::TriangleVec tris(scene->n_tris);
#pragma omp parallel for
    for (int64_t i = 0; i < (int64_t)scene->n_tris; i++) {
        const vt_tri_in &in = scene->tris[i];
        tris[i] = vt::Triangle(in.p[0], in.p[1], in.p[2], in.material, in.uvs, in.one_sided != 0);
    }
    vt::HostBvh bvh;
    vt::build_bvh_ploc(tris, bvh);
    if (collapse) vt::collapse_leaves(bvh);
    if (nodes) {
        if (*node_count < bvh.nodes.size()) throw std::runtime_error("node buffer too small");
        std::memcpy(nodes, bvh.nodes.data(), bvh.nodes.size() * sizeof(vt_node));
        if (prim_indices) std::memcpy(prim_indices, bvh.prim_indices.data(), bvh.prim_indices.size() * sizeof(uint64_t));
    }
    *node_count = bvh.nodes.size();
    return 0;
    VT_CATCH(1)
}

int vt_flatten_bvh(const vt_node *nodes, uint64_t node_count, const uint64_t *prim_indices, uint64_t n_tris,
                   uint32_t bfs_pairs, void *pairs_out, uint32_t *leaf_order_out, uint32_t *root_leaf_count,
                   uint32_t *max_depth) {
    VT_TRY
    if (!nodes || !prim_indices) throw std::runtime_error("null argument");
    vt::HostBvh bvh;
    bvh.nodes.assign(nodes, nodes + node_count);
    bvh.prim_indices.assign(prim_indices, prim_indices + n_tris);
    vt::FlatBvh flat;
    std::string err;
    if (!vt::flatten_bvh(bvh, n_tris, bfs_pairs, flat, err)) throw std::runtime_error(err);
    if (pairs_out) std::memcpy(pairs_out, flat.pairs.data(), flat.pairs.size() * sizeof(VtPair));
    if (leaf_order_out) std::memcpy(leaf_order_out, flat.leaf_order.data(), flat.leaf_order.size() * sizeof(uint32_t));
    if (root_leaf_count) *root_leaf_count = flat.root_leaf_count;
    if (max_depth) *max_depth = flat.max_depth;
    return 0;
    VT_CATCH(1)
}

int vt_vtf_read_info(const uint8_t *file, uint64_t size, vt_vtf_info *info) {
    VT_TRY
    if (!file || !info) throw std::runtime_error("null argument");
    vt::VtfInfo(file, size, info);
    return 0;
    VT_CATCH(1)
}

int vt_vtf_decode(const uint8_t *file, uint64_t size, uint32_t frame, uint32_t face, uint8_t *rgba_out, uint64_t capacity,
                  vt_vtf_info *info_or_null) {
    VT_TRY
    if (!file) throw std::runtime_error("null argument");
    vt::VtfDecode(file, size, frame, face, rgba_out, capacity, info_or_null);
    return 0;
    VT_CATCH(1)
}

int vt_mdl_read_info(const vt_mdl_files *files, vt_mdl_info *info) {
    VT_TRY
    if (!files || !info) throw std::runtime_error("null argument");
    vt::MdlInfo(files, info);
    return 0;
    VT_CATCH(1)
}

int vt_mdl_bodygroup_values(const vt_mdl_files *files, uint32_t bodygroup, uint32_t *n_values) {
    VT_TRY
    if (!files || !n_values) throw std::runtime_error("null argument");
    *n_values = vt::MdlBodygroupValues(files, bodygroup);
    return 0;
    VT_CATCH(1)
}

int vt_mdl_mesh_triangles(const vt_mdl_files *files, uint32_t bodygroup, uint32_t value, vt_tri_in *tris, vt_tri_skin *skin, uint64_t *n_tris) {
    VT_TRY
    if (!files || !n_tris) throw std::runtime_error("null argument");
    *n_tris = vt::MdlMeshTriangles(files, bodygroup, value, tris, skin, tris ? *n_tris : 0);
    return 0;
    VT_CATCH(1)
}

int vt_mdl_bind_matrices(const vt_mdl_files *files, float *out16) {
    VT_TRY
    if (!files || !out16) throw std::runtime_error("null argument");
    vt::MdlBindMatrices(files, out16);
    return 0;
    VT_CATCH(1)
}

int vt_mdl_material_index(const vt_mdl_files *files, uint32_t skin, uint32_t material_id, int32_t *index) {
    VT_TRY
    if (!files || !index) throw std::runtime_error("null argument");
    *index = vt::MdlMaterialIndex(files, skin, material_id);
    return 0;
    VT_CATCH(1)
}

int vt_mdl_material_path(const vt_mdl_files *files, uint32_t material_id, uint32_t dir, char *out, uint64_t capacity) {
    VT_TRY
    if (!files || !out || !capacity) throw std::runtime_error("null argument");
    const std::string s = vt::MdlMaterialPath(files, material_id, dir);
    if (s.size() + 1 > capacity) throw std::runtime_error("mdl: path buffer too small");
    std::memcpy(out, s.c_str(), s.size() + 1);
    return 0;
    VT_CATCH(1)
}

int vt_bsp_read_info(const uint8_t *file, uint64_t size, vt_bsp_info *info) {
    VT_TRY
    if (!file || !info) throw std::runtime_error("null argument");
    vt::BspInfo(file, size, info);
    return 0;
    VT_CATCH(1)
}

int vt_bsp_triangles(const uint8_t *file, uint64_t size, vt_tri_in *tris, float *binormals_or_null, int16_t *texinfo_or_null, uint64_t *n_tris) {
    VT_TRY
    if (!file || !n_tris) throw std::runtime_error("null argument");
    *n_tris = vt::BspTriangles(file, size, tris, binormals_or_null, texinfo_or_null, tris ? *n_tris : 0);
    return 0;
    VT_CATCH(1)
}

int vt_bsp_get_material(const uint8_t *file, uint64_t size, uint32_t material, vt_bsp_material *out) {
    VT_TRY
    if (!file || !out) throw std::runtime_error("null argument");
    vt::BspMaterial(file, size, material, out);
    return 0;
    VT_CATCH(1)
}

int vt_bsp_get_static_prop(const uint8_t *file, uint64_t size, uint32_t index, vt_bsp_static_prop *out) {
    VT_TRY
    if (!file || !out) throw std::runtime_error("null argument");
    vt::BspStaticProp(file, size, index, out);
    return 0;
    VT_CATCH(1)
}

int vt_accel_render_diffuse_wave_begin(vt_accel *a, const vt_ray *rays, uint64_t n, uint32_t spp, uint64_t seed, float weight, float *framebuffer_rgb) {
    VT_TRY
    if (!a) throw std::runtime_error("null argument");
    if (n == 0) throw std::runtime_error("render_diffuse_wave_begin: empty frame");
    a->impl.RenderDiffuseWaveBegin(rays, n, spp, seed, weight, framebuffer_rgb, false, true);
    return 0;
    VT_CATCH(1)
}

int vt_accel_render_diffuse_wave_wait(vt_accel *a) {
    VT_TRY
    if (!a) throw std::runtime_error("null argument");
    a->impl.RenderDiffuseWaveWait();
    return 0;
    VT_CATCH(1)
}

uint32_t vt_quad_plane_offset(void) { return (uint32_t)VT_QUAD_OFFSET; }

int vt_build_quads(const vt_node *nodes, uint64_t node_count, const uint64_t *prim_indices, uint64_t n_tris, void *quads_out,
                   uint64_t *n_quads, uint32_t *leaf_order_out, uint32_t *root_leaf_count, uint32_t *max_stack) {
    VT_TRY
    if (!nodes || !prim_indices || !n_quads) throw std::runtime_error("null argument");
    vt::HostBvh bvh;
    bvh.nodes.assign(nodes, nodes + node_count);
    bvh.prim_indices.assign(prim_indices, prim_indices + n_tris);
    vt::QuadBvh q;
    std::string err;
    if (!vt::build_quads(bvh, n_tris, q, err)) throw std::runtime_error(err);
    if (quads_out) {
        if (*n_quads < q.quads.size()) throw std::runtime_error("quad buffer too small");
        std::memcpy(quads_out, q.quads.data(), q.quads.size() * sizeof(VtQuad));
        if (leaf_order_out) std::memcpy(leaf_order_out, q.leaf_order.data(), q.leaf_order.size() * sizeof(uint32_t));
    }
    *n_quads = q.quads.size();
    if (root_leaf_count) *root_leaf_count = q.root_leaf_count;
    if (max_stack) *max_stack = q.max_stack;
    return 0;
    VT_CATCH(1)
}

int vt_skin_triangles(vt_tri_in *tris, const vt_tri_skin *skin, uint64_t n, const float *bones, const float *binds, uint32_t n_bones) {
    VT_TRY
    vt::SkinTriangles(tris, skin, n, bones, binds, n_bones);
    return 0;
    VT_CATCH(1)
}

int vt_compact_pairs(const void *pairs, uint64_t n_pairs, void *cpairs_out) {
    VT_TRY
    if (!pairs || !cpairs_out) throw std::runtime_error("null argument");
    std::vector<VtPair> in((const VtPair *)pairs, (const VtPair *)pairs + n_pairs);
    std::vector<VtCPair> out;
    std::string err;
    if (!vt::compact_pairs(in, out, err)) throw std::runtime_error(err);
    std::memcpy(cpairs_out, out.data(), out.size() * sizeof(VtCPair));
    return 0;
    VT_CATCH(1)
}

}  // extern "C"
