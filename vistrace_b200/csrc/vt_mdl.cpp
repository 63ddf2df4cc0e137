// vt_mdl.cpp — host-side ingestion of Source-engine models: .mdl + .vvd + .vtx -> the vt_tri_in / vt_tri_skin records of one
// body-group mesh, the bind matrices, the skin table.
//
// Restates, from the published studiomdl v48 / VVD v4 / VTX v7 file layouts, what the reference does with them:
//   * libs/MDLParser/source/MDLParser.cpp:33-60, VVDParser.cpp:33-83, VTXParser.cpp:33-56 — which files are accepted (ids, versions,
//     the checksum that ties the three files together, LoD counts) and the VVD fix-up table that assembles the root-LoD vertices;
//   * source/objects/Model.cpp:11-128 (Mesh::Mesh) — LoD 0, triangle-LIST strips only (strips are "nyi" there as well), vertex =
//     VVD[origMeshVertId + mesh.vertsOffset + model.vertsOffset / 48], normals / tangents through glm::normalize, a non-finite
//     tangent replaced by normalize(e1), per-vertex bone weights;
//   * Model.cpp:242-254 (bind matrices from Bone::poseToBone), :349-357 (skin table), :256-274 (material names).
// Unlike the reference (which trusts its files: every offset is dereferenced unchecked), every offset, count and index is
// checked against the file sizes here: a malformed file is an error, never an out-of-bounds read.
//
// One deliberate difference, documented rather than copied: Model.cpp:96-99 tests `tri.tangents[j]` — a field that has not been
// assigned yet at that point (indeterminate memory) — to decide whether to replace the NORMAL by the face normal.  With finite
// garbage (what the stack holds in practice: the previous iteration's tangent) the branch is never taken; this restatement never
// takes it.  tests/test_mdl.py compares against the compiled reference and would show a divergence.
#include <cmath>
#include <cstring>
#include <string>

#include "vt_host.h"

namespace vt {

namespace {

constexpr int32_t kIdst = 'I' + ('D' << 8) + ('S' << 16) + ('T' << 24);
constexpr int32_t kIdsv = 'I' + ('D' << 8) + ('S' << 16) + ('V' << 24);
constexpr int kMaxLods = 8;

// Fixed offsets of the v48 studiohdr_t (408 bytes) and the records it points to; all little-endian, unaligned.
constexpr size_t kMdlHeaderSize = 408, kBoneSize = 216, kBonePoseToBone = 96, kTextureSize = 64, kBodyPartSize = 16, kModelSize = 148, kMeshSize = 116;
constexpr size_t kVvdHeaderSize = 64, kVvdVertexSize = 48, kVvdTangentSize = 16, kVvdFixupSize = 12;
constexpr size_t kVtxHeaderSize = 36, kVtxBodyPartSize = 8, kVtxModelSize = 8, kVtxLodSize = 12, kVtxMeshSize = 9, kVtxStripGroupSize = 25, kVtxStripSize = 27,
                 kVtxVertexSize = 9;

struct Span {
    const uint8_t *p = nullptr;
    uint64_t n = 0;
    // bounds-checked view of [off, off + len)
    const uint8_t *at(int64_t off, uint64_t len, const char *what) const {
        if (off < 0 || (uint64_t)off > n || len > n - (uint64_t)off) throw std::runtime_error(std::string("mdl: ") + what + " runs past the end of the file");
        return p + off;
    }
    int32_t i32(int64_t off, const char *what) const {
        int32_t v;
        std::memcpy(&v, at(off, 4, what), 4);
        return v;
    }
    float f32(int64_t off, const char *what) const {
        float v;
        std::memcpy(&v, at(off, 4, what), 4);
        return v;
    }
};

struct Files {
    Span mdl, vvd, vtx;
    int32_t checksum = 0;
    int32_t n_root_verts = 0;
    // root-LoD vertices and tangents as the VVD loader assembles them (fix-ups applied), VVDParser.cpp:57-79
    std::vector<uint8_t> verts, tangents;
};

uint64_t checked_count(int32_t v, const char *what) {
    if (v < 0) throw std::runtime_error(std::string("mdl: negative ") + what);
    return (uint64_t)v;
}

void open_files(const vt_mdl_files *f, Files &F) {
    if (!f || !f->mdl || !f->vvd || !f->vtx || !f->mdl_size || !f->vvd_size || !f->vtx_size) throw std::runtime_error("mdl: null or empty file");
    F.mdl = Span{f->mdl, f->mdl_size}, F.vvd = Span{f->vvd, f->vvd_size}, F.vtx = Span{f->vtx, f->vtx_size};
    // MDLParser.cpp:49-51
    F.mdl.at(0, kMdlHeaderSize, "studio header");
    if (F.mdl.i32(0, "id") != kIdst) throw std::runtime_error("mdl: not an IDST file");
    if (F.mdl.i32(4, "version") > 48) throw std::runtime_error("mdl: versions above 48 are not supported (strip records grow in v49)");
    F.checksum = F.mdl.i32(8, "checksum");
    // VVDParser.cpp:37-47
    F.vvd.at(0, kVvdHeaderSize, "vvd header");
    if (F.vvd.i32(0, "vvd id") != kIdsv || F.vvd.i32(4, "vvd version") != 4) throw std::runtime_error("vvd: not an IDSV version 4 file");
    if (F.vvd.i32(8, "vvd checksum") != F.checksum) throw std::runtime_error("vvd: checksum does not match the mdl");
    const uint64_t n_verts = checked_count(F.vvd.i32(16, "numLoDVertices[0]"), "vertex count");
    const uint64_t n_fixups = checked_count(F.vvd.i32(16 + 4 * kMaxLods, "numFixups"), "fix-up count");
    if (n_verts > (1u << 26) || n_fixups > (1u << 24)) throw std::runtime_error("vvd: absurd vertex or fix-up count");
    const int64_t fixup_off = F.vvd.i32(20 + 4 * kMaxLods, "fixupTableOffset"), vert_off = F.vvd.i32(24 + 4 * kMaxLods, "vertexDataOffset"),
                  tan_off = F.vvd.i32(28 + 4 * kMaxLods, "tangentDataOffset");
    if (kVvdHeaderSize + kVvdFixupSize * n_fixups + (kVvdTangentSize + kVvdVertexSize) * n_verts > F.vvd.n) throw std::runtime_error("vvd: file too small for its vertex count");
    F.n_root_verts = (int32_t)n_verts;
    F.verts.resize(n_verts * kVvdVertexSize);
    F.tangents.resize(n_verts * kVvdTangentSize);
    if (n_fixups == 0) {
        if (n_verts) {
            std::memcpy(F.tangents.data(), F.vvd.at(tan_off, n_verts * kVvdTangentSize, "tangent data"), n_verts * kVvdTangentSize);
            std::memcpy(F.verts.data(), F.vvd.at(vert_off, n_verts * kVvdVertexSize, "vertex data"), n_verts * kVvdVertexSize);
        }
    } else {
        uint64_t out = 0;
        for (uint64_t k = 0; k < n_fixups; k++) {
            const int64_t fo = fixup_off + (int64_t)(k * kVvdFixupSize);
            const int32_t lod = F.vvd.i32(fo, "fix-up"), src = F.vvd.i32(fo + 4, "fix-up"), cnt = F.vvd.i32(fo + 8, "fix-up");
            if (lod < 0) continue;  // root LoD is 0
            if (src < 0 || cnt < 0 || out + (uint64_t)cnt > n_verts) throw std::runtime_error("vvd: fix-up table overruns the root LoD vertex count");
            if (cnt == 0) continue;
            std::memcpy(F.tangents.data() + out * kVvdTangentSize, F.vvd.at(tan_off + (int64_t)src * (int64_t)kVvdTangentSize, (uint64_t)cnt * kVvdTangentSize, "tangent data"),
                        (uint64_t)cnt * kVvdTangentSize);
            std::memcpy(F.verts.data() + out * kVvdVertexSize, F.vvd.at(vert_off + (int64_t)src * (int64_t)kVvdVertexSize, (uint64_t)cnt * kVvdVertexSize, "vertex data"),
                        (uint64_t)cnt * kVvdVertexSize);
            out += (uint64_t)cnt;
        }
        if (out < n_verts) {  // the reference leaves the rest of its malloc'ed arrays uninitialised; zero here
            std::memset(F.tangents.data() + out * kVvdTangentSize, 0, (n_verts - out) * kVvdTangentSize);
            std::memset(F.verts.data() + out * kVvdVertexSize, 0, (n_verts - out) * kVvdVertexSize);
        }
    }
    // VTXParser.cpp:39-53
    F.vtx.at(0, kVtxHeaderSize, "vtx header");
    if (F.vtx.i32(0, "vtx version") != 7) throw std::runtime_error("vtx: not a version 7 file");
    if (F.vtx.i32(16, "vtx checksum") != F.checksum) throw std::runtime_error("vtx: checksum does not match the mdl");
    const int32_t vtx_lods = F.vtx.i32(20, "vtx numLoDs");
    const uint64_t nbp = checked_count(F.vtx.i32(28, "vtx numBodyParts"), "body part count");
    const int64_t bpo = F.vtx.i32(32, "vtx bodyPartOffset");
    for (uint64_t i = 0; i < nbp; i++) {
        const int64_t bp = bpo + (int64_t)(i * kVtxBodyPartSize);
        const uint64_t nm = checked_count(F.vtx.i32(bp, "vtx body part"), "model count");
        const int64_t mo = bp + F.vtx.i32(bp + 4, "vtx body part");
        for (uint64_t j = 0; j < nm; j++)
            if (F.vtx.i32(mo + (int64_t)(j * kVtxModelSize), "vtx model") != vtx_lods) throw std::runtime_error("vtx: a model's LoD count differs from the header's");
    }
}

struct MeshRef {  // one (bodygroup, value): the mdl model record and the vtx LoD-0 record that belong together
    int64_t mdl_model = 0, vtx_lod = 0;
    uint64_t n_meshes = 0;
};

uint64_t bodygroup_count(const Files &F) { return checked_count(F.mdl.i32(232, "bodypartCount"), "body part count"); }

MeshRef find_mesh(const Files &F, uint32_t bodygroup, uint32_t value, uint32_t *n_values) {
    const uint64_t nbg = bodygroup_count(F);
    if (bodygroup >= nbg) throw std::runtime_error("mdl: bodygroup index out of range");
    if ((uint64_t)checked_count(F.vtx.i32(28, "vtx numBodyParts"), "body part count") <= bodygroup) throw std::runtime_error("vtx: fewer body parts than the mdl");
    const int64_t bp = F.mdl.i32(236, "bodypartOffset") + (int64_t)bodygroup * (int64_t)kBodyPartSize;
    F.mdl.at(bp, kBodyPartSize, "body part");
    const uint64_t n_models = checked_count(F.mdl.i32(bp + 4, "modelsCount"), "model count");
    if (n_values) *n_values = (uint32_t)n_models;
    MeshRef r;
    if (value >= n_models) {
        if (n_values) return r;
        throw std::runtime_error("mdl: bodygroup value out of range");
    }
    r.mdl_model = bp + F.mdl.i32(bp + 12, "modelsOffset") + (int64_t)value * (int64_t)kModelSize;
    F.mdl.at(r.mdl_model, kModelSize, "model");
    r.n_meshes = checked_count(F.mdl.i32(r.mdl_model + 72, "meshesCount"), "mesh count");
    const int64_t vbp = F.vtx.i32(32, "vtx bodyPartOffset") + (int64_t)bodygroup * (int64_t)kVtxBodyPartSize;
    if (checked_count(F.vtx.i32(vbp, "vtx body part"), "model count") <= value) throw std::runtime_error("vtx: fewer models than the mdl body part");
    const int64_t vmodel = vbp + F.vtx.i32(vbp + 4, "vtx body part") + (int64_t)value * (int64_t)kVtxModelSize;
    if (F.vtx.i32(vmodel, "vtx model") < 1) throw std::runtime_error("vtx: model without LoD 0");
    r.vtx_lod = vmodel + F.vtx.i32(vmodel + 4, "vtx model");  // LoD 0
    F.vtx.at(r.vtx_lod, kVtxLodSize, "vtx LoD");
    if (checked_count(F.vtx.i32(r.vtx_lod, "vtx LoD"), "mesh count") < r.n_meshes) throw std::runtime_error("vtx: LoD 0 has fewer meshes than the mdl model");
    return r;
}

void normalize3(const float in[3], float out[3]) {  // glm::normalize: v * (1 / sqrt(dot(v, v))), func_geometric.inl:88
    const float d = in[0] * in[0] + in[1] * in[1] + in[2] * in[2];
    const float s = 1.0f / std::sqrt(d);
    out[0] = in[0] * s, out[1] = in[1] * s, out[2] = in[2] * s;
}

// Mesh::Mesh, source/objects/Model.cpp:11-128.  tris == nullptr: count only.
uint64_t mesh_triangles(const Files &F, const MeshRef &r, vt_tri_in *tris, vt_tri_skin *skin, uint64_t capacity) {
    uint64_t n_out = 0;
    const int64_t model_verts_off = F.mdl.i32(r.mdl_model + 84, "model vertsOffset"), model_tan_off = F.mdl.i32(r.mdl_model + 88, "model tangentsOffset");
    const int64_t meshes = r.mdl_model + F.mdl.i32(r.mdl_model + 76, "meshesOffset");
    const int64_t vtx_meshes = r.vtx_lod + F.vtx.i32(r.vtx_lod + 4, "vtx LoD");
    for (uint64_t mi = 0; mi < r.n_meshes; mi++) {
        const int64_t mesh = meshes + (int64_t)(mi * kMeshSize);
        F.mdl.at(mesh, kMeshSize, "mesh");
        const int32_t material = F.mdl.i32(mesh, "mesh material");
        const int64_t mesh_verts_off = F.mdl.i32(mesh + 12, "mesh vertsOffset");
        const int64_t vmesh = vtx_meshes + (int64_t)(mi * kVtxMeshSize);
        const uint64_t n_groups = checked_count(F.vtx.i32(vmesh, "vtx mesh"), "strip group count");
        const int64_t groups = vmesh + F.vtx.i32(vmesh + 4, "vtx mesh");
        for (uint64_t gi = 0; gi < n_groups; gi++) {
            const int64_t sg = groups + (int64_t)(gi * kVtxStripGroupSize);
            F.vtx.at(sg, kVtxStripGroupSize, "strip group");
            const uint64_t sg_verts = checked_count(F.vtx.i32(sg, "strip group"), "vertex count"), sg_indices = checked_count(F.vtx.i32(sg + 8, "strip group"), "index count");
            const int64_t sg_vert_off = sg + F.vtx.i32(sg + 4, "strip group"), sg_index_off = sg + F.vtx.i32(sg + 12, "strip group");
            F.vtx.at(sg_vert_off, sg_verts * kVtxVertexSize, "strip group vertices");
            F.vtx.at(sg_index_off, sg_indices * 2, "strip group indices");
            const uint64_t n_strips = checked_count(F.vtx.i32(sg + 16, "strip group"), "strip count");
            const int64_t strips = sg + F.vtx.i32(sg + 20, "strip group");
            for (uint64_t si = 0; si < n_strips; si++) {
                const int64_t st = strips + (int64_t)(si * kVtxStripSize);
                F.vtx.at(st, kVtxStripSize, "strip");
                const uint8_t flags = *F.vtx.at(st + 18, 1, "strip flags");
                if (!(flags & 0x01)) continue;  // IS_TRILIST only; IS_TRISTRIP is "nyi" in the reference too (Model.cpp:33-35, 120-122)
                const int64_t n_idx = F.vtx.i32(st, "strip"), idx_off = F.vtx.i32(st + 4, "strip");
                if (n_idx < 0 || idx_off < 0 || (uint64_t)(idx_off + n_idx) > sg_indices) throw std::runtime_error("vtx: a strip addresses indices past its strip group");
                // the reference's loop `for (i = off; i < n + off; i += 3)` reads i + 1, i + 2 even when n is not a multiple of 3; such a tail is dropped here
                for (int64_t i = idx_off; i + 2 < idx_off + n_idx; i += 3) {
                    if (!tris) {
                        n_out++;
                        continue;
                    }
                    if (n_out >= capacity) throw std::runtime_error("mdl: triangle buffer too small");
                    vt_tri_in &t = tris[n_out];
                    vt_tri_skin &sk = skin[n_out];
                    std::memset(&t, 0, sizeof(t));
                    std::memset(&sk, 0, sizeof(sk));
                    float tang_raw[3][3];
                    uint8_t vtx_bones[3];
                    const uint8_t *vv[3];
                    for (int j = 0; j < 3; j++) {
                        uint16_t index;
                        std::memcpy(&index, F.vtx.p + sg_index_off + 2 * (i + j), 2);
                        if (index >= sg_verts) throw std::runtime_error("vtx: index past the strip group's vertices");
                        const uint8_t *vx = F.vtx.p + sg_vert_off + (int64_t)index * (int64_t)kVtxVertexSize;
                        vtx_bones[j] = vx[3];
                        uint16_t orig;
                        std::memcpy(&orig, vx + 4, 2);
                        // Mesh::GetVertexIndex / GetTangentIndex, libs/MDLParser/source/Structs.h:546-553
                        const int64_t vi = (int64_t)orig + mesh_verts_off + model_verts_off / (int64_t)kVvdVertexSize;
                        const int64_t ti = (int64_t)orig + mesh_verts_off + model_tan_off / (int64_t)kVvdTangentSize;
                        if (vi < 0 || vi >= F.n_root_verts || ti < 0 || ti >= F.n_root_verts) throw std::runtime_error("mdl: vertex index past the vvd's root LoD");
                        vv[j] = F.verts.data() + vi * (int64_t)kVvdVertexSize;
                        std::memcpy(t.p[j], vv[j] + 16, 12);       // pos
                        std::memcpy(t.uvs[j], vv[j] + 40, 8);      // texCoord
                        float nrm[3];
                        std::memcpy(nrm, vv[j] + 28, 12);
                        normalize3(nrm, t.normals[j]);             // Model.cpp:95 (the replacement at :96-98 never fires, see the header)
                        std::memcpy(tang_raw[j], F.tangents.data() + ti * (int64_t)kVvdTangentSize, 12);
                    }
                    float e1[3];
                    for (int k = 0; k < 3; k++) e1[k] = t.p[0][k] - t.p[1][k];  // Triangle ctor, Primitives.h:82
                    for (int j = 0; j < 3; j++) {
                        normalize3(tang_raw[j], t.tangents[j]);  // :100
                        if (!(std::isfinite(t.tangents[j][0]) && std::isfinite(t.tangents[j][1]) && std::isfinite(t.tangents[j][2]))) normalize3(e1, t.tangents[j]);  // :101-103
                        if (vtx_bones[j] > 0) {  // :105-111
                            sk.num_bones[j] = vtx_bones[j];
                            std::memcpy(sk.weights[j], vv[j], 12);
                            for (int b = 0; b < 3; b++) sk.bone_ids[j][b] = (int8_t)vv[j][12 + b];
                        } else {  // :112-116
                            sk.num_bones[j] = 1;
                            sk.weights[j][0] = 1.f;
                            sk.bone_ids[j][0] = 0;
                        }
                    }
                    t.material = (uint32_t)material;  // model-local material id; the skin table maps it (vt_mdl_material_index)
                    t.ent_idx = 0;
                    t.one_sided = 0;  // model triangles are two-sided (Model.cpp:88)
                    n_out++;
                }
            }
        }
    }
    return n_out;
}

}  // namespace

void MdlInfo(const vt_mdl_files *f, vt_mdl_info *info) {
    Files F;
    open_files(f, F);
    std::memset(info, 0, sizeof(*info));
    info->version = (uint32_t)F.mdl.i32(4, "version");
    info->n_bodygroups = (uint32_t)bodygroup_count(F);
    info->n_bones = (uint32_t)checked_count(F.mdl.i32(156, "boneCount"), "bone count");
    info->n_materials = (uint32_t)checked_count(F.mdl.i32(204, "textureCount"), "material count");
    info->n_material_dirs = (uint32_t)checked_count(F.mdl.i32(212, "textureDirCount"), "material directory count");
    info->n_skin_refs = (uint32_t)checked_count(F.mdl.i32(220, "skinRefCount"), "skin reference count");
    info->n_skin_families = (uint32_t)checked_count(F.mdl.i32(224, "skinFamilyCount"), "skin family count");
    info->n_vertices = (uint32_t)F.n_root_verts;
}

uint32_t MdlBodygroupValues(const vt_mdl_files *f, uint32_t bodygroup) {
    Files F;
    open_files(f, F);
    uint32_t n = 0;
    find_mesh(F, bodygroup, 0xFFFFFFFFu, &n);
    return n;
}

uint64_t MdlMeshTriangles(const vt_mdl_files *f, uint32_t bodygroup, uint32_t value, vt_tri_in *tris, vt_tri_skin *skin, uint64_t capacity) {
    Files F;
    open_files(f, F);
    const MeshRef r = find_mesh(F, bodygroup, value, nullptr);
    if (tris && !skin) throw std::runtime_error("mdl: skin buffer is null");
    return mesh_triangles(F, r, tris, skin, capacity);
}

// Model.cpp:242-254: glm::mat4, column-major, from the 3 x 4 row-major poseToBone of every bone
void MdlBindMatrices(const vt_mdl_files *f, float *out16) {
    Files F;
    open_files(f, F);
    const uint64_t n = checked_count(F.mdl.i32(156, "boneCount"), "bone count");
    const int64_t bones = F.mdl.i32(160, "boneOffset");
    for (uint64_t i = 0; i < n; i++) {
        const int64_t m = bones + (int64_t)(i * kBoneSize) + (int64_t)kBonePoseToBone;
        float r[3][4];
        std::memcpy(r, F.mdl.at(m, 48, "bone"), 48);
        float *o = out16 + 16 * i;
        for (int c = 0; c < 4; c++) {
            o[4 * c + 0] = r[0][c], o[4 * c + 1] = r[1][c], o[4 * c + 2] = r[2][c];
            o[4 * c + 3] = c == 3 ? 1.f : 0.f;
        }
    }
}

// Model::GetMaterialIdx, Model.cpp:349-357 (out of range -> 0) over MDL::GetMaterialIdx = skin table [family * skinRefCount + material]
int32_t MdlMaterialIndex(const vt_mdl_files *f, uint32_t skin, uint32_t material_id) {
    Files F;
    open_files(f, F);
    const uint64_t families = checked_count(F.mdl.i32(224, "skinFamilyCount"), "skin family count"), refs = checked_count(F.mdl.i32(220, "skinRefCount"), "skin reference count");
    const uint64_t n_mats = checked_count(F.mdl.i32(204, "textureCount"), "material count");
    if (skin >= families || material_id >= n_mats) return 0;
    if (material_id >= refs) throw std::runtime_error("mdl: material id past the skin reference table");
    int16_t v;
    std::memcpy(&v, F.mdl.at(F.mdl.i32(228, "skinRefOffset") + (int64_t)(2 * ((uint64_t)skin * refs + material_id)), 2, "skin table"), 2);
    return v;
}

// directory `dir` + name of material `material_id` (Model.cpp:256-274 tries the directories in order until the .vmt exists)
std::string MdlMaterialPath(const vt_mdl_files *f, uint32_t material_id, uint32_t dir) {
    Files F;
    open_files(f, F);
    const uint64_t n_mats = checked_count(F.mdl.i32(204, "textureCount"), "material count"), n_dirs = checked_count(F.mdl.i32(212, "textureDirCount"), "material directory count");
    if (material_id >= n_mats || dir >= n_dirs) throw std::runtime_error("mdl: material or directory index out of range");
    auto cstr = [&](int64_t off) {
        std::string s;
        for (;; off++) {
            const char ch = (char)*F.mdl.at(off, 1, "string");
            if (!ch) break;
            s.push_back(ch);
            if (s.size() > 4096) throw std::runtime_error("mdl: unterminated string");
        }
        return s;
    };
    const int64_t tex = F.mdl.i32(208, "textureOffset") + (int64_t)material_id * (int64_t)kTextureSize;
    const std::string name = cstr(tex + F.mdl.i32(tex, "texture name"));
    const int64_t dir_off = F.mdl.i32(F.mdl.i32(216, "textureDirOffset") + 4 * (int64_t)dir, "texture directory");
    return cstr(dir_off) + name;
}

}  // namespace vt
