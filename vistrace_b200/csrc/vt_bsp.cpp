// vt_bsp.cpp — host-side ingestion of Source-engine maps: .bsp (VBSP 19-21) -> the world triangles, the material list and the
// static-prop placements the reference's World object is built from.
//
// Restates, from the published VBSP lump layouts, what the reference does with such a file:
//   * libs/BSPParser/FileFormat/Parser.cpp:11-73 — which lumps must be present (offset > 0, length a multiple of the record size,
//     count within the engine limit of FileFormat/Limits.h) and BSPParser.cpp:527-578 (the lump list, versions 19-21);
//   * BSPParser.cpp:45-91 + BSPParser.h:118-187 — the game lump and the static-prop lump (versions 4 / 5 / 6 only);
//   * BSPParser.cpp:184-523 (Triangulate) — worldspawn faces only, nodraw / skip / trigger faces dropped, polygons fanned from
//     their first vertex, displacement faces cut into 2 * 4^power triangles, flat normals + tangent frames from the texture axes
//     (:12-43), UVs from the texture vectors divided by the texture size (:144-164);
//   * libs/BSPParser/Displacements/*.cpp — displacement vertices (Displacements.cpp:35-65), per-vertex normals from the adjacent
//     grid triangles (TBNGen.cpp:36-123), tangent frames (:125-154), UVs (UVGen.cpp:6-39) and the three smoothing passes across
//     neighbouring displacements, in the reference's order T-junctions, corners, edges (NormalBlending.cpp:132-312, the sub-edge
//     walk of SubEdgeIterator.cpp:62-242);
//   * source/objects/AccelStruct.cpp:236-414 (World::World) — one material per distinct texture PATH in order of first use, its
//     surface flags those of the texinfo that introduced it, every world triangle one-sided, entity 0.
// All arithmetic is float, in the reference's operation order (compiled -ffp-contract=off like the oracle), so the records equal the
// reference's bit for bit (tests/test_bsp.py compares against the compiled BSPMap).
//
// Unlike the reference, which dereferences most indices unchecked, every offset, count and index is checked against the file here:
// where the reference would read out of bounds (a face range past the face lump, a displacement whose base face is not a
// quadrilateral, a neighbour index past the displacement list, more than four corner neighbours, a model name without a
// terminator ...) the file is rejected.  Where the reference itself declares a file invalid, so does this code.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <unordered_map>
#include <vector>

#include "vt_host.h"

namespace vt {

namespace {

[[noreturn]] void fail(const std::string &what) { throw std::runtime_error("bsp: " + what); }

// ---- BSPStructs::Vector arithmetic (libs/BSPParser/FileFormat/Vector.cpp): component-wise float, nothing fused
struct Vec {
    float x = 0.f, y = 0.f, z = 0.f;
};
inline Vec operator+(Vec a, Vec b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec operator-(Vec a, Vec b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec operator*(Vec a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec operator/(Vec a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline float dot(Vec a, Vec b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec cross(Vec a, Vec b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline Vec normalised(Vec a) {  // Vector::Normalise: three divisions by sqrtf(dot)
    const float len = sqrtf(dot(a, a));
    return {a.x / len, a.y / len, a.z / len};
}
inline Vec lerp(Vec a, Vec b, float t) { return {a.x + (b.x - a.x) * t, a.y + (b.y - a.y) * t, a.z + (b.z - a.z) * t}; }  // VectorLerp

// ---- file access: little-endian, unaligned, bounds-checked once per lump
enum LumpId { L_PLANES = 1, L_TEXDATA = 2, L_VERTICES = 3, L_TEXINFO = 6, L_FACES = 7, L_EDGES = 12, L_SURFEDGES = 13, L_MODELS = 14, L_DISPINFO = 26,
              L_DISP_VERTS = 33, L_GAME = 35, L_STRING_DATA = 43, L_STRING_TABLE = 44 };
constexpr size_t kHeaderSize = 1036, kLumpDirOffset = 8, kLumpEntry = 16;
constexpr int32_t kVbsp = 'V' + ('B' << 8) + ('S' << 16) + ('P' << 24);
constexpr int32_t kStaticProps = ('s' << 24) | ('p' << 16) | ('r' << 8) | 'p';
constexpr uint32_t kSurfDropped = 0x80u | 0x200u | 0x40u;  // SURF::NODRAW | SKIP | TRIGGER (BSPParser.cpp:93-107)
constexpr int kMaxPower = 4;                               // MAX_MAP_DISP_POWER; the reference does not test it, the format does not exceed it

struct Records {
    const uint8_t *p = nullptr;
    size_t n = 0, stride = 0;
    const uint8_t *rec(size_t i) const { return p + i * stride; }
    template <class T> T get(size_t i, size_t off) const {
        T v;
        std::memcpy(&v, p + i * stride + off, sizeof(T));
        return v;
    }
    Vec vec(size_t i, size_t off) const { return {get<float>(i, off), get<float>(i, off + 4), get<float>(i, off + 8)}; }
};

struct Map {
    const uint8_t *file = nullptr;
    uint64_t size = 0;
    int32_t version = 0;
    Records verts, planes, edges, surfedges, faces, texinfos, texdatas, string_table, string_data, models, dispinfos, dispverts;
    // static props
    uint32_t sprp_version = 0;
    Records sprp_dict, sprp_props;

    // lump directory entry -> records; the acceptance rules of ParseLumpBase / GetLumpPtr (FileFormat/Parser.cpp:11-73)
    Records lump(int id, size_t stride, size_t max, const char *name) const {
        int32_t off, len;
        std::memcpy(&off, file + kLumpDirOffset + (size_t)id * kLumpEntry, 4);
        std::memcpy(&len, file + kLumpDirOffset + (size_t)id * kLumpEntry + 4, 4);
        if (len < 0 || (size_t)len % stride != 0) fail(std::string(name) + " lump: length is not a multiple of the record size");
        if (off <= 0) fail(std::string(name) + " lump is absent");
        if ((uint64_t)off + (uint64_t)len > size) fail(std::string(name) + " lump runs past the end of the file");
        Records r{file + off, (size_t)len / stride, stride};
        if (r.n > max) fail(std::string(name) + " lump holds more records than the engine allows");
        return r;
    }
};

// BSPParser.h:118-187 (ParseStaticPropLump): dictionary, leaf list and props, each behind its own int32 count; the three must fill
// the game lump exactly
void parse_static_props(Map &M, int64_t off, int64_t len, uint32_t version) {
    const size_t prop_size = version == 4 ? 56 : version == 5 ? 60 : 64;
    if (off < 0 || len < 0 || (uint64_t)off + (uint64_t)len > M.size) fail("static prop lump runs past the end of the file");
    const uint8_t *p = M.file + off;
    auto count_at = [&](uint64_t at) {
        if (at + 4 > (uint64_t)len) fail("static prop lump is truncated");
        int32_t v;
        std::memcpy(&v, p + at, 4);
        if (v < 0) fail("static prop lump: negative count");
        return (uint64_t)v;
    };
    if ((uint64_t)len < 12) fail("static prop lump is truncated");
    uint64_t at = 0;
    const uint64_t n_dict = count_at(at);
    at += 4;
    if (at + n_dict * 128 + 8 > (uint64_t)len) fail("static prop dictionary runs past its lump");
    M.sprp_dict = Records{p + at, (size_t)n_dict, 128};
    at += n_dict * 128;
    const uint64_t n_leaves = count_at(at);
    at += 4;
    if (at + n_leaves * 2 + 4 > (uint64_t)len) fail("static prop leaf list runs past its lump");
    at += n_leaves * 2;
    const uint64_t n_props = count_at(at);
    at += 4;
    if (at + n_props * prop_size != (uint64_t)len) fail("static prop lump size does not match its counts");
    M.sprp_props = Records{p + at, (size_t)n_props, prop_size};
    M.sprp_version = version;
}

void open_map(const uint8_t *file, uint64_t size, Map &M) {
    if (!file || !size) fail("null or empty file");
    if (size < kHeaderSize) fail("file is shorter than the header");
    M.file = file, M.size = size;
    int32_t ident;
    std::memcpy(&ident, file, 4);
    std::memcpy(&M.version, file + 4, 4);
    if (ident != kVbsp) fail("not a VBSP file");
    if (M.version < 19 || M.version > 21) fail("only versions 19 to 21 are supported");  // BSPParser.cpp:536
    // the lumps BSPMap::BSPMap requires, with the limits of FileFormat/Limits.h
    M.verts = M.lump(L_VERTICES, 12, 65536, "vertex");
    M.planes = M.lump(L_PLANES, 20, 65536, "plane");
    M.edges = M.lump(L_EDGES, 4, 256000, "edge");
    M.surfedges = M.lump(L_SURFEDGES, 4, 512000, "surfedge");
    M.faces = M.lump(L_FACES, 56, 65536, "face");
    M.texinfos = M.lump(L_TEXINFO, 72, 12288, "texinfo");
    M.texdatas = M.lump(L_TEXDATA, 32, 2048, "texdata");
    M.string_table = M.lump(L_STRING_TABLE, 4, 65536, "texdata string table");
    M.string_data = M.lump(L_STRING_DATA, 1, 256000, "texdata string data");
    M.models = M.lump(L_MODELS, 48, 1024, "model");
    M.dispinfos = M.lump(L_DISPINFO, 176, 2048, "dispinfo");
    M.dispverts = M.lump(L_DISP_VERTS, 20, 2048 * 17 * 17, "dispvert");
    // game lump directory (BSPParser.cpp:45-91): int32 count + 16-byte entries {id, flags u16, version u16, offset, length}
    const Records game = M.lump(L_GAME, 1, ~(size_t)0, "game");
    if (game.n < 4) fail("game lump is truncated");
    const int32_t n_game = game.get<int32_t>(0, 0);
    if (n_game < 0 || (uint64_t)n_game * 16 + 4 > game.n) fail("game lump directory runs past its lump");
    for (int32_t i = 0; i < n_game; i++) {
        const size_t e = 4 + (size_t)i * 16;
        if (game.get<int32_t>(0, e) != kStaticProps) continue;
        const uint32_t version = game.get<uint16_t>(0, e + 6);
        if (version < 4 || version > 6) fail("static prop lump version " + std::to_string(version) + " is not supported (4, 5 or 6)");
        parse_static_props(M, game.get<int32_t>(0, e + 8), game.get<int32_t>(0, e + 12), version);
    }
    if (M.models.n == 0) fail("no worldspawn model");
}

// ---- faces
struct Face {
    uint16_t plane;
    int32_t first_edge;
    int16_t n_edges, texinfo, dispinfo;
};
Face face_at(const Map &M, size_t i) {
    return {M.faces.get<uint16_t>(i, 0), M.faces.get<int32_t>(i, 4), M.faces.get<int16_t>(i, 8), M.faces.get<int16_t>(i, 10), M.faces.get<int16_t>(i, 12)};
}
// worldspawn's face range (BSPParser.cpp:189-190), checked
void world_faces(const Map &M, size_t &first, size_t &count) {
    const int32_t f = M.models.get<int32_t>(0, 40), n = M.models.get<int32_t>(0, 44);
    if (f < 0 || n < 0 || (uint64_t)f + (uint64_t)n > M.faces.n) fail("worldspawn's faces run past the face lump");
    first = (size_t)f, count = (size_t)n;
}
// a face Triangulate emits (BSPParser.cpp:195-198, 329-333): drawn, at least a triangle, a texinfo inside the lump
bool face_is_emitted(const Map &M, const Face &f) {
    if (f.texinfo < 0 || (size_t)f.texinfo >= M.texinfos.n) return false;
    if (M.texinfos.get<uint32_t>((size_t)f.texinfo, 64) & kSurfDropped) return false;
    return f.n_edges >= 3;
}
int checked_power(const Map &M, size_t disp) {
    const int32_t power = M.dispinfos.get<int32_t>(disp, 20);
    if (power < 0 || power > kMaxPower) fail("displacement power out of range");
    return power;
}
uint64_t count_triangles(const Map &M) {
    size_t first, count;
    world_faces(M, first, count);
    uint64_t n = 0;
    for (size_t i = first; i < first + count; i++) {
        const Face f = face_at(M, i);
        if (!face_is_emitted(M, f)) continue;
        if (f.dispinfo < 0) {
            n += (uint64_t)(f.n_edges - 2);
        } else {
            if ((size_t)f.dispinfo >= M.dispinfos.n) fail("face names a displacement past the dispinfo lump");  // BSPParser.cpp:204
            const uint64_t cells = 1ull << checked_power(M, (size_t)f.dispinfo);
            n += cells * cells * 2;
        }
    }
    if (n == 0) fail("the map has no drawable world triangles");  // BSPParser.cpp:210
    // a hostile face table can name the same surfedges over and over (65536 faces x 32767 edges): the reference would try to malloc
    // ~300 GB for it; no compiled map comes near this bound
    if (n > (1ull << 26)) fail("more world triangles than a compiled map can hold");
    return n;
}

// BSPMap::GetSurfEdgeVerts (BSPParser.cpp:166-182): the edge a surfedge names, reversed when the surfedge is negative
bool surfedge_verts(const Map &M, int64_t index, Vec *a, Vec *b) {
    if (index < 0 || (uint64_t)index >= M.surfedges.n) return false;
    const int32_t e = M.surfedges.get<int32_t>((size_t)index, 0);
    const int64_t ae = e < 0 ? -(int64_t)e : (int64_t)e;
    if ((uint64_t)ae >= M.edges.n) return false;
    uint16_t ia = M.edges.get<uint16_t>((size_t)ae, 0), ib = M.edges.get<uint16_t>((size_t)ae, 2);
    if (ia >= M.verts.n || ib >= M.verts.n) return false;
    if (e < 0) std::swap(ia, ib);
    *a = M.verts.vec(ia, 0);
    if (b) *b = M.verts.vec(ib, 0);
    return true;
}

// BSPMap::CalcUVs (BSPParser.cpp:144-164)
bool texture_uv(const Map &M, int32_t texinfo, Vec pos, float *uv) {
    if (texinfo < 0 || (size_t)texinfo >= M.texinfos.n) return false;
    const int32_t td = M.texinfos.get<int32_t>((size_t)texinfo, 68);
    if (td < 0 || (size_t)td >= M.texdatas.n) return false;
    float s[4], t[4];
    std::memcpy(s, M.texinfos.rec((size_t)texinfo), 16);
    std::memcpy(t, M.texinfos.rec((size_t)texinfo) + 16, 16);
    uv[0] = s[0] * pos.x + s[1] * pos.y + s[2] * pos.z + s[3];
    uv[1] = t[0] * pos.x + t[1] * pos.y + t[2] * pos.z + t[3];
    uv[0] /= (float)M.texdatas.get<int32_t>((size_t)td, 16);
    uv[1] /= (float)M.texdatas.get<int32_t>((size_t)td, 20);
    return true;
}

// tangent frame of a surface with normal n from the texture axes (CalcTangentBinormal, BSPParser.cpp:25-43; the same statements per
// displacement vertex in TBNGen.cpp:137-152)
void tangent_frame(const Map &M, int32_t texinfo, size_t plane, Vec n, Vec &t, Vec &b) {
    float tv[8];
    std::memcpy(tv, M.texinfos.rec((size_t)texinfo), 32);
    const Vec s_axis{tv[0], tv[1], tv[2]}, t_axis{tv[4], tv[5], tv[6]};
    b = normalised(t_axis);
    t = normalised(cross(n, b));
    b = normalised(cross(t, n));
    if (dot(M.planes.vec(plane, 0), cross(s_axis, t_axis)) > 0.0f) t = t * -1.f;
}

// ---- displacements
struct Grid {
    int power = 0, side = 0;  // side = 2^power + 1 vertices per edge
    size_t info = 0;          // index in the dispinfo lump
    std::vector<Vec> pos, nrm, tan, bin;
    std::vector<float> uv, alpha;
};
struct SubNeighbour {
    uint16_t index;
    uint8_t orientation, span, neighbour_span;
    bool valid() const { return index != 0xFFFF; }
};
SubNeighbour sub_neighbour(const Map &M, size_t disp, int edge, int sub) {
    const size_t off = 48 + (size_t)edge * 12 + (size_t)sub * 6;
    return {M.dispinfos.get<uint16_t>(disp, off), M.dispinfos.get<uint8_t>(disp, off + 2), M.dispinfos.get<uint8_t>(disp, off + 3),
            M.dispinfos.get<uint8_t>(disp, off + 4)};
}
Grid &neighbour_grid(std::vector<Grid> &grids, uint32_t index) {
    if (index >= grids.size()) fail("displacement neighbour index past the dispinfo lump");
    return grids[index];
}

// corner c of a grid (CornerToVertIdx, NormalBlending.cpp:50-69): 0 lower left, 1 upper left, 2 upper right, 3 lower right
inline void corner_xy(int side, int c, int &x, int &y) {
    x = (c == 2 || c == 3) ? side - 1 : 0;
    y = (c == 1 || c == 2) ? side - 1 : 0;
}
inline int corner_vertex(const Grid &g, int c) {
    int x, y;
    corner_xy(g.side, c, x, y);
    return y * g.side + x;
}
// middle vertex of edge e (GetEdgeMidPoint, :71-99): 0 left, 1 top, 2 right, 3 bottom
inline int edge_mid_vertex(const Grid &g, int e) {
    const int end = g.side - 1, mid = g.side / 2;
    const int x = e == 1 || e == 3 ? mid : (e == 2 ? end : 0);
    const int y = e == 0 || e == 2 ? mid : (e == 1 ? end : 0);
    return y * g.side + x;
}
// the corner of g within 0.1 units of p, or -1 (FindNeighborCornerVert, :101-122)
int matching_corner(const Grid &g, Vec p) {
    int best = 0;
    float best_d = 1e24;
    for (int c = 0; c < 4; c++) {
        const Vec delta = g.pos[(size_t)corner_vertex(g, c)] - p;
        const float d = sqrtf(dot(delta, delta));
        if (d < best_d) best = c, best_d = d;
    }
    return best_d <= 0.1f ? best : -1;
}

// Displacements.cpp:35-65 (positions, alphas), TBNGen.cpp:36-123 (normals), :125-154 (frames), UVGen.cpp:6-39 (uvs) for dispinfo d
void build_grid(const Map &M, size_t d, Grid &g) {
    g.info = d;
    g.power = checked_power(M, d);
    g.side = (1 << g.power) + 1;
    const int side = g.side;
    const size_t n = (size_t)side * side;
    const uint16_t map_face = M.dispinfos.get<uint16_t>(d, 36);
    if (map_face >= M.faces.n) fail("displacement names a face past the face lump");
    const Face f = face_at(M, map_face);
    if (f.n_edges != 4) fail("displacement base face is not a quadrilateral");
    if (f.texinfo < 0 || (size_t)f.texinfo >= M.texinfos.n) fail("displacement base face has no texinfo");
    if (f.plane >= M.planes.n) fail("displacement base face names a plane past the plane lump");
    const int32_t vert_start = M.dispinfos.get<int32_t>(d, 12);
    if (vert_start < 0 || (uint64_t)vert_start + n > M.dispverts.n) fail("displacement vertices run past the dispvert lump");

    // the base quad, rotated so that the corner nearest to startPosition comes first (BSPParser.cpp:248-282)
    Vec quad[4], corners[4];
    int first = 0;
    float first_d2 = std::numeric_limits<float>::max();
    const Vec start = M.dispinfos.vec(d, 0);
    for (int k = 0; k < 4; k++) {
        if (!surfedge_verts(M, (int64_t)f.first_edge + k, &quad[k], nullptr)) fail("displacement base face names an edge or vertex out of range");
        const Vec dv = start - quad[k];
        const float d2 = dot(dv, dv);
        if (d2 < first_d2) first = k, first_d2 = d2;
    }
    for (int k = 0; k < 4; k++) corners[k] = quad[(k + first) % 4];

    g.pos.resize(n), g.nrm.resize(n), g.tan.resize(n), g.bin.resize(n), g.uv.resize(n * 2), g.alpha.resize(n);
    const float oo = 1.0f / (float)(side - 1);
    {  // bilinear grid over the quad + the stored offset along each vertex's direction; alpha / 255 clamped to [0, 1]
        const Vec step0 = (corners[1] - corners[0]) * oo, step1 = (corners[2] - corners[3]) * oo;
        for (int i = 0; i < side; i++) {
            const Vec end0 = step0 * (float)i + corners[0], end1 = step1 * (float)i + corners[3];
            const Vec seg_step = (end1 - end0) * oo;
            for (int j = 0; j < side; j++) {
                const size_t v = (size_t)i * side + j, dv = (size_t)vert_start + v;
                g.pos[v] = end0 + seg_step * (float)j + M.dispverts.vec(dv, 0) * M.dispverts.get<float>(dv, 12);
                const float a = M.dispverts.get<float>(dv, 16) / 255.f;
                g.alpha[v] = a < 0.f ? 0.f : (1.f < a ? 1.f : a);  // std::clamp: a NaN stays a NaN
            }
        }
    }
    {  // vertex normal = mean of the unit normals of the grid triangles in the (up to four) cells around it, two triangles per cell,
       // cells taken in the order upper right, upper left, lower left, lower right (TBNGen.cpp:36-104)
        auto P = [&](int col, int row) { return g.pos[(size_t)col * side + row]; };  // verts[indexCol * postSpacing + indexRow]
        auto tri_normal = [&](Vec a0, Vec a1, Vec origin) { return normalised(cross(a1 - origin, a0 - origin)); };
        for (int col = 0; col < side; col++) {
            for (int row = 0; row < side; row++) {
                const bool left = row - 1 >= 0, top = col + 1 <= side - 1, right = row + 1 <= side - 1, bottom = col - 1 >= 0;
                Vec sum{};
                int cnt = 0;
                auto cell = [&](int c, int r) {  // the cell whose lower-left vertex is (c, r)
                    sum = sum + tri_normal(P(c + 1, r), P(c, r + 1), P(c, r));
                    sum = sum + tri_normal(P(c + 1, r), P(c + 1, r + 1), P(c, r + 1));
                    cnt += 2;
                };
                if (top && right) cell(col, row);
                if (left && top) cell(col, row - 1);
                if (left && bottom) cell(col - 1, row - 1);
                if (right && bottom) cell(col - 1, row);
                g.nrm[(size_t)col * side + row] = sum / (float)cnt;
            }
        }
    }
    for (size_t v = 0; v < n; v++) tangent_frame(M, f.texinfo, f.plane, g.nrm[v], g.tan[v], g.bin[v]);
    {  // UVs: the same bilinear grid over the texture coordinates of the four corners
        float cuv[4][2];
        for (int k = 0; k < 4; k++)
            if (!texture_uv(M, f.texinfo, corners[k], cuv[k])) fail("displacement base face names texture data out of range");
        const float s0[2] = {(cuv[1][0] - cuv[0][0]) * oo, (cuv[1][1] - cuv[0][1]) * oo}, s1[2] = {(cuv[2][0] - cuv[3][0]) * oo, (cuv[2][1] - cuv[3][1]) * oo};
        for (int i = 0; i < side; i++) {
            const float e0[2] = {s0[0] * (float)i + cuv[0][0], s0[1] * (float)i + cuv[0][1]}, e1[2] = {s1[0] * (float)i + cuv[3][0], s1[1] * (float)i + cuv[3][1]};
            const float st[2] = {(e1[0] - e0[0]) * oo, (e1[1] - e0[1]) * oo};
            for (int j = 0; j < side; j++) {
                const size_t v = ((size_t)i * side + j) * 2;
                g.uv[v] = e0[0] + st[0] * (float)j;
                g.uv[v + 1] = e0[1] + st[1] * (float)j;
            }
        }
    }
}

// every displacement that touches d: corner neighbours first, then the edge sub-neighbours (GetAllNeighbours, NormalBlending.cpp:25-48)
std::vector<uint32_t> touching(const Map &M, size_t d) {
    std::vector<uint32_t> out;
    for (int c = 0; c < 4; c++) {
        const size_t off = 96 + (size_t)c * 10;
        const uint8_t n = M.dispinfos.get<uint8_t>(d, off + 8);
        if (n > 4) fail("displacement corner lists more than four neighbours");
        for (int i = 0; i < n; i++) out.push_back(M.dispinfos.get<uint16_t>(d, off + 2 * (size_t)i));
    }
    for (int e = 0; e < 4; e++)
        for (int s = 0; s < 2; s++) {
            const SubNeighbour sn = sub_neighbour(M, d, e, s);
            if (sn.valid()) out.push_back(sn.index);
        }
    return out;
}

// pass 1 (BlendTJuncs, :202-247): where two neighbours meet at the middle of an edge, the three vertices share the mean frame
void blend_t_junctions(const Map &M, std::vector<Grid> &grids) {
    for (Grid &g : grids)
        for (int e = 0; e < 4; e++) {
            const SubNeighbour a = sub_neighbour(M, g.info, e, 0), b = sub_neighbour(M, g.info, e, 1);
            if (!a.valid() || !b.valid()) continue;
            const int mid = edge_mid_vertex(g, e);
            Grid &ga = neighbour_grid(grids, a.index), &gb = neighbour_grid(grids, b.index);
            const int ca = matching_corner(ga, g.pos[(size_t)mid]), cb = matching_corner(gb, g.pos[(size_t)mid]);
            if (ca == -1 || cb == -1) continue;
            const size_t va = (size_t)corner_vertex(ga, ca), vb = (size_t)corner_vertex(gb, cb);
            const Vec t = (g.tan[(size_t)mid] + (ga.tan[va] + gb.tan[vb])) / 3.f;
            const Vec bn = (g.bin[(size_t)mid] + (ga.bin[va] + gb.bin[vb])) / 3.f;
            const Vec n = (g.nrm[(size_t)mid] + (ga.nrm[va] + gb.nrm[vb])) / 3.f;
            g.tan[(size_t)mid] = ga.tan[va] = gb.tan[vb] = t;
            g.bin[(size_t)mid] = ga.bin[va] = gb.bin[vb] = bn;
            g.nrm[(size_t)mid] = ga.nrm[va] = gb.nrm[vb] = n;
        }
}

// pass 2 (BlendCorners, :124-200): every corner takes the mean frame of all touching displacements that have a corner there
void blend_corners(const Map &M, std::vector<Grid> &grids) {
    for (Grid &g : grids) {
        const std::vector<uint32_t> nb = touching(M, g.info);
        std::vector<int> nb_vertex(nb.size());
        for (int c = 0; c < 4; c++) {
            const size_t v = (size_t)corner_vertex(g, c);
            int divisor = 1;
            Vec t = g.tan[v], b = g.bin[v], n = g.nrm[v];
            for (size_t k = 0; k < nb.size(); k++) {
                const Grid &h = neighbour_grid(grids, nb[k]);
                const int hc = matching_corner(h, g.pos[v]);
                nb_vertex[k] = hc == -1 ? -1 : corner_vertex(h, hc);
                if (hc == -1) continue;
                t = t + h.tan[(size_t)nb_vertex[k]], b = b + h.bin[(size_t)nb_vertex[k]], n = n + h.nrm[(size_t)nb_vertex[k]];
                divisor++;
            }
            t = t / (float)divisor, b = b / (float)divisor, n = n / (float)divisor;
            g.tan[v] = t, g.bin[v] = b, g.nrm[v] = n;
            for (size_t k = 0; k < nb.size(); k++) {
                if (nb_vertex[k] == -1) continue;
                Grid &h = grids[nb[k]];
                h.tan[(size_t)nb_vertex[k]] = t, h.bin[(size_t)nb_vertex[k]] = b, h.nrm[(size_t)nb_vertex[k]] = n;
            }
        }
    }
}

// The part of edge `edge` a sub-neighbour with span code `span` covers, as grid coordinates {x, y} of its two ends
// (SetupSpan, SubEdgeIterator.cpp:62-83): 0 corner to corner, 1 corner to midpoint, 2 midpoint to corner.
void span_ends(int power, int edge, uint8_t span, int (&a)[2], int (&b)[2]) {
    const int side = (1 << power) + 1, free_dim = (edge & 1) ? 0 : 1;  // edges 0 / 2 run along y, 1 / 3 along x
    corner_xy(side, edge, a[0], a[1]);
    corner_xy(side, (edge + 1) & 3, b[0], b[1]);
    const bool from_far_corner = edge == 2 || edge == 3;
    if (span == 1) (from_far_corner ? a : b)[free_dim] = side / 2;
    else if (span == 2) (from_far_corner ? b : a)[free_dim] = side / 2;
}

// pass 3 (BlendEdges, :249-305 + the walk of SubEdgeIterator.cpp:85-242): along every edge shared with a sub-neighbour the paired
// vertices take the mean frame; vertices of the finer side that have no partner are interpolated between their neighbours
void blend_edges(const Map &M, std::vector<Grid> &grids) {
    static const int kEdgeAtFarSide[4] = {0, 1, 1, 0};
    for (Grid &g : grids) {
        const int side = g.side;
        for (int e = 0; e < 4; e++)
            for (int s = 0; s < 2; s++) {
                const SubNeighbour sn = sub_neighbour(M, g.info, e, s);
                if (!sn.valid()) continue;
                Grid &h = neighbour_grid(grids, sn.index);
                const int edge_dim = e & 1, free_dim = !edge_dim;
                // first vertex of the walk on this side and its image on the neighbour (TransformIntoSubNeighbor, :85-123)
                int my[2], nbv[2];
                my[edge_dim] = kEdgeAtFarSide[e] * (side - 1);
                my[free_dim] = (side / 2) * s;
                {
                    int src_a[2], src_b[2], dst_a[2], dst_b[2];
                    span_ends(g.power, e, sn.span, src_a, src_b);
                    const int nb_edge = (e + 2 + sn.orientation) & 3;
                    span_ends(h.power, nb_edge, sn.span, dst_b, dst_a);  // the neighbour's edge runs the other way
                    const int run = src_b[free_dim] - src_a[free_dim];
                    if (run == 0) fail("displacement sub-edge of zero length");
                    const int fixed = ((my[free_dim] - src_a[free_dim]) * (1 << 16)) / run;
                    if (fixed < 0 || fixed > (1 << 16)) fail("displacement neighbour span does not cover the shared edge");
                    const int nb_dim = nb_edge & 1;
                    nbv[nb_dim] = dst_a[nb_dim];
                    nbv[!nb_dim] = dst_a[!nb_dim] + ((dst_b[!nb_dim] - dst_a[!nb_dim]) * fixed) / (1 << 16);
                    if (nbv[0] < 0 || nbv[0] >= h.side || nbv[1] < 0 || nbv[1] >= h.side) fail("displacement neighbour vertex out of range");
                }
                // step on this side and, rotated by the neighbour's orientation, on the other (SetupEdgeIncrements, :141-199)
                int inc[2], tmp[2], nb_inc[2];
                inc[edge_dim] = tmp[edge_dim] = 0;
                if (h.power > g.power) inc[free_dim] = 1, tmp[free_dim] = 1 << (h.power - g.power);
                else inc[free_dim] = 1 << (g.power - h.power), tmp[free_dim] = 1;
                switch (sn.orientation) {
                    case 0: nb_inc[0] = tmp[0], nb_inc[1] = tmp[1]; break;
                    case 1: nb_inc[0] = tmp[1], nb_inc[1] = -tmp[0]; break;
                    case 2: nb_inc[0] = -tmp[0], nb_inc[1] = -tmp[1]; break;
                    default: nb_inc[0] = -tmp[1], nb_inc[1] = tmp[0]; break;
                }
                // bTouchCorners: the walk ends one step later, i.e. includes the vertex at the far end of the span
                const int end = (sn.span == 1 ? side >> 1 : side - 1) + inc[free_dim];
                int prev[2] = {my[0], my[1]};
                for (;;) {
                    my[0] += inc[0], my[1] += inc[1], nbv[0] += nb_inc[0], nbv[1] += nb_inc[1];
                    if (!(my[free_dim] < end)) break;
                    if (my[0] < 0 || my[0] >= side || my[1] < 0 || my[1] >= side) fail("displacement edge walk leaves the grid");
                    const size_t v = (size_t)my[1] * side + my[0];
                    if (!(my[free_dim] + inc[free_dim] >= end)) {  // not the last vertex: corners were settled by pass 2
                        if (nbv[0] < 0 || nbv[0] >= h.side || nbv[1] < 0 || nbv[1] >= h.side) fail("displacement edge walk leaves the neighbour's grid");
                        const size_t w = (size_t)nbv[1] * h.side + nbv[0];
                        const Vec t = (g.tan[v] + h.tan[w]) / 2.f, b = (g.bin[v] + h.bin[w]) / 2.f, n = (g.nrm[v] + h.nrm[w]) / 2.f;
                        g.tan[v] = t, g.bin[v] = b, g.nrm[v] = n;
                        h.tan[w] = t, h.bin[w] = b, h.nrm[w] = n;
                    }
                    const int from = prev[free_dim], to = my[free_dim];
                    for (int tween = from + 1; tween < to; tween++) {
                        // RemapVal(tween, from, to, 0, 1)
                        const float A = (float)from, B = (float)to, val = (float)tween;
                        const float pct = A == B ? (val >= B ? 1.f : 0.f) : 0.f + (1.f - 0.f) * (val - A) / (B - A);
                        const size_t p = (size_t)prev[1] * side + prev[0];
                        const Vec t = normalised(lerp(g.tan[p], g.tan[v], pct)), b = normalised(lerp(g.bin[p], g.bin[v], pct)), n = normalised(lerp(g.nrm[p], g.nrm[v], pct));
                        int at[2];
                        at[edge_dim] = my[edge_dim], at[free_dim] = tween;
                        const size_t q = (size_t)at[1] * side + at[0];
                        g.tan[q] = t, g.bin[q] = b, g.nrm[q] = n;
                    }
                    prev[0] = my[0], prev[1] = my[1];
                }
            }
    }
}

std::vector<Grid> build_displacements(const Map &M) {
    std::vector<Grid> grids(M.dispinfos.n);
    for (size_t d = 0; d < grids.size(); d++) build_grid(M, d, grids[d]);
    blend_t_junctions(M, grids);
    blend_corners(M, grids);
    blend_edges(M, grids);
    return grids;
}

// BSPMap::GetTexture (BSPParser.cpp:583-607), checked: texinfo -> {flags, texdata -> reflectivity, size, path}
struct Texture {
    uint32_t flags;
    Vec reflectivity;
    int32_t width, height;
    std::string path;
};
Texture texture_of(const Map &M, int32_t texinfo) {
    if (texinfo < 0 || (size_t)texinfo >= M.texinfos.n) fail("texture index out of bounds");
    const int32_t td = M.texinfos.get<int32_t>((size_t)texinfo, 68);
    if (td < 0 || (size_t)td >= M.texdatas.n) fail("texdata index out of bounds");
    const int32_t name = M.texdatas.get<int32_t>((size_t)td, 12);
    if (name < 0 || (size_t)name >= M.string_table.n) fail("texdata string table index out of bounds");
    const int32_t off = M.string_table.get<int32_t>((size_t)name, 0);
    if (off < 0 || (size_t)off >= M.string_data.n) fail("texture name starts past the string data");
    const char *s = reinterpret_cast<const char *>(M.string_data.p) + off;
    const size_t len = strnlen(s, M.string_data.n - (size_t)off);
    if (len == M.string_data.n - (size_t)off) fail("texture name is not terminated");
    return {M.texinfos.get<uint32_t>((size_t)texinfo, 64), M.texdatas.vec((size_t)td, 0), M.texdatas.get<int32_t>((size_t)td, 16), M.texdatas.get<int32_t>((size_t)td, 20),
            std::string(s, len)};
}

// materials in order of first use over the emitted triangles (World::World, AccelStruct.cpp:250-404: materialIds / world.materials)
struct MaterialList {
    std::vector<int32_t> first_texinfo;                 // the texinfo that introduced material m
    std::unordered_map<std::string, uint32_t> by_path;
    std::vector<int64_t> of_texinfo;                    // texinfo -> material (-1: not seen yet)
    uint32_t index_of(const Map &M, int32_t texinfo) {
        if (of_texinfo.empty()) of_texinfo.assign(M.texinfos.n, -1);
        int64_t &slot = of_texinfo[(size_t)texinfo];
        if (slot < 0) {
            const std::string path = texture_of(M, texinfo).path;
            auto it = by_path.find(path);
            if (it == by_path.end()) {
                it = by_path.emplace(path, (uint32_t)first_texinfo.size()).first;
                first_texinfo.push_back(texinfo);
            }
            slot = it->second;
        }
        return (uint32_t)slot;
    }
};
MaterialList list_materials(const Map &M) {
    MaterialList L;
    size_t first, count;
    world_faces(M, first, count);
    for (size_t i = first; i < first + count; i++) {
        const Face f = face_at(M, i);
        if (face_is_emitted(M, f)) L.index_of(M, f.texinfo);
    }
    return L;
}

void put(float *dst, Vec v) { dst[0] = v.x, dst[1] = v.y, dst[2] = v.z; }

// BSPMap::Triangulate's second pass (BSPParser.cpp:325-520): every emitted triangle of the world, in file order, handed to
// emit(p, n, t, b, uv, alpha, texinfo)
template <class Emit> void for_each_triangle(const Map &M, const std::vector<Grid> &grids, Emit &&emit) {
    size_t first, count;
    world_faces(M, first, count);
    for (size_t i = first; i < first + count; i++) {
        const Face f = face_at(M, i);
        if (!face_is_emitted(M, f)) continue;
        if (f.dispinfo < 0) {
            // polygon: a fan around its first vertex (BSPParser.cpp:338-420); a root the file does not resolve stays at the origin
            // as in the reference, which ignores that return value
            if (f.plane >= M.planes.n) fail("face names a plane past the plane lump");
            Vec root{};
            float root_uv[2];
            surfedge_verts(M, f.first_edge, &root, nullptr);
            if (!texture_uv(M, f.texinfo, root, root_uv)) fail("face names texture data out of range");
            for (int64_t e = (int64_t)f.first_edge + 1; e < (int64_t)f.first_edge + f.n_edges - 1; e++) {
                Vec p[3] = {root, {}, {}};
                if (!surfedge_verts(M, e, &p[1], &p[2])) fail("face names an edge or vertex out of range");
                float uv[3][2] = {{root_uv[0], root_uv[1]}, {}, {}};
                texture_uv(M, f.texinfo, p[1], uv[1]);
                texture_uv(M, f.texinfo, p[2], uv[2]);
                const Vec n = normalised(cross(p[2] - p[0], p[1] - p[0]));  // CalcNormal, BSPParser.cpp:12-23
                Vec t, b;
                tangent_frame(M, f.texinfo, f.plane, n, t, b);
                const Vec nn[3] = {n, n, n}, tt[3] = {t, t, t}, bb[3] = {b, b, b};
                const float a[3] = {1.f, 1.f, 1.f};
                emit(p, nn, tt, bb, uv, a, f.texinfo);
            }
        } else {
            // displacement: two triangles per cell, cells column by column (BSPParser.cpp:421-493)
            const Grid &g = grids[(size_t)f.dispinfo];
            const int cells = g.side - 1, side = g.side;
            for (int x = 0; x < cells; x++)
                for (int y = 0; y < cells; y++) {
                    const int a = y * side + x, b = (y + 1) * side + x, c = (y + 1) * side + (x + 1), d = y * side + (x + 1);
                    for (int tri = 0; tri < 2; tri++) {
                        const int idx[3] = {a, tri == 0 ? b : c, tri == 0 ? c : d};
                        Vec p[3], n[3], t[3], bn[3];
                        float uv[3][2], al[3];
                        for (int k = 0; k < 3; k++) {
                            const size_t v = (size_t)idx[k];
                            p[k] = g.pos[v], n[k] = g.nrm[v], t[k] = g.tan[v], bn[k] = g.bin[v];
                            uv[k][0] = g.uv[2 * v], uv[k][1] = g.uv[2 * v + 1];
                            al[k] = g.alpha[v];
                        }
                        emit(p, n, t, bn, uv, al, f.texinfo);
                    }
                }
        }
    }
}

}  // namespace

void BspInfo(const uint8_t *file, uint64_t size, vt_bsp_info *info) {
    Map M;
    open_map(file, size, M);
    std::memset(info, 0, sizeof(*info));
    info->version = (uint32_t)M.version;
    info->n_tris = count_triangles(M);
    info->n_materials = (uint32_t)list_materials(M).first_texinfo.size();
    info->n_texinfos = (uint32_t)M.texinfos.n;
    info->n_displacements = (uint32_t)M.dispinfos.n;
    info->n_static_props = (uint32_t)M.sprp_props.n;
    info->static_props_version = M.sprp_version;
    // BSPMap::IsValid: the whole triangulation has to go through (displacement smoothing, every edge, vertex and texture reference)
    const std::vector<Grid> grids = build_displacements(M);
    uint64_t n = 0;
    for_each_triangle(M, grids, [&](const Vec (&)[3], const Vec (&)[3], const Vec (&)[3], const Vec (&)[3], const float (&)[3][2], const float (&)[3], int16_t) { n++; });
    if (n != info->n_tris) fail("internal: triangle count mismatch");
}

// BSPMap::Triangulate with clockwise winding (the constructor's default, which World::World uses) + World::World's Triangle records
uint64_t BspTriangles(const uint8_t *file, uint64_t size, vt_tri_in *tris, float *binormals, int16_t *texinfos, uint64_t capacity) {
    Map M;
    open_map(file, size, M);
    const uint64_t n_tris = count_triangles(M);
    if (!tris) return n_tris;
    if (capacity < n_tris) fail("triangle buffer too small");
    const std::vector<Grid> grids = build_displacements(M);
    MaterialList mats;
    uint64_t out = 0;
    auto emit = [&](const Vec (&p)[3], const Vec (&n)[3], const Vec (&t)[3], const Vec (&b)[3], const float (&uv)[3][2], const float (&a)[3], int16_t texinfo) {
        vt_tri_in &r = tris[out];
        std::memset(&r, 0, sizeof(r));
        for (int k = 0; k < 3; k++) {
            put(r.p[k], p[k]), put(r.normals[k], n[k]), put(r.tangents[k], t[k]);
            r.uvs[k][0] = uv[k][0], r.uvs[k][1] = uv[k][1];
            r.alphas[k] = a[k];
            if (binormals) put(binormals + out * 9 + 3 * (size_t)k, b[k]);
        }
        r.material = mats.index_of(M, texinfo);
        r.ent_idx = 0;    // the world entity (AccelStruct.cpp:228-232, 414)
        r.one_sided = 1;  // AccelStruct.cpp:400-404: the world is back-face culled
        if (texinfos) texinfos[out] = texinfo;
        out++;
    };
    for_each_triangle(M, grids, emit);
    return out;
}

void BspMaterial(const uint8_t *file, uint64_t size, uint32_t material, vt_bsp_material *out) {
    Map M;
    open_map(file, size, M);
    const MaterialList L = list_materials(M);
    if (material >= L.first_texinfo.size()) fail("material index out of range");
    const Texture t = texture_of(M, L.first_texinfo[material]);
    std::memset(out, 0, sizeof(*out));
    if (t.path.size() + 1 > sizeof(out->path)) fail("texture path longer than the record holds");
    out->surf_flags = t.flags;
    out->texinfo = L.first_texinfo[material];
    out->width = t.width, out->height = t.height;
    put(out->reflectivity, t.reflectivity);
    std::memcpy(out->path, t.path.c_str(), t.path.size() + 1);
}

// BSPMap::GetStaticProp (BSPParser.h:189-207, BSPParser.cpp:620-632)
void BspStaticProp(const uint8_t *file, uint64_t size, uint32_t index, vt_bsp_static_prop *out) {
    Map M;
    open_map(file, size, M);
    if (index >= M.sprp_props.n) fail("static prop index out of bounds");
    const uint16_t type = M.sprp_props.get<uint16_t>(index, 24);
    if (type >= M.sprp_dict.n) fail("static prop dictionary index out of bounds");
    std::memset(out, 0, sizeof(*out));
    put(out->pos, M.sprp_props.vec(index, 0));
    put(out->ang, M.sprp_props.vec(index, 12));
    out->skin = M.sprp_props.get<int32_t>(index, 32);
    const char *name = reinterpret_cast<const char *>(M.sprp_dict.rec(type));
    if (strnlen(name, 128) == 128) fail("static prop model name is not terminated");
    std::memcpy(out->model, name, strnlen(name, 128) + 1);
}

}  // namespace vt
