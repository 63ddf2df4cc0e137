// ASan driver for the host-only ingestion code: each corpus file is read into an EXACT-size heap buffer so that any read past the
// file ends in a red zone.  usage: driver {bsp|vtf|mdl} files...   (mdl: triples mdl vvd vtx)
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>
#include "vt_host.h"
static std::unique_ptr<uint8_t[]> slurp(const char *p, uint64_t &n) {
    std::ifstream f(p, std::ios::binary | std::ios::ate);
    n = f ? (uint64_t)f.tellg() : 0;
    std::unique_ptr<uint8_t[]> b(new uint8_t[n ? n : 1]);
    if (n) { f.seekg(0); f.read((char *)b.get(), n); }
    return b;
}
int main(int argc, char **argv) {
    std::string kind = argv[1];
    uint64_t ok = 0, bad = 0;
    for (int i = 2; i < argc; i += (kind == "mdl" ? 3 : 1)) {
        try {
            if (kind == "vtf") {
                uint64_t n; auto b = slurp(argv[i], n);
                vt_vtf_info info; vt::VtfInfo(n ? b.get() : nullptr, n, &info);
                if (info.supported) {
                    std::unique_ptr<uint8_t[]> out(new uint8_t[info.rgba_bytes ? info.rgba_bytes : 1]);
                    for (uint32_t fr = 0; fr < std::min<uint32_t>(info.frames, 2); fr++)
                        for (uint32_t fc = 0; fc < std::min<uint32_t>(info.faces, 2); fc++) vt::VtfDecode(b.get(), n, fr, fc, out.get(), info.rgba_bytes, nullptr);
                }
            } else if (kind == "bsp") {
                uint64_t n; auto b = slurp(argv[i], n);
                vt_bsp_info info; vt::BspInfo(n ? b.get() : nullptr, n, &info);
                uint64_t nt = vt::BspTriangles(b.get(), n, nullptr, nullptr, nullptr, 0);
                std::unique_ptr<vt_tri_in[]> tris(new vt_tri_in[nt ? nt : 1]);
                std::unique_ptr<float[]> bino(new float[nt ? nt * 9 : 1]);
                std::unique_ptr<int16_t[]> ti(new int16_t[nt ? nt : 1]);
                vt::BspTriangles(b.get(), n, tris.get(), bino.get(), ti.get(), nt);
                for (uint32_t k = 0; k < std::min<uint32_t>(info.n_materials, 8); k++) { vt_bsp_material m; vt::BspMaterial(b.get(), n, k, &m); }
                for (uint32_t k = 0; k < std::min<uint32_t>(info.n_static_props, 8); k++) { vt_bsp_static_prop p; vt::BspStaticProp(b.get(), n, k, &p); }
            } else {
                uint64_t n0, n1, n2; auto a = slurp(argv[i], n0); auto b = slurp(argv[i + 1], n1); auto c = slurp(argv[i + 2], n2);
                vt_mdl_files f{n0 ? a.get() : nullptr, n0, n1 ? b.get() : nullptr, n1, n2 ? c.get() : nullptr, n2};
                vt_mdl_info info; vt::MdlInfo(&f, &info);
                for (uint32_t bg = 0; bg < std::min<uint32_t>(info.n_bodygroups, 4); bg++) {
                    uint32_t nv = vt::MdlBodygroupValues(&f, bg);
                    for (uint32_t v = 0; v < std::min<uint32_t>(nv, 4); v++) {
                        uint64_t nt = vt::MdlMeshTriangles(&f, bg, v, nullptr, nullptr, 0);
                        std::unique_ptr<vt_tri_in[]> tris(new vt_tri_in[nt ? nt : 1]);
                        std::unique_ptr<vt_tri_skin[]> skin(new vt_tri_skin[nt ? nt : 1]);
                        vt::MdlMeshTriangles(&f, bg, v, tris.get(), skin.get(), nt);
                    }
                }
                std::unique_ptr<float[]> binds(new float[info.n_bones ? info.n_bones * 16 : 1]);
                vt::MdlBindMatrices(&f, binds.get());
                for (uint32_t s = 0; s < std::min<uint32_t>(info.n_skin_families, 3); s++)
                    for (uint32_t m = 0; m < std::min<uint32_t>(info.n_skin_refs, 4); m++) vt::MdlMaterialIndex(&f, s, m);
                for (uint32_t m = 0; m < std::min<uint32_t>(info.n_materials, 4); m++)
                    for (uint32_t d = 0; d < std::min<uint32_t>(info.n_material_dirs, 2); d++) vt::MdlMaterialPath(&f, m, d);
            }
            ok++;
        } catch (const std::exception &) { bad++; }
    }
    std::printf("%s: accepted %llu rejected %llu\n", kind.c_str(), (unsigned long long)ok, (unsigned long long)bad);
    return 0;
}
