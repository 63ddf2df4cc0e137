/*
 * vistrace_b200_ivtf.hpp — C++ helper on the HOST side of the boundary: any IVTFTexture implementation -> vt_texture.
 *
 * The C ABI (vistrace_b200.h) cannot take a C++ interface pointer, and Material holds `const IVTFTexture*`
 * (source/objects/Material.h:81...), which may be an extension's own texture class
 * (include/vistrace/IVTFTexture.h:22-143).  This header flattens such a texture into the RGBA8888 mip chain the device
 * samples, through the PUBLIC interface only — GetMIPLevels / GetWidth / GetHeight / GetPixel
 * (IVTFTexture.h:44-97) — in VTF order: smallest mip first (libs/VTFParser/VTFParser.cpp:44-78).
 *
 * Header-only and templated on the texture type, so this repository does not depend on the reference's headers; the
 * reference-side binding includes <vistrace/IVTFTexture.h> first and calls vt::decode_ivtf_texture(material.baseTexture, ...).
 * Compiled and exercised by oracle/ref_binding.cpp (tests/test_gpu_parity.py::test_compiled_reference_side_binding).
 *
 * Exactness: every byte b satisfies b / 255.f == the channel GetPixel returned when the source format has 8 bits per channel
 * (what VTFTexture produces for RGBA8888 / DXTn / the other 8-bit formats, FileFormat/Parser.cpp:157-262); channels outside
 * [0, 1] (the reference's 16-bit packed formats) are clamped — use vt_vtf_read_info().supported to detect those files.
 */
#ifndef VISTRACE_B200_IVTF_HPP
#define VISTRACE_B200_IVTF_HPP

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "vistrace_b200.h"

namespace vt {

/* storage receives the texels and must outlive the returned vt_texture (its rgba points into it).
 * texture_flags: VTF TEXTURE_FLAGS (VT_TEXFLAG_CLAMPS / CLAMPT) — not part of IVTFTexture, pass what the loader knows. */
template <class IVTF>
inline vt_texture decode_ivtf_texture(const IVTF *tex, std::vector<uint8_t> &storage, uint32_t texture_flags = 0, uint16_t frame = 0,
                                      uint8_t face = 0) {
    const uint16_t mips = tex->GetMIPLevels();
    storage.clear();
    for (int m = (int)mips - 1; m >= 0; m--) {
        const uint16_t w = tex->GetWidth((uint8_t)m), h = tex->GetHeight((uint8_t)m);
        for (uint16_t y = 0; y < h; y++)
            for (uint16_t x = 0; x < w; x++) {
                const auto p = tex->GetPixel(x, y, 0, (uint8_t)m, frame, face);
                const float c[4] = {p.r, p.g, p.b, p.a};
                for (float v : c) storage.push_back((uint8_t)std::lround((v < 0.f ? 0.f : (v > 1.f ? 1.f : v)) * 255.f));
            }
    }
    vt_texture t;
    std::memset(&t, 0, sizeof(t));
    t.width = tex->GetWidth(0);
    t.height = tex->GetHeight(0);
    t.mip_count = mips;
    t.flags = texture_flags;
    t.rgba = storage.data();
    t.nbytes = storage.size();
    return t;
}

}  // namespace vt

#endif /* VISTRACE_B200_IVTF_HPP */
