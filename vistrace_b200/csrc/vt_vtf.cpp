// vt_vtf.cpp — VTF file -> texel mip chain (the vt_texture the alpha-test and TraceResult kernels sample): RGBA8888, or wide texels.
//
// SURVEY.md §8 f4 (the data format on the input side of the path): the reference reads textures through
// libs/VTFParser — header (FileFormat/Parser.cpp:99-121, FileFormat/Structs.h:21-75), image-data location incl. the
// 7.3+ resource dictionary and the low-res thumbnail (Parser.cpp:123-155), DXT1/3/5 decompressed to RGBA8888 once at
// load (VTFParser.cpp:26-84, DXTn/DXT1.cpp, DXT3.cpp, DXT5.cpp — VTFLib's decoders), every other format interpreted per
// sample by ParsePixel (Parser.cpp:157-298).  This file restates exactly that, once, at ingestion: the output bytes b
// satisfy  b / 255.f == the float ParsePixel / the decompressor would hand VTFTexture::Sample  for every channel, so
// the device's manual bilinear filter over RGBA8888 (vt_math.cuh) sees the reference's texel values bit for bit.
//
// The 16-bit formats (RGB565, BGR565, BGRX5551, BGRA5551, BGRA4444: Parser.cpp:191-196,238-262; RGBA16161616(F): :281-294) are
// not representable in 8 bits per channel under the reference's arithmetic — its shifts are not masked, so green of an RGB565
// texel is ((b0 << 5) + ((b1 & 0xE0) >> 3)) / 255.f, up to 32.1 — and are decoded to WIDE texels instead: four 16-bit numerators
// per texel plus a divisor code per channel (255, 65535 or 1; vt_texture.texel_layout), numerator / divisor being exactly the float
// ParsePixel returns.  P8 has no case in ParsePixel and reads as VTFPixel{} = opaque black (Parser.cpp:295-297).
//
// Quirks kept: 5/6-bit endpoints widen by a plain shift (<< 3, << 2: white is 248/252/248, DXT1.cpp:31-39); DXT1's
// 3-colour mode still derives colour 3 as (c0 + 2 c1 + 1) / 3 with alpha 0 (DXT1.cpp:62-65); DXT3 alpha nibbles are
// replicated (a | a << 4, DXT3.cpp:73-74); DXT5's second 24-bit alpha word is read from byte 3 of the mask
// (DXT5.cpp:113); BGRX8888 keeps byte 3 as alpha (Parser.cpp:229-236); formats without alpha get 1.0, A8 gets black.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "vt_host.h"

namespace vt {

namespace {

enum Format : int32_t {  // FileFormat/Enums.h:5-35
    F_NONE = -1, F_RGBA8888 = 0, F_ABGR8888, F_RGB888, F_BGR888, F_RGB565, F_I8, F_IA88, F_P8, F_A8, F_RGB888_BLUESCREEN,
    F_BGR888_BLUESCREEN, F_ARGB8888, F_BGRA8888, F_DXT1, F_DXT3, F_DXT5, F_BGRX8888, F_BGR565, F_BGRX5551, F_BGRA4444,
    F_DXT1_ONEBITALPHA, F_BGRA5551, F_UV88, F_UVWQ8888, F_RGBA16161616F, F_RGBA16161616, F_UVLX8888
};

// bytes per pixel of the uncompressed formats (Parser.cpp:6-34); 0 = compressed / unsupported here
uint32_t bytes_per_pixel(int32_t f) {
    switch (f) {
        case F_RGBA8888: case F_ABGR8888: case F_ARGB8888: case F_BGRA8888: case F_BGRX8888: case F_UVWQ8888: case F_UVLX8888: return 4;
        case F_RGB888: case F_BGR888: case F_RGB888_BLUESCREEN: case F_BGR888_BLUESCREEN: return 3;
        case F_RGB565: case F_BGR565: case F_BGRX5551: case F_BGRA5551: case F_BGRA4444: case F_IA88: case F_UV88: return 2;
        case F_I8: case F_A8: case F_P8: return 1;
        case F_RGBA16161616F: case F_RGBA16161616: return 8;
        default: return 0;
    }
}
bool is_dxt(int32_t f) { return f == F_DXT1 || f == F_DXT1_ONEBITALPHA || f == F_DXT3 || f == F_DXT5; }

// CalcImageSize for one mip (Parser.cpp:44-68)
uint64_t image_size(uint32_t w, uint32_t h, uint32_t d, int32_t f) {
    if (is_dxt(f)) {
        if (w < 4 && w > 0) w = 4;
        if (h < 4 && h > 0) h = 4;
        return (uint64_t)((w + 3) / 4) * ((h + 3) / 4) * ((f == F_DXT3 || f == F_DXT5) ? 16 : 8) * d;
    }
    return (uint64_t)w * h * d * bytes_per_pixel(f);
}

struct Header {  // the fields of VTFHeader this path reads (FileFormat/Structs.h:21-75, #pragma pack(1))
    uint32_t version_minor, header_size, flags, num_resources;
    uint16_t width, height, frames, first_frame, depth;
    int32_t format, low_format;
    uint8_t mips, low_w, low_h;
    uint64_t data_offset;
};

template <typename T>
T rd(const uint8_t *p, size_t off) {
    T v;
    std::memcpy(&v, p + off, sizeof(T));
    return v;
}

// ParseHeader + the offset logic of ParseImageData (Parser.cpp:99-155)
Header parse_header(const uint8_t *p, uint64_t size) {
    if (!p || size < 16) throw std::runtime_error("vtf: file shorter than the base header");
    if (std::memcmp(p, "VTF\0", 4) != 0) throw std::runtime_error("vtf: bad signature");
    const uint32_t major = rd<uint32_t>(p, 4), minor = rd<uint32_t>(p, 8);
    if (major != 7 || minor > 5) throw std::runtime_error("vtf: unsupported version (7.0 - 7.5)");
    Header h{};
    h.version_minor = minor;
    h.header_size = rd<uint32_t>(p, 12);
    if (h.header_size > size || h.header_size > 80 + 8 * 32) throw std::runtime_error("vtf: bad header size");
    // the reference copies headerSize bytes over a zeroed struct: fields past the end of a short header read as zero
    uint8_t hdr[80 + 8 * 32] = {0};
    std::memcpy(hdr, p, h.header_size);
    h.width = rd<uint16_t>(hdr, 16);
    h.height = rd<uint16_t>(hdr, 18);
    h.flags = rd<uint32_t>(hdr, 20);
    h.frames = rd<uint16_t>(hdr, 24);
    h.first_frame = rd<uint16_t>(hdr, 26);
    h.format = rd<int32_t>(hdr, 52);
    h.mips = rd<uint8_t>(hdr, 56);
    h.low_format = rd<int32_t>(hdr, 57);
    h.low_w = rd<uint8_t>(hdr, 61);
    h.low_h = rd<uint8_t>(hdr, 62);
    h.depth = minor < 2 ? (uint16_t)1 : rd<uint16_t>(hdr, 63);
    h.num_resources = minor < 3 ? 0u : rd<uint32_t>(hdr, 68);
    if (h.format == F_NONE) throw std::runtime_error("vtf: no high-resolution image");
    if (h.num_resources > 0) {
        if (h.num_resources > 32) throw std::runtime_error("vtf: more than 32 resources");
        uint32_t off = 0;
        for (uint32_t i = 0; i < h.num_resources; i++) {  // resource dictionary at byte 80: {tag[3], flags, data}
            const uint8_t *e = hdr + 80 + 8 * i;
            if (e[0] == 0x30 && e[1] == 0 && e[2] == 0) {
                if (off != 0) throw std::runtime_error("vtf: two high-resolution image resources");
                off = rd<uint32_t>(e, 4);
            }
        }
        h.data_offset = off;
    } else {
        uint64_t low = 0;
        if (h.low_format != F_NONE) low = image_size(h.low_w, h.low_h, 1, h.low_format);
        h.data_offset = h.header_size + low;
    }
    return h;
}

uint32_t face_count(const Header &h) {  // GetFaceCount (Parser.cpp:93-97)
    if (!(h.flags & 0x00004000u)) return 1;
    return (h.first_frame != 0xffff && h.version_minor < 5) ? 7 : 6;
}

struct C8 {
    uint8_t r, g, b, a;
};

void endpoints(const uint8_t *blk, C8 c[4], bool dxt1) {
    const uint16_t c0 = rd<uint16_t>(blk, 0), c1 = rd<uint16_t>(blk, 2);
    // Colour565 bit-fields: blue = bits 0-4, green = 5-10, red = 11-15; widened by plain shifts (DXT1.cpp:31-39)
    c[0] = {(uint8_t)(((c0 >> 11) & 31) << 3), (uint8_t)(((c0 >> 5) & 63) << 2), (uint8_t)((c0 & 31) << 3), 0xFF};
    c[1] = {(uint8_t)(((c1 >> 11) & 31) << 3), (uint8_t)(((c1 >> 5) & 63) << 2), (uint8_t)((c1 & 31) << 3), 0xFF};
    if (!dxt1 || c0 > c1) {  // four-colour block; DXT3/5 always (DXT3.cpp:44-52, DXT5.cpp:50-58)
        c[2] = {(uint8_t)((2 * c[0].r + c[1].r + 1) / 3), (uint8_t)((2 * c[0].g + c[1].g + 1) / 3), (uint8_t)((2 * c[0].b + c[1].b + 1) / 3), 0xFF};
        c[3] = {(uint8_t)((c[0].r + 2 * c[1].r + 1) / 3), (uint8_t)((c[0].g + 2 * c[1].g + 1) / 3), (uint8_t)((c[0].b + 2 * c[1].b + 1) / 3), 0xFF};
    } else {  // three-colour block: colour 3 is transparent but keeps the interpolated rgb (DXT1.cpp:57-65)
        c[2] = {(uint8_t)((c[0].r + c[1].r) / 2), (uint8_t)((c[0].g + c[1].g) / 2), (uint8_t)((c[0].b + c[1].b) / 2), 0xFF};
        c[3] = {(uint8_t)((c[0].r + 2 * c[1].r + 1) / 3), (uint8_t)((c[0].g + 2 * c[1].g + 1) / 3), (uint8_t)((c[0].b + 2 * c[1].b + 1) / 3), 0x00};
    }
}

// DecompressDXT1 / DXT3 / DXT5 (DXTn/*.cpp) for one w x h image into tightly packed RGBA8888
void decode_dxt(const uint8_t *src, uint8_t *dst, uint32_t w, uint32_t h, int32_t f) {
    const bool dxt1 = f == F_DXT1 || f == F_DXT1_ONEBITALPHA;
    for (uint32_t y = 0; y < h; y += 4)
        for (uint32_t x = 0; x < w; x += 4) {
            const uint8_t *ablk = src;
            if (!dxt1) src += 8;
            C8 c[4];
            endpoints(src, c, dxt1);
            const uint32_t bits = rd<uint32_t>(src, 4);
            src += 8;
            uint8_t alphas[8];
            uint32_t abits[2] = {0, 0};
            if (f == F_DXT5) {
                alphas[0] = ablk[0];
                alphas[1] = ablk[1];
                if (alphas[0] > alphas[1]) {  // 8-alpha block (DXT5.cpp:78-86)
                    for (int i = 1; i < 7; i++) alphas[i + 1] = (uint8_t)(((7 - i) * alphas[0] + i * alphas[1] + 3) / 7);
                } else {  // 6-alpha block (DXT5.cpp:87-96)
                    for (int i = 1; i < 5; i++) alphas[i + 1] = (uint8_t)(((5 - i) * alphas[0] + i * alphas[1] + 2) / 5);
                    alphas[6] = 0x00;
                    alphas[7] = 0xFF;
                }
                abits[0] = ablk[2] | (ablk[3] << 8) | (ablk[4] << 16);  // rows 0-1: 24 bits from mask byte 0
                abits[1] = ablk[5] | (ablk[6] << 8) | (ablk[7] << 16);  // rows 2-3: 24 bits from mask byte 3
            }
            for (uint32_t j = 0, k = 0; j < 4; j++)
                for (uint32_t i = 0; i < 4; i++, k++) {
                    if (x + i >= w || y + j >= h) continue;
                    uint8_t *o = dst + ((size_t)(y + j) * w + (x + i)) * 4;
                    const C8 &col = c[(bits >> (2 * k)) & 3];
                    o[0] = col.r, o[1] = col.g, o[2] = col.b;
                    if (dxt1) {
                        o[3] = col.a;
                    } else if (f == F_DXT3) {  // explicit 4-bit alpha, replicated (DXT3.cpp:67-79)
                        const uint16_t row = rd<uint16_t>(ablk, 2 * j);
                        const uint8_t a = (row >> (4 * i)) & 0x0F;
                        o[3] = (uint8_t)(a | (a << 4));
                    } else {
                        o[3] = alphas[(abits[j >> 1] >> (3 * ((j & 1) * 4 + i))) & 7];
                    }
                }
        }
}

// ParsePixel (Parser.cpp:157-298) for the formats whose channels are 8-bit integers over 255
void convert_pixel(const uint8_t *p, int32_t f, uint8_t *o) {
    switch (f) {
        case F_RGBA8888: case F_UVWQ8888: case F_UVLX8888: o[0] = p[0], o[1] = p[1], o[2] = p[2], o[3] = p[3]; break;
        case F_ABGR8888: o[0] = p[3], o[1] = p[2], o[2] = p[1], o[3] = p[0]; break;
        case F_RGB888: case F_RGB888_BLUESCREEN: o[0] = p[0], o[1] = p[1], o[2] = p[2], o[3] = 255; break;
        case F_BGR888: case F_BGR888_BLUESCREEN: o[0] = p[2], o[1] = p[1], o[2] = p[0], o[3] = 255; break;
        case F_I8: o[0] = o[1] = o[2] = p[0], o[3] = 255; break;
        case F_IA88: o[0] = o[1] = o[2] = p[0], o[3] = p[1]; break;
        case F_A8: o[0] = o[1] = o[2] = 0, o[3] = p[0]; break;
        case F_ARGB8888: o[0] = p[1], o[1] = p[2], o[2] = p[3], o[3] = p[0]; break;
        case F_BGRA8888: case F_BGRX8888: o[0] = p[2], o[1] = p[1], o[2] = p[0], o[3] = p[3]; break;
        case F_UV88: o[0] = p[0], o[1] = p[1], o[2] = 0, o[3] = 255; break;
        case F_P8: o[0] = o[1] = o[2] = 0, o[3] = 255; break;  // no case in ParsePixel: VTFPixel{} (Parser.cpp:295-297)
        default: break;
    }
}

// ParsePixel for the formats whose channels leave [0, 255] / 255: the integer NUMERATORS of its expressions, evaluated in int as
// C++ does (integer promotion, no masking where the reference does not mask); the divisors are in wide_layout()
void convert_pixel_wide(const uint8_t *p, int32_t f, uint16_t *o) {
    const int b0 = p[0], b1 = p[1];
    switch (f) {
        case F_RGB565:  // :191-196
            o[0] = (uint16_t)(b0 & 0xF8), o[1] = (uint16_t)((b0 << 5) + ((b1 & 0xE0) >> 3)), o[2] = (uint16_t)((b1 & 0x1F) << 3), o[3] = 255;
            break;
        case F_BGR565:  // :238-243
            o[0] = (uint16_t)((b1 & 0x1F) << 3), o[1] = (uint16_t)((b0 << 5) + ((b1 & 0xE0) >> 3)), o[2] = (uint16_t)(b0 & 0xF8), o[3] = 255;
            break;
        case F_BGRX5551: case F_BGRA5551:  // :244-251: alpha is static_cast<float>(b1 & 1), divisor 1
            o[0] = (uint16_t)((b1 & 0x3E) << 2), o[1] = (uint16_t)((b0 << 5) + ((b1 & 0xC0) >> 3)), o[2] = (uint16_t)(b0 & 0xF8), o[3] = (uint16_t)(b1 & 1);
            break;
        case F_BGRA4444:  // :252-258
            o[0] = (uint16_t)(b1 & 0xF0), o[1] = (uint16_t)(b0 << 4), o[2] = (uint16_t)(b0 & 0xF0), o[3] = (uint16_t)(b1 << 4);
            break;
        case F_RGBA16161616: case F_RGBA16161616F:  // :281-294: both read as four uint16 over 65535
            std::memcpy(o, p, 8);
            break;
        default: break;
    }
}

// vt_texture.texel_layout of a format: 0 = RGBA8888, else VT_TEXEL_WIDE | divisor code of channel c in bits 2c, 2c + 1
uint32_t wide_layout(int32_t f) {
    switch (f) {
        case F_RGB565: case F_BGR565: case F_BGRA4444: return VT_TEXEL_WIDE;                                  // every channel over 255
        case F_BGRX5551: case F_BGRA5551: return VT_TEXEL_WIDE | (VT_TEXEL_DIV_1 << 6);                       // alpha over 1
        case F_RGBA16161616: case F_RGBA16161616F: return VT_TEXEL_WIDE | (VT_TEXEL_DIV_65535 * 0x55u);       // every channel over 65535
        default: return 0;
    }
}

bool representable(int32_t f) {
    switch (f) {
        case F_RGBA8888: case F_UVWQ8888: case F_UVLX8888: case F_ABGR8888: case F_RGB888: case F_RGB888_BLUESCREEN: case F_BGR888:
        case F_BGR888_BLUESCREEN: case F_I8: case F_IA88: case F_A8: case F_ARGB8888: case F_BGRA8888: case F_BGRX8888: case F_UV88:
        case F_DXT1: case F_DXT1_ONEBITALPHA: case F_DXT3: case F_DXT5: case F_P8: return true;
        default: return wide_layout(f) != 0;
    }
}

}  // namespace

void VtfInfo(const uint8_t *file, uint64_t size, vt_vtf_info *out) {
    const Header h = parse_header(file, size);
    if (h.width == 0 || h.height == 0 || h.mips == 0 || h.mips > 16 || h.frames == 0) throw std::runtime_error("vtf: empty image");
    std::memset(out, 0, sizeof(*out));
    out->width = h.width;
    out->height = h.height;
    out->mip_count = h.mips;
    out->flags = h.flags;
    out->format = h.format;
    out->frames = h.frames;
    out->faces = face_count(h);
    out->depth = h.depth ? h.depth : 1;
    out->supported = representable(h.format) ? 1 : 0;
    out->texel_layout = wide_layout(h.format);
    uint64_t n = 0;
    for (uint32_t m = 0; m < h.mips; m++) n += (uint64_t)std::max(1, h.width >> m) * std::max(1, h.height >> m) * (out->texel_layout ? 8 : 4);
    out->rgba_bytes = n;
}

void VtfDecode(const uint8_t *file, uint64_t size, uint32_t frame, uint32_t face, uint8_t *rgba, uint64_t capacity, vt_vtf_info *info_out) {
    vt_vtf_info info;
    VtfInfo(file, size, &info);
    if (info_out) *info_out = info;
    if (!info.supported)
        throw std::runtime_error("vtf: image format " + std::to_string(info.format) +
                                 " is not a format the reference's ParsePixel reads");
    if (frame >= info.frames || face >= info.faces) throw std::runtime_error("vtf: frame or face out of range");
    if (!rgba || capacity < info.rgba_bytes) throw std::runtime_error("vtf: output buffer too small");
    const Header h = parse_header(file, size);
    // whole image block must be inside the file (Parser.cpp:148)
    uint64_t total = 0;
    for (uint32_t m = 0; m < h.mips; m++)
        total += image_size(std::max(1, h.width >> m), std::max(1, h.height >> m), std::max<uint32_t>(1u, info.depth >> m), h.format);
    // untrusted header: every factor is checked against the file size BEFORE it is multiplied in, so the product cannot wrap
    // (16384 x 16384 x depth 65535 x 65535 frames is ~2^65 bytes)
    if (h.data_offset > size || total > size - h.data_offset) throw std::runtime_error("vtf: image data runs past the end of the file");
    const uint64_t copies = (uint64_t)h.frames * info.faces;  // <= 65535 * 7
    if (total != 0 && copies > (size - h.data_offset) / total) throw std::runtime_error("vtf: image data runs past the end of the file");
    total *= copies;
    // storage order: mips smallest first; inside a mip frames -> faces -> z slices (VTFParser.cpp:44-78,178-205)
    const uint8_t *src = file + h.data_offset;
    uint8_t *dst = rgba;
    for (int m = (int)h.mips - 1; m >= 0; m--) {
        const uint32_t w = std::max(1, h.width >> m), hh = std::max(1, h.height >> m), d = std::max<uint32_t>(1u, info.depth >> m);
        const uint64_t slice = image_size(w, hh, 1, h.format);
        const uint8_t *img = src + ((uint64_t)frame * info.faces + face) * slice * d;  // z slice 0
        if (img < file || slice > size || (uint64_t)(img - file) > size - slice) throw std::runtime_error("vtf: image data runs past the end of the file");
        if (is_dxt(h.format)) {
            decode_dxt(img, dst, w, hh, h.format);
        } else {
            const uint32_t bpp = bytes_per_pixel(h.format);
            if (info.texel_layout) {
                for (uint64_t i = 0; i < (uint64_t)w * hh; i++) {
                    uint16_t px[4];
                    convert_pixel_wide(img + i * bpp, h.format, px);
                    std::memcpy(dst + i * 8, px, 8);
                }
            } else {
                for (uint64_t i = 0; i < (uint64_t)w * hh; i++) convert_pixel(img + i * bpp, h.format, dst + i * 4);
            }
        }
        src += slice * d * h.frames * info.faces;
        dst += (uint64_t)w * hh * (info.texel_layout ? 8 : 4);
    }
}

}  // namespace vt
