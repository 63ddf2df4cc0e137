// vt_math.cuh — device arithmetic in the reference's exact operation order.
//
// This translation unit is compiled with --fmad=false: `a * b + c` is two IEEE-754
// roundings exactly as in a host build with -ffp-contract=off, and the ONLY fused
// operation is the explicit fmaf() in the slab test, where the reference calls
// bvh::fast_multiply_add (libs/bvh/include/bvh/utilities.hpp:44-54).  Division and
// sqrt are the IEEE round-to-nearest versions (nvcc defaults -prec-div/-prec-sqrt=true).
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "vt_device.h"

#define VT_DEV __device__ __forceinline__

struct V3 {
    float x, y, z;
};
struct V2 {
    float x, y;
};
struct Px {
    float r, g, b, a;
};

VT_DEV V3 mk3(float x, float y, float z) { return V3{x, y, z}; }
VT_DEV V3 ld3(const float *p) { return V3{p[0], p[1], p[2]}; }
VT_DEV V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
VT_DEV V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
VT_DEV V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
VT_DEV V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
VT_DEV V3 neg(V3 a) { return V3{-a.x, -a.y, -a.z}; }

// bvh::dot — libs/bvh/include/bvh/vector.hpp:134-141: sum = a0*b0; sum += a1*b1; sum += a2*b2
VT_DEV float bvh_dot(V3 a, V3 b) {
    float s = a.x * b.x;
    s += a.y * b.y;
    s += a.z * b.z;
    return s;
}
// bvh::cross — vector.hpp:159-167: r[i] = a[j]*b[k] - a[k]*b[j]
VT_DEV V3 bvh_cross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// glm::dot(vec3) — libs/glm/glm/detail/func_geometric.inl:48-54
VT_DEV float glm_dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// glm::cross — func_geometric.inl:68-78
VT_DEV V3 glm_cross(V3 x, V3 y) { return V3{x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y}; }
// glm::normalize = v * (1 / sqrt(dot(v, v))) — func_geometric.inl:88, func_exponential.inl:138
VT_DEV V3 glm_normalize(V3 v) { return v * (1.0f / sqrtf(glm_dot(v, v))); }
// glm::mix / gtx lerp — detail/func_common.inl:104-111: x*(1-a) + y*a
VT_DEV float glm_lerp(float x, float y, float a) { return x * (1.0f - a) + y * a; }
VT_DEV V3 glm_lerp(V3 x, V3 y, float a) { return x * (1.0f - a) + y * a; }
// glm::min / max / clamp — func_common.inl: min(a,b) = b<a ? b : a ; max(a,b) = a<b ? b : a
VT_DEV float glm_min(float a, float b) { return (b < a) ? b : a; }
VT_DEV float glm_max(float a, float b) { return (a < b) ? b : a; }
VT_DEV float glm_clamp(float x, float lo, float hi) { return glm_min(glm_max(x, lo), hi); }
// glm::smoothstep — func_common.inl:564-570
VT_DEV float glm_smoothstep(float e0, float e1, float x) {
    float t = glm_clamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}

// TransformTexcoord — source/Utils.h:65-72; glm::dot(vec4) = (x+y) + (z+w) of the products
VT_DEV V2 transform_texcoord(V2 tc, const float *m, float scale) {
    float x = (tc.x * m[0] + tc.y * m[1]) + (m[2] + m[3]);
    float y = (tc.x * m[4] + tc.y * m[5]) + (m[6] + m[7]);
    return V2{x * scale, y * scale};
}

VT_DEV float std_clampf(float v, float lo, float hi) { return (v < lo) ? lo : (hi < v) ? hi : v; }
VT_DEV int std_clampi(int v, int lo, int hi) { return (v < lo) ? lo : (hi < v) ? hi : v; }
// intmod — libs/VTFParser/VTFParser.cpp:9-11
VT_DEV int vtf_intmod(int a, int b) { return (a % b + b) % b; }

// Corner addressing + fractional weights of VTFTexture::SampleBilinear
// (libs/VTFParser/VTFParser.cpp:207-309), shared by the alpha-only and RGBA paths.
struct BilinearTaps {
    uint32_t o00, o10, o01, o11;  // texel indices inside the chain: corners[xOff][yOff]
    float uF, vF, uFi, vFi;
};

VT_DEV BilinearTaps bilinear_taps(const VtDevTexture &t, float u, float v, uint32_t mip) {
    uint32_t width = t.width >> mip, height = t.height >> mip;
    if (width < 1) width = 1;
    if (height < 1) height = 1;
    const bool clampX = (t.flags & 0x4u) != 0, clampY = (t.flags & 0x8u) != 0;  // TEXTURE_FLAGS CLAMPS / CLAMPT
    if (clampX) u = std_clampf(u, 0.f, 0.9999f); else u -= floorf(u);  // :250-258
    if (clampY) v = std_clampf(v, 0.f, 0.9999f); else v -= floorf(v);
    u = u * (float)width - 0.5f;  // :261-262
    v = v * (float)height - 0.5f;
    int x = (int)floorf(u), y = (int)floorf(v);
    BilinearTaps b;
    b.uF = u - (float)x;
    b.vF = v - (float)y;
    b.uFi = 1.f - b.uF;
    b.vFi = 1.f - b.vF;
    int x0, x1, y0, y1;
    if (clampX) {
        x0 = std_clampi(x, 0, (int)width - 1);
        x1 = std_clampi(x + 1, 0, (int)width - 1);
    } else {
        x0 = vtf_intmod(x, (int)width);
        x1 = vtf_intmod(x + 1, (int)width);
    }
    if (clampY) {
        y0 = std_clampi(y, 0, (int)height - 1);
        y1 = std_clampi(y + 1, 0, (int)height - 1);
    } else {
        y0 = vtf_intmod(y, (int)height);
        y1 = vtf_intmod(y + 1, (int)height);
    }
    const uint32_t base = t.mip_offset[mip];
    b.o00 = base + ((uint32_t)y0 * width + (uint32_t)x0);
    b.o10 = base + ((uint32_t)y0 * width + (uint32_t)x1);
    b.o01 = base + ((uint32_t)y1 * width + (uint32_t)x0);
    b.o11 = base + ((uint32_t)y1 * width + (uint32_t)x1);
    return b;
}

// ParsePixel(RGBA8888): channel / 255.f  (libs/VTFParser/FileFormat/Parser.cpp:161-168)
VT_DEV float u8f(uint32_t c) { return (float)c / 255.f; }

VT_DEV float bilerp(float c00, float c10, float c01, float c11, const BilinearTaps &b) {
    return (c00 * b.uFi + c10 * b.uF) * b.vFi + (c01 * b.uFi + c11 * b.uF) * b.vF;  // :295-308
}

VT_DEV uint32_t ld_texel(const uint8_t *texels, uint64_t base, uint32_t index) {
    return __ldg(reinterpret_cast<const uint32_t *>(texels + base) + index);
}
// Wide texels (vt_texture.texel_layout != 0): four uint16 numerators; channel = numerator / divisor, the divisor (255, 65535 or 1)
// named by a 2-bit code per channel — the reference's ParsePixel expressions for the 16-bit formats (FileFormat/Parser.cpp:190-196,
// 238-262, 281-294), which an RGBA8888 texel cannot hold.  Warp-uniform per texture in practice, rare: kept out of line of the byte path.
VT_DEV uint2 ld_texel_wide(const uint8_t *texels, uint64_t base, uint32_t index) {
    return __ldg(reinterpret_cast<const uint2 *>(texels + base) + index);
}
VT_DEV float wide_divisor(uint32_t layout, int c) {
    const uint32_t code = (layout >> (2 * c)) & 3u;
    return code == 0u ? 255.f : (code == 1u ? 65535.f : 1.f);
}
VT_DEV float wide_channel(uint2 p, int c) { return (float)((c < 2 ? p.x : p.y) >> ((c & 1) * 16) & 0xFFFFu); }

VT_DEV Px sample_bilinear(const VtDevTexture &t, const uint8_t *texels, float u, float v, uint32_t mip) {
    BilinearTaps b = bilinear_taps(t, u, v, mip);
    Px r;
    if (t.layout) {
        const uint2 q00 = ld_texel_wide(texels, t.base, b.o00), q10 = ld_texel_wide(texels, t.base, b.o10);
        const uint2 q01 = ld_texel_wide(texels, t.base, b.o01), q11 = ld_texel_wide(texels, t.base, b.o11);
        float ch[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const float d = wide_divisor(t.layout, c);
            ch[c] = bilerp(wide_channel(q00, c) / d, wide_channel(q10, c) / d, wide_channel(q01, c) / d, wide_channel(q11, c) / d, b);
        }
        r.r = ch[0], r.g = ch[1], r.b = ch[2], r.a = ch[3];
        return r;
    }
    uint32_t p00 = ld_texel(texels, t.base, b.o00), p10 = ld_texel(texels, t.base, b.o10);
    uint32_t p01 = ld_texel(texels, t.base, b.o01), p11 = ld_texel(texels, t.base, b.o11);
    r.r = bilerp(u8f(p00 & 255u), u8f(p10 & 255u), u8f(p01 & 255u), u8f(p11 & 255u), b);
    r.g = bilerp(u8f((p00 >> 8) & 255u), u8f((p10 >> 8) & 255u), u8f((p01 >> 8) & 255u), u8f((p11 >> 8) & 255u), b);
    r.b = bilerp(u8f((p00 >> 16) & 255u), u8f((p10 >> 16) & 255u), u8f((p01 >> 16) & 255u), u8f((p11 >> 16) & 255u), b);
    r.a = bilerp(u8f(p00 >> 24), u8f(p10 >> 24), u8f(p01 >> 24), u8f(p11 >> 24), b);
    return r;
}

// Alpha channel only, mip 0: what the alpha test consumes (source/objects/Primitives.h:203).
VT_DEV float sample_alpha_mip0(const VtDevTexture &t, const uint8_t *texels, float u, float v) {
    BilinearTaps b = bilinear_taps(t, u, v, 0);
    if (t.layout) {
        const float d = wide_divisor(t.layout, 3);
        return bilerp((float)(ld_texel_wide(texels, t.base, b.o00).y >> 16) / d, (float)(ld_texel_wide(texels, t.base, b.o10).y >> 16) / d,
                      (float)(ld_texel_wide(texels, t.base, b.o01).y >> 16) / d, (float)(ld_texel_wide(texels, t.base, b.o11).y >> 16) / d, b);
    }
    uint32_t p00 = ld_texel(texels, t.base, b.o00), p10 = ld_texel(texels, t.base, b.o10);
    uint32_t p01 = ld_texel(texels, t.base, b.o01), p11 = ld_texel(texels, t.base, b.o11);
    return bilerp(u8f(p00 >> 24), u8f(p10 >> 24), u8f(p01 >> 24), u8f(p11 >> 24), b);
}

// VTFTexture::Sample — libs/VTFParser/VTFParser.cpp:311-330
VT_DEV Px tex_sample(const VtDevTexture &t, const uint8_t *texels, float u, float v, float mipLevel) {
    mipLevel = std_clampf(mipLevel, 0.f, (float)(t.mips - 1));
    float mipHigh = floorf(mipLevel), mipLow = ceilf(mipLevel);
    Px high = sample_bilinear(t, texels, u, v, (uint32_t)(uint8_t)mipHigh);
    if (mipLow == mipHigh) return high;
    Px low = sample_bilinear(t, texels, u, v, (uint32_t)(uint8_t)mipLow);
    float fract = mipLevel - mipHigh, fractInv = 1.f - fract;
    return Px{low.r * fract + high.r * fractInv, low.g * fract + high.g * fractInv, low.b * fract + high.b * fractInv,
              low.a * fract + high.a * fractInv};
}

// Counter-based random numbers of the ray generators: a hash of (slot, dimension, seed) — the reference's Sampler is a sequential
// mt19937 (source/objects/Sampler.cpp:5-20) and cannot be evaluated in parallel.  Host + device, so that a caller (and the parity
// tests) can reproduce the numbers a kernel drew: vt_sample_uniform01 in the C ABI.
__host__ __device__ __forceinline__ uint32_t vt_mix32(uint32_t h) {
    h ^= h >> 16;
    h *= 0x7feb352du;
    h ^= h >> 15;
    h *= 0x846ca68bu;
    h ^= h >> 16;
    return h;
}
__host__ __device__ __forceinline__ float vt_uniform01(unsigned long long slot, uint32_t dim, unsigned long long seed) {
    uint32_t h = vt_mix32((uint32_t)slot ^ vt_mix32((uint32_t)(slot >> 32) + 0x9e3779b9u * (dim + 1u)));
    h = vt_mix32(h ^ (uint32_t)seed ^ vt_mix32((uint32_t)(seed >> 32) + dim));
    return (float)(h >> 8) * (1.0f / 16777216.0f);
}
