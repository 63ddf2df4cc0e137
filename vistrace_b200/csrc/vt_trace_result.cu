// vt_trace_result.cu — K2: eager, batched TraceResult for sm_100a.
//
// One thread per (ray, hit) record.  Replaces the constructor and every lazy getter of
// TraceResult (source/objects/TraceResult.cpp:45-262): barycentric interpolation of position,
// normal, tangent, binormal and UV, the normal-map / blend / detail / MRAO texture paths, the
// grazing-angle normal fix and the cone-footprint LOD.  The lazy flags of the reference only
// cache; evaluating everything once is order-independent and gives the same values.
//
// HBM traffic per hit: 16 B hit + 32 B ray + one 176-byte VtTriAttr gather + the material and
// entity records (L2-resident, a few KB in total) + 4 texels per sampled texture, 128 B out.
#include "vt_kernels.h"
#include "vt_math.cuh"

namespace {

// TextureCombine — source/objects/TraceResult.cpp:11-43 (DetailBlendMode order: Material.h:12-26)
VT_DEV Px texture_combine(Px base, Px det, uint32_t mode, float bf) {
    Px r = base;
    switch (mode) {
    case 0:  // DecalModulate
        r.r = base.r * glm_lerp(1.f, 2.f * det.r, bf);
        r.g = base.g * glm_lerp(1.f, 2.f * det.g, bf);
        r.b = base.b * glm_lerp(1.f, 2.f * det.b, bf);
        r.a = base.a * 1.f;
        return r;
    case 1:  // Additive
    case 5:  // UnlitAdditive
    case 6:  // UnlitAdditiveThresholdFade
        r.r = base.r + bf * det.r;
        r.g = base.g + bf * det.g;
        r.b = base.b + bf * det.b;
        r.a = base.a + 0.f;
        return r;
    case 2: {  // TranslucentDetail
        const float blend = bf * det.a;
        r.r = glm_lerp(base.r, det.r, blend);
        r.g = glm_lerp(base.g, det.g, blend);
        r.b = glm_lerp(base.b, det.b, blend);
        return r;
    }
    case 3:  // BlendFactorFade
        r.r = glm_lerp(base.r, det.r, bf);
        r.g = glm_lerp(base.g, det.g, bf);
        r.b = glm_lerp(base.b, det.b, bf);
        r.a = glm_lerp(base.a, det.a, bf);
        return r;
    case 4: {  // TranslucentBase
        const float blend = bf * (1.f - base.a);
        r.r = glm_lerp(base.r, det.r, blend);
        r.g = glm_lerp(base.g, det.g, blend);
        r.b = glm_lerp(base.b, det.b, blend);
        r.a = det.a;
        return r;
    }
    case 7: {  // TwoPatternDecalModulate
        const float dc = glm_lerp(det.r, det.a, base.a);
        const float m = glm_lerp(1.f, 2.f * dc, bf);
        r.r = base.r * m;
        r.g = base.g * m;
        r.b = base.b * m;
        r.a = base.a * 1.f;
        return r;
    }
    case 8:  // Multiply
        r.r = glm_lerp(base.r, base.r * det.r, bf);
        r.g = glm_lerp(base.g, base.g * det.g, bf);
        r.b = glm_lerp(base.b, base.b * det.b, bf);
        r.a = glm_lerp(base.a, base.a * det.a, bf);
        return r;
    case 9:  // BaseMaskDetailAlpha
        r.a = glm_lerp(base.a, base.a * det.a, bf);
        return r;
    default:  // SSBump, SSBumpAlbedo: not implemented by the reference either
        return base;
    }
}

// TriUVInfoToTexLOD — source/Utils.h:75-78
VT_DEV float tex_lod(const VtDevTexture &t, V2 info) { return info.x + 0.5f * log2f((float)((int)t.width * (int)t.height) * info.y); }

VT_DEV void st4(float *dst, float a, float b, float c, float d) { *reinterpret_cast<float4 *>(dst) = make_float4(a, b, c, d); }

#ifndef VT_K2_MIN_BLOCKS
#define VT_K2_MIN_BLOCKS 1
#endif
__global__ void __launch_bounds__(128, VT_K2_MIN_BLOCKS)
k_trace_result(const VtSceneView S, const vt_ray *__restrict__ rays, const vt_hit *__restrict__ hits,
               const float *__restrict__ cones, vt_attr *__restrict__ attrs, unsigned long long n,
               const uint32_t *__restrict__ queue, const unsigned long long *__restrict__ queue_count) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (queue) {  // wave compaction: only the slots a generator listed get a TraceResult; everything else stays untouched
        if (i >= min(n, *queue_count)) return;
        i = __ldg(queue + i);
    } else if (i >= n) {
        return;
    }
    float *out = reinterpret_cast<float *>(attrs + i);
    const float4 h = __ldg(reinterpret_cast<const float4 *>(hits) + i);
    const uint32_t prim = __float_as_uint(h.w);
    if (prim == VT_MISS || prim >= S.n_tris) {  // Traverse returned nil (source/objects/AccelStruct.cpp:837)
        for (int k = 0; k < 7; k++) st4(out + 4 * k, 0.f, 0.f, 0.f, 0.f);
        st4(out + 28, 0.f, 0.f, 0.f, __uint_as_float(VT_MISS));
        return;
    }
    const float dist = h.x, bu = h.y, bv = h.z;
    const float4 rb = __ldg(reinterpret_cast<const float4 *>(rays + i) + 1);
    float coneWidth = -1.f, coneAngle = -1.f;  // AccelStruct.cpp:795-799 defaults
    if (cones) {
        coneWidth = __ldg(cones + 2 * i);
        coneAngle = __ldg(cones + 2 * i + 1);
    }

    // gather the triangle's attribute record (11 x 16 B)
    VtTriAttr ta;
    {
        const float4 *src = reinterpret_cast<const float4 *>(S.attrs + prim);
        float4 *dst = reinterpret_cast<float4 *>(&ta);
#pragma unroll
        for (int k = 0; k < 11; k++) dst[k] = __ldg(src + k);
    }
    const VtDevMaterial &mat = S.mats[ta.material];
    const VtDevEntity &ent = S.ents[ta.ent_idx];
    const uint8_t *texels = S.texels;
    const VtDevTexture &baseTexture = S.texs[mat.base_texture >= 0 ? (uint32_t)mat.base_texture : S.fallback_tex];

    // ---- constructor, TraceResult.cpp:45-86
    const V3 wo = neg(glm_normalize(mk3(rb.x, rb.y, rb.z)));  // AccelStruct.cpp:826, TraceResult.cpp:56
    const bool mipOverride = (coneWidth < 0.f || coneAngle <= 0.f);  // :53
    V3 vN[3], vT[3], vB[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        vN[k] = ld3(ta.normals[k]);
        vT[k] = ld3(ta.tangents[k]);
        vB[k] = glm_cross(vT[k], vN[k]);  // :61
    }
    const V3 v0 = ld3(ta.p0), v1 = ld3(ta.p0) - ld3(ta.e1), v2 = ld3(ta.p0) + ld3(ta.e2);  // :65-68
    const V3 uvw = mk3(bu, bv, 1.f - bu - bv);                                             // :70
    const V3 gN = ld3(ta.nNorm);
    float blendFactor = uvw.z * ta.alphas[0] + uvw.x * ta.alphas[1] + uvw.y * ta.alphas[2];  // :73
    const V2 texUV{uvw.z * ta.uvs[0][0] + uvw.x * ta.uvs[1][0] + uvw.y * ta.uvs[2][0],
                   uvw.z * ta.uvs[0][1] + uvw.x * ta.uvs[1][1] + uvw.y * ta.uvs[2][1]};  // :74
    V3 albedo = mk3(ent.colour[0] * mat.colour[0], ent.colour[1] * mat.colour[1], ent.colour[2] * mat.colour[2]);  // :80
    float alpha = ent.colour[3] * mat.colour[3];
    const bool hitSky = (mat.surf_flags & VT_SURF_SKY) != 0;  // :83
    const bool frontFacing = glm_dot(wo, gN) >= 0.f;          // :85

    // ---- CalcFootprint, :89-104
    V2 lodInfo{0.f, 0.f};
    if (!mipOverride) {
        coneWidth = coneAngle * dist + coneWidth;
        const float normalTerm = glm_dot(wo, gN);
        lodInfo.x = ta.lod;
        lodInfo.y = (coneWidth * coneWidth) / (normalTerm * normalTerm);
    }
#define VT_LOD(tex) (mipOverride ? 0.f : tex_lod((tex), lodInfo))

    // ---- CalcBlendFactor, :106-130
    if (mat.masked_blending) blendFactor = 0.5f;
    if (mat.blend_texture >= 0) {
        const VtDevTexture &bt = S.texs[mat.blend_texture];
        const V2 sc = transform_texcoord(texUV, mat.blend_tex_mat, mat.tex_scale);
        const Px pb = tex_sample(bt, texels, sc.x, sc.y, VT_LOD(bt));
        if (mat.masked_blending) {
            blendFactor = pb.g;
        } else {
            const float minb = glm_clamp(pb.g - pb.r, 0.f, 1.f);
            const float maxb = glm_clamp(pb.g + pb.r, 0.f, 1.f);
            blendFactor = glm_smoothstep(minb, maxb, blendFactor);
        }
    }

    // ---- GetPos, :255-262
    const V3 pos = v0 * uvw.z + v1 * uvw.x + v2 * uvw.y;

    // ---- CalcTBN, :132-187
    V3 normal = glm_normalize(vN[0] * uvw.z + vN[1] * uvw.x + vN[2] * uvw.y);
    V3 tangent = glm_normalize(vT[0] * uvw.z + vT[1] * uvw.x + vT[2] * uvw.y);
    V3 binormal = glm_normalize(vB[0] * uvw.z + vB[1] * uvw.x + vB[2] * uvw.y);
    if (mat.normal_map >= 0) {  // :140-174
        const VtDevTexture &nm = S.texs[mat.normal_map];
        V2 sc = transform_texcoord(texUV, mat.normal_map_mat, mat.tex_scale);
        Px pn = tex_sample(nm, texels, sc.x, sc.y, VT_LOD(nm));
        V3 mapped = mk3(pn.r * 2.f - 1.f, pn.g * 2.f - 1.f, pn.b * 2.f - 1.f);
        if (mat.normal_map2 >= 0) {
            const VtDevTexture &nm2 = S.texs[mat.normal_map2];
            sc = transform_texcoord(texUV, mat.normal_map_mat2, mat.tex_scale);
            pn = tex_sample(nm2, texels, sc.x, sc.y, VT_LOD(nm2));
            const V3 mapped2 = mk3(pn.r * 2.f - 1.f, pn.g * 2.f - 1.f, pn.b * 2.f - 1.f);
            mapped = glm_normalize(glm_lerp(mapped, mapped2, blendFactor));
        }
        // mat3 with columns (tangent, binormal, normal) times mappedNormal — glm type_mat3x3.inl:468-474
        V3 wn = mk3(tangent.x * mapped.x + binormal.x * mapped.y + normal.x * mapped.z,
                    tangent.y * mapped.x + binormal.y * mapped.y + normal.y * mapped.z,
                    tangent.z * mapped.x + binormal.z * mapped.y + normal.z * mapped.z);
        wn = glm_normalize(wn);
        if (isfinite(wn.x) && isfinite(wn.y) && isfinite(wn.z)) {
            normal = wn;
            tangent = glm_normalize(tangent - normal * glm_dot(tangent, normal));
            binormal = glm_cross(tangent, normal);
        }
    }
    {  // grazing-angle fix, :176-184
        const float kCosThetaThreshold = 0.1f;
        const float cosTheta = fabsf(glm_dot(wo, normal));
        if (cosTheta <= kCosThetaThreshold) {
            const float t = glm_clamp(cosTheta * (1.f / kCosThetaThreshold), 0.f, 1.f);
            normal = glm_normalize(glm_lerp(gN, normal, t));
            tangent = glm_normalize(tangent - normal * glm_dot(tangent, normal));
            binormal = glm_cross(tangent, normal);
        }
    }

    // ---- CalcShadingData, :189-253
    const V2 scaled = transform_texcoord(texUV, mat.base_tex_mat, mat.tex_scale);
    const V2 scaled2 = transform_texcoord(texUV, mat.base_tex_mat2, mat.tex_scale);
    const float baseMip = VT_LOD(baseTexture);
    Px colour = tex_sample(baseTexture, texels, scaled.x, scaled.y, baseMip);
    if (mat.base_texture2 >= 0) {
        const VtDevTexture &b2 = S.texs[mat.base_texture2];
        const Px c2 = tex_sample(b2, texels, scaled2.x, scaled2.y, VT_LOD(b2));
        colour.r = glm_lerp(colour.r, c2.r, blendFactor);
        colour.g = glm_lerp(colour.g, c2.g, blendFactor);
        colour.b = glm_lerp(colour.b, c2.b, blendFactor);
        colour.a = glm_lerp(colour.a, c2.a, blendFactor);
    }
    if (mat.detail >= 0) {
        const VtDevTexture &dt = S.texs[mat.detail];
        const V2 duv = transform_texcoord(texUV, mat.detail_mat, mat.detail_scale);
        const Px dc = tex_sample(dt, texels, duv.x, duv.y, VT_LOD(dt));
        colour = texture_combine(colour, dc, mat.detail_blend_mode, mat.detail_blend_factor);
        colour.r = glm_clamp(colour.r, 0.f, 1.f);
        colour.g = glm_clamp(colour.g, 0.f, 1.f);
        colour.b = glm_clamp(colour.b, 0.f, 1.f);
        colour.a = glm_clamp(colour.a, 0.f, 1.f);
    }
    albedo = albedo * mk3(colour.r, colour.g, colour.b);
    alpha *= colour.a;
    float metalness = 0.f, roughness = 1.f;  // TraceResult.h:46-47
    if (mat.mrao >= 0) {
        const VtDevTexture &mt = S.texs[mat.mrao];
        const Px pm = tex_sample(mt, texels, scaled.x, scaled.y, VT_LOD(mt));
        float mr = pm.r, mg = pm.g;
        if (mat.mrao2 >= 0) {
            const VtDevTexture &mt2 = S.texs[mat.mrao2];
            const Px pm2 = tex_sample(mt2, texels, scaled2.x, scaled2.y, VT_LOD(mt2));
            mr = glm_lerp(mr, pm2.r, blendFactor);
            mg = glm_lerp(mg, pm2.g, blendFactor);
        }
        metalness = mr;
        roughness = mg;
    }
#undef VT_LOD

    const uint32_t flags = (frontFacing ? VT_ATTR_FRONT_FACING : 0u) | (hitSky ? VT_ATTR_HIT_SKY : 0u) |
                           (mat.water ? VT_ATTR_HIT_WATER : 0u);
    // vt_attr, eight 16-byte stores
    st4(out + 0, pos.x, pos.y, pos.z, dist);
    st4(out + 4, normal.x, normal.y, normal.z, alpha);
    st4(out + 8, tangent.x, tangent.y, tangent.z, metalness);
    st4(out + 12, binormal.x, binormal.y, binormal.z, roughness);
    st4(out + 16, gN.x, gN.y, gN.z, baseMip);
    st4(out + 20, albedo.x, albedo.y, albedo.z, __uint_as_float(ent.id));
    st4(out + 24, uvw.x, uvw.y, uvw.z, __uint_as_float(ta.material));
    st4(out + 28, texUV.x, texUV.y, __uint_as_float(flags), __uint_as_float(prim));
}

}  // namespace

cudaError_t vt_launch_trace_result(const VtSceneView &S, const vt_ray *rays, const vt_hit *hits, const float *cones,
                                   vt_attr *attrs, uint64_t n, cudaStream_t stream, const uint32_t *queue, const unsigned long long *queue_count) {
    if (n == 0) return cudaSuccess;
    if ((queue == nullptr) != (queue_count == nullptr)) return cudaErrorInvalidValue;
    const unsigned block = 128;
    const unsigned long long grid = (n + block - 1) / block;  // queued: blocks past *queue_count exit at once
    k_trace_result<<<(unsigned)grid, block, 0, stream>>>(S, rays, hits, cones, attrs, (unsigned long long)n, queue, queue_count);
    return cudaGetLastError();
}
