// vt_raygen.cu — K3: secondary-ray generation on the device (the step either side of the hot path).
//
// The reference has no batched ray generator: GLua scripts call vistrace.CalcRayOrigin
// (source/VisTrace.cpp:1478-1519) and the BSDF sampler (source/libraries/BSDF.cpp:69-77,
// hemisphere_cos) per hit and feed the result back into accel:Traverse.  For a wavefront that
// round trip would leave the GPU idle, so the "primary + diffuse" workload spawns its bounce rays
// here, straight from the vt_attr records K2 wrote:
//
//   direction = T * (sinTheta cos phi) + B' * (sinTheta sin phi) + N' * z,
//               z = sqrt(r1), sinTheta = sqrt(1 - r1), phi = 2 pi r2          (hemisphere_cos)
//   origin    = CalcRayOrigin(pos, geometric normal on the viewer's side)
//
// N', B' are the shading normal / binormal flipped to the viewer's side for back-face hits.
// r1, r2 come from a counter-based hash of (slot, dimension, seed) — the reference's Sampler is a
// sequential mt19937 (source/objects/Sampler.cpp:5-20) and cannot be evaluated in parallel.
// Slot j = i * spp + s belongs to primary ray i (the hash counter is slot_offset + j, so a batch
// traced in tiles draws the same numbers as the batch traced whole); hits that spawn nothing (miss, sky) leave a
// MASKED slot (tmax < 0) that K1 reports as a miss without counting it as an invalid ray.
//
// Ray queue: a generator can also LIST the slots it filled (queue[pos] = slot, pos from a device counter, one
// warp-aggregated atomic per warp) and write the miss record of every masked slot itself.  K1 then walks the queue
// only — masked slots never occupy a lane — and the hit buffer still comes out complete, in slot order.  The order of
// the queue depends on warp scheduling; the hit buffer does not (rays are independent).
#include "vt_kernels.h"
#include "vt_math.cuh"

namespace {

VT_DEV float uniform01(unsigned long long slot, uint32_t dim, unsigned long long seed) { return vt_uniform01(slot, dim, seed); }

// vistrace.CalcRayOrigin — source/VisTrace.cpp:1495-1517, one component
VT_DEV float ray_origin_1(float pos, float nrm) {
    const float origin = 1.f / 32.f, fScale = 1.f / 65536.f, iScale = 256.f;
    const int iOff = (int)(nrm * iScale);
    const float iPos = __int_as_float(__float_as_int(pos) + (pos < 0.f ? -iOff : iOff));
    return fabsf(pos) < origin ? pos + nrm * fScale : iPos;
}

// Block-level tail shared by the generators: count the spawned rays, append their slots to the queue, write the
// miss record of the masked ones.  Must be reached by every thread of the (256-thread) block.  ONE atomic per block and
// counter: a warp-aggregated atomic per warp is 260 k same-address atomics for a 1080p x 4 spp wave — they serialise
// in L2 and cost as much as the generator itself (measured: K3 0.08 -> 0.16 ms).
VT_DEV void publish_slot(bool in_range, bool spawned, unsigned long long slot, unsigned long long *live, uint32_t *queue,
                         unsigned long long *queue_count, vt_hit *miss_hits) {
    __shared__ unsigned s_warp[8];
    __shared__ unsigned long long s_base;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(0xffffffffu, spawned);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    unsigned before = 0, total = 0;
#pragma unroll
    for (unsigned w = 0; w < 8; w++) {
        const unsigned c = s_warp[w];
        before += w < warp ? c : 0u;
        total += c;
    }
    if (threadIdx.x == 0 && total) {
        if (live) atomicAdd(live, (unsigned long long)total);
        if (queue) s_base = atomicAdd(queue_count, (unsigned long long)total);
    }
    if (!queue) return;
    __syncthreads();
    if (spawned) queue[s_base + before + __popc(m & ((1u << lane) - 1u))] = (uint32_t)slot;
    else if (in_range) reinterpret_cast<float4 *>(miss_hits)[slot] = make_float4(0.f, 0.f, 0.f, __uint_as_float(VT_MISS));
}

__global__ void __launch_bounds__(256)
k_bounce_rays(const vt_attr *__restrict__ attrs, unsigned long long n, uint32_t spp, unsigned long long seed,
              unsigned long long slot_offset, vt_ray *__restrict__ out, unsigned long long *__restrict__ live,
              uint32_t *__restrict__ queue, unsigned long long *__restrict__ queue_count, vt_hit *__restrict__ miss_hits,
              const VtSlotMap map, const uint32_t *__restrict__ in_queue, const unsigned long long *__restrict__ in_count) {
    unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long total = n * spp;
    if (in_queue) {  // wave compaction: thread j handles sample j % spp of parent in_queue[j / spp]; whole blocks past the end leave at once
        total = min(n, *in_count) * spp;
        if ((unsigned long long)blockIdx.x * blockDim.x >= total) return;
        if (j < total) j = (unsigned long long)__ldg(in_queue + j / spp) * spp + j % spp;
        else j = ~0ull;
    }
    bool spawned = false;
    if (j < (in_queue ? ~0ull : total)) {
        const unsigned long long i = j / spp;
        // counter of the random-number hash: slot of the GLOBAL pixel (VtSlotMap, vt_kernels.h); identity map: slot_offset + j
        unsigned long long ctr = slot_offset + j;
        if (map.tile) {
            const unsigned long long li = map.local_base + i;
            const unsigned long long gi = ((li / map.tile) * map.stride + map.phase) * map.tile + li % map.tile;
            ctr = slot_offset + gi * spp + (j - i * spp);
        }
        const float4 *a = reinterpret_cast<const float4 *>(attrs + i);
        const float4 q7 = __ldg(a + 7);  // tex_uv, flags, prim
        const uint32_t flags = __float_as_uint(q7.z), prim = __float_as_uint(q7.w);
        float4 ro = make_float4(0.f, 0.f, 0.f, 0.f), rd = make_float4(0.f, 0.f, 0.f, -1.f);  // masked slot
        if (prim != VT_MISS && !(flags & VT_ATTR_HIT_SKY)) {
            const float4 q0 = __ldg(a), q1 = __ldg(a + 1), q2 = __ldg(a + 2), q3 = __ldg(a + 3), q4 = __ldg(a + 4);
            const float sgn = (flags & VT_ATTR_FRONT_FACING) ? 1.f : -1.f;
            const V3 pos = mk3(q0.x, q0.y, q0.z);
            const V3 N = mk3(q1.x, q1.y, q1.z) * sgn, T = mk3(q2.x, q2.y, q2.z), B = mk3(q3.x, q3.y, q3.z) * sgn;
            const V3 gN = mk3(q4.x, q4.y, q4.z) * sgn;
            const float r1 = uniform01(ctr, 0, seed), r2 = uniform01(ctr, 1, seed);
            const float z = sqrtf(r1), sinTheta = sqrtf(1.f - r1), phi = 6.2831853071795864769f * r2;
            float sp, cp;
            sincosf(phi, &sp, &cp);
            const V3 d = T * (sinTheta * cp) + B * (sinTheta * sp) + N * z;
            if (isfinite(d.x) && isfinite(d.y) && isfinite(d.z) && (d.x != 0.f || d.y != 0.f || d.z != 0.f)) {
                ro = make_float4(ray_origin_1(pos.x, gN.x), ray_origin_1(pos.y, gN.y), ray_origin_1(pos.z, gN.z), 0.f);
                rd = make_float4(d.x, d.y, d.z, FLT_MAX);
                spawned = true;
            }
        }
        float4 *o = reinterpret_cast<float4 *>(out + j);
        o[0] = ro;
        o[1] = rd;
    }
    if (live || queue) publish_slot(in_queue ? (j != ~0ull) : (j < total), spawned, j, live, queue, queue_count, miss_hits);
}

// Batched SampleBSDF restricted to the diffuse lobe (source/libraries/BSDF.cpp:770-825 with activeLobes = LobeType::DiffuseReflection):
// what a GLua path tracer calls per hit between two accel:Traverse calls.  Per (hit i, sample s):
//   BSDFMaterial::PrepShadingData(albedo, metalness, roughness)                 BSDF.cpp:11-21  (dielectricInput = 1)
//   CalculateLobePDFs -> pDiffuse = (1 - metallic), normalised                  :23-56
//   lobeSelect = rnd(dim 0); entering = dot(wo, N) >= 0; incident = to_local    :780-783  (anisotropicRotation = 0: rotate() is the identity)
//   SampleDiffuse: hemisphere_cos(rnd(dim 1), rnd(dim 2)), Disney-diffuse weight :252-278, :69-77
//   weight *= (1 - metallic) / pDiffuse, pdf *= pDiffuse, scattered = from_local :788-790, :824
// wo = -normalize(ray direction), the incident direction TraceResult keeps (AccelStruct.cpp:826, TraceResult.cpp:56).  The spawned
// ray starts at CalcRayOrigin(pos, geometric normal on the side the scattered direction leaves through).  Slots that spawn
// nothing (miss, sky, lobeSelect >= pDiffuse — a fully metallic hit — or a non-finite direction) are masked; their sample record
// is what SampleBSDF leaves in BSDFSample then: zero vector, weight 0, pdf 0, lobe None.
__global__ void __launch_bounds__(256)
k_bsdf_diffuse_rays(const vt_ray *__restrict__ rays, const vt_attr *__restrict__ attrs, unsigned long long n, uint32_t spp,
                    unsigned long long seed, vt_ray *__restrict__ out, vt_bsdf_sample *__restrict__ samples,
                    unsigned long long *__restrict__ live, uint32_t *__restrict__ queue, unsigned long long *__restrict__ queue_count,
                    vt_hit *__restrict__ miss_hits) {
    const unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long total = n * spp;
    bool spawned = false;
    if (j < total) {
        const unsigned long long i = j / spp;
        const float4 *a = reinterpret_cast<const float4 *>(attrs + i);
        const float4 q7 = __ldg(a + 7);  // tex_uv, flags, prim
        const uint32_t flags = __float_as_uint(q7.z), prim = __float_as_uint(q7.w);
        float4 ro = make_float4(0.f, 0.f, 0.f, 0.f), rd = make_float4(0.f, 0.f, 0.f, -1.f);  // masked slot
        V3 world = mk3(0.f, 0.f, 0.f), weight = mk3(0.f, 0.f, 0.f);
        float pdf = 0.f;
        uint32_t lobe = VT_LOBE_NONE;
        if (prim != VT_MISS && !(flags & VT_ATTR_HIT_SKY)) {
            const float4 q0 = __ldg(a), q1 = __ldg(a + 1), q2 = __ldg(a + 2), q3 = __ldg(a + 3), q4 = __ldg(a + 4), q5 = __ldg(a + 5);
            const float4 rdir = __ldg(reinterpret_cast<const float4 *>(rays + i) + 1);
            const V3 wo = neg(glm_normalize(mk3(rdir.x, rdir.y, rdir.z)));
            const V3 pos = mk3(q0.x, q0.y, q0.z), N = mk3(q1.x, q1.y, q1.z), T = mk3(q2.x, q2.y, q2.z), B = mk3(q3.x, q3.y, q3.z);
            const V3 gN = mk3(q4.x, q4.y, q4.z);
            const float metallic = glm_clamp(q2.w, 0.f, 1.f), linearRoughness = glm_clamp(q3.w, 0.f, 1.f);
            const V3 dielectric = mk3(glm_clamp(1.f * q5.x, 0.f, 1.f), glm_clamp(1.f * q5.y, 0.f, 1.f), glm_clamp(1.f * q5.z, 0.f, 1.f));
            float pDiffuse = (1.f - metallic) * (1.f - 0.f);
            float normFactor = pDiffuse + 0.f + 0.f + 0.f;
            if (normFactor > 0.f) {
                normFactor = 1.f / normFactor;
                pDiffuse *= normFactor;
            }
            const float lobeSelect = uniform01(j, 0, seed);
            const bool entering = glm_dot(wo, N) >= 0.f;
            const V3 Ns = entering ? N : neg(N);
            const V3 incident = mk3(glm_dot(wo, T), glm_dot(wo, B), glm_dot(wo, Ns));
            V3 scattered = mk3(0.f, 0.f, 0.f);
            if (lobeSelect < pDiffuse) {
                lobe = VT_LOBE_DIFFUSE_REFLECTION;
                const float pi = 3.14159265358979323846264338327950288f;
                const float r1 = uniform01(j, 1, seed);
                const float z = sqrtf(r1), sinTheta = sqrtf(1.f - r1), phi = 2.f * pi * uniform01(j, 2, seed);
                scattered = mk3(sinTheta * cosf(phi), sinTheta * sinf(phi), z);
                pdf = (scattered.z > 0.f) ? (scattered.z / pi) : 0.f;
                const V3 halfway = glm_normalize(incident + scattered);
                const float iDotN = incident.z, sDotH = glm_dot(scattered, halfway), sDotN = scattered.z;
                const float energyBias = glm_lerp(0.f, 0.5f, linearRoughness);
                const float energyFactor = glm_lerp(1.f, 1.f / 1.51f, linearRoughness);
                const float fd90 = energyBias + 2.f * sDotH * sDotH * linearRoughness;
                const float lightScatter = 1.f + (fd90 - 1.f) * powf(1.f - sDotN, 5.f);  // schlick_dielectric(1, cos, f90), BSDF.cpp:157-160
                const float viewScatter = 1.f + (fd90 - 1.f) * powf(1.f - iDotN, 5.f);
                weight = dielectric * lightScatter * viewScatter * energyFactor;
                weight = weight * ((1.f - metallic) * (1.f - 0.f) / pDiffuse);
                pdf *= pDiffuse;
            }
            world = T * scattered.x + B * scattered.y + Ns * scattered.z;
            if (lobe != VT_LOBE_NONE && isfinite(world.x) && isfinite(world.y) && isfinite(world.z) && (world.x != 0.f || world.y != 0.f || world.z != 0.f)) {
                const float side = glm_dot(world, gN) >= 0.f ? 1.f : -1.f;
                ro = make_float4(ray_origin_1(pos.x, gN.x * side), ray_origin_1(pos.y, gN.y * side), ray_origin_1(pos.z, gN.z * side), 0.f);
                rd = make_float4(world.x, world.y, world.z, FLT_MAX);
                spawned = true;
            }
        }
        float4 *o = reinterpret_cast<float4 *>(out + j);
        o[0] = ro;
        o[1] = rd;
        float4 *sm = reinterpret_cast<float4 *>(samples + j);
        sm[0] = make_float4(world.x, world.y, world.z, pdf);
        sm[1] = make_float4(weight.x, weight.y, weight.z, __uint_as_float(lobe));
    }
    if (live || queue) publish_slot(j < total, spawned, j, live, queue, queue_count, miss_hits);
}

// Shadow rays (config 2 / 5: "primary + shadow"): one ray per non-sky hit from CalcRayOrigin(pos, geometric normal on
// the viewer's side) along a fixed direction (a sun) with the given tmax, or toward a point light (dir = light - origin,
// tmax = 1: t is parametric, source/objects/AccelStruct.cpp:810-815).  Misses and sky hits leave masked slots.
__global__ void __launch_bounds__(256)
k_shadow_rays(const vt_attr *__restrict__ attrs, unsigned long long n, float lx, float ly, float lz, int point_light, float tmax,
              vt_ray *__restrict__ out, unsigned long long *__restrict__ live, uint32_t *__restrict__ queue,
              unsigned long long *__restrict__ queue_count, vt_hit *__restrict__ miss_hits, const uint32_t *__restrict__ in_queue,
              const unsigned long long *__restrict__ in_count) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (in_queue) {  // wave compaction: only the parents the previous wave's queue lists
        const unsigned long long total = min(n, *in_count);
        if ((unsigned long long)blockIdx.x * blockDim.x >= total) return;
        i = i < total ? (unsigned long long)__ldg(in_queue + i) : ~0ull;
    }
    bool spawned = false;
    if (i < (in_queue ? ~0ull : n)) {
        const float4 *a = reinterpret_cast<const float4 *>(attrs + i);
        const float4 q7 = __ldg(a + 7);
        const uint32_t flags = __float_as_uint(q7.z), prim = __float_as_uint(q7.w);
        float4 ro = make_float4(0.f, 0.f, 0.f, 0.f), rd = make_float4(0.f, 0.f, 0.f, -1.f);  // masked slot
        if (prim != VT_MISS && !(flags & VT_ATTR_HIT_SKY)) {
            const float4 q0 = __ldg(a), q4 = __ldg(a + 4);
            const float sgn = (flags & VT_ATTR_FRONT_FACING) ? 1.f : -1.f;
            const V3 o = mk3(ray_origin_1(q0.x, q4.x * sgn), ray_origin_1(q0.y, q4.y * sgn), ray_origin_1(q0.z, q4.z * sgn));
            const V3 d = point_light ? mk3(lx - o.x, ly - o.y, lz - o.z) : mk3(lx, ly, lz);
            ro = make_float4(o.x, o.y, o.z, 0.f);
            rd = make_float4(d.x, d.y, d.z, point_light ? 1.f : tmax);
            spawned = true;
        }
        float4 *o4 = reinterpret_cast<float4 *>(out + i);
        o4[0] = ro;
        o4[1] = rd;
    }
    if (live || queue) publish_slot(in_queue ? (i != ~0ull) : (i < n), spawned, i, live, queue, queue_count, miss_hits);
}

// Pinhole primary rays, pixel-centre sampling, row-major (index = width * j + i) — the loop of
// libs/bvh/test/benchmark.cpp:129-150.  cam = {eye, image_u, image_v, dir} (already scaled).
__global__ void __launch_bounds__(256)
k_pinhole_rays(const float *__restrict__ cam, uint32_t width, uint32_t height, vt_ray *__restrict__ out) {
    const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (unsigned long long)width * height) return;
    const uint32_t i = (uint32_t)(idx % width), j = (uint32_t)(idx / width);
    const float u = 2.f * ((float)i + 0.5f) / (float)width - 1.f;
    const float v = 2.f * ((float)j + 0.5f) / (float)height - 1.f;
    const V3 eye = mk3(cam[0], cam[1], cam[2]), iu = mk3(cam[3], cam[4], cam[5]), iv = mk3(cam[6], cam[7], cam[8]),
             dir = mk3(cam[9], cam[10], cam[11]);
    V3 d = iu * u + iv * v + dir;
    d = d * (1.0f / sqrtf(bvh_dot(d, d)));  // bvh::normalize, vector.hpp:149-153
    float4 *o = reinterpret_cast<float4 *>(out + idx);
    o[0] = make_float4(eye.x, eye.y, eye.z, 0.f);
    o[1] = make_float4(d.x, d.y, d.z, FLT_MAX);
}

// K4 — fold one diffuse wave into an RGBFFF framebuffer (tightly packed 3 x f32 per pixel, the layout
// of IRenderTarget format RGBFFF: include/vistrace/IRenderTarget.h:40, source/objects/RenderTarget.cpp:61-97):
//   fb[i] += weight * albedo_i * (number of the spp bounce rays of pixel i that escape to the sky) / spp
// A bounce ray "escapes" when it misses everything or its closest hit is a sky brush
// (TraceResult::hitSky, source/objects/TraceResult.cpp:83).  This is the per-rank partial image that the
// multi-GPU path sums with ONE collective; it is harness-level shading, not part of accel:Traverse.
__global__ void __launch_bounds__(256)
k_accumulate_sky(const VtSceneView S, const vt_attr *__restrict__ attrs, const vt_hit *__restrict__ bounce_hits,
                 unsigned long long n, uint32_t spp, float weight, float *__restrict__ fb, const VtSlotMap map, int overwrite) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // where pixel i lands: slot i of fb, or — a shard writing into the FRAME (multi-GPU, vt_group.cu) — its global pixel: the
    // frame may be another GPU's memory mapped over NVLink, so the result of the shard is delivered by these very stores
    unsigned long long o = i;
    if (map.tile) {
        const unsigned long long li = map.local_base + i;
        o = ((li / map.tile) * map.stride + map.phase) * map.tile + li % map.tile;
    }
    const float4 *a = reinterpret_cast<const float4 *>(attrs + i);
    const float4 q7 = __ldg(a + 7);
    const uint32_t flags = __float_as_uint(q7.z), prim = __float_as_uint(q7.w);
    float r = 0.f, g = 0.f, b = 0.f;
    if (prim != VT_MISS) {
        const float4 q5 = __ldg(a + 5);  // albedo, ent_id
        if (flags & VT_ATTR_HIT_SKY) {
            r = q5.x, g = q5.y, b = q5.z;  // looking straight at the sky
        } else {
            uint32_t escaped = 0;
            for (uint32_t s = 0; s < spp; s++) {
                const float4 h = __ldg(reinterpret_cast<const float4 *>(bounce_hits) + i * spp + s);
                const uint32_t bp = __float_as_uint(h.w);
                if (bp == VT_MISS || bp >= S.n_tris) {
                    escaped++;
                } else {
                    const uint32_t m = S.attrs[bp].material;
                    escaped += (S.mats[m].surf_flags & VT_SURF_SKY) ? 1u : 0u;
                }
            }
            const float vis = (float)escaped / (float)spp;
            r = q5.x * vis, g = q5.y * vis, b = q5.z * vis;
        }
    }
    if (overwrite) {
        fb[3 * o + 0] = 0.f + weight * r;  // the value an accumulation into a zeroed buffer gives, bit for bit
        fb[3 * o + 1] = 0.f + weight * g;
        fb[3 * o + 2] = 0.f + weight * b;
    } else {
        fb[3 * o + 0] += weight * r;
        fb[3 * o + 1] += weight * g;
        fb[3 * o + 2] += weight * b;
    }
}

// Cross-GPU hand-shake of the peer-memory frame (vt_group.cu): a flag word in the frame owner's memory carries the step number.
// k_flag_set publishes "everything this stream did before is done" (the stores of the preceding kernels are complete at the kernel
// boundary; the fences order the flag behind them system-wide); k_flag_wait holds a stream until all `count` flags reached `step`.
__global__ void k_flag_set(volatile uint32_t *flag, uint32_t step) {
    __threadfence_system();
    *flag = step;
    __threadfence_system();
}
__global__ void k_flag_wait(const volatile uint32_t *flags, uint32_t count, uint32_t stride, uint32_t step) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (uint32_t k = threadIdx.x; k < count; k += blockDim.x)
        while ((int32_t)(flags[(size_t)k * stride] - step) < 0) {
            __nanosleep(200);
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t - t0 > 20000000000ull) __trap();  // a peer died or never called the collective: fail loudly after 20 s instead of hanging the GPU
        }
    __threadfence_system();
}

// Path shading between two waves of vt_accel_trace_paths (harness-level, like K4: it exists so that a multi-bounce workload has
// an image to deliver; not part of accel:Traverse).  One thread per LIVE path vertex of the wave — the slots the wave's queue
// lists, or all n slots of the primary wave (queue == nullptr).  Slot = pixel, so no two threads touch the same pixel:
//   sky hit (TraceResult::hitSky, source/objects/TraceResult.cpp:83):  fb += weight * throughput * albedo, the path ends;
//   surface hit:  throughput *= albedo;  if the vertex's shadow ray reached the sun (any-hit miss)  fb += weight * throughput * sun.
__global__ void __launch_bounds__(256)
k_path_shade(const vt_attr *__restrict__ attrs, const vt_hit *__restrict__ shadow_hits, const uint32_t *__restrict__ queue,
             const unsigned long long *__restrict__ queue_count, unsigned long long n, int first_wave, float weight, float sun_r,
             float sun_g, float sun_b, float *__restrict__ throughput, float *__restrict__ fb) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (queue) {
        if (i >= min(n, *queue_count)) return;
        i = __ldg(queue + i);
    } else if (i >= n) {
        return;
    }
    const float4 *a = reinterpret_cast<const float4 *>(attrs + i);
    const float4 q7 = __ldg(a + 7);
    const uint32_t flags = __float_as_uint(q7.z), prim = __float_as_uint(q7.w);
    if (prim == VT_MISS) return;
    const float4 q5 = __ldg(a + 5);  // albedo, ent_id
    float tr = 1.f, tg = 1.f, tb = 1.f;
    if (!first_wave) tr = throughput[3 * i], tg = throughput[3 * i + 1], tb = throughput[3 * i + 2];
    tr *= q5.x, tg *= q5.y, tb *= q5.z;
    float r = 0.f, g = 0.f, b = 0.f;
    if (flags & VT_ATTR_HIT_SKY) {
        r = tr, g = tg, b = tb;
    } else {
        throughput[3 * i] = tr, throughput[3 * i + 1] = tg, throughput[3 * i + 2] = tb;
        const uint32_t occluder = __float_as_uint(__ldg(reinterpret_cast<const float4 *>(shadow_hits) + i).w);
        if (occluder == VT_MISS) r = tr * sun_r, g = tg * sun_g, b = tb * sun_b;
    }
    fb[3 * i] += weight * r;
    fb[3 * i + 1] += weight * g;
    fb[3 * i + 2] += weight * b;
}

}  // namespace

cudaError_t vt_launch_path_shade(const vt_attr *attrs, const vt_hit *shadow_hits, const uint32_t *queue, const unsigned long long *queue_count,
                                 uint64_t n, bool first_wave, float weight, const float sun_rgb[3], float *throughput, float *fb, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    if ((queue == nullptr) != (queue_count == nullptr)) return cudaErrorInvalidValue;
    const unsigned block = 256;
    k_path_shade<<<(unsigned)((n + block - 1) / block), block, 0, stream>>>(attrs, shadow_hits, queue, queue_count, n, first_wave ? 1 : 0, weight,
                                                                           sun_rgb[0], sun_rgb[1], sun_rgb[2], throughput, fb);
    return cudaGetLastError();
}

cudaError_t vt_launch_accumulate_sky(const VtSceneView &S, const vt_attr *attrs, const vt_hit *bounce_hits, uint64_t n,
                                     uint32_t spp, float weight, float *fb, cudaStream_t stream, const VtSlotMap *map, bool overwrite) {
    if (n == 0) return cudaSuccess;
    const unsigned block = 256;
    k_accumulate_sky<<<(unsigned)((n + block - 1) / block), block, 0, stream>>>(S, attrs, bounce_hits, n, spp, weight, fb, map ? *map : VtSlotMap(),
                                                                               overwrite ? 1 : 0);
    return cudaGetLastError();
}

cudaError_t vt_launch_flag_set(uint32_t *flag, uint32_t step, cudaStream_t stream) {
    k_flag_set<<<1, 1, 0, stream>>>(flag, step);
    return cudaGetLastError();
}

cudaError_t vt_launch_flag_wait(const uint32_t *flags, uint32_t count, uint32_t stride, uint32_t step, cudaStream_t stream) {
    if (count == 0) return cudaSuccess;
    k_flag_wait<<<1, 32, 0, stream>>>(flags, count, stride, step);
    return cudaGetLastError();
}

cudaError_t vt_launch_bounce_rays(const vt_attr *attrs, uint64_t n, uint32_t spp, uint64_t seed, uint64_t slot_offset,
                                  vt_ray *out, unsigned long long *live, cudaStream_t stream, uint32_t *queue,
                                  unsigned long long *queue_count, vt_hit *miss_hits, const VtSlotMap *map, const uint32_t *in_queue,
                                  const unsigned long long *in_count) {
    const unsigned long long total = (unsigned long long)n * spp;
    if (total == 0) return cudaSuccess;
    if (queue && (!queue_count || !miss_hits || total > 0xFFFFFFFFull)) return cudaErrorInvalidValue;
    if ((in_queue == nullptr) != (in_count == nullptr) || (in_queue && !queue)) return cudaErrorInvalidValue;
    const unsigned block = 256;
    k_bounce_rays<<<(unsigned)((total + block - 1) / block), block, 0, stream>>>(attrs, n, spp, seed, slot_offset, out, live, queue,
                                                                                queue_count, miss_hits, map ? *map : VtSlotMap(), in_queue, in_count);
    return cudaGetLastError();
}

cudaError_t vt_launch_bsdf_diffuse_rays(const vt_ray *rays, const vt_attr *attrs, uint64_t n, uint32_t spp, uint64_t seed, vt_ray *out,
                                        vt_bsdf_sample *samples, unsigned long long *live, cudaStream_t stream, uint32_t *queue,
                                        unsigned long long *queue_count, vt_hit *miss_hits) {
    const unsigned long long total = (unsigned long long)n * spp;
    if (total == 0) return cudaSuccess;
    if (queue && (!queue_count || !miss_hits || total > 0xFFFFFFFFull)) return cudaErrorInvalidValue;
    const unsigned block = 256;
    k_bsdf_diffuse_rays<<<(unsigned)((total + block - 1) / block), block, 0, stream>>>(rays, attrs, n, spp, seed, out, samples, live, queue,
                                                                                      queue_count, miss_hits);
    return cudaGetLastError();
}

cudaError_t vt_launch_shadow_rays(const vt_attr *attrs, uint64_t n, const float light[3], bool point_light, float tmax, vt_ray *out,
                                  unsigned long long *live, cudaStream_t stream, uint32_t *queue, unsigned long long *queue_count,
                                  vt_hit *miss_hits, const uint32_t *in_queue, const unsigned long long *in_count) {
    if (n == 0) return cudaSuccess;
    if (queue && (!queue_count || !miss_hits || n > 0xFFFFFFFFull)) return cudaErrorInvalidValue;
    if ((in_queue == nullptr) != (in_count == nullptr) || (in_queue && !queue)) return cudaErrorInvalidValue;
    const unsigned block = 256;
    k_shadow_rays<<<(unsigned)((n + block - 1) / block), block, 0, stream>>>(attrs, n, light[0], light[1], light[2], point_light ? 1 : 0,
                                                                            tmax, out, live, queue, queue_count, miss_hits, in_queue, in_count);
    return cudaGetLastError();
}

cudaError_t vt_launch_pinhole_rays(const float *cam12, uint32_t width, uint32_t height, vt_ray *out, cudaStream_t stream) {
    const unsigned long long total = (unsigned long long)width * height;
    if (total == 0) return cudaSuccess;
    const unsigned block = 256;
    k_pinhole_rays<<<(unsigned)((total + block - 1) / block), block, 0, stream>>>(cam12, width, height, out);
    return cudaGetLastError();
}
