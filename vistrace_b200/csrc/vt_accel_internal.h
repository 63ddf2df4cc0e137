// vt_accel_internal.h — shared by vt_accel.cu and vt_group.cu: the HBM-resident scene of one AccelStruct and small helpers.
// Not part of the boundary (include/vistrace_b200.h is); nothing outside vistrace_b200/csrc includes this.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdlib>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "vt_host.h"
#include "vt_kernels.h"

namespace vt {

extern thread_local std::string g_last_error;

#define VT_CUDA(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " #expr); \
    } while (0)

inline int env_int(const char *name, int def) {
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : def;
}
inline float env_float(const char *name, float def) {
    const char *v = std::getenv(name);
    return (v && *v) ? (float)std::atof(v) : def;
}

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;  // elements
    void ensure(size_t n) {
        if (n <= cap) return;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        VT_CUDA(cudaMalloc(&p, n * sizeof(T)));
        cap = n;
    }
    void upload(const T *src, size_t n) {
        ensure(n ? n : 1);
        if (n) VT_CUDA(cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice));
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    size_t bytes() const { return cap * sizeof(T); }
};

constexpr int kCounterSlots = 256;  // 16-byte {queue head, invalid rays} records, one per in-flight call

struct DeviceScene {
    DevBuf<VtPair> pairs;    // exact layout  \ one of the two is resident
    DevBuf<VtCPair> cpairs;  // compact layout  | exactly one of the three is resident
    DevBuf<VtQuad> quads;    // quad layout    /
    DevBuf<VtTriRec> tris;
    DevBuf<float> tri_uv;
    DevBuf<VtTriAttr> attrs;
    DevBuf<VtDevMaterial> mats;
    DevBuf<VtDevEntity> ents;
    DevBuf<VtDevTexture> texs;
    DevBuf<uint8_t> texels;
    DevBuf<unsigned long long> counters;
    DevBuf<unsigned long long> stat_counters;
    // staging for host-pointer calls
    DevBuf<vt_ray> s_rays;
    DevBuf<vt_hit> s_hits;
    DevBuf<uint32_t> s_ray_stats;  // vt_accel_traverse_ray_stats
    DevBuf<vt_attr> s_attrs;
    DevBuf<float> s_cones;
    DevBuf<vt_bsdf_sample> s_samples;
    DevBuf<vt_ray> s_rays2;
    // per-stream tile staging of the host-pointer diffuse wave
    struct WaveLane {
        cudaStream_t stream = nullptr;
        DevBuf<vt_ray> brays;
        DevBuf<vt_hit> hits, bhits;
        DevBuf<vt_attr> attrs;
        DevBuf<float> fb;
        DevBuf<uint32_t> queue;                // live bounce slots of the tile (ray queue)
        DevBuf<unsigned long long> queue_count;
    } lanes[8];  // VT_WAVE_LANES of them are used (default 4: measured 3.26 / 3.08 / 3.15 / 3.16 ms per e2e step with 3 / 4 / 5 / 6)
    // ray-queue scratch of the device-pointer wave, one per caller stream
    struct WaveScratch {
        DevBuf<uint32_t> queue;
        DevBuf<unsigned long long> queue_count;
    };
    std::map<cudaStream_t, WaveScratch> wave_scratch;
    std::mutex wave_mutex;
    // scratch of vt_accel_trace_paths (multi-bounce path waves with compaction): two generations of hit / TraceResult records,
    // the secondary rays of a wave, three slot queues + their counters, per-pixel path throughput
    struct PathScratch {
        DevBuf<vt_hit> hits[2], shits;
        DevBuf<vt_attr> attrs[2];
        DevBuf<vt_ray> brays, srays;
        DevBuf<uint32_t> queue[3];
        DevBuf<unsigned long long> counts;  // [0..2]: the three queue counters, [8 + 2k], [9 + 2k]: rays of wave k (bounce, shadow)
        DevBuf<float> throughput;
    } path[2];  // two: samples traced on two streams at once (VT_PATHS_SLOT1) share nothing
    DevBuf<unsigned long long> live;
    // host-pointer waves: the whole frame's rays are staged here by ONE copy stream, tile after tile, so an upload never
    // waits for the lane (stream) its tile will run on; upload_done[k] gates tile k's kernels
    DevBuf<vt_ray> wave_rays, wave_rays_b;  // two: consecutive frames in flight alternate (AccelStruct::RenderDiffuseWaveBegin)
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> upload_done;
    // K5 (device refit) state, built on the first refit of a resident quad hierarchy
    DevBuf<vt_tri_in> refit_in;
    DevBuf<uint32_t> refit_parent, refit_n_inner, refit_slot_of, refit_arrive, refit_error;
    DevBuf<uint32_t> refit_leaf_quad;    // leaf slot -> the quad that holds the leaf (ranged refit starts there)
    DevBuf<uint32_t> refit_range_state;  // 3 x n_quads, all zero between calls: epoch stamps | dirty-child counts | arrivals
    uint32_t refit_epoch = 1;            // += 3 per vt_accel_refit_range
    DevBuf<float> refit_qbox;  // 6 floats per quad
    DevBuf<double> refit_cost;  // device scalar of k_refit_cost
    double refit_cost_built = 0.0, refit_cost_now = 0.0;  // node-area sums: as built (taken when the refit state is prepared) / after the last refit
    bool refit_ready = false;
    VtSceneView view{};
    VtLaunchConfig cfg;
    std::atomic<uint32_t> next_slot{0};
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;

    ~DeviceScene() {
        pairs.release();
        cpairs.release();
        quads.release();
        tris.release();
        tri_uv.release();
        attrs.release();
        mats.release();
        ents.release();
        texs.release();
        texels.release();
        counters.release();
        stat_counters.release();
        s_rays.release();
        s_hits.release();
        s_attrs.release();
        s_cones.release();
        s_samples.release();
        s_rays2.release();
        live.release();
        for (PathScratch &ps : path) {
            for (int i = 0; i < 2; i++) ps.hits[i].release(), ps.attrs[i].release();
            for (int i = 0; i < 3; i++) ps.queue[i].release();
            ps.shits.release(), ps.brays.release(), ps.srays.release(), ps.counts.release(), ps.throughput.release();
        }
        wave_rays.release();
        wave_rays_b.release();
        for (cudaEvent_t e : upload_done) cudaEventDestroy(e);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        refit_in.release();
        refit_parent.release();
        refit_n_inner.release();
        refit_slot_of.release();
        refit_leaf_quad.release();
        refit_range_state.release();
        refit_arrive.release();
        refit_error.release();
        refit_qbox.release();
        refit_cost.release();
        for (auto &l : lanes) {
            l.brays.release();
            l.hits.release();
            l.bhits.release();
            l.attrs.release();
            l.fb.release();
            l.queue.release();
            l.queue_count.release();
            if (l.stream) cudaStreamDestroy(l.stream);
        }
        for (auto &kv : wave_scratch) {
            kv.second.queue.release();
            kv.second.queue_count.release();
        }
        if (own_stream) cudaStreamDestroy(own_stream);
    }
    uint64_t scene_bytes() const {
        return pairs.bytes() + cpairs.bytes() + quads.bytes() + tris.bytes() + tri_uv.bytes() + attrs.bytes() + mats.bytes() + ents.bytes() + texs.bytes() +
               texels.bytes();
    }
};

}  // namespace vt

// the opaque handle of include/vistrace_b200.h
struct vt_accel {
    vt::AccelStruct impl;
    explicit vt_accel(int device) : impl(device) {}
};

#define VT_TRY try {
#define VT_CATCH(ret)                     \
    }                                     \
    catch (const std::exception &e) {     \
        vt::g_last_error = e.what();      \
        return ret;                       \
    }                                     \
    catch (...) {                         \
        vt::g_last_error = "unknown error"; \
        return ret;                       \
    }
