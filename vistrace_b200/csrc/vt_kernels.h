// vt_kernels.h — launch interface between the host objects and the CUDA kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vistrace_b200.h"
#include "vt_device.h"

#ifndef VT_TRAVERSE_BLOCK
#define VT_TRAVERSE_BLOCK 128
#endif
#ifndef VT_TRAVERSE_MIN_BLOCKS
#define VT_TRAVERSE_MIN_BLOCKS 8
#endif

// A/B builds only: the quad kernel stages the first S.n_smem_pairs quads (breadth-first prefix, quads_top_first) in shared memory
#ifndef VT_SMEM_QUADS_BUILD
#define VT_SMEM_QUADS_BUILD 0
#endif
#ifndef VT_COMPACT_MIN_BLOCKS
#define VT_COMPACT_MIN_BLOCKS 9
#endif

struct VtLaunchConfig {
    int persistent = 1;         // 1: machine-sized grid pulling rays from a counter; 0: one ray per thread
    int grid = 0;               // CTAs for the persistent launch (SMs x resident CTAs)
    int refill_threshold = 24;  // refill a warp when <= this many of its lanes still own a ray
    int tri_threshold = 10;     // run a triangle round when >= this many lanes have a candidate queued (6 / 8 / 10: 3.39 / 3.47 / 3.51 Grays/s)
    int key_mid = 0;            // quad layout: order a node's children by entry + exit instead of entry (vt_traverse.cu: VT_KEY_MID); set per scene
    int tail_share = 1;         // quantised layouts, closest hit: idle lanes take pending sub-trees of long rays once the ray queue is dry (VT_TAIL_SHARE)
};

// K1 — closest hit (or any hit) for n rays.  counters[0] = ray queue head (must be 0 on entry),
// counters[1] += rays rejected by the argument rules; with stats: counters[2] += traversal steps,
// counters[3] += triangle tests (SingleRayTraverser::Statistics, single_ray_traverser.hpp:132-135).
// Ray queue (both or neither): queue[0, min(n, *queue_count)) names the slots of rays/hits to trace — what a
// generator (K3) listed as live; every other slot is left untouched.
cudaError_t vt_launch_traverse(const VtSceneView &S, const vt_ray *rays, vt_hit *hits, uint64_t n, bool any_hit,
                               unsigned long long *counters, const VtLaunchConfig &cfg, cudaStream_t stream, bool stats = false,
                               const uint32_t *queue = nullptr, const unsigned long long *queue_count = nullptr);
cudaError_t vt_traverse_occupancy(int *blocks_per_sm, size_t smem_bytes, int layout);

// K2 — eager TraceResult for n (ray, hit) records; cones = n x {coneWidth, coneAngle} or nullptr.
// Queue (both or neither): only the slots queue[0, min(n, *queue_count)) get a record — wave compaction for multi-bounce paths.
cudaError_t vt_launch_trace_result(const VtSceneView &S, const vt_ray *rays, const vt_hit *hits, const float *cones,
                                   vt_attr *attrs, uint64_t n, cudaStream_t stream, const uint32_t *queue = nullptr,
                                   const unsigned long long *queue_count = nullptr);

// K3 — secondary-ray generation: spp cosine-weighted bounce rays per (non-sky) hit into slot i*spp+s,
// masked slots (tmax < 0) elsewhere; *live += spawned rays.  And a pinhole primary-ray generator.
// Ray queue (optional, all three or none): queue[pos] = slot of every ray spawned, pos handed out from *queue_count
// (must be 0 on entry), and miss_hits[slot] = a miss record for every masked slot — so the traversal that follows
// only has to visit the queue and the hit buffer is complete in slot order all the same.
// Pixel map of a SHARD of a frame (multi-GPU, vt_group.cu): the frame is cut into tiles of `tile` pixels, the shard owns every
// `stride`-th tile starting at tile `phase`, stored compactly; local pixel i (after adding local_base, the offset of the batch
// inside the shard) is global pixel ((i / tile) * stride + phase) * tile + i % tile.  The random-number counter of a bounce ray is
// taken from the GLOBAL pixel, so a frame traced in shards draws the same numbers — and gives the same image bit for bit — as the
// frame traced whole.  tile == 0: identity.
struct VtSlotMap {
    unsigned long long local_base = 0;
    unsigned long long tile = 0;
    uint32_t stride = 1, phase = 0;
};
cudaError_t vt_launch_bounce_rays(const vt_attr *attrs, uint64_t n, uint32_t spp, uint64_t seed, uint64_t slot_offset,
                                  vt_ray *out, unsigned long long *live, cudaStream_t stream, uint32_t *queue = nullptr,
                                  unsigned long long *queue_count = nullptr, vt_hit *miss_hits = nullptr, const VtSlotMap *map = nullptr,
                                  const uint32_t *in_queue = nullptr, const unsigned long long *in_count = nullptr);
// in_queue / in_count (generators, both or neither): wave compaction — the generator visits only the parents the PREVIOUS wave's
// queue lists (attrs[in_queue[k]], k < min(n, *in_count)) instead of all n slots; slots it does not visit are not written at all
// (no masked ray, no miss record), so everything downstream must go through the out queue.
cudaError_t vt_launch_shadow_rays(const vt_attr *attrs, uint64_t n, const float light[3], bool point_light, float tmax, vt_ray *out,
                                  unsigned long long *live, cudaStream_t stream, uint32_t *queue = nullptr,
                                  unsigned long long *queue_count = nullptr, vt_hit *miss_hits = nullptr,
                                  const uint32_t *in_queue = nullptr, const unsigned long long *in_count = nullptr);
// K3c — batched SampleBSDF, diffuse lobe (source/libraries/BSDF.cpp:770-825): spp samples per non-sky hit; out rays + BSDFSample records.
cudaError_t vt_launch_bsdf_diffuse_rays(const vt_ray *rays, const vt_attr *attrs, uint64_t n, uint32_t spp, uint64_t seed, vt_ray *out,
                                        vt_bsdf_sample *samples, unsigned long long *live, cudaStream_t stream, uint32_t *queue = nullptr,
                                        unsigned long long *queue_count = nullptr, vt_hit *miss_hits = nullptr);
cudaError_t vt_launch_pinhole_rays(const float *cam12, uint32_t width, uint32_t height, vt_ray *out, cudaStream_t stream);

// K4 — fb[i] += weight * albedo_i * (escaped bounce rays of pixel i) / spp, RGBFFF framebuffer.
// map (optional): pixel i of the batch is written to the GLOBAL pixel the shard map gives (VtSlotMap) — fb is then the frame, which
// may live on another GPU (peer memory over NVLink); overwrite: store instead of accumulate (same value as += into a zeroed buffer).
cudaError_t vt_launch_accumulate_sky(const VtSceneView &S, const vt_attr *attrs, const vt_hit *bounce_hits, uint64_t n,
                                     uint32_t spp, float weight, float *fb, cudaStream_t stream, const VtSlotMap *map = nullptr,
                                     bool overwrite = false);
// Flag words for the cross-GPU hand-shake of the peer-memory frame: set after everything enqueued before on `stream`; wait until all
// flags[k * stride], k < count, have reached `step` (wrap-around safe).
cudaError_t vt_launch_flag_set(uint32_t *flag, uint32_t step, cudaStream_t stream);
cudaError_t vt_launch_flag_wait(const uint32_t *flags, uint32_t count, uint32_t stride, uint32_t step, cudaStream_t stream);

// Path shading between the waves of vt_accel_trace_paths: one thread per live vertex (queue-listed slots, or all n when queue is null).
cudaError_t vt_launch_path_shade(const vt_attr *attrs, const vt_hit *shadow_hits, const uint32_t *queue, const unsigned long long *queue_count,
                                 uint64_t n, bool first_wave, float weight, const float sun_rgb[3], float *throughput, float *fb, cudaStream_t stream);

// K5 — device-side refit of the resident quad hierarchy (vt_refit.cu).  prepare: once per resident hierarchy
// (parent / inner-child count per quad, original triangle -> leaf slot).  refit_tris: Triangle constructor over the
// caller's new vertices into the resident triangle / attribute records.  refit_quads: bottom-up boxes + requantisation;
// qbox = n_quads x 24 bytes of scratch, *error != 0 afterwards when a box could not be held on the float grid.
cudaError_t vt_launch_refit_prepare(const VtSceneView &S, uint32_t *parent, uint32_t *n_inner, uint32_t *slot_of, uint32_t *leaf_quad,
                                    cudaStream_t stream);
cudaError_t vt_launch_refit_tris(const VtSceneView &S, const vt_tri_in *in, uint32_t first, uint32_t count, const uint32_t *slot_of,
                                 cudaStream_t stream);  // in[j] = new vertices of original triangle first + j
// *sum = sum of the half surface areas of the n_quads boxes in qbox (as left by vt_launch_refit_quads): the rebuild trigger's measure.
cudaError_t vt_launch_refit_cost(const void *qbox, uint32_t n_quads, double *sum, cudaStream_t stream);
cudaError_t vt_launch_refit_quads(const VtSceneView &S, const uint32_t *parent, const uint32_t *n_inner, uint32_t *arrive, void *qbox,
                                  unsigned int *error, cudaStream_t stream);
// The same for original triangles [first, first + count) only: the quads that hold them and their ancestors.  stamp / kids / arrive:
// n_quads words each, zero before the first call and left zero by every call; epoch: odd, > 0, += 3 per call.  qbox must hold the
// boxes of a previous vt_launch_refit_quads pass (untouched children are read from it).
cudaError_t vt_launch_refit_quads_range(const VtSceneView &S, uint32_t first, uint32_t count, const uint32_t *slot_of, const uint32_t *leaf_quad,
                                        const uint32_t *parent, uint32_t *stamp, uint32_t *kids, uint32_t *arrive, void *qbox, uint32_t epoch,
                                        unsigned int *error, cudaStream_t stream);
