// vt_host.h — host-side C++ objects: the mirror of the reference's AccelStruct / Triangle /
// TraceResult for the accel:Traverse path.  Same names, argument meaning and error behaviour as
// source/objects/AccelStruct.h:61-86, Primitives.h:43-217 and TraceResult.h:13-111, minus the Lua
// and game-engine ingestion (which needs a running game) — replaced by the headless Populate().
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <thread>
#include <memory>
#include <mutex>
#include <type_traits>
#include <utility>
#include <deque>
#include <vector>

#include "../../include/vistrace_b200.h"
#include "vt_device.h"

struct CUevent_st;  // cudaEvent_t is `CUevent_st *`; this header is also read by translation units without the CUDA headers

namespace vt {

// Allocator whose value-construction is DEFAULT-initialisation: vector<T>(n) / resize(n) of a trivially constructible T touches
// no memory.  The containers of a 5 M-triangle scene are ~3 GB; zero-filling them on one thread before the parallel loops
// overwrite every byte was a quarter of the rebuild latency, and it put every page on the NUMA node of that one thread.
template <class T>
struct DefaultInitAllocator : std::allocator<T> {
    template <class U>
    struct rebind {
        using other = DefaultInitAllocator<U>;
    };
    DefaultInitAllocator() = default;
    template <class U>
    DefaultInitAllocator(const DefaultInitAllocator<U> &) {}
    template <class U>
    void construct(U *p) noexcept(std::is_nothrow_default_constructible<U>::value) {
        ::new (static_cast<void *>(p)) U;
    }
    template <class U, class... Args>
    void construct(U *p, Args &&...args) {
        ::new (static_cast<void *>(p)) U(std::forward<Args>(args)...);
    }
};
template <class T>
using RawVector = std::vector<T, DefaultInitAllocator<T>>;

// TriangleBackfaceCull<float> (source/objects/Primitives.h:43-73) without the skinning scratch data.  Trivially default
// constructible on purpose (see DefaultInitAllocator): the six-argument constructor sets every field.
struct Triangle {
    float p0[3], e1[3], e2[3], n[3], nNorm[3];
    bool oneSided;
    uint32_t material;
    uint16_t entIdx;
    float normals[3][3];
    float tangents[3][3];
    float uvs[3][2];
    float alphas[3];
    float lod;

    Triangle() = default;
    // Triangle(p0, p1, p2, material, uvs, oneSided) — Primitives.h:75-89
    Triangle(const float p0_[3], const float p1[3], const float p2[3], uint32_t material_, const float uvs_[3][2],
             bool oneSided_ = false);
    void ComputeNormalAndLoD();  // Primitives.h:91-102
};

using TriangleVec = RawVector<Triangle>;
using Entity = vt_entity;      // source/objects/AccelStruct.h:33-40 (id, colour)
using Material = vt_material;  // source/objects/Material.h:74-125 (subset on the path)

struct HostBvh {  // bvh::Bvh<float>: nodes, primitive_indices, node_count (bvh.hpp:96-99)
    RawVector<vt_node> nodes;
    RawVector<uint64_t> prim_indices;
};

struct FlatBvh {
    std::vector<VtPair> pairs;
    std::vector<uint32_t> leaf_order;  // leaf-order slot -> original triangle index
    uint32_t root_leaf_count = 0;
    uint32_t max_depth = 0;
};

// sweep_below: builder-quality option — nodes of at most that many primitives are split by the exact SAH sweep over sorted
// centroids (bvh::SweepSahBuilder's evaluation, libs/bvh/include/bvh/sweep_sah_builder.hpp) instead of 16 bins; 0 = binned only.
void build_bvh(const TriangleVec &tris, HostBvh &out, int max_leaf, float trav_cost, uint32_t sweep_below = 0);
// Builder-quality option (vt_bvh_reinsert.cpp): reinsertion optimisation of a finished hierarchy, the product's version of
// bvh::ParallelReinsertionOptimizer (libs/bvh/include/bvh/parallel_reinsertion_optimizer.hpp).  `iterations` passes, each moving
// up to ~`fraction` of the nodes (largest boxes first) to the position where the sum of inner-node areas grows least; leaves keep
// their primitive ranges.  The tree is left as it was when the result would be deeper than 60 or no better.  false: malformed tree.
bool reinsert_optimize(HostBvh &bvh, int iterations, float fraction = 0.05f, double *area_before = nullptr, double *area_after = nullptr,
                       uint64_t *moves = nullptr);
// The reference's own hierarchy rebuilt from its algorithm (vt_bvh_ploc.cpp): bvh::LocallyOrderedClusteringBuilder<BVH, uint32_t>
// (search radius 14, 30-bit Morton codes) and bvh::LeafCollapser, the sequence of source/objects/AccelStruct.cpp:762-770.
void build_bvh_ploc(const TriangleVec &tris, HostBvh &out);
bool collapse_leaves(HostBvh &bvh);
// bvh::HierarchyRefitter over moved geometry of unchanged topology (hierarchy_refitter.hpp:20-31): node boxes only.
bool refit_bvh(const TriangleVec &tris, HostBvh &bvh, std::string &err);
bool flatten_bvh(const HostBvh &bvh, uint64_t n_tris, uint32_t bfs_pairs, FlatBvh &out, std::string &err);
// bvh::Bvh<float> form -> 4-wide quantised nodes (vt_device.h: VtQuad) in depth-first order + the leaf-order
// permutation of the triangles.  False (with a reason) when the tree cannot be held: a leaf of more than 15
// triangles, non-finite bounds, or a worst-case traversal stack deeper than VT_STACK_SIZE.
struct QuadBvh {
    RawVector<VtQuad> quads;
    RawVector<uint32_t> leaf_order;
    uint32_t root_leaf_count = 0;
    uint32_t max_stack = 0;  // worst-case number of pending references
    double sibling_overlap = 0.0;  // of the binary tree (vt_bvh_collapse.cpp): what the collapse rule and the kernel's child order are chosen by
    bool greedy_collapse = false;
};
bool build_quads(const HostBvh &bvh, uint64_t n_tris, QuadBvh &out, std::string &err);
uint32_t quads_top_first(QuadBvh &qb, uint32_t top);  // experiment: breadth-first prefix for shared-memory staging (VT_SMEM_QUADS)
// Which binary nodes become the children of each WIDE node when the binary tree is collapsed to `width` children per node
// (vt_bvh_collapse.cpp): the SAH-optimal choice by dynamic programming over the subtree costs, after Ylitie, Karras, Laine,
// "Efficient Incoherent Ray Traversal on GPUs Through Compressed Wide BVHs" (HPG 2017), section 3.1 — cost(n, i) = cheapest way to
// present subtree n as at most i roots — instead of greedily adopting the children of the largest child, which leaves wide
// nodes partly empty near the leaves (measured at 320 k triangles: 2.98 -> 3.46 of 4 quad slots used, 20 % fewer quads).  VT_COLLAPSE=greedy selects the old rule.
struct CollapsePlan {
    int width = 0;
    std::vector<uint8_t> split;  // [node * width + i]: how many of i + 1 roots go to the left child (0: use one root fewer)
    bool greedy = false;
    double sibling_overlap = 0.0;  // of the binary tree the plan was made for (what VT_COLLAPSE=auto decided on)
    // children of the wide node that replaces binary inner node `ni` -> kids[0 .. return value)
    int children(const HostBvh &bvh, uint32_t ni, uint32_t *kids) const;
};
bool plan_collapse(const HostBvh &bvh, int width, CollapsePlan &plan, std::string &err);
double sibling_overlap(const HostBvh &bvh);  // area-weighted overlap of sibling boxes, 0 .. ~0.5 (vt_bvh_collapse.cpp)
// 64-byte pairs in depth-first order -> 32-byte conservative compact pairs (vt_device.h: VtCPair)
bool compact_pairs(const std::vector<VtPair> &pairs, std::vector<VtCPair> &out, std::string &err);

// SkinTriangle over a batch (source/objects/AccelStruct.cpp:66-108): bakes bone * bind * weight into world space.
void SkinTriangles(vt_tri_in *tris, const vt_tri_skin *skin, uint64_t n, const float *bones, const float *binds, uint32_t n_bones);

// VTF file -> RGBA8888 mip chain (vt_vtf.cpp; libs/VTFParser restated).  Throw std::runtime_error on malformed / unsupported files.
void VtfInfo(const uint8_t *file, uint64_t size, vt_vtf_info *out);
void VtfDecode(const uint8_t *file, uint64_t size, uint32_t frame, uint32_t face, uint8_t *rgba, uint64_t capacity, vt_vtf_info *info_out);

// Source-engine model files -> triangles (vt_mdl.cpp; libs/MDLParser + source/objects/Model.cpp restated).  Throw std::runtime_error on
// malformed / unsupported files.
void MdlInfo(const vt_mdl_files *f, vt_mdl_info *info);
uint32_t MdlBodygroupValues(const vt_mdl_files *f, uint32_t bodygroup);
uint64_t MdlMeshTriangles(const vt_mdl_files *f, uint32_t bodygroup, uint32_t value, vt_tri_in *tris, vt_tri_skin *skin, uint64_t capacity);
void MdlBindMatrices(const vt_mdl_files *f, float *out16);
int32_t MdlMaterialIndex(const vt_mdl_files *f, uint32_t skin, uint32_t material_id);
std::string MdlMaterialPath(const vt_mdl_files *f, uint32_t material_id, uint32_t dir);

// Source-engine map file -> world triangles, materials, static props (vt_bsp.cpp; libs/BSPParser + World::World restated).  Throw
// std::runtime_error on malformed / unsupported files.
void BspInfo(const uint8_t *file, uint64_t size, vt_bsp_info *info);
uint64_t BspTriangles(const uint8_t *file, uint64_t size, vt_tri_in *tris, float *binormals, int16_t *texinfos, uint64_t capacity);
void BspMaterial(const uint8_t *file, uint64_t size, uint32_t material, vt_bsp_material *out);
void BspStaticProp(const uint8_t *file, uint64_t size, uint32_t index, vt_bsp_static_prop *out);

struct DeviceScene;  // HBM-resident copy, vt_accel.cu

// Eager TraceResult for one hit (source/objects/TraceResult.h:54-111): the batched path fills
// vt_attr records; this wrapper gives the single-ray Traverse the reference's getter surface.
class TraceResult {
    vt_attr a;

public:
    explicit TraceResult(const vt_attr &attr) : a(attr), distance(attr.distance), entIdx(attr.ent_id), submatIdx(attr.submat_idx) {
        hitSky = (attr.flags & VT_ATTR_HIT_SKY) != 0;
        frontFacing = (attr.flags & VT_ATTR_FRONT_FACING) != 0;
    }
    float distance;
    uint32_t entIdx;
    uint32_t submatIdx;
    bool hitSky = false;
    bool frontFacing = false;
    const float *GetPos() const { return a.pos; }
    const float *GetNormal() const { return a.normal; }
    const float *GetTangent() const { return a.tangent; }
    const float *GetBinormal() const { return a.binormal; }
    const float *GetGeometricNormal() const { return a.geometric_normal; }
    const float *GetAlbedo() const { return a.albedo; }
    const float *GetBarycentric() const { return a.uvw; }
    const float *GetTexUV() const { return a.tex_uv; }
    float GetAlpha() const { return a.alpha; }
    float GetMetalness() const { return a.metalness; }
    float GetRoughness() const { return a.roughness; }
    float GetBaseMIPLevel() const { return a.base_mip; }
    bool HitWater() const { return (a.flags & VT_ATTR_HIT_WATER) != 0; }
    const vt_attr &Raw() const { return a; }
};

class AccelStruct {
    int mDevice;
    bool mAccelBuilt = false;
    int mWantLayout = VT_LAYOUT_QUAD;  // layout requested for the next Populate (VT_LAYOUT_*)
    int mLayout = VT_LAYOUT_EXACT;     // node layout resident on the device
    mutable HostBvh mAccel;
    mutable bool mBvhStale = false;
    mutable std::mutex mBvhMutex;  // guards the lazy host-side refit in Bvh()  // a device-side refit moved the boxes; the host copy is refitted on demand
    TriangleVec mTriangles;
    std::vector<Entity> mEntities;
    std::vector<Material> mMaterials;
    DeviceScene *mpDevice = nullptr;
    uint64_t mInvalidRays = 0;
    uint64_t mLaunches = 0;
    // host-pointer diffuse waves in flight (RenderDiffuseWaveBegin / Wait): per frame one "lane finished" event per wave lane
    struct WaveFrame {
        struct TileTrace {
            uint64_t base, m;
            int lane;
            ::CUevent_st *ev[6];  // cudaEvent_t; after H2D, K1 primary, K2+K3, K1 bounce, K4, D2H (VT_WAVE_TRACE)
        };
        ::CUevent_st *done[8] = {};  // cudaEvent_t (this header is also read by translation units without the CUDA headers)
        int n_lanes = 0;
        std::vector<TileTrace> tiles;
        ::CUevent_st *ev_begin = nullptr;
    };
    std::deque<WaveFrame> mWaveFrames;
    uint64_t mWaveFrameCount = 0;
    double mRefitRebuildRatio = 0.0;  // > 0: vt_accel_refit rebuilds from scratch once RefitQuality() exceeds it
    uint64_t mRebuilds = 0;

    // per-triangle TraceResult inputs (VtTriAttr, original order): independent of the hierarchy, so Ingest writes them in the same
    // pass as the Triangle constructor and a helper thread uploads them (0.9 GB at 5 M triangles) WHILE the hierarchy is being built
    RawVector<VtTriAttr> mAttrStage;
    std::thread mAttrUpload;
    std::string mAttrUploadError;
    void JoinAttrUpload();
    bool mReplica = false;  // device arrays copied from another handle (vt_group.cu): no host containers, no refit / get_bvh

    void Ingest(const vt_scene &scene, bool stage_attrs = true);
    void Upload(const vt_scene &scene);
    friend class Group;

public:
    // What a replica on another GPU needs besides the bytes of the ten device arrays (multi-GPU, vt_group.cu): the hierarchy is
    // built and flattened ONCE and its device image is copied GPU to GPU (cudaMemcpyPeer in one process, ncclBroadcast across
    // processes); every GPU then holds the whole scene (replicate, never partition).
    struct ReplicaImage {
        uint64_t bytes[10];  // pairs, cpairs, quads, tris, tri_uv, attrs, mats, ents, texs, texels
        uint32_t n_pairs, n_tris, root_leaf_count, n_smem_pairs, has_alphatest, fallback_tex, key_mid;
        int32_t layout;
        uint32_t n_materials;
    };
    void ExportReplica(ReplicaImage &img, const void *bufs[10]) const;
    void AllocReplica(const ReplicaImage &img, void *bufs[10]);  // the caller fills bufs[i] (img.bytes[i] each), then the handle is usable
    bool IsReplica() const { return mReplica; }

    explicit AccelStruct(int device);
    ~AccelStruct();
    AccelStruct(const AccelStruct &) = delete;
    AccelStruct &operator=(const AccelStruct &) = delete;

    // Headless PopulateAccel (source/objects/AccelStruct.cpp:533-776): containers, build, upload.
    void Populate(const vt_scene &scene);
    // Same, with a hierarchy built by the caller (e.g. the reference's PLOC + LeafCollapser).
    void PopulateWithBvh(const vt_scene &scene, const vt_node *nodes, uint64_t node_count, const uint64_t *prim_indices);

    // `accel:Rebuild` for moved geometry of unchanged topology (same triangle count and order, e.g. props that moved):
    // keeps the hierarchy's structure and refits its boxes bottom-up (bvh::HierarchyRefitter, hierarchy_refitter.hpp:20-31)
    // instead of rebuilding, then re-derives the resident layout.  Throws if nothing was built or the count differs.
    void Refit(const vt_scene &scene);
    // The same for a contiguous range of the triangle array (one moved entity): only those records go up.  Quad layout only.
    void RefitRange(const vt_tri_in *tris, uint64_t first, uint64_t count);

    // Batched Traverse (the entry the north star adds behind the same object).
    void TraverseBatch(const vt_ray *rays, uint64_t n, vt_hit *hits, vt_attr *attrs, const float *cones, uint32_t flags,
                       void *stream);
    // SingleRayTraverser::Statistics (single_ray_traverser.hpp:132-135,158-163) summed over the batch:
    // traversal steps (pair visits) and primitive intersections.  Synchronous, closest hit.
    void TraverseStats(const vt_ray *rays, uint64_t n, uint32_t flags, uint64_t *steps, uint64_t *tests, uint32_t *per_ray = nullptr);
    void TraceResultBatch(const vt_ray *rays, const vt_hit *hits, uint64_t n, vt_attr *attrs, const float *cones,
                          uint32_t flags, void *stream);

    // Secondary-ray generation from TraceResult records (device or host pointers as per flags):
    // spp cosine-weighted bounce rays per non-sky hit, slot i*spp+s; *live_out = rays spawned.
    void BounceRays(const vt_attr *attrs, uint64_t n, uint32_t spp, uint64_t seed, vt_ray *out_rays, uint64_t *live_out,
                    uint32_t flags, void *stream);
    // One shadow ray per non-sky hit: origin = CalcRayOrigin(pos, geometric normal), toward a directional or point light.
    void ShadowRays(const vt_attr *attrs, uint64_t n, const float light[3], bool point_light, float tmax, vt_ray *out_rays,
                    uint64_t *live_out, uint32_t flags, void *stream);
    // Batched SampleBSDF, diffuse lobe only (source/libraries/BSDF.cpp:770-825): spp samples per non-sky hit from the TraceResult
    // records and the rays that produced them; out_rays[i*spp+s] + samples[i*spp+s].  Host or device pointers as per flags; the
    // queue arguments (device pointers only, all or none) list the spawned slots as the queued generators do.
    void SampleBsdfRays(const vt_ray *rays, const vt_attr *attrs, uint64_t n, uint32_t spp, uint64_t seed, vt_ray *out_rays,
                        vt_bsdf_sample *samples, uint64_t *live_out, uint32_t *queue, uint64_t *queue_count, vt_hit *miss_hits,
                        uint32_t flags, void *stream);
    // Ray-queue variants (DEVICE pointers only, enqueued on `stream`): the generator also lists the slots it filled —
    // queue[0, *queue_count), *queue_count zeroed here — and writes the miss record of every masked slot into
    // miss_hits; TraverseQueued then traces the listed slots only.  spp == 0 selects the shadow-ray generator.
    // in_queue / in_count (both or neither): wave compaction — only the parents the previous wave's queue lists are visited.
    void BounceRaysQueued(const vt_attr *attrs, uint64_t n, uint32_t spp, uint64_t seed, vt_ray *out_rays, uint32_t *queue,
                          uint64_t *queue_count, vt_hit *miss_hits, void *stream, const uint32_t *in_queue = nullptr,
                          const uint64_t *in_count = nullptr);
    void ShadowRaysQueued(const vt_attr *attrs, uint64_t n, const float light[3], bool point_light, float tmax, vt_ray *out_rays,
                          uint32_t *queue, uint64_t *queue_count, vt_hit *miss_hits, void *stream, const uint32_t *in_queue = nullptr,
                          const uint64_t *in_count = nullptr);
    void TraverseQueued(const vt_ray *rays, const uint32_t *queue, const uint64_t *queue_count, uint64_t capacity, vt_hit *hits,
                        vt_attr *attrs, uint32_t flags, void *stream);
    // "primary + diffuse" wave in one call: traverse the primary rays, build their TraceResults, spawn
    // spp bounce rays per hit on the device and traverse those.  Host pointers are processed in tiles
    // on several streams so the PCIe copies overlap the kernels.
    void TraceDiffuseWave(const vt_ray *rays, uint64_t n, uint32_t spp, uint64_t seed, vt_hit *hits, vt_attr *attrs,
                          vt_ray *bounce_rays, vt_hit *bounce_hits, uint64_t *live_out, uint32_t flags, void *stream);

    // Path waves with compaction (BASELINE config 5): per primary ray a path of up to `bounces` diffuse bounces, a shadow ray toward
    // the sun at every vertex; after the primary wave every kernel visits only the paths that are still alive.  DEVICE pointers.
    void TracePaths(const vt_ray *rays, uint64_t n, uint32_t bounces, const float sun_dir[3], const float sun_rgb[3], uint64_t seed,
                    float weight, float *fb, uint64_t *ray_counts, bool compact, void *stream, int slot = 0);

    // The same wave with the framebuffer as its only result: HOST rays in, HOST RGBFFF image out
    // (fb[i] = weight * albedo_i * escaped fraction of pixel i's bounce rays), tiled over streams like TraceDiffuseWave.
    void RenderDiffuseWave(const vt_ray *rays, uint64_t n, uint32_t spp, uint64_t seed, float weight, float *fb, uint64_t *live_out);
    // the same, split: enqueue a frame / wait for the oldest frame in flight (at most two): vt_accel_render_diffuse_wave_begin / _wait
    void RenderDiffuseWaveBegin(const vt_ray *rays, uint64_t n, uint32_t spp, uint64_t seed, float weight, float *fb, bool count_live, bool pipelined);
    void RenderDiffuseWaveWait();
    void DrainWaveFrames();

    // Fold a diffuse wave into an RGBFFF framebuffer (device pointers only): see k_accumulate_sky.
    void AccumulateSky(const vt_attr *attrs, const vt_hit *bounce_hits, uint64_t n, uint32_t spp, float weight, float *fb,
                       void *stream);

    // Single-ray Traverse with the reference's argument rules and messages
    // (source/objects/AccelStruct.cpp:778-838).  Returns nullptr on a miss; the caller owns the result.
    TraceResult *Traverse(const float origin[3], const float direction[3], float tMin = 0.f,
                          float tMax = 3.402823466e+38f, float coneWidth = -1.f, float coneAngle = -1.f);

    const Material &GetMaterial(size_t i) const { return mMaterials[i]; }  // AccelStruct.cpp:840-843
    const TriangleVec &Triangles() const { return mTriangles; }
    const HostBvh &Bvh() const;  // boxes refreshed lazily after a device-side refit
    bool Built() const { return mAccelBuilt; }
    int Layout() const { return mLayout; }
    void SetLayout(int layout) { mWantLayout = layout; }
    // node-area sum of the resident quad hierarchy relative to the sum right after the build (1 = as built); rebuild trigger
    double RefitQuality() const;
    void SetRefitRebuildRatio(double r) { mRefitRebuildRatio = r; }
    uint64_t Rebuilds() const { return mRebuilds; }
    uint64_t InvalidRays() const { return mInvalidRays; }
    uint64_t Launches() const { return mLaunches; }
    uint64_t DeviceBytes() const;
    int Device() const { return mDevice; }
};

}  // namespace vt
