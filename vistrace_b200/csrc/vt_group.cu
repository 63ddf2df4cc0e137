// vt_group.cu — multi-GPU ray sharding behind the C ABI (BASELINE.json north_star: "the BVH is
// replicated per GPU and ray batches (image tiles / sample indices) are sharded across the 8 x B200 box, with only a final
// framebuffer or hit-buffer gather over NVLink via NCCL").
//
// Two ways to form a group, one implementation:
//   * vt_group_create(devices, n)            ONE process drives n local GPUs — what a C++ host like the reference (a single
//                                            game process, source/objects/AccelStruct.h:61-86) would do.  A worker thread per GPU
//                                            enqueues that GPU's work; results are written by every GPU straight into the caller's
//                                            host buffers over its own PCIe link; the scene image is copied GPU to GPU
//                                            (cudaMemcpyPeer over NVLink).  No NCCL needed.
//   * vt_group_create_rank(dev, r, w, id)    ONE process per GPU (torchrun / MPI launchers; bench.py --gpus N).  The processes
//                                            share a 128-byte ncclUniqueId; the scene image is ncclBroadcast from rank 0 (the only
//                                            rank that builds); hit-buffer slices are gathered on rank 0 with grouped ncclSend /
//                                            ncclRecv over NVLink; FRAMES are delivered through peer memory: rank 0's frame is mapped
//                                            into every rank (CUDA IPC) and each rank's K4 stores its finished pixels straight into it —
//                                            the shading kernel IS the gather, flag words in rank 0's memory do the hand-shake, no
//                                            collective kernel competes with the persistent traversal grids for SMs
//                                            (VT_GROUP_GATHER=nccl selects the ncclSend / ncclRecv gather instead).  NCCL is bound at run time (dlopen of
//                                            libnccl.so.2 — the copy PyTorch already loaded when there is one), so the library has
//                                            no link-time dependency on it.
//
// There is NO exchange step during traversal: rays are independent and the scene is read-only.  A frame is cut into tiles of
// `tile` pixels dealt round-robin to the ranks (sky tiles and dense tiles mix, so the shards balance); a shard is stored
// compactly on its GPU, and the random-number counter of a bounce ray comes from its GLOBAL pixel (VtSlotMap), so the image
// of a sharded frame equals the image of the same frame traced on one GPU bit for bit.
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only: every NCCL function is reached through dlsym

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <thread>

#include "vt_accel_internal.h"

namespace vt {

// ------------------------------------------------------------------------------------------------ NCCL, bound at run time
struct NcclApi {
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclReduce) Reduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    void *handle = nullptr;

    static NcclApi &get() {
        static NcclApi api;
        static std::once_flag once;
        std::call_once(once, [] {
            const char *names[] = {std::getenv("VT_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
            for (const char *n : names) {
                if (!n || !*n) continue;
                api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
                if (api.handle) break;
            }
            if (!api.handle) return;
#define VT_SYM(name) api.name = reinterpret_cast<decltype(api.name)>(dlsym(api.handle, "nccl" #name))
            VT_SYM(GetUniqueId);
            VT_SYM(CommInitRank);
            VT_SYM(CommDestroy);
            VT_SYM(GroupStart);
            VT_SYM(GroupEnd);
            VT_SYM(Send);
            VT_SYM(Recv);
            VT_SYM(Broadcast);
            VT_SYM(Reduce);
            VT_SYM(AllGather);
            VT_SYM(GetErrorString);
            VT_SYM(GetVersion);
#undef VT_SYM
        });
        if (!api.handle || !api.GetUniqueId || !api.CommInitRank || !api.Send || !api.Recv || !api.Broadcast || !api.GroupStart || !api.GroupEnd)
            throw std::runtime_error("NCCL is not available (libnccl.so.2 could not be loaded; set VT_NCCL_LIB): a multi-process vt_group needs it");
        return api;
    }
};

#define VT_NCCL(expr)                                                                                          \
    do {                                                                                                       \
        ncclResult_t _r = (expr);                                                                              \
        if (_r != ncclSuccess)                                                                                 \
            throw std::runtime_error(std::string("NCCL error: ") + vt::NcclApi::get().GetErrorString(_r) + " at " #expr); \
    } while (0)

// ------------------------------------------------------------------------------------------------ shard geometry
// Frame of n records cut into tiles of `tile` records, dealt round-robin: rank r owns tiles r, r + world, ...  Only the last
// tile of the frame can be short, and it is the last local tile of its owner.
struct ShardGeom {
    uint64_t n = 0, tile = 1;
    uint32_t world = 1, rank = 0;
    uint64_t n_tiles() const { return (n + tile - 1) / tile; }
    uint64_t local_tiles() const {
        const uint64_t nt = n_tiles();
        return nt > rank ? (nt - rank + world - 1) / world : 0;
    }
    uint64_t tile_records(uint64_t g) const { return std::min(tile, n - g * tile); }
    uint64_t local_records(uint64_t upto_local_tile) const {  // records in local tiles [0, upto)
        if (upto_local_tile == 0) return 0;
        return (upto_local_tile - 1) * tile + tile_records((upto_local_tile - 1) * world + rank);
    }
    uint64_t local_records() const { return local_records(local_tiles()); }
    ShardGeom of(uint32_t r) const {
        ShardGeom g = *this;
        g.rank = r;
        return g;
    }
};

// Copy local tiles [j0, j1) of a shard between its COMPACT buffer (local tile j at record j * tile) and the FRAME buffer
// (global tile g at record g * tile): one strided 2-D copy for the full tiles, one plain copy for a short last tile.
static void copy_tiles(const ShardGeom &g, uint64_t j0, uint64_t j1, size_t rec, void *compact, void *frame, bool to_frame,
                       cudaMemcpyKind kind, cudaStream_t st) {
    if (j1 <= j0) return;
    const uint64_t last_g = (j1 - 1) * g.world + g.rank;
    const bool ragged = g.tile_records(last_g) != g.tile;
    const uint64_t full = (j1 - j0) - (ragged ? 1 : 0);
    char *c = static_cast<char *>(compact) + j0 * g.tile * rec;
    char *f = static_cast<char *>(frame) + (j0 * g.world + g.rank) * g.tile * rec;
    const size_t row = g.tile * rec;
    if (full) {
        if (to_frame) VT_CUDA(cudaMemcpy2DAsync(f, row * g.world, c, row, row, full, kind, st));
        else VT_CUDA(cudaMemcpy2DAsync(c, row, f, row * g.world, row, full, kind, st));
    }
    if (ragged) {
        char *c2 = c + full * row, *f2 = f + full * row * g.world;
        const size_t bytes = g.tile_records(last_g) * rec;
        if (to_frame) VT_CUDA(cudaMemcpyAsync(f2, c2, bytes, kind, st));
        else VT_CUDA(cudaMemcpyAsync(c2, f2, bytes, kind, st));
    }
}

// ------------------------------------------------------------------------------------------------ the group
class Group {
public:
    struct Member {
        int device = 0;
        uint32_t rank = 0;  // global rank of this member
        std::unique_ptr<vt_accel> accel;
        ncclComm_t comm = nullptr;
        cudaStream_t stream = nullptr;  // collectives and result copies
        cudaEvent_t lane_done[8] = {};
        cudaEvent_t caller_done = nullptr;
        DevBuf<float> fb_local;  // compact shard image (3 floats per local pixel)
        DevBuf<float> fb_stage;  // rank 0 of a multi-process group: the other ranks' compact shards, back to back
        DevBuf<vt_hit> hit_stage;
        DevBuf<vt_attr> attr_stage;
        DevBuf<unsigned char> header;
        // peer-memory frame (multi-process groups): rank 0 owns `frame` = n x 12 bytes of RGBFFF pixels in GLOBAL pixel order + flag
        // words; every other rank maps it over NVLink (CUDA IPC) and its K4 stores finished pixels straight into it
        // Two of them (VT_GROUP_FRAME_SLOT1): calls on different slots share nothing — own frame, own flags, own scratch (the wave lane
        // of the same index) — so a caller that issues consecutive frames on two streams has two frames in flight.
        struct PeerFrame {
            DevBuf<unsigned char> frame;
            void *peer_frame = nullptr;
            uint64_t frame_pixels = 0;
            uint32_t step = 0;
        } pf[2];
        uint64_t live = 0;
        // worker thread (single-process groups with several GPUs)
        std::thread th;
        std::mutex mu;
        std::condition_variable cv;
        std::function<void()> job;
        bool has_job = false, quit = false;
        std::exception_ptr err;
    };

private:
    std::vector<std::unique_ptr<Member>> mMembers;
    std::deque<cudaEvent_t> mSharedFrames;  // VT_GROUP_SHARED_HOST_FRAME | VT_GROUP_ASYNC frames in flight (at most two)
    uint64_t mSharedFrameCount = 0;
    uint32_t mWorld = 1;
    bool mMultiProcess = false;
    bool mPeerFrameUnavailable = false;  // CUDA IPC could not map rank 0's frame on some rank: every rank uses the NCCL gather instead
    uint64_t mLaunches = 0;

    static void worker_loop(Member *m) {
        for (;;) {
            std::function<void()> job;
            {
                std::unique_lock<std::mutex> lock(m->mu);
                m->cv.wait(lock, [&] { return m->has_job || m->quit; });
                if (m->quit) return;
                job = m->job;
            }
            try {
                job();
            } catch (...) {
                m->err = std::current_exception();
            }
            {
                std::lock_guard<std::mutex> lock(m->mu);
                m->has_job = false;
            }
            m->cv.notify_all();
        }
    }

    // run fn(member) for every local member: inline for one, on the members' worker threads for several
    void run_all(const std::function<void(Member &)> &fn) {
        if (mMembers.size() == 1) {
            fn(*mMembers[0]);
            return;
        }
        for (auto &mp : mMembers) {
            Member *m = mp.get();
            {
                std::lock_guard<std::mutex> lock(m->mu);
                m->err = nullptr;
                m->job = [m, &fn] { fn(*m); };
                m->has_job = true;
            }
            m->cv.notify_all();
        }
        std::exception_ptr first;
        for (auto &mp : mMembers) {
            Member *m = mp.get();
            std::unique_lock<std::mutex> lock(m->mu);
            m->cv.wait(lock, [&] { return !m->has_job; });
            if (m->err && !first) first = m->err;
        }
        if (first) std::rethrow_exception(first);
    }

    void init_member(Member &m) {
        m.accel.reset(new vt_accel(m.device));
        VT_CUDA(cudaSetDevice(m.device));
        VT_CUDA(cudaStreamCreateWithFlags(&m.stream, cudaStreamNonBlocking));
        for (auto &e : m.lane_done) VT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        VT_CUDA(cudaEventCreateWithFlags(&m.caller_done, cudaEventDisableTiming));
    }

public:
    Group(const int *devices, int n) {
        if (n <= 0 || !devices) throw std::runtime_error("vt_group_create: need at least one device");
        mWorld = (uint32_t)n;
        for (int i = 0; i < n; i++) {
            for (int k = 0; k < i; k++)
                if (devices[k] == devices[i]) throw std::runtime_error("vt_group_create: a device is listed twice");
            mMembers.emplace_back(new Member());
            mMembers.back()->device = devices[i];
            mMembers.back()->rank = (uint32_t)i;
            init_member(*mMembers.back());
        }
        for (auto &a : mMembers)  // direct GPU-to-GPU copies for the scene image where the fabric allows it
            for (auto &b : mMembers) {
                if (a->device == b->device) continue;
                int can = 0;
                cudaDeviceCanAccessPeer(&can, a->device, b->device);
                if (can) {
                    cudaSetDevice(a->device);
                    cudaError_t e = cudaDeviceEnablePeerAccess(b->device, 0);
                    if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                }
            }
        if (n > 1)
            for (auto &m : mMembers) m->th = std::thread(worker_loop, m.get());
    }

    Group(int device, int rank, int world, const uint8_t id[128]) {
        if (world <= 0 || rank < 0 || rank >= world) throw std::runtime_error("vt_group_create_rank: rank / world out of range");
        mWorld = (uint32_t)world;
        mMultiProcess = true;
        mMembers.emplace_back(new Member());
        Member &m = *mMembers.back();
        m.device = device;
        m.rank = (uint32_t)rank;
        init_member(m);
        if (world > 1) {
            if (!id) throw std::runtime_error("vt_group_create_rank: null ncclUniqueId");
            NcclApi &nccl = NcclApi::get();
            ncclUniqueId uid;
            static_assert(sizeof(uid) == 128, "ncclUniqueId is 128 bytes");
            std::memcpy(&uid, id, sizeof(uid));
            VT_CUDA(cudaSetDevice(device));
            VT_NCCL(nccl.CommInitRank(&m.comm, world, uid, rank));
        }
    }

    ~Group() {
        for (auto &mp : mMembers) {
            Member &m = *mp;
            if (m.th.joinable()) {
                {
                    std::lock_guard<std::mutex> lock(m.mu);
                    m.quit = true;
                }
                m.cv.notify_all();
                m.th.join();
            }
            cudaSetDevice(m.device);
            cudaDeviceSynchronize();
            for (cudaEvent_t e : mSharedFrames) cudaEventDestroy(e);  // frames begun with VT_GROUP_ASYNC and never awaited
            mSharedFrames.clear();
            for (auto &P : m.pf) {
                if (P.peer_frame && m.rank != 0) cudaIpcCloseMemHandle(P.peer_frame);
                P.frame.release();
            }
            if (m.comm) NcclApi::get().CommDestroy(m.comm);
            m.fb_local.release();
            m.fb_stage.release();
            m.hit_stage.release();
            m.attr_stage.release();
            m.header.release();
            for (auto &e : m.lane_done)
                if (e) cudaEventDestroy(e);
            if (m.caller_done) cudaEventDestroy(m.caller_done);
            if (m.stream) cudaStreamDestroy(m.stream);
            m.accel.reset();
        }
    }

    uint32_t World() const { return mWorld; }
    uint32_t Rank() const { return mMembers[0]->rank; }
    size_t LocalMembers() const { return mMembers.size(); }
    vt_accel *Accel(size_t i) { return i < mMembers.size() ? mMembers[i]->accel.get() : nullptr; }
    uint64_t Launches() const {
        uint64_t n = mLaunches;
        for (auto &m : mMembers) n += m->accel->impl.Launches();
        return n;
    }
    uint64_t FrameTile() const { return (uint64_t)std::max(32, env_int("VT_GROUP_TILE", 8192)); }

    // ---- build once, replicate the device image
    void Populate(const vt_scene *scene) {
        Member &root = *mMembers[0];
        AccelStruct::ReplicaImage img;
        const void *src[10];
        if (!mMultiProcess) {
            if (!scene) throw std::runtime_error("vt_group_populate: null scene");
            root.accel->impl.Populate(*scene);
            if (mMembers.size() == 1) return;
            root.accel->impl.ExportReplica(img, src);
            for (size_t k = 1; k < mMembers.size(); k++) {
                Member &m = *mMembers[k];
                void *dst[10];
                m.accel->impl.SetLayout(root.accel->impl.Layout());
                m.accel->impl.AllocReplica(img, dst);
                VT_CUDA(cudaSetDevice(m.device));
                for (int i = 0; i < 10; i++)
                    if (img.bytes[i]) VT_CUDA(cudaMemcpyPeerAsync(dst[i], m.device, src[i], root.device, img.bytes[i], m.stream));
            }
            for (size_t k = 1; k < mMembers.size(); k++) {
                VT_CUDA(cudaSetDevice(mMembers[k]->device));
                VT_CUDA(cudaStreamSynchronize(mMembers[k]->stream));
            }
            return;
        }
        // one process per GPU: rank 0 builds, everybody else receives the image over NCCL
        if (root.rank == 0) {
            if (!scene) throw std::runtime_error("vt_group_populate: rank 0 needs the scene");
            root.accel->impl.Populate(*scene);
            if (mWorld == 1) return;
            root.accel->impl.ExportReplica(img, src);
        }
        NcclApi &nccl = NcclApi::get();
        VT_CUDA(cudaSetDevice(root.device));
        root.header.ensure(sizeof(img));
        if (root.rank == 0) VT_CUDA(cudaMemcpyAsync(root.header.p, &img, sizeof(img), cudaMemcpyHostToDevice, root.stream));
        VT_NCCL(nccl.Broadcast(root.header.p, root.header.p, sizeof(img), ncclUint8, 0, root.comm, root.stream));
        VT_CUDA(cudaMemcpyAsync(&img, root.header.p, sizeof(img), cudaMemcpyDeviceToHost, root.stream));
        VT_CUDA(cudaStreamSynchronize(root.stream));
        void *dst[10];
        if (root.rank == 0) {
            for (int i = 0; i < 10; i++) dst[i] = const_cast<void *>(src[i]);
        } else {
            root.accel->impl.AllocReplica(img, dst);
        }
        for (int i = 0; i < 10; i++)
            if (img.bytes[i]) VT_NCCL(nccl.Broadcast(dst[i], dst[i], img.bytes[i], ncclUint8, 0, root.comm, root.stream));
        VT_CUDA(cudaStreamSynchronize(root.stream));
    }

    // ---- batched Traverse over contiguous, balanced slices of the ray array
    static void slice(uint64_t n, uint32_t rank, uint32_t world, uint64_t &b, uint64_t &e) {
        const uint64_t base = n / world, rem = n % world;
        b = rank * base + std::min<uint64_t>(rank, rem);
        e = b + base + (rank < rem ? 1 : 0);
    }

    void Traverse(const vt_ray *rays, uint64_t n, vt_hit *hits, vt_attr *attrs, uint32_t flags) {
        if (flags & VT_TRAVERSE_DEVICE_PTRS) throw std::runtime_error("vt_group_traverse: host pointers only");
        if (n == 0) return;
        if (!rays || !hits) throw std::runtime_error("vt_group_traverse: rays and hits must not be null");
        if (!mMultiProcess) {  // every GPU reads its slice from, and writes it back to, the caller's host arrays
            run_all([&](Member &m) {
                uint64_t b, e;
                slice(n, m.rank, mWorld, b, e);
                m.accel->impl.TraverseBatch(rays + b, e - b, hits + b, attrs ? attrs + b : nullptr, nullptr, flags, nullptr);
            });
            return;
        }
        Member &m = *mMembers[0];
        uint64_t b, e;
        slice(n, m.rank, mWorld, b, e);
        if (mWorld == 1) {
            m.accel->impl.TraverseBatch(rays, n, hits, attrs, nullptr, flags, nullptr);
            return;
        }
        // one process per GPU: own slice on the device, slices gathered on rank 0 over NVLink, one download there
        NcclApi &nccl = NcclApi::get();
        VT_CUDA(cudaSetDevice(m.device));
        DeviceScene &D = *m.accel->impl.mpDevice;
        const bool root = m.rank == 0;
        D.s_rays.ensure(std::max<uint64_t>(1, e - b));
        m.hit_stage.ensure(root ? n : std::max<uint64_t>(1, e - b));
        if (attrs) m.attr_stage.ensure(root ? n : std::max<uint64_t>(1, e - b));
        vt_hit *d_hits = m.hit_stage.p + (root ? b : 0);
        vt_attr *d_attrs = attrs ? m.attr_stage.p + (root ? b : 0) : nullptr;
        if (e > b) {
            VT_CUDA(cudaMemcpyAsync(D.s_rays.p, rays + b, (e - b) * sizeof(vt_ray), cudaMemcpyHostToDevice, m.stream));
            m.accel->impl.TraverseBatch(D.s_rays.p, e - b, d_hits, d_attrs, nullptr, flags | VT_TRAVERSE_DEVICE_PTRS, m.stream);
        }
        VT_NCCL(nccl.GroupStart());
        if (root) {
            for (uint32_t r = 1; r < mWorld; r++) {
                uint64_t rb, re;
                slice(n, r, mWorld, rb, re);
                if (re == rb) continue;
                VT_NCCL(nccl.Recv(m.hit_stage.p + rb, (re - rb) * sizeof(vt_hit), ncclUint8, (int)r, m.comm, m.stream));
                if (attrs) VT_NCCL(nccl.Recv(m.attr_stage.p + rb, (re - rb) * sizeof(vt_attr), ncclUint8, (int)r, m.comm, m.stream));
            }
        } else if (e > b) {
            VT_NCCL(nccl.Send(d_hits, (e - b) * sizeof(vt_hit), ncclUint8, 0, m.comm, m.stream));
            if (attrs) VT_NCCL(nccl.Send(d_attrs, (e - b) * sizeof(vt_attr), ncclUint8, 0, m.comm, m.stream));
        }
        VT_NCCL(nccl.GroupEnd());
        if (root) {
            VT_CUDA(cudaMemcpyAsync(hits, m.hit_stage.p, n * sizeof(vt_hit), cudaMemcpyDeviceToHost, m.stream));
            if (attrs) VT_CUDA(cudaMemcpyAsync(attrs, m.attr_stage.p, n * sizeof(vt_attr), cudaMemcpyDeviceToHost, m.stream));
        } else if (e > b) {  // a non-root process still gets its own slice back
            VT_CUDA(cudaMemcpyAsync(hits + b, d_hits, (e - b) * sizeof(vt_hit), cudaMemcpyDeviceToHost, m.stream));
            if (attrs) VT_CUDA(cudaMemcpyAsync(attrs + b, d_attrs, (e - b) * sizeof(vt_attr), cudaMemcpyDeviceToHost, m.stream));
        }
        VT_CUDA(cudaStreamSynchronize(m.stream));
    }

    // ---- the "primary + diffuse" frame, sharded by tiles
    // One member's share.  Host rays: the shard is cut into chunks that run on the handle's wave lanes, uploads on its copy
    // stream, so copies overlap kernels as in AccelStruct::RenderDiffuseWave.  Device rays (compact shard): ONE chunk on `stream`.
    // The compact shard image is left in m.fb_local; with fb_host it is also written to the caller's frame (tile-strided D2H).
    // after_chunk(j0, j1, stream): called once the kernels of local tiles [j0, j1) are enqueued on `stream` — the hook the
    // multi-process path uses to ship a finished chunk to rank 0 while later chunks are still being traced.  The chunk
    // schedule is computed from geometry `sched` (rank 0's, the longest shard) so that every rank cuts at the same tile indices.
    void member_render(Member &m, const ShardGeom &g, const vt_ray *rays, bool rays_on_device, uint32_t spp, uint64_t seed, float weight,
                       float *fb_host, bool count_live, cudaStream_t caller_stream, const ShardGeom *sched = nullptr,
                       const std::function<void(uint64_t, uint64_t, cudaStream_t)> &after_chunk = nullptr, float *frame_target = nullptr,
                       const uint32_t *consumed_flag = nullptr, uint32_t step = 0, int slot = 0, bool staging_alt = false, bool pipelined = false) {
        AccelStruct &A = m.accel->impl;
        if (!A.Built()) throw std::runtime_error("vt_group: populate the group first");
        VT_CUDA(cudaSetDevice(m.device));
        DeviceScene &D = *A.mpDevice;
        const uint64_t lt = g.local_tiles(), L = g.local_records();
        const uint64_t lt_sched = sched ? sched->local_tiles() : lt;  // >= lt
        m.live = 0;
        if (L == 0 && !after_chunk) return;
        if (L * spp > 0xFFFFFFFFull) throw std::runtime_error("vt_group: more than 2^32 bounce slots per GPU");
        m.fb_local.ensure(std::max<uint64_t>(1, L * 3));
        D.live.ensure(1);
        auto next_counter = [&]() { return D.counters.p + 2 * (D.next_slot.fetch_add(1) % kCounterSlots); };
        // device-resident shards may also be cut into chunks on several lanes (VT_GROUP_DEV_LANES > 1): a small shard's K1 launches
        // are latency-bound (the slowest ray's dependent chain, ~0.1 ms), so one chunk's primary wave overlapping another chunk's
        // bounce wave keeps the SMs busier; the lanes are fenced against the caller's stream on both sides
        const int dev_lanes = std::max(1, std::min(8, env_int("VT_GROUP_DEV_LANES", 1)));
        const int n_lanes = rays_on_device ? dev_lanes : std::max(1, std::min(8, env_int("VT_WAVE_LANES", 4)));
        const bool on_caller = rays_on_device && n_lanes == 1;  // everything goes straight onto the caller's stream
        for (int i = 0; i < n_lanes; i++)
            if (!D.lanes[i].stream) VT_CUDA(cudaStreamCreateWithFlags(&D.lanes[i].stream, cudaStreamNonBlocking));
        if (count_live) {
            cudaStream_t s0 = rays_on_device ? caller_stream : D.lanes[0].stream;
            VT_CUDA(cudaMemsetAsync(D.live.p, 0, sizeof(unsigned long long), s0));
            if (!rays_on_device) VT_CUDA(cudaStreamSynchronize(s0));
        }
        if (rays_on_device && !on_caller) {  // lanes start after whatever the caller enqueued before this call
            VT_CUDA(cudaEventRecord(m.caller_done, caller_stream));
            for (int i = 0; i < n_lanes; i++) VT_CUDA(cudaStreamWaitEvent(D.lanes[i].stream, m.caller_done, 0));
        }
        VtLaunchConfig cfg = D.cfg;
        if (n_lanes > 1) {  // several chunks in flight: leave CTA slots for the next chunk's kernels (see RenderDiffuseWave)
            const int per_sm = env_int(rays_on_device ? "VT_GROUP_DEV_CTAS_PER_SM" : "VT_WAVE_CTAS_PER_SM", 6);
            if (per_sm > 0) cfg.grid = std::min(D.cfg.grid, D.sm_count * per_sm);
        }
        const uint64_t chunk_records = (uint64_t)std::max(1, env_int("VT_WAVE_TILE", 1 << 19));
        const uint64_t dev_chunks = (uint64_t)std::max(n_lanes, env_int("VT_GROUP_DEV_CHUNKS", n_lanes));
        const uint64_t chunk_tiles_max = rays_on_device ? std::max<uint64_t>(1, (lt_sched + dev_chunks - 1) / dev_chunks) : std::max<uint64_t>(1, chunk_records / g.tile);
        // first chunk: small when the frame runs on its own (a short upload before the first kernel); with another frame in flight the
        // fill is hidden and fewer, larger chunks win (AccelStruct::RenderDiffuseWaveBegin, tools/e2e_chunk_probe.py)
        uint64_t first_default = chunk_records / 8;
        if (pipelined) first_default = L <= chunk_records ? std::max<uint64_t>(1, L / 2) : (L >= 4 * chunk_records ? chunk_records / 2 : chunk_records / 8);
        uint64_t chunk_tiles = rays_on_device ? chunk_tiles_max
                                              : std::max<uint64_t>(1, std::min(chunk_tiles_max, (uint64_t)std::max(1, env_int("VT_WAVE_FIRST", (int)first_default)) / g.tile));
        if (!rays_on_device) {
            (staging_alt ? D.wave_rays_b : D.wave_rays).ensure(std::max<uint64_t>(1, L));
            if (!D.copy_stream) VT_CUDA(cudaStreamCreateWithFlags(&D.copy_stream, cudaStreamNonBlocking));
        }
        size_t n_uploads = 0;
        int li = 0, chunks_since_fence = 0;
        bool used[8] = {};
        for (uint64_t s0 = 0, s1 = 0; s0 < lt_sched; s0 = s1, li = (li + 1) % n_lanes, chunk_tiles = std::min(chunk_tiles_max, chunk_tiles * 2)) {
            s1 = std::min(lt_sched, s0 + chunk_tiles);
            if (lt_sched - s1 < chunk_tiles / 2) s1 = lt_sched;  // no small tail chunk
            const uint64_t j0 = std::min(s0, lt), j1 = std::min(s1, lt);  // this rank's part of the chunk (may be empty at the end)
            DeviceScene::WaveLane &l = D.lanes[on_caller ? slot : li];  // a device-resident call on the caller's stream: the slot's own scratch
            cudaStream_t st = on_caller ? caller_stream : l.stream;
            used[li] = true;
            if (j1 <= j0) {
                if (after_chunk) after_chunk(s0, s1, st);
                continue;
            }
            const uint64_t cb = j0 * g.tile, mpix = g.local_records(j1) - cb;
            l.hits.ensure(mpix);
            l.attrs.ensure(mpix);
            l.brays.ensure(mpix * spp);
            l.bhits.ensure(mpix * spp);
            l.queue.ensure(mpix * spp);
            l.queue_count.ensure(1);
            if (++chunks_since_fence >= kCounterSlots / 2 - 16) {
                for (int i = 0; i < n_lanes; i++) VT_CUDA(cudaStreamSynchronize(D.lanes[i].stream));
                chunks_since_fence = 0;
            }
            unsigned long long *c0 = next_counter(), *c1 = next_counter();
            VT_CUDA(cudaMemsetAsync(c0, 0, 16, st));
            VT_CUDA(cudaMemsetAsync(c1, 0, 16, st));
            VT_CUDA(cudaMemsetAsync(l.queue_count.p, 0, sizeof(unsigned long long), st));
            float *d_fb = m.fb_local.p + cb * 3;
            if (!frame_target) VT_CUDA(cudaMemsetAsync(d_fb, 0, mpix * 3 * sizeof(float), st));
            const vt_ray *d_rays;
            if (rays_on_device) {
                d_rays = rays + cb;
            } else {
                if (n_uploads == D.upload_done.size()) {
                    cudaEvent_t ev;
                    VT_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                    D.upload_done.push_back(ev);
                }
                cudaEvent_t done = D.upload_done[n_uploads++];
                vt_ray *staging = (staging_alt ? D.wave_rays_b : D.wave_rays).p;  // consecutive frames in flight alternate
                copy_tiles(g, j0, j1, sizeof(vt_ray), staging, const_cast<vt_ray *>(rays), false, cudaMemcpyHostToDevice, D.copy_stream);
                VT_CUDA(cudaEventRecord(done, D.copy_stream));
                VT_CUDA(cudaStreamWaitEvent(st, done, 0));
                d_rays = staging + cb;
            }
            VtSlotMap map;
            map.local_base = cb, map.tile = g.tile, map.stride = g.world, map.phase = g.rank;
            VT_CUDA(vt_launch_traverse(D.view, d_rays, l.hits.p, mpix, false, c0, cfg, st));
            VT_CUDA(vt_launch_trace_result(D.view, d_rays, l.hits.p, nullptr, l.attrs.p, mpix, st));
            VT_CUDA(vt_launch_bounce_rays(l.attrs.p, mpix, spp, seed, 0, l.brays.p, count_live ? D.live.p : nullptr, st, l.queue.p, l.queue_count.p,
                                          l.bhits.p, &map));
            VT_CUDA(vt_launch_traverse(D.view, l.brays.p, l.bhits.p, mpix * spp, false, c1, cfg, st, false, l.queue.p, l.queue_count.p));
            if (frame_target) {
                // K4 delivers the shard: its stores go to the pixels' places in the frame on rank 0 (peer memory over NVLink for the
                // other ranks) — no staging buffer, no collective kernel, no de-interleave.  The frame of the PREVIOUS step must have
                // been consumed by rank 0 first (flag in rank 0's memory).
                VT_CUDA(vt_launch_flag_wait(consumed_flag, 1, 1, step - 1, st));
                VT_CUDA(vt_launch_accumulate_sky(D.view, l.attrs.p, l.bhits.p, mpix, spp, weight, frame_target, st, &map, true));
                A.mLaunches += 1;
            } else {
                VT_CUDA(vt_launch_accumulate_sky(D.view, l.attrs.p, l.bhits.p, mpix, spp, weight, d_fb, st));
            }
            A.mLaunches += 5;
            if (fb_host) copy_tiles(g, j0, j1, 3 * sizeof(float), m.fb_local.p, fb_host, true, cudaMemcpyDeviceToHost, st);
            if (after_chunk) after_chunk(s0, s1, st);
        }
        if (!on_caller) {  // the stream that carries the gather (the member's own, or the caller's) continues after every lane
            cudaStream_t after = rays_on_device ? caller_stream : m.stream;
            for (int i = 0; i < n_lanes; i++)
                if (used[i]) {
                    VT_CUDA(cudaEventRecord(m.lane_done[i], D.lanes[i].stream));
                    VT_CUDA(cudaStreamWaitEvent(after, m.lane_done[i], 0));
                }
        }
    }

    void member_finish(Member &m, bool count_live, cudaStream_t st) {
        VT_CUDA(cudaSetDevice(m.device));
        DeviceScene &D = *m.accel->impl.mpDevice;
        unsigned long long v = 0;
        if (count_live && D.live.p) VT_CUDA(cudaMemcpyAsync(&v, D.live.p, sizeof(v), cudaMemcpyDeviceToHost, st));
        VT_CUDA(cudaStreamSynchronize(st));
        m.live = v;
    }

    static constexpr uint32_t kMaxChunks = 64;
    static uint64_t frame_flag_offset(uint64_t pixels) { return (pixels * 12 + 255) / 256 * 256; }

    // Collective: make rank 0's frame hold `pixels` pixels and map it into every other rank (CUDA IPC over NVLink).
    void ensure_peer_frame(uint64_t pixels, int slot) {
        Member &m = *mMembers[0];
        Member::PeerFrame &P = m.pf[slot];
        if (P.peer_frame && pixels <= P.frame_pixels) return;
        NcclApi &nccl = NcclApi::get();
        VT_CUDA(cudaSetDevice(m.device));
        VT_CUDA(cudaDeviceSynchronize());
        if (P.peer_frame && m.rank != 0) VT_CUDA(cudaIpcCloseMemHandle(P.peer_frame));
        P.peer_frame = nullptr;
        m.header.ensure(256);
        // nobody maps the old frame any more once everybody has passed this collective
        VT_NCCL(nccl.Broadcast(m.header.p, m.header.p, 4, ncclUint8, 0, m.comm, m.stream));
        VT_CUDA(cudaStreamSynchronize(m.stream));
        cudaIpcMemHandle_t handle;
        std::memset(&handle, 0, sizeof(handle));
        const uint64_t bytes = frame_flag_offset(pixels) + ((uint64_t)mWorld * kMaxChunks + 64) * sizeof(uint32_t);
        if (m.rank == 0) {
            P.frame.release();
            P.frame.ensure(bytes);
            VT_CUDA(cudaMemset(P.frame.p, 0, bytes));
            VT_CUDA(cudaIpcGetMemHandle(&handle, P.frame.p));
            VT_CUDA(cudaMemcpy(m.header.p, &handle, sizeof(handle), cudaMemcpyHostToDevice));
        }
        static_assert(sizeof(handle) <= 256, "IPC handle fits the header buffer");
        VT_NCCL(nccl.Broadcast(m.header.p, m.header.p, sizeof(handle), ncclUint8, 0, m.comm, m.stream));
        VT_CUDA(cudaMemcpyAsync(&handle, m.header.p, sizeof(handle), cudaMemcpyDeviceToHost, m.stream));
        VT_CUDA(cudaStreamSynchronize(m.stream));
        unsigned char ok = 1;
        if (m.rank == 0) {
            P.peer_frame = P.frame.p;
        } else if (cudaIpcOpenMemHandle(&P.peer_frame, handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            P.peer_frame = nullptr;
            ok = 0;
        }
        // all ranks must agree before anybody waits on a flag a peer could never set: gather one status byte per rank
        m.header.ensure(std::max<size_t>(256, mWorld));
        VT_CUDA(cudaMemcpyAsync(m.header.p + m.rank, &ok, 1, cudaMemcpyHostToDevice, m.stream));
        if (!nccl.AllGather) throw std::runtime_error("ncclAllGather not found");
        VT_NCCL(nccl.AllGather(m.header.p + m.rank, m.header.p, 1, ncclUint8, m.comm, m.stream));
        std::vector<unsigned char> all(mWorld, 0);
        VT_CUDA(cudaMemcpyAsync(all.data(), m.header.p, mWorld, cudaMemcpyDeviceToHost, m.stream));
        VT_CUDA(cudaStreamSynchronize(m.stream));
        for (unsigned char v : all)
            if (!v) mPeerFrameUnavailable = true;
        if (mPeerFrameUnavailable) {
            if (P.peer_frame && m.rank != 0) cudaIpcCloseMemHandle(P.peer_frame);
            P.peer_frame = nullptr;
            return;
        }
        P.frame_pixels = pixels;
        P.step = 0;
    }

    // Multi-process frame through peer memory: every rank's K4 stores its finished pixels into rank 0's frame; flag words say which
    // chunk of which rank has landed; rank 0 downloads (or hands over) each stretch of the frame as soon as all ranks have delivered it.
    bool render_peer(Member &m, const ShardGeom &g0, const vt_ray *rays, bool dev_ptrs, uint32_t spp, uint64_t seed, float weight, float *fb,
                     uint64_t *live_out, cudaStream_t stream, int slot) {
        const ShardGeom g = g0.of(m.rank), sched = g0.of(0);
        const bool root = m.rank == 0;
        Member::PeerFrame &P = m.pf[slot];
        if (!mPeerFrameUnavailable) ensure_peer_frame(g0.n, slot);
        if (mPeerFrameUnavailable) return false;
        const uint32_t step = ++P.step;
        unsigned char *base = static_cast<unsigned char *>(P.peer_frame);
        float *frame = reinterpret_cast<float *>(base);
        uint32_t *flags = reinterpret_cast<uint32_t *>(base + frame_flag_offset(P.frame_pixels));
        uint32_t *consumed = flags + (uint64_t)mWorld * kMaxChunks;
        cudaStream_t st = dev_ptrs ? stream : m.stream;
        uint32_t chunk = 0;
        int ev = 0;
        auto landed = [&](uint64_t s0, uint64_t s1, cudaStream_t lane) {
            if (chunk >= kMaxChunks) throw std::runtime_error("vt_group: more than 64 chunks per shard (raise VT_WAVE_TILE)");
            // this rank's tiles of the chunk are in the frame: say so (a rank without tiles in the chunk says so as well)
            VT_CUDA(vt_launch_flag_set(flags + (uint64_t)m.rank * kMaxChunks + chunk, step, lane));
            mLaunches++;
            if (root) {
                if (lane != st) {
                    cudaEvent_t done = m.lane_done[ev++ % 8];
                    VT_CUDA(cudaEventRecord(done, lane));
                    VT_CUDA(cudaStreamWaitEvent(st, done, 0));
                }
                VT_CUDA(vt_launch_flag_wait(flags + chunk, mWorld, kMaxChunks, step, st));
                mLaunches++;
                // the chunk's tiles of ALL ranks are one contiguous stretch of the frame: global tiles [s0 * W, s1 * W)
                const uint64_t p0 = std::min(g0.n, s0 * mWorld * g0.tile), p1 = std::min(g0.n, s1 * mWorld * g0.tile);
                if (p1 > p0 && fb != frame)
                    VT_CUDA(cudaMemcpyAsync(fb + p0 * 3, frame + p0 * 3, (p1 - p0) * 3 * sizeof(float), dev_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
            }
            chunk++;
        };
        member_render(m, g, rays, dev_ptrs, spp, seed, weight, nullptr, live_out != nullptr, st, &sched, landed, frame, consumed, step, slot);
        if (root) {  // the frame may be overwritten by the next step once everything above has run
            VT_CUDA(vt_launch_flag_set(consumed, step, st));
            mLaunches++;
        }
        if (dev_ptrs) return true;
        member_finish(m, live_out != nullptr, st);
        if (live_out) *live_out = m.live;
        return true;
    }

    // rays / fb: HOST frame arrays (flags = 0; a non-root process of a multi-process group only reads its own tiles of `rays`
    // and receives only its own tiles of `fb`), or with VT_TRAVERSE_DEVICE_PTRS (multi-process groups): rays = this rank's
    // COMPACT shard resident on its GPU (vt_group_shard gives its size), fb = full-frame device image, complete on rank 0 —
    // then everything is enqueued on `stream` and the call does not synchronise.
    void RenderDiffuseWave(const vt_ray *rays, uint64_t n, uint32_t spp, uint64_t seed, float weight, float *fb, uint64_t *live_out,
                           uint32_t flags, cudaStream_t stream) {
        if (live_out) *live_out = 0;
        if (n == 0) return;
        if (spp == 0) throw std::runtime_error("vt_group_render_diffuse_wave: spp must be positive");
        const bool dev_ptrs = (flags & VT_TRAVERSE_DEVICE_PTRS) != 0;
        if (dev_ptrs && live_out) throw std::runtime_error("vt_group_render_diffuse_wave: live_out must be NULL with device pointers (the call does not synchronise)");
        ShardGeom g0;
        g0.n = n, g0.tile = FrameTile(), g0.world = mWorld;
        if (!mMultiProcess) {
            if (dev_ptrs) throw std::runtime_error("vt_group_render_diffuse_wave: device pointers need a multi-process group (one rank per GPU)");
            if (!rays || !fb) throw std::runtime_error("vt_group_render_diffuse_wave: rays and framebuffer must not be null");
            run_all([&](Member &m) {
                const ShardGeom g = g0.of(m.rank);
                member_render(m, g, rays, false, spp, seed, weight, fb, live_out != nullptr, nullptr);
                member_finish(m, live_out != nullptr, m.stream);
            });
            if (live_out)
                for (auto &m : mMembers) *live_out += m->live;
            return;
        }
        Member &m = *mMembers[0];
        const ShardGeom g = g0.of(m.rank);
        const bool root = m.rank == 0;
        if (!rays && g.local_records()) throw std::runtime_error("vt_group_render_diffuse_wave: rays must not be null");
        if (root && !fb) throw std::runtime_error("vt_group_render_diffuse_wave: rank 0 needs the framebuffer");
        VT_CUDA(cudaSetDevice(m.device));
        cudaStream_t st = dev_ptrs ? stream : m.stream;
        if (!dev_ptrs && mWorld > 1 && env_int("VT_GROUP_NO_GATHER", 0) != 0) {  // diagnosis only: every rank keeps its own tiles, nothing is gathered
            member_render(m, g, rays, false, spp, seed, weight, fb, live_out != nullptr, st);
            member_finish(m, live_out != nullptr, st);
            if (live_out) *live_out = m.live;
            return;
        }
        if (!dev_ptrs && (flags & VT_GROUP_SHARED_HOST_FRAME)) {
            // `fb` is the SAME host memory in every process (a shared mapping, pinned by each): every rank lands its own tiles over its
            // own PCIe link — no gather at all, N links instead of rank 0's one — and a one-byte ncclAllGather behind the copies is the
            // barrier that tells every rank the frame is complete
            if (!fb) throw std::runtime_error("vt_group_render_diffuse_wave: the shared framebuffer must not be null");
            const bool async = (flags & VT_GROUP_ASYNC) != 0;
            if (async && live_out) throw std::runtime_error("vt_group_render_diffuse_wave: live_out must be NULL with VT_GROUP_ASYNC");
            if (mSharedFrames.size() >= (async ? 2u : 1u)) throw std::runtime_error("vt_group_render_diffuse_wave: frames are still in flight (vt_group_wait_frame)");
            member_render(m, g, rays, false, spp, seed, weight, fb, live_out != nullptr, st, nullptr, nullptr, nullptr, nullptr, 0, 0, (mSharedFrameCount++ & 1) != 0, async);
            if (mWorld > 1) {
                NcclApi &nccl = NcclApi::get();
                if (!nccl.AllGather) throw std::runtime_error("ncclAllGather not found");
                m.header.ensure(std::max<size_t>(256, mWorld));
                VT_NCCL(nccl.AllGather(m.header.p + m.rank, m.header.p, 1, ncclUint8, m.comm, st));
                mLaunches++;
            }
            if (async) {  // the frame is awaited by vt_group_wait_frame: this rank's copies and the barrier behind them are all on `st`
                cudaEvent_t done;
                VT_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming | cudaEventBlockingSync));
                VT_CUDA(cudaEventRecord(done, st));
                mSharedFrames.push_back(done);
                return;
            }
            member_finish(m, live_out != nullptr, st);
            if (live_out) *live_out = m.live;
            return;
        }
        if (mWorld > 1) {
            const char *gather = std::getenv("VT_GROUP_GATHER");
            if (!(gather && std::string(gather) == "nccl")) {  // default: K4 stores into rank 0's frame over NVLink (peer memory)
                const int slot = (flags & VT_GROUP_FRAME_SLOT1) ? 1 : 0;
                if (slot && !dev_ptrs) throw std::runtime_error("vt_group_render_diffuse_wave: VT_GROUP_FRAME_SLOT1 needs device pointers (host-pointer calls are synchronous)");
                if (slot && env_int("VT_GROUP_DEV_LANES", 1) > 1) throw std::runtime_error("vt_group_render_diffuse_wave: frame slots and VT_GROUP_DEV_LANES > 1 exclude each other");
                if (render_peer(m, g0, rays, dev_ptrs, spp, seed, weight, fb, live_out, stream, slot)) return;
                if (slot) throw std::runtime_error("vt_group_render_diffuse_wave: VT_GROUP_FRAME_SLOT1 needs the peer-memory frame");
            }
        }
        const bool pipelined = !dev_ptrs && mWorld > 1 && env_int("VT_GROUP_PIPELINE", 1) != 0;
        if (pipelined) {
            // host pointers, several ranks: every finished chunk is shipped at once — rank r sends its tiles of the chunk, rank 0
            // receives them into its staging area and downloads them — while the lanes go on tracing the next chunk, so only the last
            // chunk's gather + download is exposed instead of the whole frame's (rank 0 lands 12 bytes per pixel through ONE PCIe link)
            NcclApi &nccl = NcclApi::get();
            const ShardGeom sched = g0.of(0);
            std::vector<uint64_t> stage_off(mWorld, 0);
            if (root) {
                uint64_t total = 0;
                for (uint32_t r = 1; r < mWorld; r++) stage_off[r] = total, total += g0.of(r).local_records();
                m.fb_stage.ensure(std::max<uint64_t>(1, total * 3));
            }
            int ev = 0;
            auto ship = [&](uint64_t s0, uint64_t s1, cudaStream_t lane) {
                cudaEvent_t done = m.lane_done[ev++ % 8];
                VT_CUDA(cudaEventRecord(done, lane));
                VT_CUDA(cudaStreamWaitEvent(m.stream, done, 0));
                VT_NCCL(nccl.GroupStart());
                if (root) {
                    for (uint32_t r = 1; r < mWorld; r++) {
                        const ShardGeom gr = g0.of(r);
                        const uint64_t a = std::min(s0, gr.local_tiles()), b = std::min(s1, gr.local_tiles());
                        if (b > a)
                            VT_NCCL(nccl.Recv(m.fb_stage.p + (stage_off[r] + a * gr.tile) * 3, (gr.local_records(b) - a * gr.tile) * 3, ncclFloat32, (int)r, m.comm, m.stream));
                    }
                } else {
                    const uint64_t a = std::min(s0, g.local_tiles()), b = std::min(s1, g.local_tiles());
                    if (b > a) VT_NCCL(nccl.Send(m.fb_local.p + a * g.tile * 3, (g.local_records(b) - a * g.tile) * 3, ncclFloat32, 0, m.comm, m.stream));
                }
                VT_NCCL(nccl.GroupEnd());
                mLaunches++;
                if (root)
                    for (uint32_t r = 1; r < mWorld; r++) {
                        const ShardGeom gr = g0.of(r);
                        const uint64_t a = std::min(s0, gr.local_tiles()), b = std::min(s1, gr.local_tiles());
                        copy_tiles(gr, a, b, 3 * sizeof(float), m.fb_stage.p + stage_off[r] * 3, fb, true, cudaMemcpyDeviceToHost, m.stream);
                    }
            };
            member_render(m, g, rays, false, spp, seed, weight, fb, live_out != nullptr, st, &sched, ship);
            member_finish(m, live_out != nullptr, st);
            if (live_out) *live_out = m.live;
            return;
        }
        // own shard; with host pointers the shard image also goes straight to this process's host frame while later chunks still run
        member_render(m, g, rays, dev_ptrs, spp, seed, weight, (!dev_ptrs && fb) ? fb : nullptr, live_out != nullptr, st);
        if (mWorld > 1) {
            NcclApi &nccl = NcclApi::get();
            if (root) {
                uint64_t total = 0;
                for (uint32_t r = 1; r < mWorld; r++) total += g0.of(r).local_records();
                m.fb_stage.ensure(std::max<uint64_t>(1, total * 3));
            }
            VT_NCCL(nccl.GroupStart());
            if (root) {
                uint64_t off = 0;
                for (uint32_t r = 1; r < mWorld; r++) {
                    const uint64_t c = g0.of(r).local_records();
                    if (c) VT_NCCL(nccl.Recv(m.fb_stage.p + off * 3, c * 3, ncclFloat32, (int)r, m.comm, st));
                    off += c;
                }
            } else if (g.local_records()) {
                VT_NCCL(nccl.Send(m.fb_local.p, g.local_records() * 3, ncclFloat32, 0, m.comm, st));
            }
            VT_NCCL(nccl.GroupEnd());
            mLaunches++;
            if (root) {  // de-interleave the received shards into the frame: tile-strided copies, device -> host or device -> device
                uint64_t off = 0;
                for (uint32_t r = 1; r < mWorld; r++) {
                    const ShardGeom gr = g0.of(r);
                    copy_tiles(gr, 0, gr.local_tiles(), 3 * sizeof(float), m.fb_stage.p + off * 3, fb, true,
                               dev_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st);
                    off += gr.local_records();
                }
            }
        }
        if (dev_ptrs) {
            if (root) copy_tiles(g, 0, g.local_tiles(), 3 * sizeof(float), m.fb_local.p, fb, true, cudaMemcpyDeviceToDevice, st);
            return;
        }
        member_finish(m, live_out != nullptr, st);
        if (live_out) *live_out = m.live;
    }

    // the oldest frame begun with VT_GROUP_ASYNC is complete in the shared host frame on every rank
    void WaitFrame() {
        if (mSharedFrames.empty()) throw std::runtime_error("vt_group_wait_frame: no frame in flight");
        Member &m = *mMembers[0];
        VT_CUDA(cudaSetDevice(m.device));
        cudaEvent_t done = mSharedFrames.front();
        mSharedFrames.pop_front();
        const cudaError_t e = cudaEventSynchronize(done);
        cudaEventDestroy(done);
        VT_CUDA(e);
    }

    // sum of per-rank device buffers on rank 0 (sample-index sharding: every rank holds a partial image of the whole frame)
    void ReduceDevice(float *buf, uint64_t count, cudaStream_t stream) {
        if (!mMultiProcess) throw std::runtime_error("vt_group_reduce_device: multi-process groups only");
        if (mWorld == 1 || count == 0) return;
        Member &m = *mMembers[0];
        VT_CUDA(cudaSetDevice(m.device));
        NcclApi &nccl = NcclApi::get();
        if (!nccl.Reduce) throw std::runtime_error("ncclReduce not found");
        VT_NCCL(nccl.Reduce(buf, buf, count, ncclFloat32, ncclSum, 0, m.comm, stream));
        mLaunches++;
    }

    // every rank contributes bytes_per_rank bytes at buf + rank * bytes_per_rank; afterwards every rank holds all of buf.  What a
    // sample-index-sharded frame does with its primary rays: each rank uploads 1 / N of them over its own PCIe link and the rest
    // arrives over NVLink instead of every rank pulling the whole array through PCIe switches it shares with its neighbours.
    void AllGatherDevice(void *buf, uint64_t bytes_per_rank, cudaStream_t stream) {
        if (!mMultiProcess) throw std::runtime_error("vt_group_all_gather_device: multi-process groups only");
        if (mWorld == 1 || bytes_per_rank == 0) return;
        Member &m = *mMembers[0];
        VT_CUDA(cudaSetDevice(m.device));
        NcclApi &nccl = NcclApi::get();
        if (!nccl.AllGather) throw std::runtime_error("ncclAllGather not found");
        VT_NCCL(nccl.AllGather(static_cast<const char *>(buf) + (uint64_t)m.rank * bytes_per_rank, buf, bytes_per_rank, ncclUint8, m.comm, stream));
        mLaunches++;
    }

    ShardGeom Geom(uint64_t n, uint32_t rank) const {
        ShardGeom g;
        g.n = n, g.tile = FrameTile(), g.world = mWorld, g.rank = rank;
        return g;
    }
};

}  // namespace vt

// ================================================================================================ C ABI
struct vt_group {
    vt::Group impl;
    vt_group(const int *devices, int n) : impl(devices, n) {}
    vt_group(int device, int rank, int world, const uint8_t id[128]) : impl(device, rank, world, id) {}
};

extern "C" {

int vt_group_unique_id(uint8_t id[128]) {
    VT_TRY
    if (!id) throw std::runtime_error("null argument");
    ncclUniqueId uid;
    VT_NCCL(vt::NcclApi::get().GetUniqueId(&uid));
    std::memcpy(id, &uid, 128);
    return 0;
    VT_CATCH(1)
}

vt_group *vt_group_create(const int *devices, int n) {
    VT_TRY
    return new vt_group(devices, n);
    VT_CATCH(nullptr)
}

vt_group *vt_group_create_rank(int device, int rank, int world, const uint8_t id[128]) {
    VT_TRY
    return new vt_group(device, rank, world, id);
    VT_CATCH(nullptr)
}

void vt_group_destroy(vt_group *g) { delete g; }

int vt_group_size(const vt_group *g) { return g ? (int)g->impl.World() : 0; }
int vt_group_rank(const vt_group *g) { return g ? (int)g->impl.Rank() : -1; }
int vt_group_local_members(const vt_group *g) { return g ? (int)g->impl.LocalMembers() : 0; }

vt_accel *vt_group_accel(vt_group *g, int local_member) { return (g && local_member >= 0) ? g->impl.Accel((size_t)local_member) : nullptr; }

int vt_group_populate(vt_group *g, const vt_scene *scene) {
    VT_TRY
    if (!g) throw std::runtime_error("null argument");
    g->impl.Populate(scene);
    return 0;
    VT_CATCH(1)
}

int vt_group_shard(const vt_group *g, uint64_t n, int rank, uint64_t *tile, uint64_t *local_count) {
    VT_TRY
    if (!g || rank < 0 || rank >= (int)g->impl.World()) throw std::runtime_error("vt_group_shard: bad argument");
    const vt::ShardGeom s = g->impl.Geom(n, (uint32_t)rank);
    if (tile) *tile = s.tile;
    if (local_count) *local_count = s.local_records();
    return 0;
    VT_CATCH(1)
}

int vt_shard_geometry(uint64_t n, int world, int rank, uint64_t tile, uint64_t *local_count, uint64_t *local_tiles) {
    VT_TRY
    if (world <= 0 || rank < 0 || rank >= world || tile == 0) throw std::runtime_error("vt_shard_geometry: bad argument");
    vt::ShardGeom s;
    s.n = n, s.tile = tile, s.world = (uint32_t)world, s.rank = (uint32_t)rank;
    if (local_count) *local_count = s.local_records();
    if (local_tiles) *local_tiles = s.local_tiles();
    return 0;
    VT_CATCH(1)
}

int vt_group_traverse(vt_group *g, const vt_ray *rays, uint64_t n, vt_hit *hits, vt_attr *attrs, uint32_t flags) {
    VT_TRY
    if (!g) throw std::runtime_error("null argument");
    g->impl.Traverse(rays, n, hits, attrs, flags);
    return 0;
    VT_CATCH(1)
}

int vt_group_render_diffuse_wave(vt_group *g, const vt_ray *rays, uint64_t n, uint32_t spp, uint64_t seed, float weight,
                                 float *framebuffer_rgb, uint64_t *live_out, uint32_t flags, void *stream) {
    VT_TRY
    if (!g) throw std::runtime_error("null argument");
    g->impl.RenderDiffuseWave(rays, n, spp, seed, weight, framebuffer_rgb, live_out, flags, (cudaStream_t)stream);
    return 0;
    VT_CATCH(1)
}

int vt_group_reduce_device(vt_group *g, float *buf, uint64_t count, void *stream) {
    VT_TRY
    if (!g) throw std::runtime_error("null argument");
    g->impl.ReduceDevice(buf, count, (cudaStream_t)stream);
    return 0;
    VT_CATCH(1)
}

int vt_group_all_gather_device(vt_group *g, void *buf, uint64_t bytes_per_rank, void *stream) {
    VT_TRY
    if (!g) throw std::runtime_error("null argument");
    g->impl.AllGatherDevice(buf, bytes_per_rank, (cudaStream_t)stream);
    return 0;
    VT_CATCH(1)
}

int vt_group_wait_frame(vt_group *g) {
    VT_TRY
    if (!g) throw std::runtime_error("null argument");
    g->impl.WaitFrame();
    return 0;
    VT_CATCH(1)
}

uint64_t vt_group_launch_count(const vt_group *g) { return g ? g->impl.Launches() : 0; }

}  // extern "C"
