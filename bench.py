#!/usr/bin/env python
"""bench.py — headline benchmark of the accel:Traverse hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[2], the one the metric is quoted on): a 5 005 460-triangle closed
scene (1582^2-quad displaced terrain inside a room), one 1920x1080 pinhole primary wave and 4 spp
cosine-sampled diffuse bounce rays per hit.  One "step" = the whole frame:

    K1 closest-hit(primary) -> K2 TraceResult -> K3 bounce-ray generation -> K1 closest-hit(bounce) -> K4 framebuffer

Rays are counted individually (primary + spawned bounce rays).  `value` is measured with every input
already resident in HBM (CUDA events on the launching stream); `e2e` is the same frame through the
C ABI with HOST (pinned) buffers, host<->device copies inside the timed region.

N > 1 (one process per GPU under torchrun) is STRONG scaling of that one frame through the native group
(vt_group_*, vistrace_b200/csrc/vt_group.cu): rank 0 builds the hierarchy, its device image is
ncclBroadcast to the other GPUs, the frame is cut into tiles dealt round-robin to the ranks so that no ray
is traced twice, and the framebuffer shards are gathered on rank 0 over NVLink (ncclSend / ncclRecv) —
the image equals the single-GPU image bit for bit.  The weak-scaling figure of round 1 (every rank traces
the whole frame for its own samples, one ncclReduce) is kept as the side field `weak`.

`--impl reference` times the UNMODIFIED reference (oracle/_ref/libvt_ref.so: its own PLOC + LeafCollapser
hierarchy, SingleRayTraverser, TriangleBackfaceCull::intersect, TraceResult) on ALL host cores over the
same scene and the same kind of rays, rank 0 only, whatever N is.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# torchrun exports OMP_NUM_THREADS=1 for multi-rank launches; the host side of this engine (triangle set-up, hierarchy
# build, flatten) and the reference arm are OpenMP code, so the thread count is fixed before any OpenMP runtime loads:
#   * the reference arm runs on rank 0 alone and gets EVERY host core at every N (its `cores` must not depend on N);
#   * our arm: rank 0 is the only rank that builds (the image is broadcast), so it gets every core too; the other ranks
#     keep a share for their own ray generation.
_world = int(os.environ.get("WORLD_SIZE", "1"))
_rank = int(os.environ.get("RANK", "0"))
_is_reference = any(a == "reference" or a == "--impl=reference" for a in sys.argv[1:])
if _is_reference or (_world > 1 and os.environ.get("OMP_NUM_THREADS", "1") == "1"):
    _cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(_cores if (_is_reference or _rank == 0) else max(1, _cores // _world))

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT, SPP = 1920, 1080, 4
QUADS = 1582
CAMERA = ((0.0, -330.0, 200.0), (0.0, 0.0, 10.0))
METRIC = "Mrays/s closest-hit (primary+diffuse)"
WORKLOAD = f"config3: {2 * QUADS * QUADS + 12}-tri closed terrain scene, {WIDTH}x{HEIGHT} primary + {SPP} spp cosine diffuse bounce"
NODE_BYTES = {"exact": 64, "compact": 32, "quad": 64}  # bytes one traversal step fetches, per node layout
# identical in both arms (the driver compares the two lines' config objects)
CONFIG = {"workload": WORKLOAD, "frame": f"{WIDTH}x{HEIGHT}", "spp": SPP,
          "l2": "no explicit flush: one step streams ~0.6 GB of ray/hit/attribute buffers and walks a 0.5 GB hierarchy, both > 126 MB L2"}
TRI_BYTES, RAY_BYTES, HIT_BYTES = 64, 32, 16
_OUT = sys.stdout


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_scene():
    from vistrace_b200 import scenes

    return scenes.scene_terrain_closed(QUADS)


def primary_rays():
    from vistrace_b200 import scenes

    return scenes.pinhole_rays(WIDTH, HEIGHT, *CAMERA)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed regions (resident and e2e) run.

    ONE nvidia-smi process for the whole run, started before the first timed region and killed after the last (the recipe's
    "start before, kill after"): its start-up initialises NVML, which takes driver locks for several hundred milliseconds — started
    at the top of a 56 ms host-driven region (20 e2e steps) it slowed the region it was meant to observe (e2e 2.55 -> 2.82 ms per
    step, session r4g).  A region begins only after the first sample has arrived; only samples taken inside a region are summarised."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thread, self.windows, self.failed = index, [], None, None, [], False

    def start(self):
        if self.proc or self.failed:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc, self.failed = None, True
            return
        t_end = time.perf_counter() + 5.0
        while not self.rows and self.proc.poll() is None and time.perf_counter() < t_end:  # NVML is up once the first line is out
            time.sleep(0.01)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def __enter__(self):  # re-enterable: one window per timed region
        self.start()
        self.windows.append([time.perf_counter(), None])
        return self

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.06)  # a region shorter than the sampling period still gets its sample (the GPU is busy until the sync before this)
        self.windows[-1][1] = time.perf_counter()

    def close(self):
        if self.proc:
            self.proc.terminate()
            self.thread.join(timeout=2)
            self.proc = None

    def summary(self):
        self.close()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, r in self.rows:
            if not any(w0 <= t <= (w1 if w1 is not None else t) for w0, w1 in self.windows):
                continue
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except (KeyError, ValueError):
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_ncu():
    """Per-launch counters of the dominant kernel from the committed ncu capture (profiles/traffic.json): DRAM bytes, L2
    (lts) bytes and warp instructions executed — quantities only a profiler can see; the kernel time they are divided by
    is measured live."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path))
        except ValueError:
            pass
    return {}


def measured_l2_peak():
    """L2 -> SM bandwidth for the traversal kernel's access shape (32 divergent 64-byte records per warp, L2-resident set),
    measured on this pool's B200 by tools/ubench/l2bw.cu (profiles/r2_l2bw.json)."""
    path = os.path.join(ROOT, "profiles", "r2_l2bw.json")
    try:
        res = json.load(open(path))["results"]
        best = max(r["gbs"] for r in res if r["shape"] == "gather" and r["record_bytes"] == 64 and r["set_mb"] <= 112)
        return float(best), "measured (profiles/r2_l2bw.json: 64-byte record gather, L2-resident set)"
    except (OSError, KeyError, ValueError):
        return 9250.0, "fallback (148 SMs x 1 sector/clk x 1.965 GHz)"


# ------------------------------------------------------------------------------------- reference arm
def cpu_reference_sample(scene, rays, bounce, reps=2):
    """Time the reference's own CPU path on one step's rays: primary traversal + TraceResult, bounce traversal.
    Returns the checker and its hit buffers too, so the same run doubles as the parity check of the timed step."""
    import oracle

    kind = "reference" if oracle.available("reference") else "port"
    t0 = time.time()
    if kind == "reference":
        cpu = oracle.CpuScene(scene, "reference", build_bvh=True)  # PLOC + LeafCollapser, source/objects/AccelStruct.cpp:762-770
    else:
        import vistrace_b200 as vt

        cpu = oracle.CpuScene(scene, "port", build_bvh=False)
        cpu.set_bvh(*vt.build_bvh(scene))
    build_s = time.time() - t0
    best = float("inf")
    for _ in range(reps):
        a = cpu.traverse(rays, want_attrs=True)
        b = cpu.traverse(bounce)
        best = min(best, a["seconds"] + b["seconds"])
    n = len(rays) + len(bounce)
    return {"kind": kind, "cores": cpu.max_threads, "mrays": n / best / 1e6, "seconds": best, "rays": n, "build_s": build_s,
            "cpu": cpu, "primary_hits": a["hits"], "bounce_hits": b["hits"]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    from vistrace_b200 import abi, scenes

    if not (oracle.available("reference") or oracle.available("port")):
        print(json.dumps({"impl": "reference", "unavailable": "neither oracle/_ref/libvt_ref.so nor oracle/libvt_oracle.so is built"}), file=_OUT, flush=True)
        return 0
    scene = make_scene()
    rays = primary_rays()
    kind = "reference" if oracle.available("reference") else "port"
    log(f"[reference] building the {kind} hierarchy over {scene.n_tris} triangles ...")
    if kind == "reference":
        cpu = oracle.CpuScene(scene, "reference", build_bvh=True)
    else:
        import vistrace_b200 as vt

        cpu = oracle.CpuScene(scene, "port", build_bvh=False)
        cpu.set_bvh(*vt.build_bvh(scene))
    # bounded sample of the step: every `stride`-th pixel, all of its spp bounce rays
    stride = max(1, args.ref_stride)
    sub = rays[::stride]
    first = cpu.traverse(sub, want_attrs=True)
    bounce, _ = scenes.bounce_rays(first["attrs"], spp=SPP, key=7)
    n_step = len(sub) + len(bounce)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        cpu.traverse(sub, want_attrs=True)
        cpu.traverse(bounce)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = n_step / (ms * 1e-3) / 1e6
    sample = f"every {stride}th pixel of the {WIDTH}x{HEIGHT} frame ({len(sub)} primary rays incl. TraceResult) + their {len(bounce)} bounce rays per step"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": CONFIG,
        "details": {"hierarchy": "PLOC + LeafCollapser (reference build)" if kind == "reference" else "product builder",
                    "omp_threads": cpu.max_threads, "host_cpus": os.cpu_count()},
        "cpu_baseline": {"value": round(value, 3), "unit": "Mrays/s", "cores": cpu.max_threads, "kind": kind, "sample": sample},
        "e2e": {"value": round(value, 3), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), file=_OUT, flush=True)
    return 0


# ------------------------------------------------------------------------------------------ our arm
class ResidentFrame:
    """The five-kernel step over a device-resident frame (or any batch of primary rays) on ONE GPU, through the C ABI."""

    def __init__(self, accel, torch, dev, rays, use_queue=True):
        self.accel, self.torch, self.n = accel, torch, len(rays)
        n = self.n

        def dev_bytes(nbytes):
            return torch.empty(max(1, nbytes), dtype=torch.uint8, device=dev)

        self.d_rays = torch.from_numpy(np.ascontiguousarray(rays).view(np.uint8).reshape(-1).copy()).to(dev)
        self.d_hits, self.d_attrs = dev_bytes(n * 16), dev_bytes(n * 128)
        self.d_brays, self.d_bhits = dev_bytes(n * SPP * 32), dev_bytes(n * SPP * 16)
        self.d_fb = torch.zeros(n * 3, dtype=torch.float32, device=dev)
        # ray queue: K3 lists the slots that received a bounce ray, K1 visits only those (VT_BENCH_QUEUE=0: trace every slot)
        self.use_queue = use_queue
        self.d_queue, self.d_qcount = dev_bytes(n * SPP * 4), torch.zeros(1, dtype=torch.int64, device=dev)
        self.stream = torch.cuda.current_stream()
        self.sh = self.stream.cuda_stream
        self.k1_events = []

    def primary(self):
        self.accel.traverse_device(self.d_rays.data_ptr(), self.n, self.d_hits.data_ptr(), self.d_attrs.data_ptr(), stream=self.sh)  # K1 + K2

    def bounce(self, seed, before_k1=None):
        a, n = self.accel, self.n
        if self.use_queue:
            a.bounce_rays_queued_device(self.d_attrs.data_ptr(), n, SPP, seed, self.d_brays.data_ptr(), self.d_queue.data_ptr(),
                                        self.d_qcount.data_ptr(), self.d_bhits.data_ptr(), stream=self.sh)     # K3 (+ queue, miss records)
            if before_k1:
                before_k1()
            a.traverse_queued_device(self.d_brays.data_ptr(), self.d_queue.data_ptr(), self.d_qcount.data_ptr(), n * SPP,
                                     self.d_bhits.data_ptr(), stream=self.sh)                                  # K1 (dominant)
        else:
            a.bounce_rays_device(self.d_attrs.data_ptr(), n, SPP, seed, self.d_brays.data_ptr(), stream=self.sh)
            if before_k1:
                before_k1()
            a.traverse_device(self.d_brays.data_ptr(), n * SPP, self.d_bhits.data_ptr(), stream=self.sh)

    def step(self, seed, weight, timed=False):
        torch = self.torch
        self.primary()
        e0 = e1 = None
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.bounce(seed, (lambda: e0.record(self.stream)) if timed else None)
        if timed:
            e1.record(self.stream)
            self.k1_events.append((e0, e1))
        self.accel.accumulate_sky_device(self.d_attrs.data_ptr(), self.d_bhits.data_ptr(), self.n, SPP, weight, self.d_fb.data_ptr(), stream=self.sh)  # K4

    def live_bounce_rays(self):
        """Bounce rays one step spawns (identical work every step up to the RNG seed): counted once, outside the timed region."""
        from vistrace_b200 import abi

        self.primary()
        self.torch.cuda.synchronize()
        attrs = np.frombuffer(self.d_attrs.cpu().numpy().tobytes(), abi.ATTR)[: self.n]
        return int(((attrs["prim"] != abi.VT_MISS) & ((attrs["flags"] & abi.VT_ATTR_HIT_SKY) == 0)).sum()) * SPP


def config5_side(args, torch, dist, dev, world, rank, local_rank, sync_all, max_over_ranks):
    """BASELINE.json configs[4], the world-scale path-tracing workload, as a SIDE measurement next to the headline: a 19 995 044-
    triangle scene (terrain + 143 props), 3840x2160, 16 samples per pixel, per sample primary + shadow + 3 diffuse bounces each with
    a shadow ray (vt_accel_trace_paths: wave compaction).  N > 1 shards by SAMPLE INDEX: rank r traces samples r, r + N, ... of
    every pixel and the per-rank images are summed on rank 0 with one ncclReduce (strong scaling: the frame is fixed).
    `value` = device-resident (primary rays in HBM), `e2e` = host rays up on every rank, host image down on rank 0."""
    import vistrace_b200 as vt
    from vistrace_b200 import scenes

    W5, H5, SPP5, BOUNCES = 3840, 2160, 16, 3
    t0 = time.time()
    scene = scenes.scene_terrain_closed(2980, n_props=143) if rank == 0 else None
    rays = scenes.pinhole_rays(W5, H5, *CAMERA)
    n = len(rays)
    gen_s = time.time() - t0
    t0 = time.time()
    if world == 1:
        group, accel = None, vt.Accel(local_rank).populate(scene)
    else:
        uid = torch.from_numpy(vt.group_unique_id() if rank == 0 else np.zeros(128, np.uint8)).to(dev)
        dist.broadcast(uid, src=0)
        group = vt.Group(device=local_rank, rank=rank, world=world, unique_id=uid.cpu().numpy()).populate(scene)
        accel = group.accel(0)
    populate_s = time.time() - t0
    n_tris = int(scene.n_tris) if scene is not None else None
    del scene
    sun = np.array((0.3, 0.2, 0.93), np.float32)
    sun = sun / np.linalg.norm(sun)
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream
    h_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1)).pin_memory()
    d_rays = torch.empty(n * 32, dtype=torch.uint8, device=dev)
    d_rays.copy_(h_rays)
    d_fb = torch.zeros(n * 3, dtype=torch.float32, device=dev)
    h_fb = torch.empty(n * 3, dtype=torch.float32).pin_memory() if rank == 0 else None
    my_samples = list(range(rank, SPP5, world))
    fif5 = 2 if (len(my_samples) > 1 and os.environ.get("VT_BENCH_FRAMES_IN_FLIGHT", "2") != "1") else 1
    stream2 = torch.cuda.Stream() if fif5 > 1 else None
    d_fb2 = torch.zeros(n * 3, dtype=torch.float32, device=dev) if fif5 > 1 else None
    counts = accel.trace_paths_device(d_rays.data_ptr(), n, BOUNCES, sun, (1, 1, 1), 1, 0.0, d_fb.data_ptr(), want_counts=True, stream=sh)
    rays_per_sample = int(counts.sum())  # the same for every sample up to the random directions of the bounces (counted once, outside the timed region)

    # host path at N > 1: every rank uploads only ITS 1 / N of the ray array (padded to equal chunks) and the chunks are exchanged over
    # NVLink with one ncclAllGather — 8 ranks pulling the whole 265 MB array through shared PCIe switches cost 13 ms per frame
    chunk = ((n + world - 1) // world) * 32
    d_rays_padded = torch.empty(chunk * world, dtype=torch.uint8, device=dev) if world > 1 else None
    h_rays_padded = None
    if world > 1:
        h_rays_padded = torch.zeros(chunk * world, dtype=torch.uint8).pin_memory()
        h_rays_padded[: n * 32] = h_rays

    def frame(it, host):
        rays_ptr = d_rays.data_ptr()
        if host and world == 1:
            d_rays.copy_(h_rays, non_blocking=True)
        elif host:
            d_rays_padded[rank * chunk:(rank + 1) * chunk].copy_(h_rays_padded[rank * chunk:(rank + 1) * chunk], non_blocking=True)
            group.all_gather_device(d_rays_padded.data_ptr(), chunk, stream=sh)
            rays_ptr = d_rays_padded.data_ptr()
        d_fb.zero_()
        if fif5 > 1:
            # two samples in flight (the handle's two path-scratch slots on two streams, one framebuffer each): a sample's small late
            # waves and launch tails run under the other sample's large early waves
            d_fb2.zero_()
            fork = torch.cuda.Event()
            fork.record(stream)
            stream2.wait_event(fork)
            for j, smp in enumerate(my_samples):
                k = j % 2
                accel.trace_paths_device(rays_ptr, n, BOUNCES, sun, (1, 1, 1), 1000 * it + smp, 1.0 / SPP5, (d_fb2 if k else d_fb).data_ptr(),
                                         stream=(stream2 if k else stream).cuda_stream, slot=k)
            join = torch.cuda.Event()
            join.record(stream2)
            stream.wait_event(join)
            d_fb.add_(d_fb2)
        else:
            for smp in my_samples:
                accel.trace_paths_device(rays_ptr, n, BOUNCES, sun, (1, 1, 1), 1000 * it + smp, 1.0 / SPP5, d_fb.data_ptr(), stream=sh)
        if group is not None:
            group.reduce_device(d_fb.data_ptr(), n * 3, stream=sh)
        if host and rank == 0:
            h_fb.copy_(d_fb, non_blocking=True)
        if host:
            stream.synchronize()

    frames = 2
    frame(0, False)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for it in range(frames):
        frame(1 + it, False)
    e1.record(stream)
    sync_all()
    ms = max_over_ranks(e0.elapsed_time(e1)) / frames
    frame(0, True)
    sync_all()
    t0 = time.perf_counter()
    for it in range(frames):
        frame(1 + it, True)
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0)) / frames
    rays_per_frame = rays_per_sample * SPP5
    out = {"workload": f"config5: {n_tris}-tri terrain + props, {W5}x{H5}, {SPP5} spp, per sample primary + shadow + {BOUNCES} diffuse bounces each with a shadow ray (wave compaction)",
           "scaling": "strong", "sharding": "one GPU" if world == 1 else f"by sample index over {world} ranks, one ncclReduce of the {n * 12 // 1000000} MB image per frame",
           "rays_per_frame": rays_per_frame, "value": round(rays_per_frame / ms / 1e3, 2), "unit": "Mrays/s", "ms_per_frame": round(ms, 3),
           "e2e": {"value": round(rays_per_frame / e2e_ms / 1e3, 2), "ms_per_frame": round(e2e_ms, 3), "h2d_bytes_per_frame": n * 32 if world == 1 else chunk, "d2h_bytes_per_frame": n * 12,
                   "how": "host rays up, host image down" if world == 1 else "each rank uploads 1/N of the host rays, ncclAllGather, own samples, ncclReduce, rank 0 downloads the image"},
           "samples_in_flight": fif5, "rays_per_wave_of_one_sample": [int(c) for c in counts], "scene_generation_s": round(gen_s, 1), "populate_s": round(populate_s, 1)}
    if group is not None:
        group.close()
    else:
        accel.close()
    del d_rays, d_fb, d_fb2
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    import vistrace_b200 as vt
    from vistrace_b200 import abi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — vistrace_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from vistrace_b200 import shard

    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t.item())

    def pinned(nbytes):
        return torch.empty(nbytes, dtype=torch.uint8).pin_memory()

    t0 = time.time()
    rays = primary_rays()
    n = len(rays)
    scene = make_scene() if rank == 0 else None  # only the builder needs the triangles
    gen_s = time.time() - t0
    t0 = time.time()
    group = None
    if world == 1:
        accel = vt.Accel(local_rank)
        if args.builder == "ploc":
            accel.populate(scene, bvh=vt.build_bvh_ploc(scene))
        else:
            accel.populate(scene)
    else:
        # the launcher's part of vt_group_create_rank: hand rank 0's ncclUniqueId to every process
        uid = torch.from_numpy(vt.group_unique_id() if rank == 0 else np.zeros(128, np.uint8)).to(dev)
        dist.broadcast(uid, src=0)
        group = vt.Group(device=local_rank, rank=rank, world=world, unique_id=uid.cpu().numpy())
        if args.builder == "ploc" and rank == 0:
            os.environ["VT_BUILDER"] = "ploc"
        group.populate(scene)  # rank 0 builds ONCE; the device image reaches the other GPUs by ncclBroadcast over NVLink
        accel = group.accel(0)
    populate_s = time.time() - t0
    setup_s = gen_s + populate_s
    # pin this process to the CPUs next to its GPU only NOW: the build (rank 0, every host core) is done, the pinned
    # staging buffers of the e2e path are allocated below and get first-touched on the GPU's NUMA node
    numa = shard.bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if rank == 0:
        st = accel.stats()
        log(f"[bench] scene {st['n_tris']} tris, {st['node_count']} nodes, {st['device_bytes'] / 1e6:.0f} MB resident, scene generation {gen_s:.1f}s, "
            f"populate (ingest + build + flatten + upload{' + ncclBroadcast' if world > 1 else ''}) {populate_s:.1f}s")
    layout = accel.layout
    shard_tile = group.shard(n)[0] if group else None
    use_queue = os.environ.get("VT_BENCH_QUEUE", "1") != "0"
    seed0 = 1000
    steps, warmup = args.steps, args.warmup
    clocks = ClockSampler(local_rank)
    clocks.start()  # NVML start-up happens here, not at the top of a timed region
    launch_count = (lambda: group.launch_count) if group else (lambda: accel.launch_count)
    launches = 0

    # ------------------------------------------------------------------ resident (`value`)
    # FRAMES_IN_FLIGHT consecutive steps are issued on as many streams (each frame its own buffers): every launch of K1 ends with
    # ~0.1 ms in which a few hundred long rays hold the kernel while most SMs idle (profiles/r2_tail_sharing.md); under the next
    # frame's bulk that time is not lost.  Every step still does all of its work; VT_BENCH_FRAMES_IN_FLIGHT=1 is the round-1 schedule.
    fif = max(1, min(2, int(os.environ.get("VT_BENCH_FRAMES_IN_FLIGHT", "2"))))
    main = torch.cuda.current_stream()
    streams = [torch.cuda.Stream() for _ in range(fif)] if fif > 1 else [main]

    def run_timed(step_fn):
        """K steps alternating over the streams, bracketed by events on the main stream that all of them wait for / are joined into."""
        sync_all()
        t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_begin.record(main)
        for st_ in streams:
            if st_ is not main:
                st_.wait_event(t_begin)
        for it in range(steps):
            step_fn(it, it % fif)
        for st_ in streams:
            if st_ is not main:
                ev = torch.cuda.Event()
                ev.record(st_)
                main.wait_event(ev)
        t_end.record(main)
        sync_all()
        return t_begin.elapsed_time(t_end)

    if world == 1:
        frames = []
        for st_ in streams:
            with torch.cuda.stream(st_):
                frames.append(ResidentFrame(accel, torch, dev, rays, use_queue))
        frame = frames[0]
        live = frame.live_bounce_rays()
        stream = frame.stream
        for it in range(max(warmup, fif)):
            frames[it % fif].step(seed0 + it, 1.0)
        sync_all()
        l0 = launch_count()
        for f_ in frames:
            f_.d_fb.zero_()
        with clocks:
            total_ms = run_timed(lambda it, k: frames[k].step(seed0 + warmup + it, 1.0 / max(1, steps)))
        launches += launch_count() - l0
        # the dominant kernel's own duration: a few single-frame steps right after the timed region (inside it, kernels of two frames
        # share the SMs and an event pair around one of them would time the mix)
        sync_all()
        for it in range(5):
            frame.step(seed0 + warmup + steps + it, 0.0, timed=True)
        sync_all()
        k1_ms = float(np.mean([a.elapsed_time(b) for a, b in frame.k1_events]))
    else:
        # STRONG scaling: this rank's tiles of the frame are resident on its GPU (compact shard); one call per step enqueues
        # K1 K2 K3 K1 K4 over the shard — K4's stores ARE the gather: straight into rank 0's frame over NVLink peer memory — and the
        # copy of the completed frame on rank 0.  Two frames in flight = the group's two frame slots on two streams.
        idx = group.shard_indices(n)
        d_shard = torch.from_numpy(np.ascontiguousarray(rays[idx]).view(np.uint8).reshape(-1).copy()).to(dev)
        d_fbs = [torch.zeros(n * 3, dtype=torch.float32, device=dev) for _ in range(fif)]
        stream = main
        _, live_local = group.render_diffuse_wave(rays, SPP, seed=seed0, weight=1.0)  # also the first warm-up of the host path
        live = sum_over_ranks(live_local)

        def group_step(it, k):
            group.render_diffuse_wave_device(d_shard.data_ptr(), n, SPP, seed0 + it, 1.0, d_fbs[k].data_ptr(), stream=streams[k].cuda_stream, slot=k)

        for it in range(max(warmup, fif)):
            group_step(it, it % fif)
        sync_all()
        l0 = launch_count()
        with clocks:
            total_ms = run_timed(lambda it, k: group_step(warmup + it, k))
        launches += launch_count() - l0
        k1_ms = None
    rays_per_step = n + live  # whole job: the frame is traced once, whatever N is
    ms_per_step = max_over_ranks(total_ms) / steps
    value = rays_per_step / (ms_per_step * 1e-3) / 1e6

    # ------------------------------------------------------------------ e2e: HOST rays in, HOST RGBFFF image out
    h_rays_t, h_fb_t = pinned(n * 32), pinned(n * 12)
    h_rays = h_rays_t.numpy().view(abi.RAY)
    h_rays[:] = rays
    h_fb = h_fb_t.numpy().view(np.float32).reshape(n, 3)
    e2e_steps = max(1, steps)
    e2e_async = False
    if world == 1:
        h2d_step, d2h_step = n * 32, n * 12
        if fif > 1:
            # two frames in flight through the split call: every step still uploads its rays and lands its own image in host memory
            e2e_async = True
            e2e_call = ("vt_accel_render_diffuse_wave_begin / _wait, two frames in flight: host rays in, host RGBFFF framebuffer out "
                        "(each step its own pinned framebuffer)")
            h_fb2_t = pinned(n * 12)
            h_fbs = [h_fb, h_fb2_t.numpy().view(np.float32).reshape(n, 3)]
        else:
            e2e_call = "vt_accel_render_diffuse_wave: host rays in, host RGBFFF framebuffer out"

        def e2e_step(it):
            return accel.render_diffuse_wave(h_rays, SPP, seed=seed0 + it, weight=1.0, out=h_fb)[1]
    else:
        # the host frame is ONE buffer shared by the processes of the node (POSIX shm, pinned by each): every rank's GPU lands its own
        # tiles through its own PCIe link, a one-byte ncclAllGather is the barrier; VT_BENCH_E2E_GATHER=1 measures the round-2 path
        # instead (finished pixels stored into rank 0's device frame over NVLink, rank 0 downloads the whole image through one link)
        shared_frame = None
        shared_frames = []
        if os.environ.get("VT_BENCH_E2E_GATHER", "0") == "0":
            ok = 1
            names = [f"vt_bench_frame_{os.environ.get('MASTER_PORT', '0')}_{k}" for k in range(fif)]
            try:
                if rank == 0:
                    shared_frames = [shard.SharedPinnedFrame(nm, n * 12, create=True) for nm in names]
            except (OSError, RuntimeError) as e:
                log(f"[bench] shared host frame unavailable on rank 0 ({e}): falling back to the NVLink gather")
                ok = 0
            dist.barrier()
            try:
                if rank != 0:
                    shared_frames = [shard.SharedPinnedFrame(nm, n * 12, create=False) for nm in names]
            except (OSError, RuntimeError):
                ok = 0
            ok_t = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)  # every rank must have the mapping, or nobody uses it
            if int(ok_t.item()) == 0:
                for fr in shared_frames:
                    fr.close()
                shared_frames = []
        if shared_frames:
            shared_frame = shared_frames[0]
            h_fbs = [fr.array(np.float32, (n, 3)) for fr in shared_frames]
            h_fb = h_fbs[0]
            e2e_async = fif > 1
            e2e_call = ("vt_group_render_diffuse_wave(VT_GROUP_SHARED_HOST_FRAME), one process per GPU: each rank uploads the host rays of its own "
                        "tiles, traces them and lands its tiles of the RGBFFF frame in host memory shared by all ranks (its own PCIe link); "
                        "complete on every rank after a one-byte ncclAllGather"
                        + ("; two frames in flight (VT_GROUP_ASYNC + vt_group_wait_frame), each step its own shared frame" if fif > 1 else ""))
            d2h_step = len(idx) * 12
        else:
            e2e_call = ("vt_group_render_diffuse_wave (one process per GPU): each rank uploads the host rays of its own tiles, traces them, "
                        "finished pixels stored into rank 0's frame over NVLink (peer memory), rank 0 downloads the frame")
            d2h_step = n * 12 if rank == 0 else 0                # rank 0 lands the whole image
        h2d_step = len(idx) * 32                                 # this rank's tiles (max over ranks reported below)

        def e2e_step(it):
            return group.render_diffuse_wave(h_rays, SPP, seed=seed0 + it, weight=1.0, out=h_fb, want_live=False, shared_frame=shared_frame is not None)[1]
    for it in range(min(2, warmup)):
        e2e_step(it)
    if e2e_async:
        begin = (lambda it: accel.render_diffuse_wave_begin(h_rays, SPP, seed0 + warmup + it, 1.0, h_fbs[it % 2])) if world == 1 else \
                (lambda it: group.render_diffuse_wave_begin(h_rays, SPP, seed0 + warmup + it, 1.0, h_fbs[it % 2]))
        wait = accel.render_diffuse_wave_wait if world == 1 else group.wait_frame
        for it in range(2):  # warm-up of the two-in-flight schedule (second staging buffer, tile sizes of its own)
            begin(-2 + it)
        wait()
        wait()
    sync_all()
    l0 = launch_count()
    with clocks:
        sync_all()
        t0 = time.perf_counter()
        if e2e_async:
            for it in range(e2e_steps):
                if it >= 2:
                    wait()
                begin(it)
            for it in range(min(2, e2e_steps)):
                wait()
        else:
            for it in range(e2e_steps):
                e2e_step(warmup + it)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
    launches += launch_count() - l0
    if world > 1 and shared_frame is not None:
        sync_all()
        for fr in shared_frames:
            fr.close()
    e2e_s = max_over_ranks(e2e_s)
    e2e_value = rays_per_step / (e2e_s / e2e_steps) / 1e6
    h2d_step, d2h_step = sum_over_ranks(h2d_step), sum_over_ranks(d2h_step)  # whole job, like `value`: bytes all ranks move per step

    # ------------------------------------------------------------------ side fields
    weak = None
    if world > 1:
        # round 1's figure, kept for continuity: every rank traces the WHOLE frame for its own samples, one ncclReduce per step
        frame = ResidentFrame(accel, torch, dev, rays, use_queue)
        wsteps = max(3, min(steps, 10))
        for it in range(2):
            frame.step(seed0 + it + 7919 * rank, 1.0 / world)
        sync_all()
        t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_begin.record(frame.stream)
        for it in range(wsteps):
            frame.step(seed0 + 2 + it + 7919 * rank, 1.0 / (world * wsteps))
            group.reduce_device(frame.d_fb.data_ptr(), n * 3, stream=frame.sh)
        t_end.record(frame.stream)
        sync_all()
        weak_ms = max_over_ranks(t_begin.elapsed_time(t_end)) / wsteps
        weak = {"value": round(world * rays_per_step / (weak_ms * 1e-3) / 1e6, 2), "unit": "Mrays/s", "ms_per_step": round(weak_ms, 4),
                "what": "every rank traces the whole frame for its own 4 samples per pixel; per-rank images summed on rank 0 with one ncclReduce"}
    hits_variant = None
    if world == 1:
        # the same wave returning every hit record instead of the image (vt_accel_trace_diffuse_wave)
        h_hits_t, h_bhits_t = pinned(n * 16), pinned(n * SPP * 16)
        out = {"hits": h_hits_t.numpy().view(abi.HIT), "bounce_hits": h_bhits_t.numpy().view(abi.HIT)}
        hit_steps = max(1, min(5, steps))
        accel.trace_diffuse_wave(h_rays, SPP, seed=seed0, out=out)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for it in range(hit_steps):
            res = accel.trace_diffuse_wave(h_rays, SPP, seed=seed0 + warmup + it, out=out)
        torch.cuda.synchronize()
        e2e_hits_ms = 1e3 * (time.perf_counter() - t0) / hit_steps
        hits_variant = {"call": "vt_accel_trace_diffuse_wave", "value": round((n + int(res["live_bounce"])) / (e2e_hits_ms * 1e-3) / 1e6, 2),
                        "ms_per_step": round(e2e_hits_ms, 3), "d2h_bytes_per_step": n * 16 + n * SPP * 16}

    # ------------------------------------------------------------------ rank 0: roofline, CPU baseline, parity
    roof, cpu_line, parity = None, None, None
    if world == 1 and rank == 0:
        brays_host = np.frombuffer(frame.d_brays.cpu().numpy().tobytes(), abi.RAY)[: n * SPP]
        bhits_host = np.frombuffer(frame.d_bhits.cpu().numpy().tobytes(), abi.HIT)[: n * SPP]
        phits_host = np.frombuffer(frame.d_hits.cpu().numpy().tobytes(), abi.HIT)[:n]
        live_mask = brays_host["tmax"] >= 0
        bounce_live = np.ascontiguousarray(brays_host[live_mask])
        # S (node visits) and I (triangle tests) per ray for the algorithmic byte count: the engine's own
        # SingleRayTraverser::Statistics counters (vt_accel_traverse_stats) over the bounce wave, outside the timed region
        n_steps, n_tests = accel.traverse_stats(frame.d_brays.data_ptr(), n * SPP)
        n_live = max(1, int(live_mask.sum()))
        S, I = n_steps / n_live, n_tests / n_live
        peak, peak_src = measured_peak()
        n_masked = int((~live_mask).sum())
        node_bytes = NODE_BYTES[layout]
        if use_queue:  # K1 reads one 4-byte queue entry per live ray and never touches a masked slot
            algo_bytes = n_live * (node_bytes * S + TRI_BYTES * I + RAY_BYTES + HIT_BYTES + 4)
        else:
            algo_bytes = n_live * (node_bytes * S + TRI_BYTES * I + RAY_BYTES + HIT_BYTES) + n_masked * (RAY_BYTES + HIT_BYTES)
        achieved = algo_bytes / (k1_ms * 1e-3) / 1e9
        ncu = recorded_ncu()
        same_layout = ncu.get("layout") == layout
        sm_hz = (clocks.summary().get("sm_mhz") or 1965.0) * 1e6
        # three roofs for the same launch: algorithmic bytes against the HBM copy peak (the contract's figure: > physical, the tree is
        # served by L1/L2), bytes that actually crossed L2 -> SM against the measured L2 gather peak, and warp instructions issued
        # against the issue rate of 148 SMs x 4 schedulers.  `bound` names the largest fraction: the unit the kernel is closest to.
        l2_peak, l2_src = measured_l2_peak()
        l2 = issue = None
        if same_layout and ncu.get("k_traverse_bounce_lts_bytes_per_launch"):
            b = float(ncu["k_traverse_bounce_lts_bytes_per_launch"])
            l2 = {"bytes_per_launch": int(b), "achieved": round(b / (k1_ms * 1e-3) / 1e9, 1), "peak": l2_peak, "unit": "GB/s",
                  "frac": round(b / (k1_ms * 1e-3) / 1e9 / l2_peak, 4), "peak_source": l2_src, "bytes_source": "ncu lts__t_bytes.sum (profiles/traffic.json)"}
        if same_layout and ncu.get("k_traverse_bounce_warp_inst_per_launch"):
            wi = float(ncu["k_traverse_bounce_warp_inst_per_launch"])
            ipeak = 148 * 4 * sm_hz / 1e9
            issue = {"warp_inst_per_launch": int(wi), "achieved": round(wi / (k1_ms * 1e-3) / 1e9, 1), "peak": round(ipeak, 1), "unit": "G warp-inst/s",
                     "frac": round(wi / (k1_ms * 1e-3) / 1e9 / ipeak, 4), "peak_source": "148 SMs x 4 schedulers x SM clock under load",
                     "thread_inst_per_warp_inst": ncu.get("k_traverse_bounce_thread_inst_per_warp_inst"),
                     "inst_source": "ncu smsp__inst_executed.sum (profiles/traffic.json)"}
        dram = ncu.get("k_traverse_bounce_dram_bytes_per_launch") if same_layout else None
        fracs = {"hbm (physical DRAM traffic)": (dram / (k1_ms * 1e-3) / 1e9 / peak) if dram else 0.0,
                 "l2": l2["frac"] if l2 else 0.0, "issue": issue["frac"] if issue else 0.0}
        bound = max(fracs, key=fracs.get) if any(fracs.values()) else "hbm"
        hbm_algo = {"achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": int(algo_bytes),
                    "what": "SURVEY section 8d's figure: node_bytes * S + 64 * I + ray + hit bytes per ray over the kernel time, against the HBM copy peak; it "
                            "exceeds what DRAM physically moves (`traffic`) because the hierarchy is served by L1/L2"}
        common = {"traffic": dram, "kernel": "k_traverse (closest hit, bounce wave" + (", ray queue)" if use_queue else ")"), "kernel_ms": round(k1_ms, 4),
                  "kernel_ms_source": "CUDA events around the launch, mean over 5 single-frame steps right after the timed region" if fif > 1 else "CUDA events around the launch inside the timed region",
                  "node_layout": layout, "node_bytes": node_bytes, "node_visits_per_ray": round(S, 2), "tri_tests_per_ray": round(I, 2),
                  "physical_hbm_frac": round(fracs["hbm (physical DRAM traffic)"], 4), "hbm_algorithmic": hbm_algo, "l2": l2, "issue": issue}
        if bound == "issue":  # `bound` = the roof the kernel is closest to, with that roof's own achieved / peak
            roof = {"bound": "issue", "achieved": issue["achieved"], "peak": issue["peak"], "unit": issue["unit"], "frac": issue["frac"], **common}
        elif bound == "l2":
            roof = {"bound": "l2", "achieved": l2["achieved"], "peak": l2["peak"], "unit": "GB/s", "frac": l2["frac"], **common}
        else:  # no ncu counters for this layout: only the algorithmic figure can be computed live
            roof = {"bound": "hbm", "achieved": hbm_algo["achieved"], "peak": peak, "unit": "GB/s", "frac": hbm_algo["frac"], **common}
        if not args.no_cpu:
            try:
                cpu = cpu_reference_sample(scene, rays, bounce_live)
                cpu_line = {"value": round(cpu["mrays"], 3), "unit": "Mrays/s", "cores": cpu["cores"], "kind": cpu["kind"],
                            "sample": f"one full step on the host: {len(rays)} primary rays incl. TraceResult + {len(bounce_live)} bounce rays, best of 2, "
                                      f"{'reference PLOC+LeafCollapser hierarchy' if cpu['kind'] == 'reference' else 'product hierarchy'} (build {cpu['build_s']:.1f}s excluded)"}
                # parity of the timed step's own rays: the engine's hit records against the checker's (its own tree), every ray
                from vistrace_b200.report import classify_hits, merge_reports

                chk = cpu["cpu"].tri_intersect
                parity = merge_reports([classify_hits(phits_host, cpu["primary_hits"], rays, chk),
                                        classify_hits(np.ascontiguousarray(bhits_host[live_mask]), cpu["bounce_hits"], bounce_live, chk)])
                parity["checker"] = f"{cpu['kind']} traversal on its own hierarchy, every ray of one step (primary + bounce)"
            except Exception as e:  # the checker is optional for the number itself
                log(f"[bench] cpu_baseline leg unavailable: {e}")
    # ------------------------------------------------------------------ side measurement: BASELINE configs[4] (every rank takes part)
    config5 = None
    if args.config5 != "off":
        if group:
            group.close()
            group = None
        else:
            accel.close()
        torch.cuda.empty_cache()
        try:
            config5 = config5_side(args, torch, dist, dev, world, rank, local_rank, sync_all, max_over_ranks)
        except Exception as e:  # a side field must never cost the headline line
            log(f"[bench] config5 side measurement failed on rank {rank}: {e}")
    if rank != 0:
        if group:
            group.close()
        if world > 1:
            dist.destroy_process_group()
        return 0

    line = {
        "metric": METRIC, "value": round(value, 2), "unit": "Mrays/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": CONFIG,
        "details": {"rays_per_step": rays_per_step, "numa_node": numa, "setup_s": round(setup_s, 1), "populate_s": round(populate_s, 2),
                    "parallelism": "one GPU" if world == 1 else f"hierarchy built once and replicated (ncclBroadcast), frame cut into {shard_tile}-pixel tiles dealt round-robin to {world} ranks, "
                                   "no ray traced twice, every rank's shading kernel stores its finished pixels straight into rank 0's frame over NVLink (peer memory)",
                    "frames_in_flight": fif,
                    "schedule": (f"{fif} consecutive steps in flight on {fif} streams (each its own buffers / frame slot): a launch's tail of long rays runs under the next "
                                 "frame's bulk; every step does all of its work" if fif > 1 else "one step at a time on one stream"),
                    "hierarchy": ("reference-identical PLOC + LeafCollapser (vt_build_bvh_ploc)" if args.builder == "ploc" else "product builder (binned SAH)")
                                 + f", {layout} node layout"},
        "clocks": clocks.summary(),
        "e2e": {"value": round(e2e_value, 2), "unit": "Mrays/s", "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": d2h_step,
                "ms_per_step": round(1e3 * e2e_s / e2e_steps, 3), "call": e2e_call, "all_hit_records_variant": hits_variant},
        "gpu_launches": int(launches),
    }
    if weak:
        line["weak"] = weak
    if config5:
        line["config5"] = config5
    if roof:
        line["roofline"] = roof
    if cpu_line:
        line["cpu_baseline"] = cpu_line
    if parity:
        line["parity"] = parity
    print(json.dumps(line), file=_OUT, flush=True)
    if group:
        group.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    # stdout carries exactly ONE JSON line: libraries that chat on fd 1 (NCCL prints its version there) go to stderr
    global _OUT
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-stride", type=int, default=1, help="reference arm: trace every n-th pixel per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--config5", default="on", choices=["on", "off"], help="side measurement of BASELINE configs[4] (20 M triangles, 4K, 16 spp path waves)")
    ap.add_argument("--builder", default="product", choices=["product", "ploc"],
                    help="hierarchy of our arm: the product's binned-SAH builder (default, the headline) or the bit-identical "
                         "restatement of the reference's PLOC + LeafCollapser build (the tree the reference arm traverses)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
