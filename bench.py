#!/usr/bin/env python
"""bench.py — headline benchmark of the accel:Traverse hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[2], the one the metric is quoted on): a 5 005 460-triangle closed
scene (1582^2-quad displaced terrain inside a room), one 1920x1080 pinhole primary wave and 4 spp
cosine-sampled diffuse bounce rays per hit.  One "step" = the whole wave:

    K1 closest-hit(primary) -> K2 TraceResult -> K3 bounce-ray generation -> K1 closest-hit(bounce) -> K4 framebuffer

Rays are counted individually (primary + spawned bounce rays).  `value` is measured with every input
already resident in HBM (CUDA events on the launching stream); `e2e` is the same wave through the
C ABI with HOST (pinned) buffers, host<->device copies inside the timed region.  With N > 1 (one
process per GPU under torchrun) the hierarchy is replicated, every rank traces its own 4 samples per
pixel of the same frame (weak scaling, no data-path collective) and the per-rank framebuffers are
summed with ONE NCCL reduce per step.

`--impl reference` times the UNMODIFIED reference (oracle/_ref/libvt_ref.so: its own PLOC + LeafCollapser
hierarchy, SingleRayTraverser, TriangleBackfaceCull::intersect, TraceResult) on the host cores over the
same scene and the same kind of rays, rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# torchrun exports OMP_NUM_THREADS=1 for multi-rank launches; the host side of this engine (triangle set-up, hierarchy
# build, flatten) is OpenMP code, so give every rank its share of the host cores before any OpenMP runtime loads.
_world = int(os.environ.get("WORLD_SIZE", "1"))
if _world > 1 and os.environ.get("OMP_NUM_THREADS", "1") == "1":
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // _world))

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT, SPP = 1920, 1080, 4
QUADS = 1582
CAMERA = ((0.0, -330.0, 200.0), (0.0, 0.0, 10.0))
METRIC = "Mrays/s closest-hit (primary+diffuse)"
WORKLOAD = f"config3: {2 * QUADS * QUADS + 12}-tri closed terrain scene, {WIDTH}x{HEIGHT} primary + {SPP} spp cosine diffuse bounce"
NODE_BYTES = {"exact": 64, "compact": 32, "quad": 64}  # bytes one traversal step fetches, per node layout
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
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed regions (resident and e2e) run."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):  # re-enterable: rows accumulate over every timed region
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
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


def recorded_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/), or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get("k_traverse_bounce_dram_bytes_per_launch")
        except ValueError:
            pass
    return None


# ------------------------------------------------------------------------------------- reference arm
def cpu_reference_sample(scene, rays, bounce, reps=2):
    """Time the reference's own CPU path on one step's rays: primary traversal + TraceResult, bounce traversal."""
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
    return {"kind": kind, "cores": cpu.max_threads, "mrays": n / best / 1e6, "seconds": best, "rays": n, "build_s": build_s}


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
        "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "hierarchy": "PLOC + LeafCollapser (reference build)" if kind == "reference" else "product builder"},
        "cpu_baseline": {"value": round(value, 3), "unit": "Mrays/s", "cores": cpu.max_threads, "kind": kind, "sample": sample},
        "e2e": {"value": round(value, 3), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), file=_OUT, flush=True)
    return 0


# ------------------------------------------------------------------------------------------ our arm
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

    numa = shard.bind_to_gpu_numa_node(local_rank)  # before any pinned buffer or OpenMP thread exists
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    t0 = time.time()
    scene = make_scene()
    rays = primary_rays()
    n = len(rays)
    accel = vt.Accel(local_rank)
    build = vt.build_bvh_ploc if args.builder == "ploc" else vt.build_bvh
    if world > 1:  # the hierarchy is built once (rank 0) and replicated over NCCL; every GPU holds the whole scene
        accel.populate(scene, bvh=shard.replicate_bvh(build(scene) if rank == 0 else None, device=dev))
    elif args.builder == "ploc":
        accel.populate(scene, bvh=build(scene))
    else:
        accel.populate(scene)
    st = accel.stats()
    if rank == 0:
        log(f"[bench] scene {st['n_tris']} tris, {st['node_count']} nodes, {st['device_bytes'] / 1e6:.0f} MB resident, setup {time.time() - t0:.1f}s")

    def dev_bytes(nbytes):
        return torch.empty(nbytes, dtype=torch.uint8, device=dev)

    d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1)).to(dev)
    d_hits, d_attrs = dev_bytes(n * 16), dev_bytes(n * 128)
    d_brays, d_bhits = dev_bytes(n * SPP * 32), dev_bytes(n * SPP * 16)
    d_fb = torch.zeros(n * 3, dtype=torch.float32, device=dev)
    # ray queue: K3 lists the slots that received a bounce ray, K1 visits only those (VT_BENCH_QUEUE=0: trace every slot)
    use_queue = os.environ.get("VT_BENCH_QUEUE", "1") != "0"
    d_queue, d_qcount = dev_bytes(n * SPP * 4), torch.zeros(1, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream
    seed0 = 1000 + rank * 7919  # every rank traces its own samples of the frame

    ev_pairs = []

    def bounce_wave(seed, before_k1=None):
        if use_queue:
            accel.bounce_rays_queued_device(d_attrs.data_ptr(), n, SPP, seed, d_brays.data_ptr(), d_queue.data_ptr(), d_qcount.data_ptr(),
                                            d_bhits.data_ptr(), stream=sh)                                     # K3 (+ queue, miss records)
            if before_k1:
                before_k1()
            accel.traverse_queued_device(d_brays.data_ptr(), d_queue.data_ptr(), d_qcount.data_ptr(), n * SPP, d_bhits.data_ptr(), stream=sh)
        else:
            accel.bounce_rays_device(d_attrs.data_ptr(), n, SPP, seed, d_brays.data_ptr(), stream=sh)          # K3
            if before_k1:
                before_k1()
            accel.traverse_device(d_brays.data_ptr(), n * SPP, d_bhits.data_ptr(), stream=sh)                  # K1 (dominant)

    def step(it, timed):
        seed = seed0 + it
        accel.traverse_device(d_rays.data_ptr(), n, d_hits.data_ptr(), d_attrs.data_ptr(), stream=sh)          # K1 + K2
        e0 = e1 = None
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        bounce_wave(seed, (lambda: e0.record(stream)) if timed else None)
        if timed:
            e1.record(stream)
            ev_pairs.append((e0, e1))
        accel.accumulate_sky_device(d_attrs.data_ptr(), d_bhits.data_ptr(), n, SPP, 1.0 / (world * max(1, args.steps)), d_fb.data_ptr(), stream=sh)  # K4
        if world > 1:
            dist.reduce(d_fb, dst=0, op=dist.ReduceOp.SUM)  # the one collective: per-rank partial images -> rank 0

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # live bounce rays per step (identical work every step up to the RNG seed): count once, outside the timed region
    accel.traverse_device(d_rays.data_ptr(), n, d_hits.data_ptr(), d_attrs.data_ptr(), stream=sh)
    torch.cuda.synchronize()
    attrs_host = np.frombuffer(d_attrs.cpu().numpy().tobytes(), abi.ATTR)
    live = int(((attrs_host["prim"] != abi.VT_MISS) & ((attrs_host["flags"] & abi.VT_ATTR_HIT_SKY) == 0)).sum()) * SPP
    rays_per_step = n + live

    for it in range(args.warmup):
        step(it, False)
    sync_all()
    launches0 = accel.launch_count
    d_fb.zero_()
    with ClockSampler(local_rank) as clocks:
        sync_all()
        t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_begin.record(stream)
        for it in range(args.steps):
            step(args.warmup + it, True)
        t_end.record(stream)
        sync_all()
    total_ms = t_begin.elapsed_time(t_end)
    launches = accel.launch_count - launches0
    k1_ms = float(np.mean([a.elapsed_time(b) for a, b in ev_pairs]))
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * rays_per_step / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: the same step through the C ABI with HOST (pinned) buffers, copies inside the timed region.
    # The step's result is the framebuffer (as in the resident step above, which ends in K4): rays up, RGBFFF image down.
    def pinned(nbytes):
        return torch.empty(nbytes, dtype=torch.uint8).pin_memory()

    h_rays_t, h_fb_t = pinned(n * 32), pinned(n * 12)
    h_rays = h_rays_t.numpy().view(abi.RAY)
    h_rays[:] = rays
    h_fb = h_fb_t.numpy().view(np.float32).reshape(n, 3)
    e2e_steps = max(1, args.steps)
    if world == 1 or os.environ.get("VT_BENCH_E2E") == "tiled":
        e2e_call = "vt_accel_render_diffuse_wave: host rays in, host RGBFFF framebuffer out"
        h2d_step, d2h_step = n * 32, n * 12

        def e2e_step(it):
            return accel.render_diffuse_wave(h_rays, SPP, seed=seed0 + it, weight=1.0, out=h_fb)[1]
    else:
        # N GPUs: every rank uploads 1/N of the ray array and downloads 1/N of the finished image; the rest moves over
        # NVLink (one all_gather of the rays, one all_reduce of the partial images) — shard.ShardedFrame
        frame = shard.ShardedFrame(n, dev)
        h_rays_f32 = h_rays_t.view(torch.float32)
        e2e_call = "shard.ShardedFrame.step: 1/N of the host rays in per rank, all_gather, trace own samples, all_reduce, 1/N of the host RGBFFF image out per rank"
        h2d_step, d2h_step = frame.h2d_bytes, frame.d2h_bytes

        def e2e_step(it):
            def trace(d_rays_full):
                p = d_rays_full.data_ptr()
                accel.traverse_device(p, n, d_hits.data_ptr(), d_attrs.data_ptr(), stream=sh)
                bounce_wave(seed0 + it)
                d_fb.zero_()
                accel.accumulate_sky_device(d_attrs.data_ptr(), d_bhits.data_ptr(), n, SPP, 1.0 / world, d_fb.data_ptr(), stream=sh)
                return d_fb

            frame.step(h_rays_f32, trace)  # the finished chunk lands in the frame's own pinned buffer
            return live
    for it in range(min(2, args.warmup)):
        e2e_step(it)
    sync_all()
    launches_e2e0 = accel.launch_count
    with clocks:
        t0 = time.perf_counter()
        for it in range(e2e_steps):
            live_e2e = e2e_step(args.warmup + it)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
    launches += accel.launch_count - launches_e2e0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_rays = n + int(live_e2e)
    e2e_value = world * e2e_rays / (e2e_s / e2e_steps) / 1e6

    # ---- the same wave returning every hit record instead of the image (vt_accel_trace_diffuse_wave): reported beside
    # e2e at N = 1 (at N > 1 every rank would pull 166 MB per step through shared PCIe uplinks: not the sharded design)
    hits_variant = None
    if world == 1:
        h_hits_t, h_bhits_t = pinned(n * 16), pinned(n * SPP * 16)
        out = {"hits": h_hits_t.numpy().view(abi.HIT), "bounce_hits": h_bhits_t.numpy().view(abi.HIT)}
        hit_steps = max(1, min(5, args.steps))
        accel.trace_diffuse_wave(h_rays, SPP, seed=seed0, out=out)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for it in range(hit_steps):
            res = accel.trace_diffuse_wave(h_rays, SPP, seed=seed0 + args.warmup + it, out=out)
        torch.cuda.synchronize()
        e2e_hits_ms = 1e3 * (time.perf_counter() - t0) / hit_steps
        hits_variant = {"call": "vt_accel_trace_diffuse_wave", "value": round((n + int(res["live_bounce"])) / (e2e_hits_ms * 1e-3) / 1e6, 2),
                        "ms_per_step": round(e2e_hits_ms, 3), "d2h_bytes_per_step": n * 16 + n * SPP * 16}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (K1 over the bounce rays) + CPU baseline, rank 0 only
    brays_host = np.frombuffer(d_brays.cpu().numpy().tobytes(), abi.RAY)
    live_mask = brays_host["tmax"] >= 0
    bounce_live = brays_host[live_mask]
    cpu = None
    # S (node visits) and I (triangle tests) per ray for the algorithmic byte count: the engine's own
    # SingleRayTraverser::Statistics counters (vt_accel_traverse_stats) over the bounce wave, outside the timed region
    steps, tests = accel.traverse_stats(d_brays.data_ptr(), n * SPP)
    n_live_rays = max(1, int(live_mask.sum()))
    S, I = steps / n_live_rays, tests / n_live_rays
    try:
        if world == 1 and not args.no_cpu:
            cpu = cpu_reference_sample(scene, rays, bounce_live)
    except Exception as e:  # the checker is optional for the number itself
        log(f"[bench] cpu_baseline leg unavailable: {e}")
    peak, peak_src = measured_peak()
    roof = None
    if S is not None:
        n_live, n_masked = int(live_mask.sum()), int((~live_mask).sum())
        if use_queue:  # K1 reads one 4-byte queue entry per live ray and never touches a masked slot
            algo_bytes = n_live * (NODE_BYTES[accel.layout] * S + TRI_BYTES * I + RAY_BYTES + HIT_BYTES + 4)
        else:
            algo_bytes = n_live * (NODE_BYTES[accel.layout] * S + TRI_BYTES * I + RAY_BYTES + HIT_BYTES) + n_masked * (RAY_BYTES + HIT_BYTES)
        achieved = algo_bytes / (k1_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": recorded_traffic(), "kernel": "k_traverse (closest hit, bounce wave" + (", ray queue)" if use_queue else ")"), "kernel_ms": round(k1_ms, 4),
                "algorithmic_bytes_per_launch": int(algo_bytes), "node_layout": accel.layout, "node_bytes": NODE_BYTES[accel.layout],
                "node_visits_per_ray": round(S, 2), "tri_tests_per_ray": round(I, 2),
                "peak_source": peak_src}
    line = {
        "metric": METRIC, "value": round(value, 2), "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_step_per_gpu": rays_per_step, "parallelism": f"replicated hierarchy, {world} x ray/sample shard", "numa_node": numa,
                   "l2": "no explicit flush: one step streams ~0.6 GB of ray/hit/attribute buffers and walks a 0.5 GB hierarchy, both > 126 MB L2",
                   "hierarchy": ("reference-identical PLOC + LeafCollapser (vt_build_bvh_ploc)" if args.builder == "ploc" else "product builder (binned SAH)")
                                + f", {accel.layout} node layout"},
        "clocks": clocks.summary(),
        "e2e": {"value": round(e2e_value, 2), "unit": "Mrays/s", "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": d2h_step,
                "ms_per_step": round(1e3 * e2e_s / e2e_steps, 3), "call": e2e_call, "all_hit_records_variant": hits_variant},
        "gpu_launches": int(launches),
    }
    if roof:
        line["roofline"] = roof
    if cpu and "mrays" in cpu:
        line["cpu_baseline"] = {"value": round(cpu["mrays"], 3), "unit": "Mrays/s", "cores": cpu["cores"], "kind": cpu["kind"],
                                "sample": f"one full step on the host: {len(rays)} primary rays incl. TraceResult + {len(bounce_live)} bounce rays, best of 2, "
                                          f"{'reference PLOC+LeafCollapser hierarchy' if cpu['kind'] == 'reference' else 'product hierarchy'} (build {cpu['build_s']:.1f}s excluded)"}
    print(json.dumps(line), file=_OUT, flush=True)
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
