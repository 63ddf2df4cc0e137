"""Every BASELINE.json config at full size on one GPU: device-resident timings (CUDA events), the
quantised-vs-exact layout agreement at that size, and the reference's CPU path on a bounded sample.

usage: python tools/config_sweep.py [--configs 1,2,3,4,5] [--cpu-rays 200000] [--out gpurun_out/sweep.jsonl]

Not the bench (bench.py measures the headline config) and not a test: it produces the per-config table in
profiles/.  Rays are counted individually (primary + every secondary ray spawned).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vistrace_b200 as vt  # noqa: E402
from vistrace_b200 import abi, scenes  # noqa: E402

SUN = np.array((0.3, 0.2, 0.93), np.float32)
SUN = (SUN / np.sqrt((SUN * SUN).sum(dtype=np.float32))).astype(np.float32)


def dev(nbytes):
    return torch.empty(int(nbytes), dtype=torch.uint8, device="cuda")


def to_dev(a):
    return torch.from_numpy(a.view(np.uint8).reshape(-1)).cuda()


def to_host(t, dtype, n):
    return np.frombuffer(t[: n * dtype.itemsize].cpu().numpy().tobytes(), dtype)


class Wave:
    """Device buffers for one wave of n rays (+ spp secondary rays per ray)."""

    def __init__(self, n, spp=1):
        self.n, self.spp = n, spp
        self.hits, self.attrs = dev(n * 16), dev(n * 128)
        self.srays, self.shits = dev(n * 32), dev(n * 16)
        self.brays, self.bhits = dev(n * spp * 32), dev(n * spp * 16)
        self.queue, self.qcount = dev(n * spp * 4), torch.zeros(1, dtype=torch.int64, device="cuda")  # ray queue of the secondary wave


def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def live_count(t, n):
    r = to_host(t, abi.RAY, n)
    return int((r["tmax"] >= 0).sum())


def agreement(scene, bvh, rays_list):
    """Quantised layouts against the exact layout on the same hierarchy: records that differ / of which ties."""
    out = {}
    want = None
    for layout in ("exact", "quad", "compact"):
        a = vt.Accel(0, layout=layout).populate(scene, bvh=bvh)
        got = [a.traverse(r) for r in rays_list]
        a.close()
        if layout == "exact":
            want = got
            continue
        diff = ties = total = 0
        for g, w in zip(got, want):
            d = (g.view(np.uint32).reshape(-1, 4) != w.view(np.uint32).reshape(-1, 4)).any(1)
            both = (g["prim"][d] != abi.VT_MISS) & (w["prim"][d] != abi.VT_MISS)
            tie = both & (np.abs(g["t"][d].astype(np.float64) - w["t"][d]) <= 1e-6 * np.abs(w["t"][d].astype(np.float64)))
            diff, ties, total = diff + int(d.sum()), ties + int(tie.sum()), total + len(g)
        out[layout] = {"rays": total, "records_differing": diff, "of_which_ties": ties}
    return out


def cpu_reference(scene, samples, n_rays):
    """The reference's own CPU path (oracle/_ref: its PLOC + LeafCollapser tree) on a bounded sample of each ray kind."""
    try:
        import oracle

        if not oracle.available("reference"):
            return None
        t0 = time.time()
        cpu = oracle.CpuScene(scene, "reference", build_bvh=True)
        build_s = time.time() - t0
        secs = rays = 0
        for r in samples:
            sub = r[:: max(1, len(r) // n_rays)][:n_rays]
            secs += cpu.traverse(sub)["seconds"]
            rays += len(sub)
        return {"Mrays_s": round(rays / secs / 1e6, 2), "threads": cpu.max_threads, "sample_rays": rays, "build_s": round(build_s, 1)}
    except Exception as e:  # the table is still useful without this column
        return {"error": str(e)[:200]}


def run_config(cfg, args):
    t0 = time.time()
    s = torch.cuda.current_stream().cuda_stream
    if cfg == 1:
        name = "config1: 100k-tri height field, 1920x1080 primary"
        scene, (W, H), cam = scenes.scene_heightfield(224), (1920, 1080), ((0, -80, 60), (0, 0, 5))
    elif cfg == 2:
        name = "config2: 1M-tri scene of 256 skinned props, 1920x1080 primary + shadow"
        scene, (W, H), cam = scenes.scene_props_skinned(256, 63, 31, 64), (1920, 1080), ((0, -95, 40), (0, 0, 10))
    elif cfg == 3:
        name = "config3: 5M-tri closed terrain, 1920x1080 primary + 4 spp diffuse (the bench workload)"
        scene, (W, H), cam = scenes.scene_terrain_closed(1582), (1920, 1080), ((0.0, -330.0, 200.0), (0.0, 0.0, 10.0))
    elif cfg == 4:
        name = "config4: 1M-tri alpha-tested foliage (500k cards, two 256^2 VTFs), 1920x1080 primary + attrs + 1 bounce"
        scene, (W, H), cam = scenes.scene_foliage(n_cards=500000, tex_size=256, ground_quads=64), (1920, 1080), ((0, -48, 20), (0, 0, 8))
    else:
        name = "config5: 20M-tri terrain + props, 3840x2160, per sample primary + shadow + 3 bounces each with a shadow ray"
        scene, (W, H), cam = scenes.scene_terrain_closed(2980, n_props=143), (3840, 2160), ((0.0, -330.0, 200.0), (0.0, 0.0, 10.0))
    gen_s = time.time() - t0
    t0 = time.time()
    bvh = vt.build_bvh(scene)
    build_s = time.time() - t0
    rays = scenes.pinhole_rays(W, H, *cam)
    n = len(rays)
    accel = vt.Accel(0).populate(scene, bvh=bvh)
    d_rays = to_dev(rays)
    spp = 4 if cfg == 3 else 1
    w = Wave(n, spp)
    res = {"config": cfg, "name": name, "n_tris": scene.n_tris, "layout": accel.layout, "ray_queue": not args.no_queue, "gen_s": round(gen_s, 1), "build_s": round(build_s, 1),
           "device_MB": round(accel.stats()["device_bytes"] / 1e6)}
    P = lambda t: t.data_ptr()
    Q = not args.no_queue  # ray queue: generators list the live slots, traversal visits only those

    def shadow_wave(src_attrs, wv):
        if Q:
            accel.shadow_rays_queued_device(P(src_attrs), n, SUN, P(wv.srays), P(wv.queue), P(wv.qcount), P(wv.shits), stream=s)
            accel.traverse_queued_device(P(wv.srays), P(wv.queue), P(wv.qcount), n, P(wv.shits), any_hit=True, stream=s)
        else:
            accel.shadow_rays_device(P(src_attrs), n, SUN, P(wv.srays), stream=s)
            accel.traverse_device(P(wv.srays), n, P(wv.shits), any_hit=True, stream=s)

    def bounce_wave_k(src, k_spp, seed, d_hits, d_attrs=None):
        if Q:
            accel.bounce_rays_queued_device(P(src.attrs), n, k_spp, seed, P(src.brays), P(src.queue), P(src.qcount), P(d_hits), stream=s)
            accel.traverse_queued_device(P(src.brays), P(src.queue), P(src.qcount), n * k_spp, P(d_hits), P(d_attrs) if d_attrs is not None else None,
                                         stream=s)
        else:
            accel.bounce_rays_device(P(src.attrs), n, k_spp, seed, P(src.brays), stream=s)
            accel.traverse_device(P(src.brays), n * k_spp, P(d_hits), P(d_attrs) if d_attrs is not None else None, stream=s)

    stages = {}
    samples = [rays]
    # primary (+ TraceResult)
    stages["primary K1"] = (timed(lambda: accel.traverse_device(P(d_rays), n, P(w.hits), stream=s)), n)
    stages["primary K1+K2"] = (timed(lambda: accel.traverse_device(P(d_rays), n, P(w.hits), P(w.attrs), stream=s)), n)
    total_ms, total_rays = stages["primary K1+K2"][0], n
    if cfg in (2, 5):
        accel.shadow_rays_device(P(w.attrs), n, SUN, P(w.srays), stream=s)
        torch.cuda.synchronize()
        ls = live_count(w.srays, n)
        ms = timed(lambda: shadow_wave(w.attrs, w))
        stages["shadow K3b+K1 any-hit"] = (ms, ls)
        total_ms, total_rays = total_ms + ms, total_rays + ls
        sr = to_host(w.srays, abi.RAY, n)
        samples.append(sr[sr["tmax"] >= 0])
    if cfg in (3, 4):
        accel.bounce_rays_device(P(w.attrs), n, spp, 17, P(w.brays), stream=s)
        torch.cuda.synchronize()
        lb = live_count(w.brays, n * spp)
        ms = timed(lambda: bounce_wave_k(w, spp, 17, w.bhits))
        stages[f"bounce K3+K1 ({spp} spp)"] = (ms, lb)
        total_ms, total_rays = total_ms + ms, total_rays + lb
        br = to_host(w.brays, abi.RAY, n * spp)
        samples.append(br[br["tmax"] >= 0])
    if cfg == 5:
        # path waves: bounce k from the hits of wave k-1, each followed by its shadow rays; w2 ping-pongs with w
        w2 = Wave(n, 1)
        src, dst = w, w2
        for k in range(3):
            accel.bounce_rays_device(P(src.attrs), n, 1, 100 + k, P(src.brays), stream=s)
            torch.cuda.synchronize()
            lb = live_count(src.brays, n)

            def bounce_wave(src=src, dst=dst, k=k):
                bounce_wave_k(src, 1, 100 + k, dst.hits, dst.attrs)
                shadow_wave(dst.attrs, dst)

            ms = timed(bounce_wave, reps=2)
            ls = live_count(dst.srays, n)
            stages[f"bounce {k + 1}: K3+K1+K2, shadow K3b+K1 any-hit"] = (ms, lb + ls)
            total_ms, total_rays = total_ms + ms, total_rays + lb + ls
            if k == 0:
                br = to_host(src.brays, abi.RAY, n)
                samples.append(br[br["tmax"] >= 0])
            src, dst = dst, src
    res["stages"] = {k: {"ms": round(ms, 3), "rays": r, "Mrays_s": round(r / ms / 1e3, 1)} for k, (ms, r) in stages.items()}
    res["wave"] = {"ms": round(total_ms, 3), "rays": total_rays, "Mrays_s": round(total_rays / total_ms / 1e3, 1)}
    steps, tests = accel.traverse_stats(P(d_rays), n)
    res["primary_node_visits_tri_tests_per_ray"] = [round(steps / n, 2), round(tests / n, 2)]
    accel.close()
    del w
    torch.cuda.empty_cache()
    if not args.no_agreement:
        res["layout_agreement_vs_exact"] = agreement(scene, bvh, [rr[:: max(1, len(rr) // 2000000)] for rr in samples])
    if args.cpu_rays:
        res["cpu_reference"] = cpu_reference(scene, samples, args.cpu_rays)
        if res["cpu_reference"] and "Mrays_s" in res["cpu_reference"]:
            res["gpu_over_cpu"] = round(res["wave"]["Mrays_s"] / res["cpu_reference"]["Mrays_s"], 1)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4,5")
    ap.add_argument("--cpu-rays", type=int, default=200000)
    ap.add_argument("--no-agreement", action="store_true")
    ap.add_argument("--no-queue", action="store_true", help="trace every slot of a secondary wave instead of the generator's ray queue")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    for cfg in (int(c) for c in args.configs.split(",")):
        r = run_config(cfg, args)
        line = json.dumps(r)
        print(line, flush=True)
        if args.out:
            with open(args.out, "a") as f:
                f.write(line + "\n")


if __name__ == "__main__":
    main()
