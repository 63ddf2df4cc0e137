"""Why small K1 launches are slow (what one rank of an 8-GPU strong-scaling run sees), measured on ONE GPU.

1. per-ray chain lengths (vt_accel_traverse_ray_stats) of rank 0's shard of the bench frame and of its bounce wave;
2. K1 over the shard with the longest rays removed (is the launch bound by its slowest rays?);
3. the device-resident group step over lane / grid / chunk settings.
usage: python tools/small_shard_probe.py [--world 8] [--quads 1582] [--part stats,tail,sweep] [--k1-only]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vistrace_b200 as vt  # noqa: E402
from vistrace_b200 import abi, scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--quads", type=int, default=1582)
ap.add_argument("--part", default="stats,tail,sweep")
ap.add_argument("--k1-only", action="store_true", help="three primary + bounce K1 launches over the shard and exit (for ncu)")
ap.add_argument("--reps", type=int, default=7)
args = ap.parse_args()
parts = set(args.part.split(","))
SPP = 4

scene = scenes.scene_terrain_closed(args.quads)
rays = scenes.pinhole_rays(1920, 1080, (0.0, -330.0, 200.0), (0.0, 0.0, 10.0))
group = vt.Group(device=0, rank=0, world=1).populate(scene)
accel = group.accel(0)
idx = vt.shard_indices(len(rays), args.world, 0, 8192)
shard = np.ascontiguousarray(rays[idx])
n = len(shard)
stream = torch.cuda.current_stream()
s = stream.cuda_stream


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()).cuda()


def empty(nbytes):
    return torch.empty(max(1, nbytes), dtype=torch.uint8, device="cuda")


def time_k1(d_rays, count, d_hits, reps=args.reps):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        accel.traverse_device(d_rays.data_ptr(), count, d_hits.data_ptr(), stream=s)
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


hits, attrs = accel.traverse(shard, want_attrs=True)
brays, live = accel.bounce_rays(attrs, SPP, seed=1003)
blive = np.ascontiguousarray(brays[brays["tmax"] >= 0])
d_shard, d_blive = dev(shard), dev(blive)
d_h1, d_h2 = empty(n * 16), empty(len(blive) * 16)

if args.k1_only:
    for _ in range(3):
        accel.traverse_device(d_shard.data_ptr(), n, d_h1.data_ptr(), stream=s)
        accel.traverse_device(d_blive.data_ptr(), len(blive), d_h2.data_ptr(), stream=s)
    torch.cuda.synchronize()
    sys.exit(0)

out = {"world": args.world, "primary_rays": n, "bounce_rays": int(len(blive))}
if "stats" in parts or "tail" in parts:
    ps, pt = accel.traverse_ray_stats(shard)
    bs, bt = accel.traverse_ray_stats(blive)
    for name, st, tt in (("primary", ps, pt), ("bounce", bs, bt)):
        rounds = st.astype(np.int64) + tt  # node steps + triangle tests: the ray's dependent chain
        out[name + "_chain"] = {"mean_steps": round(float(st.mean()), 2), "mean_tests": round(float(tt.mean()), 2),
                                **{f"p{q}": int(np.percentile(rounds, q)) for q in (50, 90, 99, 99.9)}, "max": int(rounds.max()),
                                "max_steps": int(st.max())}
if "tail" in parts:
    t_full_p, t_full_b = time_k1(d_shard, n, d_h1), time_k1(d_blive, len(blive), d_h2)
    out["k1_ms"] = {"primary": round(t_full_p, 4), "bounce": round(t_full_b, 4)}
    for name, src, st, tt, d_h in (("primary", shard, ps, pt, d_h1), ("bounce", blive, bs, bt, d_h2)):
        rounds = st.astype(np.int64) + tt
        for q in (99.9, 99, 95):
            keep = rounds <= np.percentile(rounds, q)
            sub = np.ascontiguousarray(src[keep])
            out["k1_ms"][f"{name}_without_top_{round(100 - q, 1)}pct"] = round(time_k1(dev(sub), len(sub), d_h), 4)
        order = np.argsort(-rounds, kind="stable")  # longest first: the tail starts at time zero
        out["k1_ms"][f"{name}_longest_first"] = round(time_k1(dev(src[order]), len(src), d_h), 4)
    # the same number of rays, every ray a copy of a median one: the launch without any tail
    med = shard[np.argsort(ps.astype(np.int64) + pt)[n // 2]]
    out["k1_ms"]["primary_all_median_ray"] = round(time_k1(dev(np.repeat(med[None], n, 0).reshape(-1)), n, d_h1), 4)
print(json.dumps(out), flush=True)

if "sweep" in parts:
    d_fb = torch.zeros(n * 3, dtype=torch.float32, device="cuda")
    h_rays_t = torch.empty(n * 32, dtype=torch.uint8).pin_memory()
    h_rays = h_rays_t.numpy().view(abi.RAY)
    h_rays[:] = shard
    h_fb = torch.empty(n * 12, dtype=torch.uint8).pin_memory().numpy().view(np.float32).reshape(n, 3)
    _, live = group.render_diffuse_wave(h_rays, SPP, seed=5, out=h_fb)
    steps = 20
    for lanes, ctas, chunks in [(1, 0, 1), (2, 6, 2), (2, 4, 2), (2, 4, 4), (4, 4, 4), (4, 2, 4), (4, 2, 8), (4, 3, 8), (8, 2, 8), (8, 1, 8), (8, 2, 16), (8, 1, 16), (8, 1, 32)]:
        os.environ["VT_GROUP_DEV_LANES"] = str(lanes)
        os.environ["VT_GROUP_DEV_CTAS_PER_SM"] = str(ctas)
        os.environ["VT_GROUP_DEV_CHUNKS"] = str(chunks)
        for _ in range(3):
            group.render_diffuse_wave_device(d_shard.data_ptr(), n, SPP, 5, 1.0, d_fb.data_ptr(), stream=s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for it in range(steps):
            group.render_diffuse_wave_device(d_shard.data_ptr(), n, SPP, 5 + it, 1.0, d_fb.data_ptr(), stream=s)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print(json.dumps({"world": args.world, "lanes": lanes, "ctas_per_sm": ctas, "chunks": chunks, "ms_per_step": round(ms, 4),
                          "Mrays_per_gpu": round((n + live) / ms / 1e3, 1)}), flush=True)
