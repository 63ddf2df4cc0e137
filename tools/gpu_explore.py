"""GPU exploration harness (not a test, not the bench): times K1/K2 over knob settings.

usage: python tools/gpu_explore.py [--quads 1582] [--res 1920x1080] [--spp 4] [--knobs a=b,c=d;...]
Knobs are the VT_* environment variables read at populate time.
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


def to_dev(a):
    return torch.from_numpy(a.view(np.uint8).reshape(-1)).cuda()


def time_traverse(accel, d_rays, n, d_hits, reps=5, any_hit=False, d_attrs=None):
    stream = torch.cuda.current_stream()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        accel.traverse_device(d_rays.data_ptr(), n, d_hits.data_ptr(), None if d_attrs is None else d_attrs.data_ptr(), any_hit=any_hit, stream=stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def _morton3(q, bits):
    """Interleave three `bits`-bit integer columns of q (n x 3) into one Morton code."""
    code = np.zeros(len(q), np.uint64)
    for b in range(bits):
        for a in range(3):
            code |= ((q[:, a].astype(np.uint64) >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + a)
    return code


def sort_key(rays, name):
    """Host-side ordering keys for the ray-coherence experiment (how much would a device-side sort buy?)."""
    o, d = rays["o"], rays["d"]
    octant = ((d[:, 0] < 0).astype(np.uint64) | ((d[:, 1] < 0).astype(np.uint64) << np.uint64(1)) | ((d[:, 2] < 0).astype(np.uint64) << np.uint64(2)))
    lo, hi = o.min(0), o.max(0)

    def oq(bits):
        return np.minimum(((o - lo) / (hi - lo + 1e-9) * (1 << bits)).astype(np.int64), (1 << bits) - 1)

    dn = d / np.maximum(np.abs(d).max(-1, keepdims=True), 1e-30)

    def dq(bits):
        return np.minimum(((dn * 0.5 + 0.5) * (1 << bits)).astype(np.int64), (1 << bits) - 1)

    idx = np.arange(len(rays), dtype=np.uint64)
    if name == "none":
        return idx
    if name == "random":
        return np.random.default_rng(1).permutation(len(rays)).astype(np.uint64)
    if name == "oct":
        return octant
    if name.startswith("blk") and name.endswith("_oct"):  # octant inside blocks of N consecutive rays
        n = np.uint64(int(name[3:-4]))
        return (idx // n) * np.uint64(8) + octant
    if name.startswith("blk") and name.endswith("_dir"):  # 3 x 3-bit direction cell inside blocks of N rays
        n = np.uint64(int(name[3:-4]))
        return (idx // n) * np.uint64(512) + _morton3(dq(3), 3)
    if name == "morton_o":
        return _morton3(oq(10), 10)
    if name == "oct_morton_o":
        return (octant << np.uint64(30)) | _morton3(oq(10), 10)
    if name == "morton_o5_dir":  # coarse origin cell, then direction cell, then fine origin
        return (_morton3(oq(5), 5) << np.uint64(24)) | (_morton3(dq(3), 3) << np.uint64(15)) | (_morton3(oq(10), 10) & np.uint64(0x7FFF))
    if name == "morton_o4_dir":
        return (_morton3(oq(4), 4) << np.uint64(24)) | (_morton3(dq(4), 4) << np.uint64(12))
    if name == "dir_morton_o":
        return (_morton3(dq(3), 3) << np.uint64(30)) | _morton3(oq(10), 10)
    raise ValueError(name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quads", type=int, default=224)
    ap.add_argument("--res", default="1920x1080")
    ap.add_argument("--spp", type=int, default=4)
    ap.add_argument("--knobs", default="")
    ap.add_argument("--props", type=int, default=0)
    ap.add_argument("--cpu", type=int, default=0, help="time the CPU reference on this many rays")
    ap.add_argument("--rebuild", action="store_true", help="rebuild the hierarchy per knob set (builder knobs: VT_MAX_LEAF, VT_TRAV_COST, ...)")
    ap.add_argument("--ref-hits", default="", help="npz path: saved on first use, compared (records differing) on later runs — A/B of library builds")
    ap.add_argument("--sorts", default="", help="comma list of host-side bounce-ray orderings to time (see sort_key)")
    args = ap.parse_args()
    w, h = map(int, args.res.split("x"))
    t0 = time.time()
    if args.quads <= 400:
        scene = scenes.scene_heightfield(args.quads)
        cam = ((0, -80, 60), (0, 0, 5))
    else:
        scene = scenes.scene_terrain_closed(args.quads, n_props=args.props)
        cam = ((0, -330, 200), (0, 0, 10))
    print(f"scene: {scene.n_tris} tris, gen {time.time()-t0:.1f}s", flush=True)
    t0 = time.time()
    bvh = vt.build_bvh(scene)
    print(f"build: {len(bvh[0])} nodes in {time.time()-t0:.2f}s", flush=True)
    rays = scenes.pinhole_rays(w, h, *cam)
    knob_sets = [dict(kv.split("=") for kv in ks.split(",") if kv) for ks in args.knobs.split(";")] if args.knobs else [{}]
    bounce = None
    for knobs in knob_sets:
        for k in list(os.environ):
            if k.startswith("VT_"):
                del os.environ[k]
        os.environ.update(knobs)
        if args.rebuild:
            t0 = time.time()
            bvh = vt.build_bvh(scene)
            print(f"rebuild {knobs}: {len(bvh[0])} nodes in {time.time()-t0:.2f}s", flush=True)
        accel = vt.Accel(0)
        t0 = time.time()
        accel.populate(scene, bvh=bvh)
        t_up = time.time() - t0
        if bounce is None:
            hits, attrs = accel.traverse(rays, want_attrs=True)
            bounce, _ = scenes.bounce_rays(attrs, spp=args.spp)
            print(f"primary hit frac {(hits['prim'] != abi.VT_MISS).mean():.3f}; bounce rays {len(bounce)}", flush=True)
            d_rays, d_bounce = to_dev(rays), to_dev(bounce)
            d_hits = torch.empty(max(len(rays), len(bounce)) * 16, dtype=torch.uint8, device="cuda")
            d_attrs = torch.empty(len(rays) * 128, dtype=torch.uint8, device="cuda")
        st_p = accel.traverse_stats(d_rays.data_ptr(), len(rays))
        st_b = accel.traverse_stats(d_bounce.data_ptr(), len(bounce))
        ms_p = time_traverse(accel, d_rays, len(rays), d_hits)
        ms_b = time_traverse(accel, d_bounce, len(bounce), d_hits)
        ms_a = time_traverse(accel, d_bounce, len(bounce), d_hits, any_hit=True)
        ms_pa = time_traverse(accel, d_rays, len(rays), d_hits, d_attrs=d_attrs)
        print(json.dumps({"knobs": knobs, "layout": accel.layout, "S_I_primary": [round(st_p[0] / len(rays), 2), round(st_p[1] / len(rays), 2)],
                          "S_I_bounce": [round(st_b[0] / len(bounce), 2), round(st_b[1] / len(bounce), 2)], "upload_s": round(t_up, 2), "primary_Mrays": round(len(rays) / ms_p / 1e3, 1),
                          "bounce_Mrays": round(len(bounce) / ms_b / 1e3, 1), "bounce_anyhit_Mrays": round(len(bounce) / ms_a / 1e3, 1),
                          "primary+attrs_Mrays": round(len(rays) / ms_pa / 1e3, 1), "ms": [round(ms_p, 3), round(ms_b, 3), round(ms_a, 3), round(ms_pa, 3)]}), flush=True)
        if args.ref_hits:
            accel.traverse_device(d_rays.data_ptr(), len(rays), d_hits.data_ptr(), None)
            torch.cuda.synchronize()
            hp = d_hits[: len(rays) * 16].cpu().numpy().copy()
            accel.traverse_device(d_bounce.data_ptr(), len(bounce), d_hits.data_ptr(), None)
            torch.cuda.synchronize()
            hb = d_hits[: len(bounce) * 16].cpu().numpy().copy()
            if os.path.exists(args.ref_hits):
                ref = np.load(args.ref_hits)
                dp = (hp.reshape(-1, 16) != ref["hp"].reshape(-1, 16)).any(1).sum()
                db = (hb.reshape(-1, 16) != ref["hb"].reshape(-1, 16)).any(1).sum()
                print(json.dumps({"vs_ref_hits": args.ref_hits, "primary_records_differing": int(dp), "bounce_records_differing": int(db)}), flush=True)
            else:
                np.savez(args.ref_hits, hp=hp, hb=hb)
                print(json.dumps({"saved_ref_hits": args.ref_hits}), flush=True)
        for name in [x for x in args.sorts.split(",") if x]:
            order = np.argsort(sort_key(bounce, name), kind="stable")
            d_sorted = to_dev(np.ascontiguousarray(bounce[order]))
            st_s = accel.traverse_stats(d_sorted.data_ptr(), len(bounce))
            ms_s = time_traverse(accel, d_sorted, len(bounce), d_hits)
            print(json.dumps({"sort": name, "bounce_Mrays": round(len(bounce) / ms_s / 1e3, 1), "ms": round(ms_s, 3),
                              "S_I": [round(st_s[0] / len(bounce), 2), round(st_s[1] / len(bounce), 2)]}), flush=True)
            del d_sorted
        accel.close()
    if args.cpu:
        import oracle

        kind = "reference" if oracle.available("reference") else "port"
        cpu = oracle.CpuScene(scene, kind, build_bvh=False)
        cpu.set_bvh(*bvh)
        for name, rr in (("primary", rays), ("bounce", bounce)):
            sub = rr[:: max(1, len(rr) // args.cpu)][: args.cpu]
            r = cpu.traverse(sub, want_stats=True)
            print(json.dumps({"cpu": kind, "rays": name, "threads": cpu.max_threads, "Mrays": round(len(sub) / r["seconds"] / 1e6, 2),
                              "steps_per_ray": round(r["steps"] / len(sub), 1), "isects_per_ray": round(r["isects"] / len(sub), 1)}), flush=True)


if __name__ == "__main__":
    main()
