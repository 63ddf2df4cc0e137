"""BASELINE config 5 on one GPU through vt_accel_trace_paths: 20 M-triangle terrain + props, 3840x2160, per sample primary + shadow +
3 diffuse bounces each with a shadow ray — with and without wave compaction.  usage: python tools/config5_paths.py [--quads 2980] [--samples 4]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vistrace_b200 as vt  # noqa: E402
from vistrace_b200 import scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quads", type=int, default=2980)
    ap.add_argument("--props", type=int, default=143)
    ap.add_argument("--res", default="3840x2160")
    ap.add_argument("--samples", type=int, default=4)
    ap.add_argument("--bounces", type=int, default=3)
    args = ap.parse_args()
    W, H = map(int, args.res.split("x"))
    t0 = time.time()
    scene = scenes.scene_terrain_closed(args.quads, n_props=args.props)
    rays = scenes.pinhole_rays(W, H, (0.0, -330.0, 200.0), (0.0, 0.0, 10.0))
    gen_s = time.time() - t0
    t0 = time.time()
    accel = vt.Accel(0).populate(scene)
    setup_s = time.time() - t0
    n = len(rays)
    d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1).copy()).cuda()
    d_fb = torch.zeros(n * 3, dtype=torch.float32, device="cuda")
    sun = np.array((0.3, 0.2, 0.93), np.float32)
    sun = sun / np.linalg.norm(sun)
    stream = torch.cuda.current_stream()
    for compact in (True, False):
        counts = accel.trace_paths_device(d_rays.data_ptr(), n, args.bounces, sun, (1, 1, 1), 1, 1.0, d_fb.data_ptr(), want_counts=True, compact=compact, stream=stream.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for s in range(args.samples):
            accel.trace_paths_device(d_rays.data_ptr(), n, args.bounces, sun, (1, 1, 1), 100 + s, 1.0 / args.samples, d_fb.data_ptr(), compact=compact, stream=stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.samples
        total = int(counts.sum())
        print(json.dumps({"n_tris": int(scene.n_tris), "res": args.res, "bounces": args.bounces, "compaction": compact, "rays_per_sample": total,
                          "rays_per_wave": [int(c) for c in counts], "ms_per_sample": round(ms, 3), "Mrays_s": round(total / ms / 1e3, 1),
                          "gen_s": round(gen_s, 1), "setup_s": round(setup_s, 1)}), flush=True)


if __name__ == "__main__":
    main()
