"""Host-pointer frames of shard size (what one rank of an N-GPU e2e step does, minus the other ranks' PCIe traffic): synchronous calls
and two frames in flight, over the tile schedule (VT_WAVE_TILE / VT_WAVE_FIRST).
usage: python tools/e2e_chunk_probe.py [--worlds 8,2,1]"""
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

ap = argparse.ArgumentParser()
ap.add_argument("--worlds", default="8,2,1")
ap.add_argument("--steps", type=int, default=30)
args = ap.parse_args()
scene = scenes.scene_terrain_closed(1582)
rays = scenes.pinhole_rays(1920, 1080, (0.0, -330.0, 200.0), (0.0, 0.0, 10.0))
accel = vt.Accel(0).populate(scene)
for world in [int(w) for w in args.worlds.split(",")]:
    sub = np.ascontiguousarray(rays[vt.shard_indices(len(rays), world, 0, 8192)]) if world > 1 else rays
    n = len(sub)
    h_rays_t = torch.empty(n * 32, dtype=torch.uint8).pin_memory()
    h_rays = h_rays_t.numpy().view(abi.RAY)
    h_rays[:] = sub
    fb_t = [torch.empty(n * 12, dtype=torch.uint8).pin_memory() for _ in range(2)]
    fbs = [t.numpy().view(np.float32).reshape(n, 3) for t in fb_t]
    for tile, first in ((1 << 19, 1 << 16), (1 << 19, 1 << 17), (1 << 19, 1 << 18), (1 << 22, 1 << 22), (1 << 18, 1 << 16)):
        os.environ["VT_WAVE_TILE"], os.environ["VT_WAVE_FIRST"] = str(tile), str(first)
        res = {"world": world, "rays": n, "tile": tile, "first": first}
        for mode in ("sync", "two_in_flight"):
            for it in range(3):
                accel.render_diffuse_wave(h_rays, 4, seed=it, weight=1.0, out=fbs[0])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if mode == "sync":
                for it in range(args.steps):
                    accel.render_diffuse_wave(h_rays, 4, seed=10 + it, weight=1.0, out=fbs[0])
            else:
                for it in range(args.steps):
                    if it >= 2:
                        accel.render_diffuse_wave_wait()
                    accel.render_diffuse_wave_begin(h_rays, 4, 10 + it, 1.0, fbs[it % 2])
                accel.render_diffuse_wave_wait()
                accel.render_diffuse_wave_wait()
            res[mode + "_ms"] = round(1e3 * (time.perf_counter() - t0) / args.steps, 4)
        print(json.dumps(res), flush=True)
