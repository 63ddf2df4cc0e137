"""Do consecutive frames overlap usefully?  The bench step (K1 K2 K3 K1 K4 over the resident 1080p frame) issued on ONE stream
against the same steps alternating between 2 / 3 streams (one frame's launch tails under the next frame's bulk).
usage: python tools/frames_in_flight_probe.py [--worlds 1,8] [--steps 24]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import vistrace_b200 as vt  # noqa: E402
from vistrace_b200 import scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--worlds", default="1,8")
ap.add_argument("--steps", type=int, default=24)
args = ap.parse_args()
dev = torch.device("cuda", 0)
scene = scenes.scene_terrain_closed(1582)
rays = scenes.pinhole_rays(1920, 1080, (0.0, -330.0, 200.0), (0.0, 0.0, 10.0))
accel = vt.Accel(0).populate(scene)
for world in [int(w) for w in args.worlds.split(",")]:
    sub = np.ascontiguousarray(rays[vt.shard_indices(len(rays), world, 0, 8192)]) if world > 1 else rays
    for n_streams in (1, 2, 3):
        streams = [torch.cuda.Stream() for _ in range(n_streams)]
        frames = []
        for s in streams:
            with torch.cuda.stream(s):
                frames.append(bench.ResidentFrame(accel, torch, dev, sub, True))
        live = frames[0].live_bounce_rays()
        for it in range(6):
            frames[it % n_streams].step(100 + it, 1.0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main = torch.cuda.current_stream()
        e0.record(main)
        for s in streams:
            s.wait_event(e0)
        for it in range(args.steps):
            frames[it % n_streams].step(200 + it, 1.0)
        for s in streams:
            ev = torch.cuda.Event()
            ev.record(s)
            main.wait_event(ev)
        e1.record(main)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        print(json.dumps({"world": world, "rays_per_step": len(sub) + live, "frames_in_flight": n_streams, "ms_per_step": round(ms, 4),
                          "Mrays_s": round((len(sub) + live) / ms / 1e3, 1)}), flush=True)
        del frames
