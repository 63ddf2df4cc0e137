"""Minimal driver for ncu captures of K1 on the bench's bounce wave.

usage: VT_LAYOUT=compact python tools/profile_k1.py [--quads 1582] [--reps 3]
k_traverse launches: #0 primary (with attrs), then `reps` launches over the bounce rays
(-k regex:k_traverse -s 2 -c 1 captures a warm bounce launch).
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vistrace_b200 as vt  # noqa: E402
from vistrace_b200 import scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--quads", type=int, default=1582)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--spp", type=int, default=4)
args = ap.parse_args()
scene = scenes.scene_terrain_closed(args.quads)
rays = scenes.pinhole_rays(1920, 1080, (0.0, -330.0, 200.0), (0.0, 0.0, 10.0))
accel = vt.Accel(0).populate(scene)
n = len(rays)
dev = lambda nbytes: torch.empty(nbytes, dtype=torch.uint8, device="cuda")
d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1)).cuda()
d_hits, d_attrs, d_brays, d_bhits = dev(n * 16), dev(n * 128), dev(n * args.spp * 32), dev(n * args.spp * 16)
s = torch.cuda.current_stream().cuda_stream
accel.traverse_device(d_rays.data_ptr(), n, d_hits.data_ptr(), d_attrs.data_ptr(), stream=s)
accel.bounce_rays_device(d_attrs.data_ptr(), n, args.spp, 1003, d_brays.data_ptr(), stream=s)
for _ in range(args.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    accel.traverse_device(d_brays.data_ptr(), n * args.spp, d_bhits.data_ptr(), stream=s)
    e1.record()
    torch.cuda.synchronize()
    print(f"layout={accel.layout} bounce K1 {e0.elapsed_time(e1):.3f} ms", flush=True)
