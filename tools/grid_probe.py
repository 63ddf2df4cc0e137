"""K1 on SMALL launches (what a rank of an 8-GPU strong-scaling run sees): persistent-grid size vs time.
VT_K1_RAYS_PER_LANE = R shrinks the grid so that every lane gets about R rays.  usage: python tools/grid_probe.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vistrace_b200 as vt  # noqa: E402
from vistrace_b200 import abi, scenes  # noqa: E402

scene = scenes.scene_terrain_closed(1582)
rays = scenes.pinhole_rays(1920, 1080, (0.0, -330.0, 200.0), (0.0, 0.0, 10.0))
bvh = vt.build_bvh(scene)
spp = 4
stream = torch.cuda.current_stream()
sh = stream.cuda_stream
for R in (0, 2, 3, 4, 6, 8, 12, 16):
    os.environ["VT_K1_RAYS_PER_LANE"] = str(R)
    accel = vt.Accel(0).populate(scene, bvh=bvh)
    for world in (8, 4, 2, 1):
        sub = np.ascontiguousarray(rays[vt.shard_indices(len(rays), world, 0, 8192)])
        n = len(sub)
        d_rays = torch.from_numpy(sub.view(np.uint8).reshape(-1).copy()).cuda()
        d_hits, d_attrs = torch.empty(n * 16, dtype=torch.uint8, device="cuda"), torch.empty(n * 128, dtype=torch.uint8, device="cuda")
        d_brays, d_bhits = torch.empty(n * spp * 32, dtype=torch.uint8, device="cuda"), torch.empty(n * spp * 16, dtype=torch.uint8, device="cuda")

        def step():
            accel.trace_diffuse_wave_device(d_rays.data_ptr(), n, spp, 5, d_hits.data_ptr(), d_attrs.data_ptr(), d_brays.data_ptr(), d_bhits.data_ptr(), stream=sh)

        def timed(fn, reps=10):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                fn()
            e1.record(stream)
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps

        ms_step = timed(step)
        ms_p = timed(lambda: accel.traverse_device(d_rays.data_ptr(), n, d_hits.data_ptr(), stream=sh))
        ms_b = timed(lambda: accel.traverse_device(d_brays.data_ptr(), n * spp, d_bhits.data_ptr(), stream=sh))
        print(json.dumps({"rays_per_lane": R, "world": world, "primary_rays": n, "K1K2K3K1_ms": round(ms_step, 4), "K1_primary_ms": round(ms_p, 4),
                          "K1_bounce_allslots_ms": round(ms_b, 4)}), flush=True)
    accel.close()
