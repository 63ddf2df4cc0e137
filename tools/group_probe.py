"""What one rank of an N-GPU strong-scaling run has to do, measured on ONE GPU: rank 0's shard of the bench frame (tiles dealt
round-robin) traced through a one-member group, device-resident and host-pointer paths, over lane / chunk settings.  NCCL's gather
is not included (the probe has one GPU); everything else a rank does per step is.
usage: python tools/group_probe.py [--worlds 1,2,4,8] [--lanes 1,2,4] [--quads 1582]"""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--worlds", default="1,2,4,8")
    ap.add_argument("--lanes", default="1,2,4")
    ap.add_argument("--quads", type=int, default=1582)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    scene = scenes.scene_terrain_closed(args.quads)
    rays = scenes.pinhole_rays(1920, 1080, (0.0, -330.0, 200.0), (0.0, 0.0, 10.0))
    group = vt.Group(device=0, rank=0, world=1).populate(scene)
    n_full, spp = len(rays), 4
    stream = torch.cuda.current_stream()
    for world in [int(w) for w in args.worlds.split(",")]:
        idx = vt.shard_indices(n_full, world, 0, 8192)
        sub = np.ascontiguousarray(rays[idx])
        n = len(sub)
        d_rays = torch.from_numpy(sub.view(np.uint8).reshape(-1).copy()).cuda()
        d_fb = torch.zeros(n * 3, dtype=torch.float32, device="cuda")
        h_rays_t = torch.empty(n * 32, dtype=torch.uint8).pin_memory()
        h_rays = h_rays_t.numpy().view(abi.RAY)
        h_rays[:] = sub
        h_fb = torch.empty(n * 12, dtype=torch.uint8).pin_memory().numpy().view(np.float32).reshape(n, 3)
        _, live = group.render_diffuse_wave(h_rays, spp, seed=5, out=h_fb)
        for lanes in [int(x) for x in args.lanes.split(",")]:
            os.environ["VT_GROUP_DEV_LANES"] = str(lanes)
            for _ in range(3):
                group.render_diffuse_wave_device(d_rays.data_ptr(), n, spp, 5, 1.0, d_fb.data_ptr(), stream=stream.cuda_stream)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for it in range(args.steps):
                group.render_diffuse_wave_device(d_rays.data_ptr(), n, spp, 5 + it, 1.0, d_fb.data_ptr(), stream=stream.cuda_stream)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            print(json.dumps({"world": world, "shard_rays": n, "rays_per_step": n + live, "path": "device", "lanes": lanes, "ms_per_step": round(ms, 4),
                              "Mrays_per_gpu": round((n + live) / ms / 1e3, 1)}), flush=True)
        for wl in (4, 2, 1):
            os.environ["VT_WAVE_LANES"] = str(wl)
            for _ in range(2):
                group.render_diffuse_wave(h_rays, spp, seed=5, out=h_fb, want_live=False)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for it in range(args.steps):
                group.render_diffuse_wave(h_rays, spp, seed=5 + it, out=h_fb, want_live=False)
            ms = 1e3 * (time.perf_counter() - t0) / args.steps
            print(json.dumps({"world": world, "shard_rays": n, "rays_per_step": n + live, "path": "host", "wave_lanes": wl, "ms_per_step": round(ms, 4),
                              "Mrays_per_gpu": round((n + live) / ms / 1e3, 1)}), flush=True)
        del os.environ["VT_WAVE_LANES"]


if __name__ == "__main__":
    main()
