"""Where the N-rank end-to-end step of the headline frame spends its time (run under torchrun on N GPUs):
the full vt_group_render_diffuse_wave, the same without the pipelined gather, without any gather, and the bare H2D / D2H copies."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vistrace_b200 as vt  # noqa: E402
from vistrace_b200 import abi, scenes  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
uid = torch.from_numpy(vt.group_unique_id() if rank == 0 else np.zeros(128, np.uint8)).to(dev)
dist.broadcast(uid, src=0)
group = vt.Group(device=local, rank=rank, world=world, unique_id=uid.cpu().numpy())
group.populate(scenes.scene_terrain_closed(1582) if rank == 0 else None)
rays = scenes.pinhole_rays(1920, 1080, (0.0, -330.0, 200.0), (0.0, 0.0, 10.0))
n = len(rays)
h_rays_t = torch.empty(n * 32, dtype=torch.uint8).pin_memory()
h_rays = h_rays_t.numpy().view(abi.RAY)
h_rays[:] = rays
h_fb = torch.empty(n * 12, dtype=torch.uint8).pin_memory().numpy().view(np.float32).reshape(n, 3)


def timed(fn, steps=20):
    for _ in range(3):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for it in range(steps):
        fn()
    torch.cuda.synchronize()
    t = torch.tensor([1e3 * (time.perf_counter() - t0) / steps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(float(t.item()), 4)


out = {"world": world}
for name, env in (("full", {}), ("no_pipeline", {"VT_GROUP_PIPELINE": "0"}), ("no_gather", {"VT_GROUP_NO_GATHER": "1"}),
                  ("one_lane", {"VT_WAVE_LANES": "1"}), ("two_lanes", {"VT_WAVE_LANES": "2"}), ("chunk_128k", {"VT_WAVE_TILE": "131072", "VT_WAVE_FIRST": "32768"})):
    os.environ.update(env)
    out[name + "_ms"] = timed(lambda: group.render_diffuse_wave(h_rays, 4, seed=5, out=h_fb, want_live=False))
    for k in env:
        del os.environ[k]
idx = group.shard_indices(n)
d = torch.empty(len(idx) * 32, dtype=torch.uint8, device=dev)
h_chunk = torch.empty(len(idx) * 32, dtype=torch.uint8).pin_memory()
out["h2d_shard_ms"] = timed(lambda: (d.copy_(h_chunk, non_blocking=True), torch.cuda.synchronize()))
d_fb = torch.empty(n * 12, dtype=torch.uint8, device=dev)
h_full = torch.empty(n * 12, dtype=torch.uint8).pin_memory()
if rank == 0:
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        h_full.copy_(d_fb, non_blocking=True)
        torch.cuda.synchronize()
    out["d2h_frame_rank0_alone_ms"] = round(1e3 * (time.perf_counter() - t0) / 20, 4)
dist.barrier()
if rank == 0:
    print(json.dumps(out), flush=True)
group.close()
dist.destroy_process_group()
