"""torchrun worker for test_two_gpu_sharded_trace_nccl: each rank owns one GPU and a replica of the scene."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vistrace_b200 as vt  # noqa: E402
from vistrace_b200 import scenes, shard  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
scene = scenes.scene_props(8, 31, 15, 16)
rays = scenes.pinhole_rays(401, 203, (0, -95, 40), (0, 0, 10))
dev = torch.device("cuda", local)
bvh = shard.replicate_bvh(vt.build_bvh(scene) if rank == 0 else None, device=dev)  # built once, replicated over NCCL
accel = vt.Accel(local).populate(scene, bvh=bvh)
full = shard.trace_sharded(lambda r: accel.traverse(r), rays, device=torch.device("cuda", local))
single = accel.traverse(rays)
assert full.tobytes() == single.tobytes(), "sharded result differs from the single-GPU result"

# shard.ShardedFrame on the GPUs: 1/N of the host rays per rank, all_gather over NVLink, every rank traces the whole
# frame for its own seed, all_reduce, 1/N of the finished image per rank == the sum of the per-seed single-GPU images
n, spp, world = len(rays), 2, dist.get_world_size()


def dbuf(nbytes):
    return torch.empty(nbytes, dtype=torch.uint8, device=dev)


d_hits, d_attrs, d_brays, d_bhits = dbuf(n * 16), dbuf(n * 128), dbuf(n * spp * 32), dbuf(n * spp * 16)
sh = torch.cuda.current_stream().cuda_stream


def image(d_rays, seed):
    fb = torch.zeros(n * 3, dtype=torch.float32, device=dev)
    accel.traverse_device(d_rays.data_ptr(), n, d_hits.data_ptr(), d_attrs.data_ptr(), stream=sh)
    accel.bounce_rays_device(d_attrs.data_ptr(), n, spp, seed, d_brays.data_ptr(), stream=sh)
    accel.traverse_device(d_brays.data_ptr(), n * spp, d_bhits.data_ptr(), stream=sh)
    accel.accumulate_sky_device(d_attrs.data_ptr(), d_bhits.data_ptr(), n, spp, 1.0 / world, fb.data_ptr(), stream=sh)
    return fb


h_rays = torch.from_numpy(rays.view(np.float32).reshape(-1).copy()).pin_memory()
frame = shard.ShardedFrame(n, dev)
b, e, chunk = frame.step(h_rays, lambda d: image(d, 100 + rank))
d_all = h_rays.to(dev)
want = sum(image(d_all, 100 + r) for r in range(world)).cpu().numpy().reshape(-1, 3)
assert (b, e) == (min(n, rank * frame.chunk), min(n, (rank + 1) * frame.chunk))
np.testing.assert_array_equal(chunk, want[b:e])
assert np.abs(want).sum() > 0
dist.barrier()
if rank == 0:
    print("MULTI_GPU_OK", len(rays))
dist.destroy_process_group()
