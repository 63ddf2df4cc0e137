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
dist.barrier()
if rank == 0:
    print("MULTI_GPU_OK", len(rays))
dist.destroy_process_group()
