"""torchrun worker for test_multi_process_group_two_ranks_nccl: one process per GPU, native vt_group in rank mode."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vistrace_b200 as vt  # noqa: E402
from vistrace_b200 import scenes  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
uid = torch.from_numpy(vt.group_unique_id() if rank == 0 else np.zeros(128, np.uint8)).to(dev)
dist.broadcast(uid, src=0)  # the launcher's job: hand rank 0's ncclUniqueId to everybody
group = vt.Group(device=local, rank=rank, world=world, unique_id=uid.cpu().numpy())
scene = scenes.scene_heightfield(64)
rays = scenes.pinhole_rays(333, 187, (0, -80, 60), (0, 0, 5))
group.populate(scene if rank == 0 else None)  # rank 0 builds; everybody else receives the device image over NCCL
n, spp = len(rays), 3

hits, attrs = group.traverse(rays, want_attrs=True)
os.environ["VT_GROUP_TILE"] = "1000"
img, live = group.render_diffuse_wave(rays, spp, seed=9, weight=0.5)
live_t = torch.tensor([live], dtype=torch.int64, device=dev)
dist.all_reduce(live_t)

# the same frame several times over (the consumed / landed flag hand-shake of the peer-memory frame), then the NCCL gather
for it in range(4):
    img_it, _ = group.render_diffuse_wave(rays, spp, seed=9, weight=0.5, want_live=False)
    if rank == 0:
        np.testing.assert_array_equal(img_it, img)
os.environ["VT_GROUP_GATHER"] = "nccl"
img_nccl, _ = group.render_diffuse_wave(rays, spp, seed=9, weight=0.5, want_live=False)
del os.environ["VT_GROUP_GATHER"]
if rank == 0:
    np.testing.assert_array_equal(img_nccl, img)

# one host frame shared by all processes (POSIX shm, pinned by each): every rank lands its own tiles, complete on every rank
from vistrace_b200 import shard  # noqa: E402
frame = shard.SharedPinnedFrame(f"vt_test_frame_{os.environ.get('MASTER_PORT', '0')}", n * 12, create=(rank == 0)) if rank == 0 else None
dist.barrier()
if rank != 0:
    frame = shard.SharedPinnedFrame(f"vt_test_frame_{os.environ.get('MASTER_PORT', '0')}", n * 12, create=False)
shared = frame.array(np.float32, (n, 3))
want = torch.from_numpy(img if rank == 0 else np.zeros((n, 3), np.float32)).to(dev)
dist.broadcast(want, src=0)  # rank 0's gathered image: what the shared frame must hold on EVERY rank when the call returns
want = want.cpu().numpy()
for it in range(3):
    if rank == 0:
        shared[:] = -1.0
    dist.barrier()
    group.render_diffuse_wave(rays, spp, seed=9, weight=0.5, out=shared, want_live=False, shared_frame=True)
    np.testing.assert_array_equal(shared, want)
    dist.barrier()
# two such frames in flight (VT_GROUP_ASYNC): frame k + 1 is begun before frame k is awaited
frame_b = shard.SharedPinnedFrame(f"vt_test_frame_b_{os.environ.get('MASTER_PORT', '0')}", n * 12, create=True) if rank == 0 else None
dist.barrier()
if rank != 0:
    frame_b = shard.SharedPinnedFrame(f"vt_test_frame_b_{os.environ.get('MASTER_PORT', '0')}", n * 12, create=False)
pair = [shared, frame_b.array(np.float32, (n, 3))]
seeds = [9, 11, 9, 11, 9]
img11 = group.render_diffuse_wave(rays, spp, seed=11, weight=0.5, want_live=False)[0]  # collective: every rank calls it; complete on rank 0
want11 = torch.from_numpy(np.ascontiguousarray(img11) if rank == 0 else np.zeros((n, 3), np.float32)).to(dev)
dist.broadcast(want11, src=0)
wants = {9: want, 11: want11.cpu().numpy()}
landed = []
for k, seed in enumerate(seeds):
    if k >= 2:
        group.wait_frame()
        landed.append(pair[k % 2].copy())
        dist.barrier()  # everybody has looked at the frame before anybody's GPU overwrites it
    group.render_diffuse_wave_begin(rays, spp, seed, 0.5, pair[k % 2])
for k in (3, 4):
    group.wait_frame()
    landed.append(pair[k % 2].copy())
for k, seed in enumerate(seeds):
    np.testing.assert_array_equal(landed[k], wants[seed], err_msg=f"frame {k} in flight")
dist.barrier()
frame_b.close()
frame.close()

# device-resident shards -> frame-sized device image on rank 0
idx = group.shard_indices(n)
d_rays = torch.from_numpy(np.ascontiguousarray(rays[idx]).view(np.uint8).reshape(-1).copy()).to(dev)
d_fb = torch.full((n * 3,), -1.0, dtype=torch.float32, device=dev)
group.render_diffuse_wave_device(d_rays.data_ptr(), n, spp, 9, 0.5, d_fb.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()

# two frames in flight: consecutive frames alternate between the group's two frame slots on two streams (different seeds, so a
# frame landing in the wrong slot or a stale hand-shake flag would show)
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
d_slot = [torch.full((n * 3,), -1.0, dtype=torch.float32, device=dev) for _ in range(2)]
torch.cuda.synchronize()
for it in range(7):
    k = it % 2
    group.render_diffuse_wave_device(d_rays.data_ptr(), n, spp, 20 + it, 0.5, d_slot[k].data_ptr(), stream=streams[k].cuda_stream, slot=k)
torch.cuda.synchronize()
dist.barrier()
slot_imgs = [d.cpu().numpy().reshape(-1, 3) for d in d_slot]  # frames 6 (slot 0) and 5 (slot 1) were the last to land

# sample-index sharding: per-rank partial images summed on rank 0 with one ncclReduce
part = torch.full((1000,), float(rank + 1), dtype=torch.float32, device=dev)
group.reduce_device(part.data_ptr(), 1000, stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()

# every rank contributes its slice of a device buffer; afterwards every rank holds the whole buffer (ncclAllGather, in place)
slice_bytes = 4096
gathered = torch.zeros(world * slice_bytes, dtype=torch.uint8, device=dev)
gathered[rank * slice_bytes:(rank + 1) * slice_bytes] = rank + 1
group.all_gather_device(gathered.data_ptr(), slice_bytes, stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
assert gathered.cpu().numpy().reshape(world, slice_bytes).tolist() == [[r + 1] * slice_bytes for r in range(world)]

single = vt.Accel(local).populate(scene)  # each rank checks what it holds against its own single-GPU run
want_hits, want_attrs = single.traverse(rays, want_attrs=True)
want_img, want_live = single.render_diffuse_wave(rays, spp, seed=9, weight=0.5)
if rank == 0:
    assert hits.tobytes() == want_hits.tobytes() and attrs.tobytes() == want_attrs.tobytes(), "gathered hit buffer differs"
    np.testing.assert_array_equal(img, want_img)
    np.testing.assert_array_equal(d_fb.cpu().numpy().reshape(-1, 3), want_img)
    for k, seed in ((0, 26), (1, 25)):
        np.testing.assert_array_equal(slot_imgs[k], single.render_diffuse_wave(rays, spp, seed=seed, weight=0.5)[0], err_msg=f"frame slot {k}")
    assert int(live_t.item()) == want_live
    assert float(part[0].item()) == sum(range(1, world + 1))
else:  # a non-root rank holds its own slice / tiles
    b = rank * (n // world) + min(rank, n % world)
    e = b + n // world + (1 if rank < n % world else 0)
    assert hits[b:e].tobytes() == want_hits[b:e].tobytes()
dist.barrier()
if rank == 0:
    print("GROUP_OK", n, int(live_t.item()))
group.close()
dist.destroy_process_group()
