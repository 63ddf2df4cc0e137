"""N > 1 host path on CPU: world_size-2 gloo processes shard a ray batch, trace their slices and gather the
hit buffer.  The tracer here is the oracle (there is no GPU in this container); on the GPU box the same
plumbing runs with the CUDA engine over NCCL (test_gpu_multi in test_gpu_parity.py, bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT, load_golden


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    import oracle
    from vistrace_b200 import shard

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene, z = load_golden("props_small")
    cpu = oracle.CpuScene(scene, "port", build_bvh=False)
    # the hierarchy exists on rank 0 only and is replicated by broadcast (as bench.py --gpus N does over NCCL)
    nodes, prims = shard.replicate_bvh((z["nodes"], z["prim_indices"]) if rank == 0 else None)
    assert nodes.tobytes() == z["nodes"].tobytes() and prims.tobytes() == z["prim_indices"].astype(np.uint64).tobytes()
    cpu.set_bvh(nodes, prims)
    rays = z["rays"][:5001]  # odd size: ragged shards
    full = shard.trace_sharded(lambda r: cpu.traverse(r, threads=1)["hits"], rays)
    # per-rank partial framebuffers summed with one reduce (the bench's collective)
    b, e = shard.shard_range(len(rays), rank, world)
    fb = torch.zeros(len(rays), dtype=torch.float64)
    fb[b:e] = torch.from_numpy(full["t"][b:e].astype(np.float64))
    dist.reduce(fb, dst=0, op=dist.ReduceOp.SUM)
    # the N-GPU end-to-end step (shard.ShardedFrame): 1/N of the rays in per rank, all_gather, every rank traces the whole
    # frame for its own sample set, all_reduce, 1/N of the finished image out per rank
    from vistrace_b200 import abi

    frame = shard.ShardedFrame(len(rays), torch.device("cpu"))
    h_rays = torch.from_numpy(np.ascontiguousarray(rays).view(np.float32).reshape(-1).copy())
    if rank != 0:
        h_rays[: frame.begin * 8] = 0  # a rank may only read its own chunk of the host array
        h_rays[frame.end * 8:] = 0

    def trace(d_rays):
        got = np.frombuffer(d_rays.numpy().tobytes(), abi.RAY)
        assert got.tobytes() == np.ascontiguousarray(rays).tobytes()  # the gathered frame is the whole ray array
        h = cpu.traverse(got, threads=1)["hits"]
        img = np.stack([h["t"], h["u"], h["v"]], 1).astype(np.float32) * np.float32(rank + 1)  # "this rank's samples"
        return torch.from_numpy(img.reshape(-1).copy())

    b, e, img = frame.step(h_rays, trace)
    np.save(os.path.join(out_dir, f"frame{rank}.npy"), np.concatenate([[b, e], img.reshape(-1)]))
    if rank == 0:
        np.savez(os.path.join(out_dir, "out.npz"), hits=full, fb=fb.numpy())
    else:
        np.save(os.path.join(out_dir, f"hits{rank}.npy"), full)
    dist.destroy_process_group()


def test_shard_ranges_cover_exactly():
    from vistrace_b200.shard import shard_range

    for n in (0, 1, 7, 8, 1000, 2073600):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


@pytest.mark.timeout(300)
def test_two_rank_gloo_shard_and_gather(built, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    scene, z = load_golden("props_small")
    out = np.load(tmp_path / "out.npz")
    want = z["hits"][:5001]  # golden: the reference's own hits for these rays
    assert out["hits"].tobytes() == want.tobytes()
    assert np.load(tmp_path / "hits1.npy").tobytes() == want.tobytes()  # every rank holds the full buffer
    np.testing.assert_allclose(out["fb"], want["t"].astype(np.float64))
    # ShardedFrame: the two chunks tile the frame and hold (1 + 2) x the single-rank image
    img = np.stack([want["t"], want["u"], want["v"]], 1).astype(np.float32)
    total = img * np.float32(1) + img * np.float32(2)
    spans = []
    for r in range(2):
        f = np.load(tmp_path / f"frame{r}.npy")
        b, e = int(f[0]), int(f[1])
        spans.append((b, e))
        np.testing.assert_array_equal(f[2:].astype(np.float32).reshape(-1, 3), total[b:e])
    assert spans == [(0, 2501), (2501, 5001)]


def _tile_worker(rank, world, port, out_dir):
    """The frame sharding of the native group (vt_group.cu: tiles dealt round-robin, compact shards, gather on rank 0) with the
    C port as the tracer and gloo as the fabric: the assembled frame must equal the frame traced whole."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    import oracle
    import vistrace_b200 as vt

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene, z = load_golden("props_small")
    cpu = oracle.CpuScene(scene, "port", build_bvh=False)
    cpu.set_bvh(z["nodes"], z["prim_indices"])
    rays = z["rays"][:5001]
    n, tile = len(rays), 512  # 10 tiles, the last one ragged (393 rays) and owned by rank 1
    idx = vt.shard_indices(n, world, rank, tile)
    mine = cpu.traverse(np.ascontiguousarray(rays[idx]), threads=1)["hits"]  # the compact shard, traced on its own
    counts = [len(vt.shard_indices(n, world, r, tile)) for r in range(world)]
    t = torch.from_numpy(mine.view(np.uint8).reshape(-1).copy())
    if rank == 0:
        frame = np.zeros(n, mine.dtype)
        frame[idx] = mine
        for r in range(1, world):
            buf = torch.empty(counts[r] * mine.dtype.itemsize, dtype=torch.uint8)
            dist.recv(buf, src=r)
            frame[vt.shard_indices(n, world, r, tile)] = np.frombuffer(buf.numpy().tobytes(), mine.dtype)
        np.save(os.path.join(out_dir, "tile_frame.npy"), frame.view(np.uint8))
    else:
        dist.send(t, dst=0)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_geometry_covers_every_record_once(built):
    import vistrace_b200 as vt

    for n in (0, 1, 511, 512, 513, 5001, 2073600):
        for world in (1, 2, 3, 8):
            for tile in (64, 512, 8192):
                parts = [vt.shard_indices(n, world, r, tile) for r in range(world)]
                allidx = np.concatenate(parts) if parts else np.zeros(0, np.int64)
                assert np.array_equal(np.sort(allidx), np.arange(n)), (n, world, tile)
                if n >= world * tile * 4:  # round-robin tiles: shards differ by at most one tile
                    assert max(map(len, parts)) - min(map(len, parts)) <= tile


@pytest.mark.timeout(300)
def test_two_rank_gloo_tile_sharded_frame(built, tmp_path):
    port = _free_port()
    mp.spawn(_tile_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    scene, z = load_golden("props_small")
    want = z["hits"][:5001]
    assert np.load(tmp_path / "tile_frame.npy").tobytes() == want.tobytes()
