"""Multi-GPU plumbing: one process per GPU, hierarchy replicated, ray batch sharded, ONE collective.

The path has no exchange step during traversal (rays are independent, the scene is read-only), so
the only communication is the final gather of the per-rank hit slices (or the sum of per-rank
framebuffers, see bench.py).  torch.distributed is used as plumbing: NCCL over NVLink on the GPUs,
gloo in the CPU tests.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import abi


def shard_range(n, rank, world):
    """Contiguous, balanced slice [begin, end) of n rays for `rank` (tile order is preserved)."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_hits(local_hits, n_total, group=None, device=None):
    """all_gather the per-rank vt_hit slices (numpy, shard_range order) into the full n_total-record array."""
    world = dist.get_world_size(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    longest = max(e - b for b, e in sizes)
    buf = np.zeros(longest, abi.HIT)
    buf[: len(local_hits)] = local_hits
    t = torch.from_numpy(buf.view(np.uint8).reshape(-1).copy())
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    full = np.zeros(n_total, abi.HIT)
    for (b, e), o in zip(sizes, outs):
        full[b:e] = np.frombuffer(o.cpu().numpy().tobytes(), abi.HIT)[: e - b]
    return full


def trace_sharded(trace_fn, rays, group=None, device=None):
    """Every rank traces its own slice of `rays` with trace_fn(rays_slice) -> hits and all ranks receive the full hit buffer."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    b, e = shard_range(len(rays), rank, world)
    return gather_hits(trace_fn(rays[b:e]), len(rays), group, device)


def replicate_bvh(bvh, group=None, device=None, src=0):
    """Rank `src` built the hierarchy (bvh = (nodes, prim_indices) in bvh::Bvh<float> form); every other rank passes
    None and receives a bit-identical replica through two broadcasts (NCCL over NVLink when `device` is a GPU) —
    the hierarchy is built ONCE per job, not once per GPU."""
    rank = dist.get_rank(group)
    dev = device if device is not None else torch.device("cpu")
    counts = torch.zeros(2, dtype=torch.int64, device=dev)
    if rank == src:
        nodes = np.ascontiguousarray(bvh[0], abi.NODE)
        prims = np.ascontiguousarray(bvh[1], np.uint64)
        counts = torch.tensor([len(nodes), len(prims)], dtype=torch.int64, device=dev)
    dist.broadcast(counts, src=src, group=group)
    n_nodes, n_prims = (int(v) for v in counts.cpu())
    if rank == src:
        t_nodes = torch.from_numpy(nodes.view(np.uint8).reshape(-1).copy()).to(dev)
        t_prims = torch.from_numpy(prims.view(np.int64).copy()).to(dev)
    else:
        t_nodes = torch.empty(n_nodes * abi.NODE.itemsize, dtype=torch.uint8, device=dev)
        t_prims = torch.empty(n_prims, dtype=torch.int64, device=dev)
    dist.broadcast(t_nodes, src=src, group=group)
    dist.broadcast(t_prims, src=src, group=group)
    if rank == src:
        return nodes, prims
    return np.frombuffer(t_nodes.cpu().numpy().tobytes(), abi.NODE), t_prims.cpu().numpy().view(np.uint64)
