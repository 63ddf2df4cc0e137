"""Multi-GPU plumbing: one process per GPU, hierarchy replicated, ray batch sharded, ONE collective.

The path has no exchange step during traversal (rays are independent, the scene is read-only), so
the only communication is the final gather of the per-rank hit slices (or the sum of per-rank
framebuffers, see bench.py).  torch.distributed is used as plumbing: NCCL over NVLink on the GPUs,
gloo in the CPU tests.
"""
import os

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


def frame_chunk(n, world):
    """Records per rank when an n-record frame is cut into `world` equal chunks (the last one may be ragged):
    rank r owns [r * chunk, min(n, (r + 1) * chunk)).  Equal chunks are what all_gather_into_tensor wants."""
    return (n + world - 1) // world


class ShardedFrame:
    """One frame traced by every rank (each its own samples), moved over the fabric instead of PCIe.

    At N GPUs the naive end-to-end step uploads the SAME primary rays N times (66 MB per GPU at 1080p) and downloads
    N partial images; GPUs that share a PCIe switch uplink then halve each other's copy bandwidth (measured: 3.7 ms
    per step at 1 GPU, 4.5 ms at 8).  Here every rank uploads only its 1/N chunk of the ray array, the chunks are
    exchanged with ONE all_gather over NVLink/NVSwitch, every rank traces the whole frame for its own sample set
    (`trace_fn`: device ray tensor -> device RGBFFF partial image), the partial images are summed with ONE
    all_reduce and every rank downloads only its 1/N chunk of the finished image.  PCIe traffic per rank and step:
    (32 B up + 12 B down) * n / N instead of * n.  No collective touches the traversal itself.

    Works with any torch.distributed backend (NCCL on the GPUs; gloo in the CPU tests, where trace_fn is the checker).
    """

    def __init__(self, n, device, group=None):
        self.n, self.device, self.group = n, device, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.chunk = frame_chunk(n, self.world)
        self.begin = min(n, self.rank * self.chunk)
        self.end = min(n, self.begin + self.chunk)
        pin = device.type == "cuda"
        self.d_rays = torch.zeros(self.chunk * self.world * 8, dtype=torch.float32, device=device)  # vt_ray = 8 floats
        self.h_fb = torch.zeros(self.chunk * 3, dtype=torch.float32, pin_memory=pin)
        self.h2d_bytes = (self.end - self.begin) * abi.RAY.itemsize
        self.d2h_bytes = (self.end - self.begin) * 12

    def step(self, h_rays, trace_fn):
        """h_rays: torch float32 view (n * 8) of the host ray array (pinned for the GPU path); only this rank's chunk is read.
        Returns (begin, end, image_chunk) — image_chunk is a host float32 array [(end - begin), 3] of the FINISHED image."""
        c, r = self.chunk * 8, self.rank
        mine = self.d_rays[r * c:(r + 1) * c]
        mine[: (self.end - self.begin) * 8].copy_(h_rays[self.begin * 8:self.end * 8], non_blocking=True)
        if self.world > 1:
            dist.all_gather_into_tensor(self.d_rays, mine, group=self.group)
        fb = trace_fn(self.d_rays[: self.n * 8])  # float32 [n * 3], this rank's samples
        if self.world > 1:
            dist.all_reduce(fb, op=dist.ReduceOp.SUM, group=self.group)
        k = (self.end - self.begin) * 3
        self.h_fb[:k].copy_(fb[self.begin * 3:self.end * 3], non_blocking=True)
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()
        return self.begin, self.end, self.h_fb[:k].numpy().reshape(-1, 3)


def bind_to_gpu_numa_node(local_rank):
    """Pin this process (and the OpenMP threads it will create) to the CPUs of the NUMA node its GPU hangs off, so that
    pinned staging buffers are first-touched next to the GPU's PCIe root.  Returns the node or None when the topology
    cannot be read (containers without /sys access, single-node hosts) — never an error."""
    try:
        props = torch.cuda.get_device_properties(local_rank)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None


class SharedPinnedFrame:
    """One host buffer mapped into every process of a one-node job (POSIX shared memory) and pinned by each of them
    (cudaHostRegister): the landing frame of vt_group_render_diffuse_wave(..., VT_GROUP_SHARED_HOST_FRAME) — every rank's GPU
    writes its own tiles into it through its own PCIe link.  `create` on exactly one process, before the others open it."""

    def __init__(self, name, nbytes, create):
        self.path = os.path.join("/dev/shm", name)
        self.nbytes = int(nbytes)
        self.owner = bool(create)
        if create:
            with open(self.path, "wb") as f:
                f.truncate(self.nbytes)
        self.map = np.memmap(self.path, dtype=np.uint8, mode="r+", shape=(self.nbytes,))
        rc = torch.cuda.cudart().cudaHostRegister(self.map.ctypes.data, self.nbytes, 0)
        if int(rc) != 0:
            raise RuntimeError(f"cudaHostRegister of the shared frame failed: {rc}")
        self.registered = True

    def array(self, dtype, shape):
        return self.map.view(dtype).reshape(shape)

    def close(self):
        if getattr(self, "registered", False):
            torch.cuda.cudart().cudaHostUnregister(self.map.ctypes.data)
            self.registered = False
        if self.owner and os.path.exists(self.path):
            os.unlink(self.path)
            self.owner = False
