"""vistrace_b200 — B200-native ray-query engine behind VisTrace's accel:Traverse.

The product is the C-ABI shared library ``libvistrace_b200.so`` (include/vistrace_b200.h):
host-side C++ objects + hand-written sm_100a CUDA kernels.  This Python package is only the
thin ctypes binding tests and bench.py use to reach it; there is no Python compute path and
no CPU fallback — a missing library or a missing GPU is an error, never a silent detour.
"""
from . import abi  # noqa: F401
from .binding import Accel, BspFile, MdlFiles, Group, group_unique_id, shard_indices, sample_uniform01, build_bvh, build_bvh_ploc, build_quads, compact_pairs, flatten_bvh, lib, library_path, quad_plane_offset, refit_bvh, skin_triangles, vtf_decode, vtf_info  # noqa: F401
