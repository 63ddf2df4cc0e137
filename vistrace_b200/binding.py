"""ctypes binding of include/vistrace_b200.h (plumbing; see package docstring)."""
import ctypes as C
import os

import numpy as np

from . import abi

HERE = os.path.dirname(os.path.abspath(__file__))


def library_path():
    # VT_LIB: alternative build of the same library (kernel tuning experiments); default = the in-tree build
    return os.environ.get("VT_LIB") or os.path.join(HERE, "libvistrace_b200.so")


_lib = None

# every symbol include/vistrace_b200.h declares: (restype, argtypes)
_vp, _u64, _u32, _i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
SYMBOLS = {
    "vt_device_count": (_i32, []),
    "vt_accel_create": (_vp, [_i32]),
    "vt_accel_destroy": (None, [_vp]),
    "vt_accel_populate": (_i32, [_vp, _vp]),
    "vt_accel_populate_with_bvh": (_i32, [_vp, _vp, _vp, _u64, _vp]),
    "vt_accel_refit": (_i32, [_vp, _vp]),
    "vt_refit_bvh": (_i32, [_vp, _vp, _u64, _vp]),
    "vt_accel_refit_range": (_i32, [_vp, _vp, _u64, _u64]),
    "vt_accel_get_bvh": (_i32, [_vp, _vp, _vp, _vp, _vp]),
    "vt_accel_traverse": (_i32, [_vp, _vp, _u64, _vp, _vp, _u32, _vp]),
    "vt_accel_traverse_stats": (_i32, [_vp, _vp, _u64, _u32, _vp, _vp]),
    "vt_accel_traverse_ray_stats": (_i32, [_vp, _vp, _u64, _u32, _vp]),
    "vt_accel_render_diffuse_wave_begin": (_i32, [_vp, _vp, _u64, _u32, _u64, C.c_float, _vp]),
    "vt_accel_render_diffuse_wave_wait": (_i32, [_vp]),
    "vt_accel_traverse_cones": (_i32, [_vp, _vp, _vp, _u64, _vp, _vp, _u32, _vp]),
    "vt_accel_trace_result": (_i32, [_vp, _vp, _vp, _u64, _vp, _u32, _vp]),
    "vt_accel_bounce_rays": (_i32, [_vp, _vp, _u64, _u32, _u64, _vp, _vp, _u32, _vp]),
    "vt_accel_sample_bsdf_rays": (_i32, [_vp, _vp, _vp, _u64, _u32, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _u32, _vp]),
    "vt_sample_uniform01": (C.c_float, [_u64, _u32, _u64]),
    "vt_accel_shadow_rays": (_i32, [_vp, _vp, _u64, _vp, _i32, C.c_float, _vp, _vp, _u32, _vp]),
    "vt_accel_bounce_rays_queued": (_i32, [_vp, _vp, _u64, _u32, _u64, _vp, _vp, _vp, _vp, _vp]),
    "vt_accel_shadow_rays_queued": (_i32, [_vp, _vp, _u64, _vp, _i32, C.c_float, _vp, _vp, _vp, _vp, _vp]),
    "vt_accel_traverse_queued": (_i32, [_vp, _vp, _vp, _vp, _u64, _vp, _vp, _u32, _vp]),
    "vt_accel_bounce_rays_requeued": (_i32, [_vp, _vp, _vp, _vp, _u64, _u32, _u64, _vp, _vp, _vp, _vp, _vp]),
    "vt_accel_shadow_rays_requeued": (_i32, [_vp, _vp, _vp, _vp, _u64, _vp, _i32, C.c_float, _vp, _vp, _vp, _vp, _vp]),
    "vt_accel_trace_diffuse_wave": (_i32, [_vp, _vp, _u64, _u32, _u64, _vp, _vp, _vp, _vp, _vp, _u32, _vp]),
    "vt_accel_render_diffuse_wave": (_i32, [_vp, _vp, _u64, _u32, _u64, C.c_float, _vp, _vp]),
    "vt_accel_trace_paths": (_i32, [_vp, _vp, _u64, _u32, _vp, _vp, _u64, C.c_float, _vp, _vp, _u32, _vp]),
    "vt_accel_accumulate_sky": (_i32, [_vp, _vp, _vp, _u64, _u32, C.c_float, _vp, _vp]),
    "vt_accel_set_layout": (_i32, [_vp, _i32]),
    "vt_accel_get_layout": (_i32, [_vp]),
    "vt_compact_pairs": (_i32, [_vp, _u64, _vp]),
    "vt_skin_triangles": (_i32, [_vp, _vp, _u64, _vp, _vp, _u32]),
    "vt_quad_plane_offset": (_u32, []),
    "vt_vtf_read_info": (_i32, [_vp, _u64, _vp]),
    "vt_vtf_decode": (_i32, [_vp, _u64, _u32, _u32, _vp, _u64, _vp]),
    "vt_mdl_read_info": (_i32, [_vp, _vp]),
    "vt_mdl_bodygroup_values": (_i32, [_vp, _u32, _vp]),
    "vt_mdl_mesh_triangles": (_i32, [_vp, _u32, _u32, _vp, _vp, _vp]),
    "vt_mdl_bind_matrices": (_i32, [_vp, _vp]),
    "vt_mdl_material_index": (_i32, [_vp, _u32, _u32, _vp]),
    "vt_mdl_material_path": (_i32, [_vp, _u32, _u32, _vp, _u64]),
    "vt_bsp_read_info": (_i32, [_vp, _u64, _vp]),
    "vt_bsp_triangles": (_i32, [_vp, _u64, _vp, _vp, _vp, _vp]),
    "vt_bsp_get_material": (_i32, [_vp, _u64, _u32, _vp]),
    "vt_bsp_get_static_prop": (_i32, [_vp, _u64, _u32, _vp]),
    "vt_build_bvh_ploc": (_i32, [_vp, _i32, _vp, _vp, _vp]),
    "vt_optimize_bvh": (_i32, [_vp, _u64, _i32, C.c_double, _vp, _vp, _vp]),
    "vt_build_quads": (_i32, [_vp, _u64, _vp, _u64, _vp, _vp, _vp, _vp, _vp]),
    "vt_accel_refit_quality": (_i32, [_vp, _vp, _vp]),
    "vt_accel_set_refit_rebuild_ratio": (_i32, [_vp, C.c_double]),
    "vt_accel_invalid_rays": (_u64, [_vp]),
    "vt_accel_launch_count": (_u64, [_vp]),
    "vt_accel_stats": (_i32, [_vp, _vp, _vp, _vp]),
    "vt_accel_get_tri_derived": (_i32, [_vp, _vp]),
    "vt_build_bvh": (_i32, [_vp, _vp, _vp, _vp]),
    "vt_flatten_bvh": (_i32, [_vp, _u64, _vp, _u64, _u32, _vp, _vp, _vp, _vp]),
    "vt_group_unique_id": (_i32, [_vp]),
    "vt_group_create": (_vp, [_vp, _i32]),
    "vt_group_create_rank": (_vp, [_i32, _i32, _i32, _vp]),
    "vt_group_destroy": (None, [_vp]),
    "vt_group_size": (_i32, [_vp]),
    "vt_group_rank": (_i32, [_vp]),
    "vt_group_local_members": (_i32, [_vp]),
    "vt_group_accel": (_vp, [_vp, _i32]),
    "vt_group_populate": (_i32, [_vp, _vp]),
    "vt_group_traverse": (_i32, [_vp, _vp, _u64, _vp, _vp, _u32]),
    "vt_group_shard": (_i32, [_vp, _u64, _i32, _vp, _vp]),
    "vt_shard_geometry": (_i32, [_u64, _i32, _i32, _u64, _vp, _vp]),
    "vt_group_render_diffuse_wave": (_i32, [_vp, _vp, _u64, _u32, _u64, C.c_float, _vp, _vp, _u32, _vp]),
    "vt_group_reduce_device": (_i32, [_vp, _vp, _u64, _vp]),
    "vt_group_all_gather_device": (_i32, [_vp, _vp, _u64, _vp]),
    "vt_group_launch_count": (_u64, [_vp]),
    "vt_group_wait_frame": (_i32, [_vp]),
    "vt_last_error": (C.c_char_p, []),
}


def lib():
    """Load libvistrace_b200.so or fail loudly (there is no fallback)."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C vistrace_b200/csrc`; vistrace_b200 has no CPU fallback"
            )
        L = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _err():
    return lib().vt_last_error().decode()


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what}: {_err()}")


def _ptr(x):
    """numpy array -> host address; int -> passed through (device pointer); None -> NULL."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return int(x)


def build_bvh(scene):
    """Host-only hierarchy build -> (nodes, prim_indices) in bvh::Bvh<float> form."""
    L = lib()
    nodes = np.zeros(max(1, 2 * scene.n_tris), abi.NODE)  # a binary tree over n leaves has at most 2n - 1 nodes: ONE build, not two
    prims = np.zeros(scene.n_tris, np.uint64)
    cap = C.c_uint64(len(nodes))
    _check(L.vt_build_bvh(C.cast(scene.ptr(), _vp), nodes.ctypes.data, C.addressof(cap), prims.ctypes.data), "vt_build_bvh")
    return nodes[: cap.value].copy(), prims


def build_bvh_ploc(scene, collapse=True):
    """Host-only: the reference's own PLOC (+ LeafCollapser) hierarchy rebuilt from its algorithm -> (nodes, prim_indices)."""
    L = lib()
    nodes = np.zeros(max(1, 2 * scene.n_tris), abi.NODE)
    prims = np.zeros(scene.n_tris, np.uint64)
    cap = C.c_uint64(len(nodes))
    _check(L.vt_build_bvh_ploc(C.cast(scene.ptr(), _vp), int(collapse), nodes.ctypes.data, C.addressof(cap), prims.ctypes.data), "vt_build_bvh_ploc")
    return nodes[: cap.value].copy(), prims


def optimize_bvh(nodes, iterations=4, fraction=0.05):
    """Host-only reinsertion optimisation -> (optimised COPY of `nodes`, inner-node area before, after, moves applied);
    prim_indices stays valid (leaves keep their ranges)."""
    out = np.ascontiguousarray(nodes, abi.NODE).copy()
    before, after, moves = C.c_double(0), C.c_double(0), C.c_uint64(0)
    _check(lib().vt_optimize_bvh(out.ctypes.data, len(out), int(iterations), float(fraction), C.addressof(before), C.addressof(after),
                                 C.addressof(moves)), "vt_optimize_bvh")
    return out, before.value, after.value, moves.value


def refit_bvh(scene, nodes, prim_indices):
    """Host-only bvh::HierarchyRefitter step: a refitted COPY of `nodes` for the (moved) triangles of `scene`."""
    out = np.ascontiguousarray(nodes, abi.NODE).copy()
    prims = np.ascontiguousarray(prim_indices, np.uint64)
    _check(lib().vt_refit_bvh(C.cast(scene.ptr(), _vp), out.ctypes.data, len(out), prims.ctypes.data), "vt_refit_bvh")
    return out


def skin_triangles(tris, skin, bones, binds):
    """Host-only SkinTriangle (source/objects/AccelStruct.cpp:66-108): returns a skinned COPY of the vt_tri_in records.
    bones/binds: [n_bones, 16] glm::mat4 (column-major); skin: abi.TRI_SKIN records or None (one-bone overload)."""
    out = np.ascontiguousarray(tris, abi.TRI_IN).copy()
    bones = np.ascontiguousarray(bones, np.float32).reshape(-1, 16)
    binds = np.ascontiguousarray(binds, np.float32).reshape(-1, 16)
    if len(bones) != len(binds):
        raise ValueError("bones and binds must have the same length")
    sk = None if skin is None else np.ascontiguousarray(skin, abi.TRI_SKIN)
    _check(lib().vt_skin_triangles(out.ctypes.data, _ptr(sk), len(out), bones.ctypes.data, binds.ctypes.data, len(bones)), "vt_skin_triangles")
    return out


PAIR = np.dtype(
    [
        ("l_bounds", np.float32, 6),
        ("l_count", np.uint32),
        ("l_first", np.uint32),
        ("r_bounds", np.float32, 6),
        ("r_count", np.uint32),
        ("r_first", np.uint32),
    ]
)


def flatten_bvh(nodes, prim_indices, bfs_pairs=384):
    """Host-only flatten -> dict(pairs, leaf_order, root_leaf_count, max_depth)."""
    L = lib()
    nodes = np.ascontiguousarray(nodes, abi.NODE)
    prim_indices = np.ascontiguousarray(prim_indices, np.uint64)
    n_pairs = max(0, (len(nodes) - 1) // 2)
    pairs = np.zeros(n_pairs, PAIR)
    order = np.zeros(len(prim_indices), np.uint32)
    rl, md = C.c_uint32(0), C.c_uint32(0)
    _check(
        L.vt_flatten_bvh(nodes.ctypes.data, len(nodes), prim_indices.ctypes.data, len(prim_indices), bfs_pairs,
                         pairs.ctypes.data, order.ctypes.data, C.addressof(rl), C.addressof(md)),
        "vt_flatten_bvh",
    )
    return {"pairs": pairs, "leaf_order": order, "root_leaf_count": rl.value, "max_depth": md.value}


CPAIR = np.dtype([("origin_adj", np.float32, 3), ("exp", np.uint8, 3), ("counts", np.uint8), ("q", np.uint8, (3, 4)), ("ref", np.uint32)])
assert PAIR.itemsize == 64 and CPAIR.itemsize == 32


def compact_pairs(pairs):
    """Host-only: depth-first 64-byte pairs -> 32-byte conservative compact pairs."""
    pairs = np.ascontiguousarray(pairs, PAIR)
    out = np.zeros(len(pairs), CPAIR)
    _check(lib().vt_compact_pairs(pairs.ctypes.data, len(pairs), out.ctypes.data), "vt_compact_pairs")
    return out


QUAD = np.dtype([("origin_adj", np.float32, 3), ("scale", np.float32, 3), ("q", np.uint8, (3, 2, 4)), ("ref", np.uint32, 4)])
assert QUAD.itemsize == 64


VTF_INFO = np.dtype([("width", np.uint32), ("height", np.uint32), ("mip_count", np.uint32), ("flags", np.uint32), ("format", np.int32),
                     ("frames", np.uint32), ("faces", np.uint32), ("depth", np.uint32), ("supported", np.uint32), ("texel_layout", np.uint32),
                     ("rgba_bytes", np.uint64)])


def vtf_info(data):
    """Header of a VTF file held in `data` (bytes): a VTF_INFO record."""
    buf = np.frombuffer(bytes(data), np.uint8)
    info = np.zeros(1, VTF_INFO)
    _check(lib().vt_vtf_read_info(buf.ctypes.data, len(buf), info.ctypes.data), "vt_vtf_read_info")
    return info[0]


def vtf_decode(data, frame=0, face=0):
    """VTF file -> (width, height, mip_count, flags, texel chain as uint8, smallest mip first, texel_layout): a SceneData texture tuple."""
    buf = np.frombuffer(bytes(data), np.uint8)
    info = vtf_info(data)
    out = np.zeros(int(info["rgba_bytes"]), np.uint8)
    _check(lib().vt_vtf_decode(buf.ctypes.data, len(buf), frame, face, out.ctypes.data, len(out), None), "vt_vtf_decode")
    return int(info["width"]), int(info["height"]), int(info["mip_count"]), int(info["flags"]), out, int(info["texel_layout"])


MDL_INFO = np.dtype([("version", np.uint32), ("n_bodygroups", np.uint32), ("n_bones", np.uint32), ("n_materials", np.uint32), ("n_material_dirs", np.uint32),
                     ("n_skin_refs", np.uint32), ("n_skin_families", np.uint32), ("n_vertices", np.uint32)])


class MdlFiles:
    """The three files of a Source-engine model (bytes) behind vt_mdl_* (host-only ingestion, include/vistrace_b200.h)."""

    class _C(C.Structure):
        _fields_ = [("mdl", _vp), ("mdl_size", _u64), ("vvd", _vp), ("vvd_size", _u64), ("vtx", _vp), ("vtx_size", _u64)]

    def __init__(self, mdl, vvd, vtx):
        self._bufs = [np.frombuffer(bytes(b), np.uint8).copy() if len(b) else np.zeros(0, np.uint8) for b in (mdl, vvd, vtx)]
        self.c = self._C(*[v for b in self._bufs for v in (b.ctypes.data if len(b) else None, len(b))])
        self.L = lib()

    def _p(self):
        return C.addressof(self.c)

    def info(self):
        out = np.zeros(1, MDL_INFO)
        _check(self.L.vt_mdl_read_info(self._p(), out.ctypes.data), "vt_mdl_read_info")
        return out[0]

    def bodygroup_values(self, bodygroup):
        n = C.c_uint32(0)
        _check(self.L.vt_mdl_bodygroup_values(self._p(), bodygroup, C.addressof(n)), "vt_mdl_bodygroup_values")
        return n.value

    def mesh_triangles(self, bodygroup, value):
        """(vt_tri_in records in model space, vt_tri_skin records) of Model::GetMesh(bodygroup, value)."""
        n = C.c_uint64(0)
        _check(self.L.vt_mdl_mesh_triangles(self._p(), bodygroup, value, None, None, C.addressof(n)), "vt_mdl_mesh_triangles")
        tris, skin = np.zeros(n.value, abi.TRI_IN), np.zeros(n.value, abi.TRI_SKIN)
        if n.value:
            _check(self.L.vt_mdl_mesh_triangles(self._p(), bodygroup, value, tris.ctypes.data, skin.ctypes.data, C.addressof(n)), "vt_mdl_mesh_triangles")
        return tris, skin

    def bind_matrices(self):
        out = np.zeros((int(self.info()["n_bones"]), 16), np.float32)
        if len(out):
            _check(self.L.vt_mdl_bind_matrices(self._p(), out.ctypes.data), "vt_mdl_bind_matrices")
        return out

    def material_index(self, skin, material_id):
        v = C.c_int32(0)
        _check(self.L.vt_mdl_material_index(self._p(), skin, material_id, C.addressof(v)), "vt_mdl_material_index")
        return v.value

    def material_path(self, material_id, directory=0):
        buf = C.create_string_buffer(4200)
        _check(self.L.vt_mdl_material_path(self._p(), material_id, directory, buf, len(buf)), "vt_mdl_material_path")
        return buf.value.decode("latin-1")


class BspFile:
    """A Source-engine map (bytes) behind vt_bsp_* (host-only ingestion, include/vistrace_b200.h)."""

    def __init__(self, data):
        self._buf = np.frombuffer(bytes(data), np.uint8).copy() if len(data) else np.zeros(0, np.uint8)
        self.L = lib()

    def _a(self):
        return (self._buf.ctypes.data if len(self._buf) else None), len(self._buf)

    def info(self):
        out = np.zeros(1, abi.BSP_INFO)
        _check(self.L.vt_bsp_read_info(*self._a(), out.ctypes.data), "vt_bsp_read_info")
        return out[0]

    def triangles(self):
        """(vt_tri_in records of the world, binormals [n, 3, 3], texinfo index per triangle)."""
        n = C.c_uint64(0)
        _check(self.L.vt_bsp_triangles(*self._a(), None, None, None, C.addressof(n)), "vt_bsp_triangles")
        tris, bino, texinfo = np.zeros(n.value, abi.TRI_IN), np.zeros((n.value, 3, 3), np.float32), np.zeros(n.value, np.int16)
        _check(self.L.vt_bsp_triangles(*self._a(), tris.ctypes.data, bino.ctypes.data, texinfo.ctypes.data, C.addressof(n)), "vt_bsp_triangles")
        return tris, bino, texinfo

    def material(self, index):
        out = np.zeros(1, abi.BSP_MATERIAL)
        _check(self.L.vt_bsp_get_material(*self._a(), index, out.ctypes.data), "vt_bsp_get_material")
        return out[0]

    def static_prop(self, index):
        out = np.zeros(1, abi.BSP_STATIC_PROP)
        _check(self.L.vt_bsp_get_static_prop(*self._a(), index, out.ctypes.data), "vt_bsp_get_static_prop")
        return out[0]


def quad_plane_offset():
    """OFFSET of the quad layout's plane decode: plane = (OFFSET + q) * scale + origin_adj."""
    return int(lib().vt_quad_plane_offset())


def build_quads(nodes, prim_indices):
    """Host-only: binary hierarchy -> dict(quads, leaf_order, root_leaf_count, max_stack) of the 4-wide layout."""
    L = lib()
    nodes = np.ascontiguousarray(nodes, abi.NODE)
    prim_indices = np.ascontiguousarray(prim_indices, np.uint64)
    args = (nodes.ctypes.data, len(nodes), prim_indices.ctypes.data, len(prim_indices))
    quads = np.zeros(max(1, len(nodes)), QUAD)  # never more quads than binary nodes: one call
    cnt = C.c_uint64(len(quads))
    order = np.zeros(len(prim_indices), np.uint32)
    rl, ms = C.c_uint32(0), C.c_uint32(0)
    _check(L.vt_build_quads(*args, quads.ctypes.data, C.addressof(cnt), order.ctypes.data, C.addressof(rl), C.addressof(ms)), "vt_build_quads")
    return {"quads": quads[: cnt.value].copy(), "leaf_order": order, "root_leaf_count": rl.value, "max_stack": ms.value}


class Accel:
    """vt_accel handle: the AccelStruct of source/objects/AccelStruct.h:61-86 bound to one GPU."""

    def __init__(self, device=0, layout=None):
        """layout: "compact" (32-byte conservative pairs, default), "exact" (the reference's nodes verbatim), or
        None = the library default (VT_LAYOUT in the environment, else compact)."""
        self.L = lib()
        self.h = self.L.vt_accel_create(device)
        if not self.h:
            raise RuntimeError(f"vt_accel_create: {_err()}")
        self.scene = None
        if layout is not None:
            _check(self.L.vt_accel_set_layout(self.h, {"exact": 0, "compact": 1, "quad": 2}[layout]), "vt_accel_set_layout")

    @classmethod
    def borrowed(cls, handle, scene=None):
        """Wrap a vt_accel* owned by something else (a vt_group member): never destroyed from here."""
        self = cls.__new__(cls)
        self.L, self.h, self.scene, self._borrowed = lib(), handle, scene, True
        return self

    @property
    def layout(self):
        """Layout resident after populate ("compact" falls back to "exact" for trees it cannot hold)."""
        return ("exact", "compact", "quad")[self.L.vt_accel_get_layout(self.h)]

    def close(self):
        if getattr(self, "h", None):
            if not getattr(self, "_borrowed", False):
                self.L.vt_accel_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def populate(self, scene, bvh=None):
        """PopulateAccel: build on the host (or take the caller's bvh=(nodes, prim_indices)) and upload."""
        self.scene = scene
        if bvh is None:
            _check(self.L.vt_accel_populate(self.h, C.cast(scene.ptr(), _vp)), "vt_accel_populate")
        else:
            nodes = np.ascontiguousarray(bvh[0], abi.NODE)
            prims = np.ascontiguousarray(bvh[1], np.uint64)
            _check(
                self.L.vt_accel_populate_with_bvh(self.h, C.cast(scene.ptr(), _vp), nodes.ctypes.data, len(nodes), prims.ctypes.data),
                "vt_accel_populate_with_bvh",
            )
        return self

    def refit(self, scene):
        """accel:Rebuild for moved geometry of unchanged topology: keep the hierarchy, refit its boxes, re-upload."""
        self.scene = scene
        _check(self.L.vt_accel_refit(self.h, C.cast(scene.ptr(), _vp)), "vt_accel_refit")
        return self

    def refit_range(self, tris, first):
        """New vertices / attributes for the triangles [first, first + len(tris)) of the populated scene (quad layout)."""
        tris = np.ascontiguousarray(tris, abi.TRI_IN)
        _check(self.L.vt_accel_refit_range(self.h, tris.ctypes.data, first, len(tris)), "vt_accel_refit_range")
        self.scene.tris[first:first + len(tris)] = tris  # keep the Python-side copy in step (tri_derived sizes, later refits)
        return self

    def refit_quality(self):
        """(node-area sum now / as built, rebuilds triggered so far)."""
        ratio, n = C.c_double(0), C.c_uint64(0)
        _check(self.L.vt_accel_refit_quality(self.h, C.addressof(ratio), C.addressof(n)), "vt_accel_refit_quality")
        return ratio.value, n.value

    def set_refit_rebuild_ratio(self, ratio):
        _check(self.L.vt_accel_set_refit_rebuild_ratio(self.h, float(ratio)), "vt_accel_set_refit_rebuild_ratio")
        return self

    def get_bvh(self):
        cnt, nt = C.c_uint64(0), C.c_uint64(0)
        _check(self.L.vt_accel_get_bvh(self.h, None, C.addressof(cnt), None, C.addressof(nt)), "vt_accel_get_bvh")
        nodes = np.zeros(cnt.value, abi.NODE)
        prims = np.zeros(nt.value, np.uint64)
        _check(self.L.vt_accel_get_bvh(self.h, nodes.ctypes.data, None, prims.ctypes.data, None), "vt_accel_get_bvh")
        return nodes, prims

    def tri_derived(self):
        out = np.zeros((self.scene.n_tris, 16), np.float32)
        _check(self.L.vt_accel_get_tri_derived(self.h, out.ctypes.data), "vt_accel_get_tri_derived")
        return out

    def traverse(self, rays, want_attrs=False, any_hit=False, cones=None):
        """Host-buffer call: numpy rays in, numpy hits (and attrs) out, synchronous."""
        rays = np.ascontiguousarray(rays, abi.RAY)
        hits = np.zeros(len(rays), abi.HIT)
        attrs = np.zeros(len(rays), abi.ATTR) if want_attrs else None
        flags = abi.VT_TRAVERSE_ANY_HIT if any_hit else 0
        if cones is None:
            rc = self.L.vt_accel_traverse(self.h, rays.ctypes.data, len(rays), hits.ctypes.data, _ptr(attrs), flags, None)
        else:
            cones = np.ascontiguousarray(cones, np.float32)
            rc = self.L.vt_accel_traverse_cones(self.h, rays.ctypes.data, cones.ctypes.data, len(rays), hits.ctypes.data,
                                                _ptr(attrs), flags, None)
        _check(rc, "vt_accel_traverse")
        return (hits, attrs) if want_attrs else hits

    def traverse_stats(self, rays, n=None):
        """(pair visits, triangle tests) summed over the batch; rays = numpy array or (device pointer, n)."""
        steps, tests = C.c_uint64(0), C.c_uint64(0)
        if isinstance(rays, np.ndarray):
            rays = np.ascontiguousarray(rays, abi.RAY)
            n, flags = len(rays), 0
        else:
            flags = abi.VT_TRAVERSE_DEVICE_PTRS
        _check(self.L.vt_accel_traverse_stats(self.h, _ptr(rays), n, flags, C.addressof(steps), C.addressof(tests)), "vt_accel_traverse_stats")
        return steps.value, tests.value

    def traverse_ray_stats(self, rays):
        """(steps, tests) per ray (quad / compact layouts), each saturating at 65535."""
        rays = np.ascontiguousarray(rays, abi.RAY)
        out = np.zeros(len(rays), np.uint32)
        _check(self.L.vt_accel_traverse_ray_stats(self.h, rays.ctypes.data, len(rays), 0, out.ctypes.data), "vt_accel_traverse_ray_stats")
        return out & 0xFFFF, out >> 16

    def traverse_device(self, d_rays, n, d_hits, d_attrs=None, any_hit=False, stream=None):
        """Device-pointer call (ints = CUDA device addresses): enqueues on `stream` and returns."""
        flags = abi.VT_TRAVERSE_DEVICE_PTRS | (abi.VT_TRAVERSE_ANY_HIT if any_hit else 0)
        _check(self.L.vt_accel_traverse(self.h, _ptr(d_rays), n, _ptr(d_hits), _ptr(d_attrs), flags, _ptr(stream)), "vt_accel_traverse")

    def trace_result(self, rays, hits):
        rays = np.ascontiguousarray(rays, abi.RAY)
        hits = np.ascontiguousarray(hits, abi.HIT)
        attrs = np.zeros(len(rays), abi.ATTR)
        _check(
            self.L.vt_accel_trace_result(self.h, rays.ctypes.data, hits.ctypes.data, len(rays), attrs.ctypes.data, 0, None),
            "vt_accel_trace_result",
        )
        return attrs

    def bounce_rays(self, attrs, spp, seed=0):
        """Host-buffer K3: (rays[n*spp] with masked slots, number spawned)."""
        attrs = np.ascontiguousarray(attrs, abi.ATTR)
        out = np.zeros(len(attrs) * spp, abi.RAY)
        live = C.c_uint64(0)
        _check(self.L.vt_accel_bounce_rays(self.h, attrs.ctypes.data, len(attrs), spp, seed, out.ctypes.data, C.addressof(live), 0, None),
               "vt_accel_bounce_rays")
        return out, live.value

    def sample_bsdf_rays(self, rays, attrs, spp, seed=0):
        """Host-buffer batched SampleBSDF (diffuse lobe): (rays[n*spp] with masked slots, BSDF_SAMPLE records[n*spp], number spawned)."""
        rays = np.ascontiguousarray(rays, abi.RAY)
        attrs = np.ascontiguousarray(attrs, abi.ATTR)
        out = np.zeros(len(attrs) * spp, abi.RAY)
        samples = np.zeros(len(attrs) * spp, abi.BSDF_SAMPLE)
        live = C.c_uint64(0)
        _check(self.L.vt_accel_sample_bsdf_rays(self.h, rays.ctypes.data, attrs.ctypes.data, len(attrs), spp, seed, out.ctypes.data, samples.ctypes.data,
                                                C.addressof(live), None, None, None, 0, None), "vt_accel_sample_bsdf_rays")
        return out, samples, live.value

    def sample_bsdf_rays_device(self, d_rays, d_attrs, n, spp, seed, d_out, d_samples, d_queue=None, d_queue_count=None, d_miss_hits=None, stream=None):
        _check(self.L.vt_accel_sample_bsdf_rays(self.h, _ptr(d_rays), _ptr(d_attrs), n, spp, seed, _ptr(d_out), _ptr(d_samples), None, _ptr(d_queue),
                                                _ptr(d_queue_count), _ptr(d_miss_hits), abi.VT_TRAVERSE_DEVICE_PTRS, _ptr(stream)), "vt_accel_sample_bsdf_rays")

    def shadow_rays(self, attrs, light, point_light=False, tmax=3.4028234663852886e38):
        """Host-buffer shadow-ray generation: (rays[n] with masked slots, number spawned)."""
        attrs = np.ascontiguousarray(attrs, abi.ATTR)
        out = np.zeros(len(attrs), abi.RAY)
        live = C.c_uint64(0)
        lv = (C.c_float * 3)(*[float(v) for v in light])
        _check(self.L.vt_accel_shadow_rays(self.h, attrs.ctypes.data, len(attrs), C.cast(lv, _vp), int(point_light), tmax, out.ctypes.data,
                                           C.addressof(live), 0, None), "vt_accel_shadow_rays")
        return out, live.value

    def shadow_rays_device(self, d_attrs, n, light, d_out, point_light=False, tmax=3.4028234663852886e38, stream=None):
        lv = (C.c_float * 3)(*[float(v) for v in light])
        _check(self.L.vt_accel_shadow_rays(self.h, _ptr(d_attrs), n, C.cast(lv, _vp), int(point_light), tmax, _ptr(d_out), None,
                                           abi.VT_TRAVERSE_DEVICE_PTRS, _ptr(stream)), "vt_accel_shadow_rays")

    def bounce_rays_device(self, d_attrs, n, spp, seed, d_out, stream=None):
        _check(self.L.vt_accel_bounce_rays(self.h, _ptr(d_attrs), n, spp, seed, _ptr(d_out), None, abi.VT_TRAVERSE_DEVICE_PTRS, _ptr(stream)),
               "vt_accel_bounce_rays")

    # ---- ray queue (device pointers): the generator lists the slots it filled, the traversal visits only those
    def bounce_rays_queued_device(self, d_attrs, n, spp, seed, d_out, d_queue, d_queue_count, d_miss_hits, stream=None):
        _check(self.L.vt_accel_bounce_rays_queued(self.h, _ptr(d_attrs), n, spp, seed, _ptr(d_out), _ptr(d_queue), _ptr(d_queue_count),
                                                  _ptr(d_miss_hits), _ptr(stream)), "vt_accel_bounce_rays_queued")

    def shadow_rays_queued_device(self, d_attrs, n, light, d_out, d_queue, d_queue_count, d_miss_hits, point_light=False,
                                  tmax=3.4028234663852886e38, stream=None):
        lv = (C.c_float * 3)(*[float(x) for x in light])
        _check(self.L.vt_accel_shadow_rays_queued(self.h, _ptr(d_attrs), n, C.cast(lv, _vp), int(point_light), tmax, _ptr(d_out),
                                                  _ptr(d_queue), _ptr(d_queue_count), _ptr(d_miss_hits), _ptr(stream)),
               "vt_accel_shadow_rays_queued")

    def traverse_queued_device(self, d_rays, d_queue, d_queue_count, capacity, d_hits, d_attrs=None, any_hit=False, stream=None, queue_attrs=False):
        """queue_attrs: TraceResult only of the slots the queue lists (wave compaction)."""
        flags = (abi.VT_TRAVERSE_ANY_HIT if any_hit else 0) | (abi.VT_TRAVERSE_QUEUE_ATTRS if queue_attrs else 0)
        _check(self.L.vt_accel_traverse_queued(self.h, _ptr(d_rays), _ptr(d_queue), _ptr(d_queue_count), capacity, _ptr(d_hits), _ptr(d_attrs),
                                               flags, _ptr(stream)), "vt_accel_traverse_queued")

    # ---- wave compaction: generators driven by the previous wave's queue
    def bounce_rays_requeued_device(self, d_attrs, d_in_queue, d_in_count, n, spp, seed, d_out, d_queue, d_queue_count, d_miss_hits, stream=None):
        _check(self.L.vt_accel_bounce_rays_requeued(self.h, _ptr(d_attrs), _ptr(d_in_queue), _ptr(d_in_count), n, spp, seed, _ptr(d_out), _ptr(d_queue),
                                                    _ptr(d_queue_count), _ptr(d_miss_hits), _ptr(stream)), "vt_accel_bounce_rays_requeued")

    def shadow_rays_requeued_device(self, d_attrs, d_in_queue, d_in_count, n, light, d_out, d_queue, d_queue_count, d_miss_hits, point_light=False,
                                    tmax=3.4028234663852886e38, stream=None):
        lv = (C.c_float * 3)(*[float(x) for x in light])
        _check(self.L.vt_accel_shadow_rays_requeued(self.h, _ptr(d_attrs), _ptr(d_in_queue), _ptr(d_in_count), n, C.cast(lv, _vp), int(point_light), tmax,
                                                    _ptr(d_out), _ptr(d_queue), _ptr(d_queue_count), _ptr(d_miss_hits), _ptr(stream)),
               "vt_accel_shadow_rays_requeued")

    def trace_diffuse_wave(self, rays, spp, seed=0, want_attrs=False, want_bounce_rays=False, out=None):
        """Host-buffer wave.  `out` may carry preallocated (e.g. pinned) numpy views: hits, bounce_hits, attrs, bounce_rays."""
        rays = rays if isinstance(rays, np.ndarray) and rays.dtype == abi.RAY and rays.flags.c_contiguous else np.ascontiguousarray(rays, abi.RAY)
        n = len(rays)
        out = dict(out or {})
        if "hits" not in out:
            out["hits"] = np.zeros(n, abi.HIT)
        if "bounce_hits" not in out:
            out["bounce_hits"] = np.zeros(n * spp, abi.HIT)
        if want_attrs and "attrs" not in out:
            out["attrs"] = np.zeros(n, abi.ATTR)
        if want_bounce_rays and "bounce_rays" not in out:
            out["bounce_rays"] = np.zeros(n * spp, abi.RAY)
        live = C.c_uint64(0)
        _check(self.L.vt_accel_trace_diffuse_wave(self.h, rays.ctypes.data, n, spp, seed, out["hits"].ctypes.data, _ptr(out.get("attrs")),
                                                  _ptr(out.get("bounce_rays")), out["bounce_hits"].ctypes.data, C.addressof(live), 0, None),
               "vt_accel_trace_diffuse_wave")
        out["live_bounce"] = live.value
        return out

    def trace_diffuse_wave_device(self, d_rays, n, spp, seed, d_hits, d_attrs, d_bounce_rays, d_bounce_hits, stream=None):
        _check(self.L.vt_accel_trace_diffuse_wave(self.h, _ptr(d_rays), n, spp, seed, _ptr(d_hits), _ptr(d_attrs), _ptr(d_bounce_rays),
                                                  _ptr(d_bounce_hits), None, abi.VT_TRAVERSE_DEVICE_PTRS, _ptr(stream)),
               "vt_accel_trace_diffuse_wave")

    def render_diffuse_wave(self, rays, spp, seed=0, weight=1.0, out=None):
        """Host rays in, host RGBFFF framebuffer out (numpy float32 [n, 3]); returns (framebuffer, bounce rays spawned)."""
        rays = rays if isinstance(rays, np.ndarray) and rays.dtype == abi.RAY and rays.flags.c_contiguous else np.ascontiguousarray(rays, abi.RAY)
        fb = np.empty((len(rays), 3), np.float32) if out is None else out
        live = C.c_uint64(0)
        _check(self.L.vt_accel_render_diffuse_wave(self.h, rays.ctypes.data, len(rays), spp, seed, weight, fb.ctypes.data, C.addressof(live)),
               "vt_accel_render_diffuse_wave")
        return fb, live.value

    def render_diffuse_wave_begin(self, rays, spp, seed, weight, out):
        """Asynchronous form: enqueue the frame (rays / out: pinned numpy arrays that stay alive) and return; at most two in flight."""
        _check(self.L.vt_accel_render_diffuse_wave_begin(self.h, rays.ctypes.data, len(rays), spp, seed, weight, out.ctypes.data),
               "vt_accel_render_diffuse_wave_begin")

    def render_diffuse_wave_wait(self):
        """Block until the oldest frame begun with render_diffuse_wave_begin is complete in its framebuffer."""
        _check(self.L.vt_accel_render_diffuse_wave_wait(self.h), "vt_accel_render_diffuse_wave_wait")

    def trace_paths_device(self, d_rays, n, bounces, sun_dir, sun_rgb, seed, weight, d_fb, want_counts=False, compact=True, stream=None, slot=0):
        """Path waves with compaction over device-resident primary rays; returns the per-wave ray counts when asked (synchronous then)."""
        sd = (C.c_float * 3)(*[float(x) for x in sun_dir])
        sc = (C.c_float * 3)(*[float(x) for x in sun_rgb])
        counts = np.zeros(2 + 2 * bounces, np.uint64) if want_counts else None
        flags = abi.VT_TRAVERSE_DEVICE_PTRS | (0 if compact else abi.VT_PATHS_NO_COMPACTION) | (abi.VT_PATHS_SLOT1 if slot else 0)
        _check(self.L.vt_accel_trace_paths(self.h, _ptr(d_rays), n, bounces, C.cast(sd, _vp), C.cast(sc, _vp), seed, weight, _ptr(d_fb), _ptr(counts),
                                           flags, _ptr(stream)), "vt_accel_trace_paths")
        return counts

    def accumulate_sky_device(self, d_attrs, d_bounce_hits, n, spp, weight, d_fb, stream=None):
        _check(self.L.vt_accel_accumulate_sky(self.h, _ptr(d_attrs), _ptr(d_bounce_hits), n, spp, weight, _ptr(d_fb), _ptr(stream)),
               "vt_accel_accumulate_sky")

    @property
    def invalid_rays(self):
        return int(self.L.vt_accel_invalid_rays(self.h))

    @property
    def launch_count(self):
        return int(self.L.vt_accel_launch_count(self.h))

    def stats(self):
        a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        _check(self.L.vt_accel_stats(self.h, C.addressof(a), C.addressof(b), C.addressof(c)), "vt_accel_stats")
        return {"n_tris": a.value, "node_count": b.value, "device_bytes": c.value}


def shard_indices(n, world, rank, tile):
    """Host-only: global record indices of `rank`'s shard of an n-record frame cut into `tile`-record tiles dealt round-robin,
    in the shard's compact order (vt_shard_geometry gives the count; the order is tile by tile)."""
    cnt, lt = C.c_uint64(0), C.c_uint64(0)
    _check(lib().vt_shard_geometry(n, world, rank, tile, C.addressof(cnt), C.addressof(lt)), "vt_shard_geometry")
    tiles = np.arange(rank, rank + lt.value * world, world, dtype=np.int64)
    idx = (tiles[:, None] * tile + np.arange(tile, dtype=np.int64)[None, :]).reshape(-1)
    idx = idx[idx < n]
    assert len(idx) == cnt.value
    return idx


def sample_uniform01(slots, dim, seed):
    """The generators' counter-based random numbers for an array of slots (host-only)."""
    L = lib()
    return np.array([L.vt_sample_uniform01(int(s), dim, seed) for s in np.asarray(slots).reshape(-1)], np.float32)


def group_unique_id():
    """128-byte ncclUniqueId (numpy uint8) for vt_group_create_rank: rank 0 calls this, the launcher distributes it."""
    out = np.zeros(128, np.uint8)
    _check(lib().vt_group_unique_id(out.ctypes.data), "vt_group_unique_id")
    return out


class Group:
    """vt_group handle: one AccelStruct resident on several GPUs (include/vistrace_b200.h, multi-GPU section).

    Group(devices=[0, 1, ...])                       one process drives the listed local GPUs
    Group(device=d, rank=r, world=w, unique_id=id)   one process per GPU; `id` from group_unique_id() on rank 0
    """

    def __init__(self, devices=None, device=None, rank=None, world=None, unique_id=None):
        self.L = lib()
        if devices is not None:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            self.h = self.L.vt_group_create(C.cast(arr, _vp), len(devices))
        else:
            uid = None if unique_id is None else np.ascontiguousarray(unique_id, np.uint8)
            self.h = self.L.vt_group_create_rank(int(device), int(rank), int(world), _ptr(uid))
        if not self.h:
            raise RuntimeError(f"vt_group_create: {_err()}")
        self.scene = None

    def close(self):
        if getattr(self, "h", None):
            self.L.vt_group_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    @property
    def world(self):
        return int(self.L.vt_group_size(self.h))

    @property
    def rank(self):
        return int(self.L.vt_group_rank(self.h))

    @property
    def local_members(self):
        return int(self.L.vt_group_local_members(self.h))

    @property
    def launch_count(self):
        return int(self.L.vt_group_launch_count(self.h))

    def accel(self, local_member=0):
        h = self.L.vt_group_accel(self.h, local_member)
        if not h:
            raise RuntimeError("vt_group_accel: no such member")
        return Accel.borrowed(h, self.scene)

    def populate(self, scene):
        """Build once, replicate the device image (multi-process groups: only rank 0 passes the scene; others pass None)."""
        self.scene = scene
        _check(self.L.vt_group_populate(self.h, None if scene is None else C.cast(scene.ptr(), _vp)), "vt_group_populate")
        return self

    def shard(self, n, rank=None):
        """(tile, local_count) of `rank`'s shard of an n-pixel frame."""
        tile, cnt = C.c_uint64(0), C.c_uint64(0)
        _check(self.L.vt_group_shard(self.h, n, self.rank if rank is None else rank, C.addressof(tile), C.addressof(cnt)), "vt_group_shard")
        return tile.value, cnt.value

    def shard_indices(self, n, rank=None):
        """Global pixel indices of `rank`'s shard, in its compact (tile) order."""
        r = self.rank if rank is None else rank
        tile, cnt = self.shard(n, r)
        tiles = np.arange(r, (n + tile - 1) // tile, self.world, dtype=np.int64)
        idx = (tiles[:, None] * tile + np.arange(tile, dtype=np.int64)[None, :]).reshape(-1)
        idx = idx[idx < n]
        assert len(idx) == cnt
        return idx

    def traverse(self, rays, want_attrs=False, any_hit=False, out=None):
        rays = rays if isinstance(rays, np.ndarray) and rays.dtype == abi.RAY and rays.flags.c_contiguous else np.ascontiguousarray(rays, abi.RAY)
        out = dict(out or {})
        hits = out.get("hits") if "hits" in out else np.zeros(len(rays), abi.HIT)
        attrs = (out.get("attrs") if "attrs" in out else np.zeros(len(rays), abi.ATTR)) if want_attrs else None
        _check(self.L.vt_group_traverse(self.h, rays.ctypes.data, len(rays), hits.ctypes.data, _ptr(attrs), abi.VT_TRAVERSE_ANY_HIT if any_hit else 0),
               "vt_group_traverse")
        return (hits, attrs) if want_attrs else hits

    def render_diffuse_wave(self, rays, spp, seed=0, weight=1.0, out=None, want_live=True, shared_frame=False):
        """Host rays (frame-sized array) in, host RGBFFF framebuffer out; returns (framebuffer, bounce rays spawned by this process).
        shared_frame: `out` is the same pinned host memory in every process of the group (VT_GROUP_SHARED_HOST_FRAME)."""
        rays = rays if isinstance(rays, np.ndarray) and rays.dtype == abi.RAY and rays.flags.c_contiguous else np.ascontiguousarray(rays, abi.RAY)
        fb = np.zeros((len(rays), 3), np.float32) if out is None else out
        live = C.c_uint64(0)
        _check(self.L.vt_group_render_diffuse_wave(self.h, rays.ctypes.data, len(rays), spp, seed, weight, fb.ctypes.data,
                                                   C.addressof(live) if want_live else None, abi.VT_GROUP_SHARED_HOST_FRAME if shared_frame else 0, None),
               "vt_group_render_diffuse_wave")
        return fb, live.value

    def render_diffuse_wave_begin(self, rays, spp, seed, weight, shared_out):
        """Enqueue one frame into the host frame all processes share (VT_GROUP_SHARED_HOST_FRAME | VT_GROUP_ASYNC); at most two in flight."""
        _check(self.L.vt_group_render_diffuse_wave(self.h, rays.ctypes.data, len(rays), spp, seed, weight, shared_out.ctypes.data, None,
                                                   abi.VT_GROUP_SHARED_HOST_FRAME | abi.VT_GROUP_ASYNC, None), "vt_group_render_diffuse_wave")

    def wait_frame(self):
        _check(self.L.vt_group_wait_frame(self.h), "vt_group_wait_frame")

    def render_diffuse_wave_device(self, d_rays_shard, n, spp, seed, weight, d_fb, stream=None, slot=0):
        """Device-resident shard in, frame-sized device image out (complete on rank 0); enqueued on `stream`.  slot 0 / 1: the group's
        two independent sets of per-frame state (VT_GROUP_FRAME_SLOT1) — alternate them over two streams for two frames in flight."""
        _check(self.L.vt_group_render_diffuse_wave(self.h, _ptr(d_rays_shard), n, spp, seed, weight, _ptr(d_fb), None,
                                                   abi.VT_TRAVERSE_DEVICE_PTRS | (abi.VT_GROUP_FRAME_SLOT1 if slot else 0), _ptr(stream)),
               "vt_group_render_diffuse_wave")

    def all_gather_device(self, d_buf, bytes_per_rank, stream=None):
        _check(self.L.vt_group_all_gather_device(self.h, _ptr(d_buf), bytes_per_rank, _ptr(stream)), "vt_group_all_gather_device")

    def reduce_device(self, d_buf, count, stream=None):
        _check(self.L.vt_group_reduce_device(self.h, _ptr(d_buf), count, _ptr(stream)), "vt_group_reduce_device")
