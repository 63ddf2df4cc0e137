"""CPU checkers for the accel:Traverse path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Two interchangeable back ends behind one Python class:

* ``kind="port"``      oracle/libvt_oracle.so — our plain-C restatement (oracle/vt_oracle.c),
                       every function citing the reference file:line it follows.
* ``kind="reference"`` oracle/_ref/libvt_ref.so — the UNMODIFIED reference compiled from
                       /root/reference by oracle/Makefile (ref_harness.cpp drives it headless).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference``
legs may import this package.  vistrace_b200 (the product) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from vistrace_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libvt_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libvt_ref.so")
REFERENCE_ROOT = "/root/reference"


def build(ref=True, quiet=True):
    """Compile the C restatement and, when /root/reference is present, the real reference."""
    out = subprocess.DEVNULL if quiet else None
    subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=out)
    if ref and os.path.isdir(REFERENCE_ROOT):
        subprocess.check_call(["make", "-C", HERE, "-j8", "ref"], stdout=out)


def available(kind):
    return os.path.exists(PORT_SO if kind == "port" else REF_SO)


_libs = {}


def _lib(kind):
    if kind in _libs:
        return _libs[kind]
    path, pre = (PORT_SO, "vto_") if kind == "port" else (REF_SO, "vtref_")
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: run `make -C oracle {'oracle' if kind == 'port' else 'ref'}`")
    lib = C.CDLL(path)
    vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
    sig = {
        "max_threads": (i32, []),
        "create": (vp, [vp, i32]),
        "destroy": (None, [vp]),
        "get_tri_derived": (None, [vp, vp]),
        "get_bvh": (None, [vp, vp, vp, vp, vp]),
        "set_bvh": (None, [vp, vp, u64, vp]),
        "refit": (None, [vp, vp]),
        "traverse": (C.c_double, [vp, vp, u64, vp, vp, i32, vp]),
        "trace_result": (None, [vp, vp, vp, u64, vp, i32]),
        "trace_result_cones": (None, [vp, vp, vp, vp, u64, vp, i32]),
        "sample": (None, [vp, i32, vp, u64, vp]),
        "node_intersect": (None, [vp, vp, vp]),
        "tri_intersect": (i32, [vp, u64, vp, vp]),
        "skin_triangles": (None, [vp, vp, u64, vp, vp, C.c_uint32, vp]),
        "sample_bsdf_diffuse": (None, [vp, vp, vp, u64, vp, vp]),
    }
    if kind != "port":  # reference only: its VTF parser is not restated in the C port (the product's decoder is checked against it)
        sig["vtf_pixels"] = (C.c_int64, [vp, u64, C.c_uint32, C.c_uint32, vp, u64])
        sig["reinsertion_optimize"] = (None, [vp, vp])
        sig["mdl_open"] = (vp, [vp, u64, vp, u64, vp, u64])
        sig["mdl_close"] = (None, [vp])
        sig["mdl_info"] = (None, [vp, vp])
        sig["mdl_bodygroup_values"] = (C.c_int32, [vp, C.c_uint32])
        sig["mdl_mesh"] = (C.c_int64, [vp, C.c_uint32, C.c_uint32, vp, vp, u64])
        sig["mdl_bind_matrices"] = (None, [vp, vp])
        sig["mdl_material_index"] = (C.c_int32, [vp, C.c_int32, C.c_int32])
        sig["mdl_material_path"] = (C.c_int32, [vp, C.c_int32, C.c_int32, vp, u64])
        sig["bsp_open"] = (vp, [vp, u64])
        sig["bsp_close"] = (None, [vp])
        sig["bsp_info"] = (None, [vp, vp])
        sig["bsp_triangles"] = (None, [vp, vp, vp, vp])
        sig["bsp_material"] = (C.c_int32, [vp, C.c_uint32, vp])
        sig["bsp_static_prop"] = (C.c_int32, [vp, C.c_int32, vp])
    ns = type("ns", (), {})()
    for name, (res, args) in sig.items():
        fn = getattr(lib, pre + name)
        fn.restype, fn.argtypes = res, args
        setattr(ns, name, fn)
    _libs[kind] = ns
    return ns


def _p(a):
    return None if a is None else a.ctypes.data


class CpuScene:
    """One scene held by a CPU checker (the reference's AccelStruct or the C port of it)."""

    def __init__(self, scene, kind="reference", build_bvh=True):
        self.kind = kind
        self.lib = _lib(kind)
        self.scene = scene  # keep the numpy buffers alive
        self.h = self.lib.create(C.cast(scene.ptr(), C.c_void_p), 1 if build_bvh else 0)
        if not self.h:
            raise RuntimeError("oracle scene creation failed")
        self.n_tris = scene.n_tris

    def close(self):
        if self.h:
            self.lib.destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    @property
    def max_threads(self):
        return self.lib.max_threads()

    def tri_derived(self):
        """(n, 16) float32: p0, e1, e2, n, nNorm, lod as the Triangle constructor derives them."""
        out = np.zeros((self.n_tris, 16), np.float32)
        self.lib.get_tri_derived(self.h, out.ctypes.data)
        return out

    def get_bvh(self):
        cnt, nt = C.c_uint64(0), C.c_uint64(0)
        self.lib.get_bvh(self.h, None, C.addressof(cnt), None, C.addressof(nt))
        nodes = np.zeros(cnt.value, abi.NODE)
        prims = np.zeros(nt.value, np.uint64)
        self.lib.get_bvh(self.h, nodes.ctypes.data, None, prims.ctypes.data, None)
        return nodes, prims

    def set_bvh(self, nodes, prim_indices):
        nodes = np.ascontiguousarray(nodes, abi.NODE)
        prim_indices = np.ascontiguousarray(prim_indices, np.uint64)
        assert len(prim_indices) == self.n_tris
        self.lib.set_bvh(self.h, nodes.ctypes.data, len(nodes), prim_indices.ctypes.data)

    def refit(self, scene):
        """Same topology, moved vertices: the checker's own bottom-up refit of its current hierarchy
        (reference kind: bvh::HierarchyRefitter itself)."""
        assert scene.n_tris == self.n_tris
        self.scene = scene
        self.lib.refit(self.h, C.cast(scene.ptr(), C.c_void_p))

    def reinsertion_optimize(self):
        """Reference kind only: bvh::ParallelReinsertionOptimizer over the current hierarchy -> (SAH cost before, after)."""
        c = np.zeros(2, np.float64)
        self.lib.reinsertion_optimize(self.h, c.ctypes.data)
        return float(c[0]), float(c[1])

    def traverse(self, rays, want_attrs=False, threads=0, want_stats=False):
        """Returns dict(hits, attrs?, seconds, steps?, isects?)."""
        rays = np.ascontiguousarray(rays, abi.RAY)
        hits = np.zeros(len(rays), abi.HIT)
        attrs = np.zeros(len(rays), abi.ATTR) if want_attrs else None
        stats = np.zeros(2, np.uint64) if want_stats else None
        sec = self.lib.traverse(self.h, rays.ctypes.data, len(rays), hits.ctypes.data, _p(attrs), threads, _p(stats))
        if sec < 0:
            raise RuntimeError("oracle traverse: acceleration structure not built")
        out = {"hits": hits, "seconds": sec}
        if want_attrs:
            out["attrs"] = attrs
        if want_stats:
            out["steps"], out["isects"] = int(stats[0]), int(stats[1])
        return out

    def time_traverse(self, rays, threads=0, reps=3):
        """Best-of-reps seconds for the traversal-only loop."""
        rays = np.ascontiguousarray(rays, abi.RAY)
        hits = np.zeros(len(rays), abi.HIT)
        best = float("inf")
        for _ in range(reps):
            best = min(best, self.lib.traverse(self.h, rays.ctypes.data, len(rays), hits.ctypes.data, None, threads, None))
        return best

    def trace_result(self, rays, hits, threads=0, cones=None):
        """TraceResult records for given hits; cones = [n, 2] {coneWidth, coneAngle} or None (mip 0)."""
        rays = np.ascontiguousarray(rays, abi.RAY)
        hits = np.ascontiguousarray(hits, abi.HIT)
        attrs = np.zeros(len(rays), abi.ATTR)
        if cones is not None:
            cones = np.ascontiguousarray(cones, np.float32).reshape(len(rays), 2)
            self.lib.trace_result_cones(self.h, rays.ctypes.data, hits.ctypes.data, cones.ctypes.data, len(rays), attrs.ctypes.data, threads)
            return attrs
        self.lib.trace_result(self.h, rays.ctypes.data, hits.ctypes.data, len(rays), attrs.ctypes.data, threads)
        return attrs

    def sample(self, tex, uvm):
        uvm = np.ascontiguousarray(uvm, np.float32).reshape(-1, 3)
        out = np.zeros((len(uvm), 4), np.float32)
        self.lib.sample(self.h, tex, uvm.ctypes.data, len(uvm), out.ctypes.data)
        return out

    def tri_intersect(self, prim, ray):
        ray = np.ascontiguousarray(ray, abi.RAY).reshape(1)
        tuv = np.zeros(3, np.float32)
        ok = self.lib.tri_intersect(self.h, prim, ray.ctypes.data, tuv.ctypes.data)
        return (bool(ok), tuv)


def node_intersect(node, ray, kind="reference"):
    """FastNodeIntersector::intersect on one node -> (entry, exit)."""
    node = np.ascontiguousarray(node, abi.NODE).reshape(1)
    ray = np.ascontiguousarray(ray, abi.RAY).reshape(1)
    out = np.zeros(2, np.float32)
    _lib(kind).node_intersect(node.ctypes.data, ray.ctypes.data, out.ctypes.data)
    return float(out[0]), float(out[1])


def skin_triangles(tris, skin, bones, binds, kind="reference"):
    """SkinTriangle (source/objects/AccelStruct.cpp:66-108) over vt_tri_in records -> float32 [n, 27] =
    {p0, e1, e2, normals[3], tangents[3]} as the reference leaves them.  skin=None: the one-bone overload."""
    tris = np.ascontiguousarray(tris, abi.TRI_IN)
    bones = np.ascontiguousarray(bones, np.float32).reshape(-1, 16)
    binds = np.ascontiguousarray(binds, np.float32).reshape(-1, 16)
    out = np.zeros((len(tris), 27), np.float32)
    _lib(kind).skin_triangles(tris.ctypes.data, None if skin is None else np.ascontiguousarray(skin, abi.TRI_SKIN).ctypes.data, len(tris),
                              bones.ctypes.data, binds.ctypes.data, len(bones), out.ctypes.data)
    return out


def vtf_pixels(data, n_texels, frame=0, face=0):
    """Reference VTFTexture over the VTF file bytes `data`: (n_texels, 4) float32 of GetPixel in storage order
    (smallest mip first), or None when the reference parser rejects the file."""
    buf = np.frombuffer(bytes(data), np.uint8)
    out = np.zeros((n_texels, 4), np.float32)
    n = _lib("reference").vtf_pixels(buf.ctypes.data, len(buf), frame, face, out.ctypes.data, n_texels)
    if n == -1:
        return None
    if n != n_texels:
        raise RuntimeError(f"vtf_pixels: reference reports {n} texels, expected {n_texels}")
    return out


def sample_bsdf_diffuse(attrs, wo, rnd, kind="reference"):
    """SampleBSDF (source/libraries/BSDF.cpp:770-825) with only the diffuse lobe active, per TraceResult record: BSDFMaterial
    prepared by PrepShadingData(albedo, metalness, roughness); wo = incident direction [n, 3]; rnd = [n, 3] the numbers the
    ISampler hands out (lobeSelect, r1, r2).  Returns (abi.BSDF_SAMPLE records, returned flags)."""
    attrs = np.ascontiguousarray(attrs, abi.ATTR)
    wo = np.ascontiguousarray(wo, np.float32).reshape(len(attrs), 3)
    rnd = np.ascontiguousarray(rnd, np.float32).reshape(len(attrs), 3)
    out = np.zeros(len(attrs), abi.BSDF_SAMPLE)
    ret = np.zeros(len(attrs), np.int32)
    _lib(kind).sample_bsdf_diffuse(attrs.ctypes.data, wo.ctypes.data, rnd.ctypes.data, len(attrs), out.ctypes.data, ret.ctypes.data)
    return out, ret


class RefModel:
    """The reference's own MDL / VVD / VTX parsers + BodyGroup / Mesh (libs/MDLParser, source/objects/Model.cpp) over file bytes."""

    def __init__(self, mdl, vvd, vtx):
        self.lib = _lib("reference")
        self._bufs = [np.frombuffer(bytes(b), np.uint8).copy() for b in (mdl, vvd, vtx)]
        args = [v for b in self._bufs for v in (b.ctypes.data if len(b) else None, len(b))]
        self.h = self.lib.mdl_open(*args)
        info = np.zeros(6, np.int32)
        self.lib.mdl_info(self.h, info.ctypes.data)
        self.valid = bool(info[0])
        self.n_bodygroups, self.n_bones, self.n_materials, self.n_skin_families, self.n_vertices = (int(v) for v in info[1:])

    def close(self):
        if self.h:
            self.lib.mdl_close(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def bodygroup_values(self, bodygroup):
        return int(self.lib.mdl_bodygroup_values(self.h, bodygroup))

    def mesh(self, bodygroup, value):
        """(records with p = {p0, e1, e2} as the reference's Triangle stores them, skin records) or None."""
        n = int(self.lib.mdl_mesh(self.h, bodygroup, value, None, None, 0))
        if n < 0:
            return None
        tris, skin = np.zeros(n, abi.TRI_IN), np.zeros(n, abi.TRI_SKIN)
        if n:
            self.lib.mdl_mesh(self.h, bodygroup, value, tris.ctypes.data, skin.ctypes.data, n)
        return tris, skin

    def bind_matrices(self):
        out = np.zeros((self.n_bones, 16), np.float32)
        if self.n_bones:
            self.lib.mdl_bind_matrices(self.h, out.ctypes.data)
        return out

    def material_index(self, skin, material_id):
        return int(self.lib.mdl_material_index(self.h, skin, material_id))

    def material_path(self, material_id, directory=0):
        buf = C.create_string_buffer(4200)
        self.lib.mdl_material_path(self.h, material_id, directory, buf, len(buf))
        return buf.value.decode("latin-1")


class RefBsp:
    """The reference's own BSPMap (libs/BSPParser: parse, Triangulate, displacement smoothing) over file bytes, plus the material
    bookkeeping of World::World (source/objects/AccelStruct.cpp:236-414) — the checker of vt_bsp_*."""

    def __init__(self, data):
        self.lib = _lib("reference")
        self._buf = np.frombuffer(bytes(data), np.uint8).copy()
        self.h = self.lib.bsp_open(self._buf.ctypes.data if len(self._buf) else None, len(self._buf))
        info = np.zeros(5, np.int64)
        self.lib.bsp_info(self.h, info.ctypes.data)
        self.valid, self.textures_ok = bool(info[0]), bool(info[1])
        self.n_tris, self.n_materials, self.n_static_props = (int(v) for v in info[2:])

    def close(self):
        if self.h:
            self.lib.bsp_close(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def triangles(self):
        """(vt_tri_in records, binormals [n, 3, 3], texinfo indices [n])."""
        n = self.n_tris if self.textures_ok else 0
        tris, bino, texinfo = np.zeros(n, abi.TRI_IN), np.zeros((n, 3, 3), np.float32), np.zeros(n, np.int16)
        if n:
            self.lib.bsp_triangles(self.h, tris.ctypes.data, bino.ctypes.data, texinfo.ctypes.data)
        return tris, bino, texinfo

    def material(self, index):
        out = np.zeros(1, abi.BSP_MATERIAL)
        return out[0] if self.lib.bsp_material(self.h, index, out.ctypes.data) == 0 else None

    def static_prop(self, index):
        out = np.zeros(1, abi.BSP_STATIC_PROP)
        return out[0] if self.lib.bsp_static_prop(self.h, index, out.ctypes.data) == 0 else None
