"""POD records of include/vistrace_b200.h as numpy dtypes / ctypes structures.

Plumbing only: lets tests and bench.py hand numpy buffers to the C ABI.  The
layouts are asserted against the header's sizes (vt_ray 32 B, vt_hit 16 B,
vt_node 32 B, vt_tri_in 152 B, vt_attr 128 B).
"""
import ctypes as C

import numpy as np

VT_MISS = 0xFFFFFFFF
VT_TEXFLAG_CLAMPS = 0x4
VT_TEXFLAG_CLAMPT = 0x8
VT_MATFLAG_ALPHATEST = 256
VT_MATFLAG_NOCULL = 8192
VT_SURF_SKY = 0x4
VT_ATTR_FRONT_FACING = 1
VT_ATTR_HIT_SKY = 2
VT_ATTR_HIT_WATER = 4
VT_TRAVERSE_DEVICE_PTRS = 1
VT_TRAVERSE_ANY_HIT = 2
VT_TRAVERSE_QUEUE_ATTRS = 4
VT_GROUP_SHARED_HOST_FRAME = 16
VT_GROUP_FRAME_SLOT1 = 32
VT_GROUP_ASYNC = 64
VT_TEXEL_WIDE, VT_TEXEL_DIV_255, VT_TEXEL_DIV_65535, VT_TEXEL_DIV_1 = 0x100, 0, 1, 2
VT_PATHS_NO_COMPACTION = 8
VT_PATHS_SLOT1 = 16
VT_LOBE_NONE, VT_LOBE_DIFFUSE_REFLECTION = 0, 1

f4, u4, i4, u2, u1 = np.float32, np.uint32, np.int32, np.uint16, np.uint8

RAY = np.dtype([("o", f4, 3), ("tmin", f4), ("d", f4, 3), ("tmax", f4)], align=False)
HIT = np.dtype([("t", f4), ("u", f4), ("v", f4), ("prim", u4)], align=False)
BSDF_SAMPLE = np.dtype([("scattered", f4, 3), ("pdf", f4), ("weight", f4, 3), ("lobe", u4)], align=False)  # vt_bsdf_sample = BSDFSample
NODE = np.dtype([("bounds", f4, 6), ("prim_count", u4), ("first", u4)], align=False)
TRI_IN = np.dtype(
    [
        ("p", f4, (3, 3)),
        ("normals", f4, (3, 3)),
        ("tangents", f4, (3, 3)),
        ("uvs", f4, (3, 2)),
        ("alphas", f4, 3),
        ("material", u4),
        ("ent_idx", u2),
        ("one_sided", u1),
        ("pad", u1),
    ],
    align=False,
)
TRI_SKIN = np.dtype([("num_bones", u1, 3), ("pad", u1), ("bone_ids", np.int8, (3, 3)), ("pad2", u1, 3), ("weights", f4, (3, 3))], align=False)
# vt_bsp_info / vt_bsp_material / vt_bsp_static_prop (include/vistrace_b200.h)
BSP_INFO = np.dtype([("version", np.uint32), ("n_materials", np.uint32), ("n_texinfos", np.uint32), ("n_displacements", np.uint32), ("n_static_props", np.uint32),
                     ("static_props_version", np.uint32), ("n_tris", np.uint64)])
BSP_MATERIAL = np.dtype([("surf_flags", np.uint32), ("texinfo", np.int32), ("width", np.int32), ("height", np.int32), ("reflectivity", f4, 3), ("path", "S260")])
BSP_STATIC_PROP = np.dtype([("pos", f4, 3), ("ang", f4, 3), ("skin", np.int32), ("model", "S128")])
ATTR = np.dtype(
    [
        ("pos", f4, 3),
        ("distance", f4),
        ("normal", f4, 3),
        ("alpha", f4),
        ("tangent", f4, 3),
        ("metalness", f4),
        ("binormal", f4, 3),
        ("roughness", f4),
        ("geometric_normal", f4, 3),
        ("base_mip", f4),
        ("albedo", f4, 3),
        ("ent_id", u4),
        ("uvw", f4, 3),
        ("submat_idx", u4),
        ("tex_uv", f4, 2),
        ("flags", u4),
        ("prim", u4),
    ],
    align=False,
)
_IDENT = (1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0)  # glm::identity<mat2x4>(): col0=(1,0,0,0), col1=(0,1,0,0)
MATERIAL = np.dtype(
    [
        ("flags", u4),
        ("surf_flags", u4),
        ("alphatest_reference", f4),
        ("tex_scale", f4),
        ("colour", f4, 4),
        ("base_tex_mat", f4, 8),
        ("base_tex_mat2", f4, 8),
        ("normal_map_mat", f4, 8),
        ("normal_map_mat2", f4, 8),
        ("blend_tex_mat", f4, 8),
        ("detail_mat", f4, 8),
        ("detail_scale", f4),
        ("detail_blend_factor", f4),
        ("detail_tint", f4, 3),
        ("base_texture", i4),
        ("base_texture2", i4),
        ("normal_map", i4),
        ("normal_map2", i4),
        ("mrao", i4),
        ("mrao2", i4),
        ("blend_texture", i4),
        ("detail", i4),
        ("detail_blend_mode", u1),
        ("masked_blending", u1),
        ("detail_alpha_mask_base_texture", u1),
        ("water", u1),
    ],
    align=False,
)
ENTITY = np.dtype([("id", u4), ("colour", f4, 4)], align=False)

assert RAY.itemsize == 32 and HIT.itemsize == 16 and NODE.itemsize == 32
assert TRI_SKIN.itemsize == 52
assert TRI_IN.itemsize == 152 and ATTR.itemsize == 128 and ENTITY.itemsize == 20
assert MATERIAL.itemsize == 280


def default_materials(n):
    """n materials with the defaults of source/objects/Material.h:74-125."""
    m = np.zeros(n, MATERIAL)
    m["alphatest_reference"] = 0.5
    m["tex_scale"] = 1.0
    m["colour"] = 1.0
    for k in ("base_tex_mat", "base_tex_mat2", "normal_map_mat", "normal_map_mat2", "blend_tex_mat", "detail_mat"):
        m[k] = _IDENT
    m["detail_scale"] = 4.0
    m["detail_blend_factor"] = 1.0
    m["detail_tint"] = 1.0
    for k in ("base_texture", "base_texture2", "normal_map", "normal_map2", "mrao", "mrao2", "blend_texture", "detail"):
        m[k] = -1
    return m


class Texture(C.Structure):
    _fields_ = [
        ("width", C.c_uint16),
        ("height", C.c_uint16),
        ("mip_count", C.c_uint16),
        ("pad", C.c_uint16),
        ("flags", C.c_uint32),
        ("texel_layout", C.c_uint32),
        ("rgba", C.c_void_p),
        ("nbytes", C.c_uint64),
    ]


class Scene(C.Structure):
    _fields_ = [
        ("tris", C.c_void_p),
        ("n_tris", C.c_uint64),
        ("materials", C.c_void_p),
        ("n_materials", C.c_uint32),
        ("entities", C.c_void_p),
        ("n_entities", C.c_uint32),
        ("textures", C.c_void_p),
        ("n_textures", C.c_uint32),
    ]


assert C.sizeof(Texture) == 32 and C.sizeof(Scene) == 64


class SceneData:
    """Owns the numpy buffers of one scene and exposes them as a vt_scene."""

    def __init__(self, tris, materials=None, entities=None, textures=()):
        self.tris = np.ascontiguousarray(tris, TRI_IN)
        self.materials = default_materials(1) if materials is None else np.ascontiguousarray(materials, MATERIAL)
        if entities is None:
            entities = np.zeros(1, ENTITY)
            entities["colour"] = 1.0
        self.entities = np.ascontiguousarray(entities, ENTITY)
        # textures: list of (width, height, mip_count, flags, uint8 array in VTF order: smallest mip first[, texel_layout]) — what
        # vtf_decode returns; texel_layout 0 (default) = RGBA8888, else wide texels (include/vistrace_b200.h: VT_TEXEL_WIDE)
        self.textures = [(int(t[0]), int(t[1]), int(t[2]), int(t[3]), np.ascontiguousarray(t[4], np.uint8).ravel(), int(t[5]) if len(t) > 5 else 0) for t in textures]
        self._tex_arr = (Texture * max(1, len(self.textures)))()
        for i, (w, h, m, fl, px, layout) in enumerate(self.textures):
            self._tex_arr[i] = Texture(w, h, m, 0, fl, layout, px.ctypes.data, px.nbytes)
        self.c = Scene(
            self.tris.ctypes.data,
            len(self.tris),
            self.materials.ctypes.data,
            len(self.materials),
            self.entities.ctypes.data,
            len(self.entities),
            C.cast(self._tex_arr, C.c_void_p).value if self.textures else None,
            len(self.textures),
        )

    @property
    def n_tris(self):
        return len(self.tris)

    def ptr(self):
        return C.byref(self.c)
