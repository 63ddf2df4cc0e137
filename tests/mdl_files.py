"""Synthetic Source-engine model files (studiomdl v48 .mdl, VVD v4, VTX v7) for the ingestion tests: written from the published
file layouts (the same layouts libs/MDLParser/source/Structs.h declares), small enough to read in a hex dump, and built to reach
every branch of the loader: several body groups with several values, several meshes per model, two strip groups, a triangle-STRIP
strip that must be ignored, vertices with 0 .. 3 bones, a NaN tangent, a zero normal, a VVD with and without a fix-up table."""
import struct

import numpy as np


class _Buf:
    def __init__(self):
        self.b = bytearray()

    def tell(self):
        return len(self.b)

    def put(self, data):
        pos = len(self.b)
        self.b += data
        return pos

    def zeros(self, n):
        return self.put(bytes(n))

    def i32(self, pos, v):
        struct.pack_into("<i", self.b, pos, int(v))

    def cstr(self, s):
        return self.put(s.encode("latin-1") + b"\0")


def make_model(seed=0, body=((2, 1), (3,)), n_bones=3, n_materials=3, n_skins=2, fixups=False, version=48, checksum=0x1234ABCD):
    """body: per body group, per value, the number of meshes of that model.  Returns dict(mdl, vvd, vtx bytes + what was put in)."""
    rng = np.random.default_rng(seed)
    f4 = np.float32
    # ---- geometry: per model a vertex pool; per mesh a slice of it; per mesh two strip groups
    models = []  # (bodygroup, value, meshes=[dict(material, vert_off, n_verts, groups=[dict(verts=[(orig, nbones)], indices, strips=[(flags, off, n)])])])
    all_verts = []  # rows of the root-LoD VVD in logical order: (weights3, bones3, numbones, pos3, normal3, uv2), tangent4
    for bg, values in enumerate(body):
        for val, n_meshes in enumerate(values):
            first_vertex = len(all_verts)
            meshes = []
            off = 0
            for mi in range(n_meshes):
                nv = int(rng.integers(6, 14))
                for _ in range(nv):
                    w = rng.uniform(0.1, 1.0, 3).astype(f4)
                    w = (w / w.sum()).astype(f4)
                    bones = rng.integers(0, n_bones, 3).astype(np.int8)
                    pos = rng.uniform(-20, 20, 3).astype(f4)
                    nrm = rng.normal(size=3).astype(f4)
                    uv = rng.uniform(0, 4, 2).astype(f4)
                    tan = np.append(rng.normal(size=3), rng.choice([-1.0, 1.0])).astype(f4)
                    all_verts.append([w, bones, int(rng.integers(1, 4)), pos, nrm, uv, tan])
                groups = []
                for gi in range(2):
                    nsv = int(rng.integers(4, 9))
                    sverts = [(int(rng.integers(0, nv)), int(rng.integers(0, 4))) for _ in range(nsv)]  # (origMeshVertId, numBones incl. 0)
                    n_tri_a, n_tri_b = int(rng.integers(1, 5)), int(rng.integers(1, 4))
                    idx = rng.integers(0, nsv, 3 * (n_tri_a + n_tri_b) + 4).astype(np.uint16)
                    strips = [(0x01, 0, 3 * n_tri_a), (0x02, 3 * n_tri_a, 4), (0x01, 3 * n_tri_a + 4, 3 * n_tri_b)]  # list, STRIP (ignored), list
                    groups.append({"verts": sverts, "indices": idx, "strips": strips})
                meshes.append({"material": int(rng.integers(0, n_materials)), "vert_off": off, "n_verts": nv, "groups": groups})
                off += nv
            models.append({"bg": bg, "val": val, "first_vertex": first_vertex, "n_verts": off, "meshes": meshes})
    # special vertices: a NaN tangent (replaced by normalize(e1)) and a zero normal (normalize gives NaN, kept)
    all_verts[1][6][:3] = np.nan
    all_verts[2][4][:] = 0.0
    g0 = models[0]["meshes"][0]["groups"][0]  # make sure the first triangle of the first strip uses them
    g0["verts"][0], g0["verts"][1] = (1, 2), (2, 0)
    g0["indices"][:3] = (0, 1, 2)
    n_verts = len(all_verts)

    # ---- VVD
    def vvd_row(v):
        w, bones, nb, pos, nrm, uv, _ = v
        return struct.pack("<3f3bB3f3f2f", *w, *bones, nb, *pos, *nrm, *uv)

    rows = [vvd_row(v) for v in all_verts]
    tans = [struct.pack("<4f", *v[6]) for v in all_verts]
    if fixups:
        # the file stores three segments in another order (plus a LoD-1-only junk segment in front); the fix-ups put them back
        cuts = sorted(rng.choice(np.arange(1, n_verts), 2, replace=False))
        segs = [(0, cuts[0]), (cuts[0], cuts[1]), (cuts[1], n_verts)]
        order = [2, 0, 1]
        junk = 5
        file_rows, file_tans, where = [bytes(48)] * junk, [bytes(16)] * junk, {}
        for s in order:
            where[s] = len(file_rows)
            file_rows += rows[segs[s][0]:segs[s][1]]
            file_tans += tans[segs[s][0]:segs[s][1]]
        fix = [(-1, 0, junk)] + [(0, where[s], segs[s][1] - segs[s][0]) for s in range(3)]  # lod -1 < root LoD: skipped
    else:
        file_rows, file_tans, fix = rows, tans, []
    vvd = _Buf()
    vvd.zeros(64)
    fix_off = vvd.tell()
    for lod, src, cnt in fix:
        vvd.put(struct.pack("<3i", lod, src, cnt))
    vert_off = vvd.put(b"".join(file_rows))
    tan_off = vvd.put(b"".join(file_tans))
    struct.pack_into("<4i8i4i", vvd.b, 0, ord("I") + (ord("D") << 8) + (ord("S") << 16) + (ord("V") << 24), 4, _s32(checksum), 1,
                     *([n_verts] + [0] * 7), len(fix), fix_off, vert_off, tan_off)

    # ---- MDL
    mdl = _Buf()
    mdl.zeros(408)
    bone_off = mdl.tell()
    pose = rng.normal(size=(n_bones, 3, 4)).astype(f4)
    for b in range(n_bones):
        rec = bytearray(216)
        struct.pack_into("<12f", rec, 96, *pose[b].reshape(-1))
        mdl.put(rec)
    tex_off = mdl.tell()
    mdl.zeros(64 * n_materials)
    names = [f"mat_{seed}_{i}" for i in range(n_materials)]
    for i, nm in enumerate(names):
        pos = mdl.cstr(nm)
        mdl.i32(tex_off + 64 * i, pos - (tex_off + 64 * i))
    dirs = ["models/props/", "models/shared\\"]
    dir_tbl = mdl.zeros(4 * len(dirs))
    for i, d in enumerate(dirs):
        mdl.i32(dir_tbl + 4 * i, mdl.cstr(d))
    skin_tbl = rng.integers(0, n_materials, (n_skins, n_materials)).astype(np.int16)
    skin_off = mdl.put(skin_tbl.tobytes())
    bp_off = mdl.zeros(16 * len(body))
    for bg, values in enumerate(body):
        bp = bp_off + 16 * bg
        model_pos = mdl.zeros(148 * len(values))
        mdl.i32(bp + 0, mdl.cstr(f"bodygroup{bg}") - bp)
        mdl.i32(bp + 4, len(values))
        mdl.i32(bp + 8, 1)
        mdl.i32(bp + 12, model_pos - bp)
        for val in range(len(values)):
            m = next(x for x in models if x["bg"] == bg and x["val"] == val)
            mp = model_pos + 148 * val
            mesh_pos = mdl.zeros(116 * len(m["meshes"]))
            mdl.i32(mp + 72, len(m["meshes"]))
            mdl.i32(mp + 76, mesh_pos - mp)
            mdl.i32(mp + 80, m["n_verts"])
            mdl.i32(mp + 84, m["first_vertex"] * 48)  # byte offsets into the vvd's vertex / tangent arrays
            mdl.i32(mp + 88, m["first_vertex"] * 16)
            for mi, mesh in enumerate(m["meshes"]):
                q = mesh_pos + 116 * mi
                mdl.i32(q + 0, mesh["material"])
                mdl.i32(q + 4, mp - q)
                mdl.i32(q + 8, mesh["n_verts"])
                mdl.i32(q + 12, mesh["vert_off"])
    struct.pack_into("<3i", mdl.b, 0, ord("I") + (ord("D") << 8) + (ord("S") << 16) + (ord("T") << 24), version, _s32(checksum))
    for off, v in ((156, n_bones), (160, bone_off), (204, n_materials), (208, tex_off), (212, len(dirs)), (216, dir_tbl), (220, n_materials), (224, n_skins),
                   (228, skin_off), (232, len(body)), (236, bp_off), (76, 0)):
        mdl.i32(off, v)
    mdl.i32(76, len(mdl.b))  # dataLength

    # ---- VTX
    vtx = _Buf()
    vtx.zeros(36)
    vbp_off = vtx.zeros(8 * len(body))
    for bg, values in enumerate(body):
        vbp = vbp_off + 8 * bg
        vmodels = vtx.zeros(8 * len(values))
        vtx.i32(vbp, len(values))
        vtx.i32(vbp + 4, vmodels - vbp)
        for val in range(len(values)):
            m = next(x for x in models if x["bg"] == bg and x["val"] == val)
            vm = vmodels + 8 * val
            lod = vtx.zeros(12)
            vtx.i32(vm, 1)
            vtx.i32(vm + 4, lod - vm)
            vmeshes = vtx.zeros(9 * len(m["meshes"]))
            vtx.i32(lod, len(m["meshes"]))
            vtx.i32(lod + 4, vmeshes - lod)
            for mi, mesh in enumerate(m["meshes"]):
                vmesh = vmeshes + 9 * mi
                sgs = vtx.zeros(25 * len(mesh["groups"]))
                vtx.i32(vmesh, len(mesh["groups"]))
                vtx.i32(vmesh + 4, sgs - vmesh)
                for gi, g in enumerate(mesh["groups"]):
                    sg = sgs + 25 * gi
                    vpos = vtx.put(b"".join(struct.pack("<3BBH3b", 0, 1, 2, nb, orig, 0, 0, 0) for orig, nb in g["verts"]))
                    ipos = vtx.put(g["indices"].tobytes())
                    spos = vtx.put(b"".join(struct.pack("<4ihB2i", n, off, 0, 0, 1, fl, 0, 0) for fl, off, n in g["strips"]))
                    for o, v in ((0, len(g["verts"])), (4, vpos - sg), (8, len(g["indices"])), (12, ipos - sg), (16, len(g["strips"])), (20, spos - sg)):
                        vtx.i32(sg + o, v)
    struct.pack_into("<2i2H6i", vtx.b, 0, 7, 24, 53, 9, 3, _s32(checksum), 1, 0, len(body), vbp_off)
    return {"mdl": bytes(mdl.b), "vvd": bytes(vvd.b), "vtx": bytes(vtx.b), "models": models, "verts": all_verts, "pose": pose, "skin_table": skin_tbl,
            "names": names, "dirs": dirs, "n_bones": n_bones, "body": body}


def _s32(v):
    return v - (1 << 32) if v >= (1 << 31) else v
