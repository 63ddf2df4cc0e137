"""Synthetic Source-engine map files (VBSP 19-21) for tests/test_bsp.py: brush polygons, displacement patches with neighbour data,
several texinfos sharing texture names, a second brush model, a static-prop game lump.  Written from the published lump layouts
(the sizes and offsets asserted in tests/test_bsp.py against the reference's own structs via the compiled checker's results).

Displacement layouts are geometrically CONSISTENT, as compiled maps are: neighbours share corner vertices exactly, the orientation
code is the rotation between the two patches' local frames, sub-neighbour 0 / 1 cover the low / high half of a long edge.  That
keeps the reference's unchecked neighbour walk inside its arrays (it has no bounds checks of its own)."""
import struct

import numpy as np

f4 = np.float32
HEADER_SIZE = 1036
L_PLANES, L_TEXDATA, L_VERTICES, L_TEXINFO, L_FACES, L_EDGES, L_SURFEDGES, L_MODELS = 1, 2, 3, 6, 7, 12, 13, 14
L_DISPINFO, L_DISP_VERTS, L_GAME, L_STRING_DATA, L_STRING_TABLE = 26, 33, 35, 43, 44
SURF_SKY, SURF_TRIGGER, SURF_NODRAW, SURF_SKIP = 0x4, 0x40, 0x80, 0x200
SPRP = (ord("s") << 24) | (ord("p") << 16) | (ord("r") << 8) | ord("p")
DPRP = (ord("d") << 24) | (ord("p") << 16) | (ord("r") << 8) | ord("p")
# world directions a patch edge can face, clockwise: west, north, east, south
DIRS = [(-1, 0), (0, 1), (1, 0), (0, -1)]


class _Mesh:
    """Vertex / edge / surfedge pools shared by all faces (edge 0 is a dummy: a surfedge cannot name it reversed)."""

    def __init__(self):
        self.verts, self.vert_ids = [], {}
        self.edges, self.edge_ids = [(0, 0)], {}
        self.surfedges = []

    def vert(self, p):
        key = tuple(np.asarray(p, f4).tolist())
        if key not in self.vert_ids:
            self.vert_ids[key] = len(self.verts)
            self.verts.append(key)
        return self.vert_ids[key]

    def loop(self, points):
        """Surfedges of a closed vertex loop -> (first surfedge, count)."""
        ids = [self.vert(p) for p in points]
        first = len(self.surfedges)
        for k, a in enumerate(ids):
            b = ids[(k + 1) % len(ids)]
            if (a, b) in self.edge_ids:
                self.surfedges.append(self.edge_ids[(a, b)])
            elif (b, a) in self.edge_ids:
                self.surfedges.append(-self.edge_ids[(b, a)])
            else:
                self.edge_ids[(a, b)] = len(self.edges)
                self.edges.append((a, b))
                self.surfedges.append(len(self.edges) - 1)
        return first, len(ids)


def _face(plane, first_edge, n_edges, texinfo, dispinfo=-1):
    return struct.pack("<HBBihhhh4sifiiiiiHHI", plane, 0, 1, first_edge, n_edges, texinfo, dispinfo, -1, b"\0\0\0\0", -1, 1.0, 0, 0, 0, 0, -1, 0, 0, 0)


def _dispinfo(start, vert_start, power, map_face, edge_nb, corner_nb):
    """edge_nb[e] = [(index, orientation, span, neighbour_span) or None] * 2; corner_nb[c] = list of indices (<= 4)."""
    out = struct.pack("<3fiiiifiHxxii", *start, vert_start, 0, power, 0, 0.0, 1, map_face, 0, 0)
    for e in range(4):
        for s in range(2):
            nb = edge_nb[e][s]
            out += struct.pack("<HBBBx", *(nb if nb is not None else (0xFFFF, 0, 0, 0)))
    for c in range(4):
        ids = list(corner_nb[c])
        out += struct.pack("<4HBx", *(ids + [0] * (4 - len(ids)))[:4], len(ids))
    out += struct.pack("<10I", *([0xFFFFFFFF] * 10))
    assert len(out) == 176
    return out


def _patch_layout(kind, rng):
    """Axis-aligned patches in the z = 0 plane: list of dict(x0, y0, size, rot, power)."""
    cells = []
    if kind == "grid":
        w, h = int(rng.integers(2, 4)), int(rng.integers(1, 3))
        for cy in range(h):
            for cx in range(w):
                cells.append(dict(x0=cx, y0=cy, size=1))
    elif kind == "tjunc":  # one double-size patch with two small ones on its east and two on its north side, a small one diagonally
        cells.append(dict(x0=0, y0=0, size=2))
        cells += [dict(x0=2, y0=0, size=1), dict(x0=2, y0=1, size=1), dict(x0=0, y0=2, size=1), dict(x0=1, y0=2, size=1), dict(x0=2, y0=2, size=1)]
    elif kind == "single":
        cells.append(dict(x0=0, y0=0, size=1))
    for c in cells:
        c["rot"] = int(rng.integers(0, 4))
        c["power"] = int(rng.integers(2, 5))
    if kind == "tjunc":
        cells[0]["power"] = int(rng.integers(3, 5))
    return cells


def _patch_neighbours(cells):
    """Edge / corner neighbour records per patch, in each patch's LOCAL edge / corner numbering (see the module docstring)."""
    def interval(c, axis):
        return (c["x0"], c["x0"] + c["size"]) if axis == 0 else (c["y0"], c["y0"] + c["size"])

    edge_nb = [[[None, None] for _ in range(4)] for _ in cells]
    corner_nb = [[[] for _ in range(4)] for _ in cells]
    for ia, a in enumerate(cells):
        for w, (dx, dy) in enumerate(DIRS):
            ea = (w - a["rot"]) % 4
            axis, along = (0, 1) if dx else (1, 0)  # the edge lies at a fixed coordinate on `axis` and runs along `along`
            fixed = interval(a, axis)[1 if (dx + dy) > 0 else 0]
            lo, hi = interval(a, along)
            touching = []
            for ib, b in enumerate(cells):
                if ib == ia:
                    continue
                if interval(b, axis)[0 if (dx + dy) > 0 else 1] != fixed:
                    continue
                blo, bhi = interval(b, along)
                if bhi <= lo or blo >= hi:
                    continue
                touching.append((ib, blo, bhi))
            # which way the local free coordinate of edge `ea` runs in the world: the local axes are the world axes rotated clockwise
            # by rot * 90 degrees; local x = R^rot(+X), local y = R^rot(+Y)
            rot = a["rot"]
            lx = [(1, 0), (0, -1), (-1, 0), (0, 1)][rot]
            ly = [(0, 1), (1, 0), (0, -1), (-1, 0)][rot]
            free_axis = ly if ea in (0, 2) else lx  # edges 0 / 2 run along local y, 1 / 3 along local x
            increasing = (free_axis[along] > 0)
            for ib, blo, bhi in touching:
                b = cells[ib]
                orientation = (a["rot"] - b["rot"]) % 4
                if (blo, bhi) == (lo, hi):
                    edge_nb[ia][ea][0] = (ib, orientation, 0, 0)
                elif bhi - blo < hi - lo:  # b covers half of a's edge
                    low_half_world = blo == lo
                    low_half_local = low_half_world if increasing else not low_half_world
                    sub = 0 if low_half_local else 1
                    edge_nb[ia][ea][sub] = (ib, orientation, 1 if sub == 0 else 2, 0)
                else:  # a covers half of b's edge
                    edge_nb[ia][ea][0] = (ib, orientation, 0, 1 if blo == lo else 2)
        # corner neighbours: patches that touch a corner point of a without sharing an edge with it
        rot = a["rot"]
        loop = [(a["x0"], a["y0"]), (a["x0"], a["y0"] + a["size"]), (a["x0"] + a["size"], a["y0"] + a["size"]), (a["x0"] + a["size"], a["y0"])]
        edge_ids = {nb[0] for e in range(4) for nb in edge_nb[ia][e] if nb is not None}
        for c in range(4):
            pt = loop[(c + rot) % 4]
            for ib, b in enumerate(cells):
                if ib == ia or ib in edge_ids:
                    continue
                bl = [(b["x0"], b["y0"]), (b["x0"], b["y0"] + b["size"]), (b["x0"] + b["size"], b["y0"] + b["size"]), (b["x0"] + b["size"], b["y0"])]
                if pt in bl and len(corner_nb[ia][c]) < 4:
                    corner_nb[ia][c].append(ib)
    return edge_nb, corner_nb


def make_map(seed=0, layout="grid", version=20, sprp_version=6, n_polys=6, n_props=3, scale=128.0, random_neighbours=False, power=None, all_drawn=False):
    """-> dict(data=bytes, n_world_tris, polys, patches, names, ...).  `random_neighbours` replaces the consistent neighbour records
    by random in-range ones (fuzz: the reference may then walk outside its arrays, which this library rejects)."""
    rng = np.random.default_rng(1000 + seed)
    mesh = _Mesh()
    planes, faces, texinfos, texdatas = [], [], [], []
    # texture names: two texdatas share a name, two texinfos share a texdata
    names = [f"nature/ground{seed}", "concrete/wall_a", "tools/toolsnodraw", "tools/toolsskybox", "Nature/Blend_Grass"]
    name_of_texdata = [0, 1, 1, 2, 3, 4]
    for nm in name_of_texdata:
        texdatas.append(struct.pack("<3fiiiii", *rng.uniform(0, 1, 3).astype(f4), nm, int(rng.choice([64, 128, 256, 512])), int(rng.choice([64, 128, 256])), 64, 64))
    tex_flags = [0, 0, 0, SURF_NODRAW, SURF_SKY, 0, SURF_TRIGGER, SURF_SKIP, 0x800]
    tex_data = [0, 1, 2, 3, 4, 5, 1, 0, 1]
    for fl, td in zip(tex_flags, tex_data):
        vecs = rng.normal(0, 1, (2, 4)).astype(f4)
        vecs[:, 3] = rng.uniform(-64, 64, 2)
        texinfos.append(struct.pack("<8f8fIi", *vecs.ravel(), *np.zeros(8, f4), fl, td))
    drawable = [i for i, fl in enumerate(tex_flags) if not fl & (SURF_NODRAW | SURF_TRIGGER | SURF_SKIP)]

    # ---- brush polygons: convex loops in random planes; consecutive ones share an edge
    polys = []
    prev_edge = None
    for k in range(n_polys):
        n = int(rng.integers(3, 7))
        centre = rng.uniform(-500, 500, 3)
        u = rng.normal(0, 1, 3)
        u /= np.linalg.norm(u)
        v = np.cross(u, rng.normal(0, 1, 3))
        v /= np.linalg.norm(v)
        ang = np.sort(rng.uniform(0, 2 * np.pi, n))
        rad = rng.uniform(20, 90, n)
        pts = [(centre + r * (np.cos(a) * u + np.sin(a) * v)).astype(f4) for a, r in zip(ang, rad)]
        if prev_edge is not None and k % 2 == 1:
            pts[0], pts[1] = prev_edge[1], prev_edge[0]  # traversed the other way round -> a negative surfedge
        prev_edge = (pts[0], pts[1])
        normal = np.cross(u, v).astype(f4)
        planes.append(struct.pack("<3ffi", *normal, f4(normal @ centre), 0))
        first, cnt = mesh.loop(pts)
        ti = int(rng.integers(0, len(texinfos))) if k else drawable[0]
        if k == 2:
            ti = -1  # no texinfo: dropped
        faces.append(_face(len(planes) - 1, first, cnt, ti))
        polys.append(dict(points=pts, texinfo=ti, n=cnt))
    # a degenerate two-edge face (dropped) and a face whose texinfo index is past the lump (dropped)
    first, cnt = mesh.loop([np.array([0, 0, 900], f4), np.array([10, 0, 900], f4)])
    faces.append(_face(0, first, cnt, drawable[0]))
    polys.append(dict(points=[], texinfo=drawable[0], n=2))
    first, cnt = mesh.loop([np.array([0, 0, 950], f4), np.array([10, 0, 950], f4), np.array([0, 10, 950], f4)])
    faces.append(_face(0, first, cnt, len(texinfos) + 3))
    polys.append(dict(points=[], texinfo=len(texinfos) + 3, n=3))

    # ---- displacement patches
    cells = _patch_layout(layout, rng) if layout else []
    if power is not None:
        for c in cells:
            c["power"] = power
    edge_nb, corner_nb = _patch_neighbours(cells) if cells else ([], [])
    dispinfos, dispverts = [], []
    planes.append(struct.pack("<3ffi", 0.0, 0.0, 1.0, 0.0, 2))
    ground_plane = len(planes) - 1
    patches = []
    for i, c in enumerate(cells):
        x0, y0, sz = c["x0"] * scale, c["y0"] * scale, c["size"] * scale
        loop = [np.array(p, f4) for p in ((x0, y0, 0), (x0, y0 + sz, 0), (x0 + sz, y0 + sz, 0), (x0 + sz, y0, 0))]
        first, cnt = mesh.loop(loop)
        ti = drawable[(i + 1) % len(drawable)] if (i != 1 or all_drawn) else 3  # patch 1 sits on a NODRAW face: smoothed with, not emitted
        faces.append(_face(ground_plane, first, cnt, ti, dispinfo=i))
        side = (1 << c["power"]) + 1
        start = loop[c["rot"]] + rng.uniform(-0.5, 0.5, 3).astype(f4)
        vecs = rng.normal(0, 1, (side, side, 3)).astype(f4)
        vecs[..., 2] = np.abs(vecs[..., 2]) + 0.5
        vecs /= np.linalg.norm(vecs, axis=-1, keepdims=True).astype(f4)
        dist = rng.uniform(0, 24, (side, side)).astype(f4)
        dist[0, :] = dist[-1, :] = dist[:, 0] = dist[:, -1] = 0  # shared borders stay on the base quad, as a compiled map keeps them welded
        alpha = rng.uniform(-40, 300, (side, side)).astype(f4)
        if random_neighbours:
            enb = [[(int(rng.integers(0, len(cells))), int(rng.integers(0, 4)), int(rng.integers(0, 3)), int(rng.integers(0, 3))) if rng.random() < 0.5 else None
                    for _ in range(2)] for _ in range(4)]
            cnb = [[int(rng.integers(0, len(cells))) for _ in range(int(rng.integers(0, 5)))] for _ in range(4)]
        else:
            enb, cnb = edge_nb[i], corner_nb[i]
        dispinfos.append(_dispinfo(start, len(dispverts), c["power"], len(faces) - 1, enb, cnb))
        for a in range(side):
            for b in range(side):
                dispverts.append(struct.pack("<3fff", *vecs[a, b], dist[a, b], alpha[a, b]))
        patches.append(dict(cell=c, texinfo=ti, face=len(faces) - 1, loop=loop, edge_nb=enb, corner_nb=cnb))
    n_world_faces = len(faces)
    # ---- a second brush model (a door, say): its faces are not part of the world
    first, cnt = mesh.loop([np.array(p, f4) for p in ((0, 0, 2000), (50, 0, 2000), (50, 50, 2000), (0, 50, 2000))])
    faces.append(_face(ground_plane, first, cnt, drawable[0]))
    models = struct.pack("<9fiii", *np.zeros(9, f4), 0, 0, n_world_faces) + struct.pack("<9fiii", *np.zeros(9, f4), 0, n_world_faces, 1)

    # ---- strings
    string_data, string_table = b"", []
    for nm in names:
        string_table.append(len(string_data))
        string_data += nm.encode() + b"\0"

    # ---- game lump: a detail-prop lump the parser skips + the static props
    prop_size = {4: 56, 5: 60, 6: 64}.get(sprp_version, 64)
    dict_names = [b"models/props/tree01.mdl", b"models/props_c17/oildrum001.mdl"]
    props = []
    sprp = struct.pack("<i", len(dict_names)) + b"".join(nm.ljust(128, b"\0") for nm in dict_names)
    leaves = [1, 2, 3]
    sprp += struct.pack("<i", len(leaves)) + struct.pack(f"<{len(leaves)}H", *leaves)
    sprp += struct.pack("<i", n_props)
    for k in range(n_props):
        pos, ang = rng.uniform(-1000, 1000, 3).astype(f4), rng.uniform(-180, 180, 3).astype(f4)
        ptype, skin = int(rng.integers(0, len(dict_names))), int(rng.integers(0, 4))
        rec = struct.pack("<3f3fHHHBBiff3f", *pos, *ang, ptype, 0, 1, 6, 0, skin, -1.0, 0.0, *pos)
        rec += {4: b"", 5: struct.pack("<f", 1.0), 6: struct.pack("<fHH", 1.0, 0, 0)}.get(sprp_version, struct.pack("<fHH", 1.0, 0, 0))
        assert len(rec) == prop_size
        sprp += rec
        props.append(dict(pos=pos, ang=ang, model=dict_names[ptype].decode(), skin=skin))
    dprp = b"\x01\x02\x03\x04" * 5

    lumps = {
        L_PLANES: b"".join(planes), L_TEXDATA: b"".join(texdatas), L_VERTICES: b"".join(struct.pack("<3f", *v) for v in mesh.verts),
        L_TEXINFO: b"".join(texinfos), L_FACES: b"".join(faces), L_EDGES: b"".join(struct.pack("<HH", *e) for e in mesh.edges),
        L_SURFEDGES: struct.pack(f"<{len(mesh.surfedges)}i", *mesh.surfedges), L_MODELS: models, L_DISPINFO: b"".join(dispinfos),
        L_DISP_VERTS: b"".join(dispverts), L_STRING_DATA: string_data, L_STRING_TABLE: struct.pack(f"<{len(string_table)}i", *string_table),
    }
    # lay the file out: header, the lumps above (4-byte aligned; an empty lump still gets an offset > 0), the game lump last
    body, directory, off = b"", {}, HEADER_SIZE
    for lid, blob in lumps.items():
        directory[lid] = (off, len(blob))
        pad = (-len(blob)) % 4
        body += blob + b"\0" * pad
        off += len(blob) + pad
    game_dir_size = 4 + 2 * 16
    dprp_off, sprp_off = off + game_dir_size, off + game_dir_size + len(dprp)
    game = struct.pack("<i", 2) + struct.pack("<iHHii", DPRP, 0, 4, dprp_off, len(dprp)) + struct.pack("<iHHii", SPRP, 0, sprp_version, sprp_off, len(sprp))
    game_blob = game + dprp + sprp
    directory[L_GAME] = (off, len(game_blob))
    body += game_blob
    header = struct.pack("<4si", b"VBSP", version)
    for lid in range(64):
        o, ln = directory.get(lid, (0, 0))
        header += struct.pack("<iiii", o, ln, 0, 0)
    header += struct.pack("<i", 1)
    assert len(header) == HEADER_SIZE
    data = header + body

    def emitted(ti):
        return 0 <= ti < len(texinfos) and not tex_flags[ti] & (SURF_NODRAW | SURF_TRIGGER | SURF_SKIP)

    n_tris = sum(p["n"] - 2 for p in polys if p["n"] >= 3 and emitted(p["texinfo"]))
    n_tris += sum(2 * (1 << p["cell"]["power"]) ** 2 for p in patches if emitted(p["texinfo"]))
    return dict(data=data, n_world_tris=n_tris, polys=polys, patches=patches, names=names, tex_flags=tex_flags, tex_data=tex_data, name_of_texdata=name_of_texdata,
                props=props, directory=directory, sprp=(sprp_off, len(sprp)), n_texinfos=len(texinfos), emitted=emitted, verts=mesh.verts)


def patch_lump(data, directory, lump, offset, payload):
    """Overwrite bytes inside a lump (offset relative to the lump start)."""
    o = directory[lump][0] + offset
    return data[:o] + payload + data[o + len(payload):]


def set_lump_entry(data, lump, offset=None, length=None):
    """Rewrite a lump's directory entry."""
    at = 8 + 16 * lump
    o, ln = struct.unpack_from("<ii", data, at)
    return data[:at] + struct.pack("<ii", o if offset is None else offset, ln if length is None else length) + data[at + 8:]
