"""Map ingestion: vt_bsp_* (vistrace_b200/csrc/vt_bsp.cpp, host only, no GPU) against the reference's own
BSPMap (oracle/_ref: libs/BSPParser — parser, Triangulate, displacement generation and smoothing) and the material bookkeeping of
World::World (source/objects/AccelStruct.cpp:236-414) on synthetic map files (tests/bsp_files.py), and against what the generator
put into the files."""
import struct

import numpy as np
import pytest

import bsp_files
from bsp_files import make_map, patch_lump, set_lump_entry

f4 = np.float32
FLOAT_FIELDS = ("p", "normals", "tangents", "uvs", "alphas")


def _same_floats(a, b):  # bit-identical; NaNs compare equal to NaNs whatever sign / payload the FPU picked
    a, b = np.ascontiguousarray(a, f4), np.ascontiguousarray(b, f4)
    na, nb = np.isnan(a), np.isnan(b)
    return np.array_equal(na, nb) and np.array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb])


def _assert_equals_reference(vt, oracle_mod, m):
    bsp = vt.BspFile(m["data"])
    info = bsp.info()
    tris, bino, texinfo = bsp.triangles()
    assert int(info["n_tris"]) == len(tris) == m["n_world_tris"]
    if not oracle_mod.available("reference"):
        return bsp, info, tris, bino, texinfo
    ref = oracle_mod.RefBsp(m["data"])
    assert ref.valid and ref.textures_ok
    assert (ref.n_tris, ref.n_materials, ref.n_static_props) == (int(info["n_tris"]), int(info["n_materials"]), int(info["n_static_props"]))
    rt, rb, rx = ref.triangles()
    for f in FLOAT_FIELDS:
        assert _same_floats(tris[f], rt[f]), f
    assert _same_floats(bino, rb), "binormals"
    np.testing.assert_array_equal(texinfo, rx)
    for f in ("material", "ent_idx", "one_sided"):
        np.testing.assert_array_equal(tris[f], rt[f])
    for k in range(ref.n_materials):
        a, b = bsp.material(k), ref.material(k)
        assert a.tobytes() == b.tobytes(), k
    for k in range(ref.n_static_props):
        assert bsp.static_prop(k).tobytes() == ref.static_prop(k).tobytes(), k
    return bsp, info, tris, bino, texinfo


@pytest.mark.parametrize("layout", ["grid", "tjunc", "single", None])
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_bsp_ingestion_equals_the_reference_parser(built, oracle_mod, seed, layout):
    import vistrace_b200 as vt

    m = make_map(seed=seed, layout=layout, version=19 + seed % 3, sprp_version=4 + seed % 3)
    bsp, info, tris, bino, texinfo = _assert_equals_reference(vt, oracle_mod, m)
    assert int(info["version"]) == 19 + seed % 3 and int(info["static_props_version"]) == 4 + seed % 3
    assert int(info["n_texinfos"]) == m["n_texinfos"] and int(info["n_displacements"]) == len(m["patches"])
    assert (tris["one_sided"] == 1).all() and (tris["ent_idx"] == 0).all()

    # ---- against the generator: brush polygons are fans around their first vertex, in file order, alphas 1
    k = 0
    for p in m["polys"]:
        if p["n"] < 3 or not m["emitted"](p["texinfo"]):
            continue
        pts = p["points"]
        for j in range(1, p["n"] - 1):
            np.testing.assert_array_equal(tris["p"][k], np.stack([pts[0], pts[j], pts[j + 1]]))
            assert texinfo[k] == p["texinfo"] and (tris["alphas"][k] == 1).all()
            # flat shading: the three vertex frames are equal, the normal is the unit normal of cross(p2 - p0, p1 - p0)
            assert _same_floats(tris["normals"][k][0], tris["normals"][k][1]) and _same_floats(tris["normals"][k][0], tris["normals"][k][2])
            k += 1
    # ---- displacement patches: 2 * 4^power triangles each, corner vertices on the base quad, alphas clamped to [0, 1]
    for p in m["patches"]:
        if not m["emitted"](p["texinfo"]):
            continue
        n = 2 * (1 << p["cell"]["power"]) ** 2
        blk = tris[k:k + n]
        assert (texinfo[k:k + n] == p["texinfo"]).all()
        a = blk["alphas"]
        assert ((a >= 0) & (a <= 1)).all()
        verts = {v.tobytes() for v in blk["p"].reshape(-1, 3)}
        assert all(c.tobytes() in verts for c in p["loop"])  # border offsets are zero: the corners stay on the base quad
        k += n
    assert k == len(tris)

    # ---- materials: one per distinct texture NAME in order of first use; flags of the texinfo that introduced it
    seen, want = {}, []
    for t in texinfo:
        name = m["names"][m["name_of_texdata"][m["tex_data"][int(t)]]]
        if name not in seen:
            seen[name] = len(want)
            want.append((name, int(t)))
    assert int(info["n_materials"]) == len(want)
    for i, (name, t) in enumerate(want):
        mat = bsp.material(i)
        assert mat["path"].decode() == name and int(mat["texinfo"]) == t and int(mat["surf_flags"]) == m["tex_flags"][t]
    np.testing.assert_array_equal(tris["material"], [seen[m["names"][m["name_of_texdata"][m["tex_data"][int(t)]]]] for t in texinfo])
    with pytest.raises(RuntimeError):
        bsp.material(len(want))

    # ---- static props
    assert int(info["n_static_props"]) == len(m["props"])
    for i, p in enumerate(m["props"]):
        sp = bsp.static_prop(i)
        np.testing.assert_array_equal(sp["pos"], p["pos"])
        np.testing.assert_array_equal(sp["ang"], p["ang"])
        assert sp["model"].decode() == p["model"] and int(sp["skin"]) == p["skin"]
    with pytest.raises(RuntimeError):
        bsp.static_prop(len(m["props"]))


def test_bsp_smoothing_welds_shared_displacement_borders(built, oracle_mod):
    """After the smoothing passes the vertices two equal-power patches share carry the same normal on both sides (what the passes
    are for) — whatever the rotation between the patches' local frames; this also checks that the generator's neighbour records
    are the geometrically consistent ones."""
    import vistrace_b200 as vt

    for seed in (5, 6, 7):
        m = make_map(seed=seed, layout="grid", power=3, all_drawn=True, n_polys=1)
        bsp, info, tris, _, texinfo = _assert_equals_reference(vt, oracle_mod, m)
        first = sum(p["n"] - 2 for p in m["polys"] if p["n"] >= 3 and m["emitted"](p["texinfo"]))
        pos = tris["p"][first:].reshape(-1, 3)
        nrm = tris["normals"][first:].reshape(-1, 3)
        patch = np.repeat(np.arange(len(m["patches"])), 3 * 2 * 64)
        keys = {}
        for i, p in enumerate(pos):
            keys.setdefault(p.tobytes(), []).append(i)
        shared = [ids for ids in keys.values() if len({int(patch[i]) for i in ids}) > 1]
        assert len(shared) >= 9
        for ids in shared:
            assert all(_same_floats(nrm[ids[0]], nrm[i]) for i in ids), (seed, pos[ids[0]])


def test_bsp_random_neighbour_records(built, oracle_mod):
    """Fuzz: random (in-range) neighbour records.  Whatever this library accepts it must triangulate exactly like the reference;
    it may reject files on which the reference walks outside its arrays (no claim there), and it must reject what the reference
    rejects."""
    import vistrace_b200 as vt

    accepted = 0
    for seed in range(24):
        m = make_map(seed=100 + seed, layout="grid", random_neighbours=True)
        bsp = vt.BspFile(m["data"])
        try:
            tris, bino, texinfo = bsp.triangles()
        except RuntimeError as e:
            msg = str(e)
            if "leaves" in msg or "zero length" in msg:
                continue  # the reference reads out of bounds / divides by zero there
            assert "span does not cover" in msg or "out of range" in msg, msg
            if oracle_mod.available("reference"):
                assert not oracle_mod.RefBsp(m["data"]).valid, msg
            continue
        accepted += 1
        if oracle_mod.available("reference"):
            ref = oracle_mod.RefBsp(m["data"])
            assert ref.valid
            rt, rb, _ = ref.triangles()
            for f in FLOAT_FIELDS:
                assert _same_floats(tris[f], rt[f]), (seed, f)
            assert _same_floats(bino, rb)
    assert accepted >= 3


def _rejected(vt, oracle_mod, data, reference_also=True):
    with pytest.raises(RuntimeError):
        vt.BspFile(data).info()
    with pytest.raises(RuntimeError):
        vt.BspFile(data).triangles()
    if reference_also and oracle_mod.available("reference"):
        assert not oracle_mod.RefBsp(data).valid


def test_bsp_malformed_files_are_rejected(built, oracle_mod):
    import vistrace_b200 as vt

    m = make_map(seed=7, layout="grid")
    data, d = m["data"], m["directory"]
    _rejected(vt, oracle_mod, b"")
    _rejected(vt, oracle_mod, data[:500])
    _rejected(vt, oracle_mod, b"XBSP" + data[4:])
    for version in (18, 22):
        _rejected(vt, oracle_mod, data[:4] + struct.pack("<i", version) + data[8:])
    # a required lump that is absent, runs past the file, or is not a whole number of records
    for lump in (bsp_files.L_VERTICES, bsp_files.L_FACES, bsp_files.L_DISPINFO, bsp_files.L_STRING_TABLE, bsp_files.L_GAME):
        _rejected(vt, oracle_mod, set_lump_entry(data, lump, offset=0))
    _rejected(vt, oracle_mod, set_lump_entry(data, bsp_files.L_FACES, length=len(data)))
    _rejected(vt, oracle_mod, set_lump_entry(data, bsp_files.L_FACES, length=d[bsp_files.L_FACES][1] - 3))
    _rejected(vt, oracle_mod, set_lump_entry(data, bsp_files.L_PLANES, length=-20))
    # static props: unsupported version, counts that do not fill the lump
    game_off = d[bsp_files.L_GAME][0]
    _rejected(vt, oracle_mod, data[:game_off + 4 + 16 + 6] + struct.pack("<H", 7) + data[game_off + 4 + 16 + 8:])
    sprp_off, sprp_len = m["sprp"]
    _rejected(vt, oracle_mod, data[:sprp_off] + struct.pack("<i", 3) + data[sprp_off + 4:])
    _rejected(vt, oracle_mod, data[:game_off] + struct.pack("<i", 1 << 20) + data[game_off + 4:])
    # a face whose edge list leaves the surfedge lump; a texinfo whose texdata index is out of range
    first_face = d[bsp_files.L_FACES][0]
    _rejected(vt, oracle_mod, data[:first_face + 4] + struct.pack("<i", 1 << 20) + data[first_face + 8:])
    ti0 = m["polys"][0]["texinfo"]
    _rejected(vt, oracle_mod, patch_lump(data, d, bsp_files.L_TEXINFO, 72 * ti0 + 68, struct.pack("<i", 99)))
    # no drawable triangle at all
    nodraw = data
    for t in range(m["n_texinfos"]):
        nodraw = patch_lump(nodraw, d, bsp_files.L_TEXINFO, 72 * t + 64, struct.pack("<I", bsp_files.SURF_NODRAW))
    _rejected(vt, oracle_mod, nodraw)

    # ---- files the reference reads out of bounds on (it does not check these): rejected here, no statement about the reference
    _rejected(vt, oracle_mod, patch_lump(data, d, bsp_files.L_MODELS, 40, struct.pack("<ii", 0, 1 << 16)), reference_also=False)      # worldspawn face range
    _rejected(vt, oracle_mod, patch_lump(data, d, bsp_files.L_DISPINFO, 36, struct.pack("<H", 60000)), reference_also=False)            # dispinfo.mapFace
    _rejected(vt, oracle_mod, patch_lump(data, d, bsp_files.L_DISPINFO, 12, struct.pack("<i", 1 << 24)), reference_also=False)          # dispinfo.dispVertStart
    _rejected(vt, oracle_mod, patch_lump(data, d, bsp_files.L_DISPINFO, 20, struct.pack("<i", 9)), reference_also=False)                # dispinfo.power
    _rejected(vt, oracle_mod, patch_lump(data, d, bsp_files.L_DISPINFO, 96 + 8, struct.pack("<B", 9)), reference_also=False)            # nine corner neighbours
    _rejected(vt, oracle_mod, patch_lump(data, d, bsp_files.L_DISPINFO, 96, struct.pack("<HHHHB", 500, 0, 0, 0, 1)), reference_also=False)  # neighbour index
    # texture name without a terminator / starting past the string data
    sd_off, sd_len = d[bsp_files.L_STRING_DATA]
    _rejected(vt, oracle_mod, data[:sd_off] + b"x" * sd_len + data[sd_off + sd_len:], reference_also=False)
    _rejected(vt, oracle_mod, patch_lump(data, d, bsp_files.L_STRING_TABLE, 0, struct.pack("<i", 1 << 20)), reference_also=False)
    # static prop naming a dictionary entry that does not exist: the map loads, the prop query fails (as BSPMap::GetStaticProp throws)
    bad_prop = data[:sprp_off + sprp_len - 64 + 24] + struct.pack("<H", 77) + data[sprp_off + sprp_len - 64 + 26:]
    b = vt.BspFile(bad_prop)
    assert int(b.info()["n_static_props"]) == len(m["props"])
    with pytest.raises(RuntimeError):
        b.static_prop(len(m["props"]) - 1)
    if oracle_mod.available("reference"):
        assert oracle_mod.RefBsp(bad_prop).static_prop(len(m["props"]) - 1) is None


def test_bsp_world_goes_through_the_engine_tree_builder(built):
    """The records are ready for vt_accel_populate: the host builder accepts them (no GPU needed for the hierarchy)."""
    import vistrace_b200 as vt
    from vistrace_b200 import abi

    m = make_map(seed=3, layout="grid")
    bsp = vt.BspFile(m["data"])
    tris, _, _ = bsp.triangles()
    n_mats = int(bsp.info()["n_materials"])
    mats = abi.default_materials(n_mats)
    for i in range(n_mats):
        mats["surf_flags"][i] = bsp.material(i)["surf_flags"]
    ents = np.zeros(1, abi.ENTITY)
    ents["colour"] = 1.0
    scene = abi.SceneData(tris, mats, ents, [])
    nodes, prims = vt.build_bvh(scene)
    assert sorted(prims.tolist()) == list(range(len(tris)))
