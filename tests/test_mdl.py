"""Model ingestion: vt_mdl_* (vistrace_b200/csrc/vt_mdl.cpp, host only, no GPU) against the reference's own
MDL / VVD / VTX parsers and its BodyGroup / Mesh constructors (oracle/_ref: libs/MDLParser, source/objects/Model.cpp) on synthetic
model files (tests/mdl_files.py), and against what the generator put into the files."""
import numpy as np
import pytest

from mdl_files import make_model

f4 = np.float32


def _derived(tris):
    """p0, e1 = p0 - p1, e2 = p2 - p0 as the Triangle constructor derives them (source/objects/Primitives.h:82)."""
    p = tris["p"]
    return p[:, 0], (p[:, 0] - p[:, 1]).astype(f4), (p[:, 2] - p[:, 0]).astype(f4)


def _same_floats(a, b):  # bit-identical, NaN patterns included up to the sign / payload the FPU picks
    a, b = np.asarray(a, f4), np.asarray(b, f4)
    return np.array_equal(a.view(np.uint32)[~np.isnan(a)], b.view(np.uint32)[~np.isnan(b)]) and np.array_equal(np.isnan(a), np.isnan(b))


@pytest.mark.parametrize("fixups", [False, True])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_mdl_ingestion_equals_the_reference_loader(built, oracle_mod, seed, fixups):
    import vistrace_b200 as vt

    m = make_model(seed=seed, body=((2, 1), (3,), (1, 2, 1)) if seed else ((2, 1), (3,)), fixups=fixups)
    files = vt.MdlFiles(m["mdl"], m["vvd"], m["vtx"])
    info = files.info()
    assert (int(info["n_bodygroups"]), int(info["n_bones"]), int(info["n_materials"]), int(info["n_skin_families"]), int(info["n_vertices"])) == \
        (len(m["body"]), m["n_bones"], len(m["names"]), len(m["skin_table"]), len(m["verts"]))
    ref = oracle_mod.RefModel(m["mdl"], m["vvd"], m["vtx"]) if oracle_mod.available("reference") else None
    if ref is not None:
        assert ref.valid
        assert (ref.n_bodygroups, ref.n_bones, ref.n_materials, ref.n_skin_families, ref.n_vertices) == \
            (int(info["n_bodygroups"]), int(info["n_bones"]), int(info["n_materials"]), int(info["n_skin_families"]), int(info["n_vertices"]))
        np.testing.assert_array_equal(files.bind_matrices().view(np.uint32), ref.bind_matrices().view(np.uint32))
    # bind matrices against the generator: glm::mat4 columns from the 3 x 4 row-major poseToBone (Model.cpp:242-254)
    for b, bind in enumerate(files.bind_matrices().reshape(-1, 4, 4)):
        np.testing.assert_array_equal(bind[:, :3], m["pose"][b].T)
        np.testing.assert_array_equal(bind[:, 3], (0, 0, 0, 1))
    nan_tangent_fixed = zero_normal_seen = False
    n_total = 0
    for bg, values in enumerate(m["body"]):
        assert files.bodygroup_values(bg) == len(values)
        for val in range(len(values)):
            tris, skin = files.mesh_triangles(bg, val)
            model = next(x for x in m["models"] if x["bg"] == bg and x["val"] == val)
            want_n = sum(n // 3 for mesh in model["meshes"] for g in mesh["groups"] for fl, off, n in g["strips"] if fl & 1)
            assert len(tris) == want_n > 0  # the triangle-STRIP strips are ignored (Model.cpp:33-35)
            n_total += want_n
            assert (tris["one_sided"] == 0).all() and (tris["alphas"] == 0).all() and (tris["ent_idx"] == 0).all()
            # positions / uvs / material / bones against what the generator wrote
            k = 0
            for mesh in model["meshes"]:
                for g in mesh["groups"]:
                    for fl, off, n in g["strips"]:
                        if not fl & 1:
                            continue
                        for i in range(off, off + n - 2, 3):
                            for j in range(3):
                                orig, nb = g["verts"][int(g["indices"][i + j])]
                                v = m["verts"][model["first_vertex"] + mesh["vert_off"] + orig]
                                np.testing.assert_array_equal(tris["p"][k, j], v[3])
                                np.testing.assert_array_equal(tris["uvs"][k, j], v[5])
                                if nb > 0:
                                    assert skin["num_bones"][k, j] == nb
                                    np.testing.assert_array_equal(skin["weights"][k, j], v[0])
                                    np.testing.assert_array_equal(skin["bone_ids"][k, j], v[1])
                                else:
                                    assert skin["num_bones"][k, j] == 1 and skin["weights"][k, j, 0] == 1 and skin["bone_ids"][k, j, 0] == 0
                                if np.isnan(v[6][:3]).any():
                                    e1 = (tris["p"][k, 0] - tris["p"][k, 1]).astype(f4)
                                    want = (e1 * (f4(1) / np.sqrt((e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]).astype(f4)))).astype(f4)
                                    np.testing.assert_array_equal(tris["tangents"][k, j], want)  # normalize(e1), Model.cpp:101-103
                                    nan_tangent_fixed = True
                                if not v[4].any():
                                    assert np.isnan(tris["normals"][k, j]).all()  # normalize(0): kept, see vt_mdl.cpp header
                                    zero_normal_seen = True
                            assert tris["material"][k] == mesh["material"]
                            k += 1
            assert k == want_n
            if ref is not None:
                assert ref.bodygroup_values(bg) == len(values)
                rt, rs = ref.mesh(bg, val)
                assert len(rt) == len(tris)
                p0, e1, e2 = _derived(tris)
                assert _same_floats(p0, rt["p"][:, 0]) and _same_floats(e1, rt["p"][:, 1]) and _same_floats(e2, rt["p"][:, 2])
                for f in ("normals", "tangents", "uvs", "alphas"):
                    assert _same_floats(tris[f], rt[f]), f
                np.testing.assert_array_equal(tris["material"], rt["material"])
                np.testing.assert_array_equal(skin["num_bones"], rs["num_bones"])
                np.testing.assert_array_equal(skin["weights"].view(np.uint32), rs["weights"].view(np.uint32))
                np.testing.assert_array_equal(skin["bone_ids"], rs["bone_ids"])
    assert nan_tangent_fixed and zero_normal_seen and n_total > 10
    for skin_id in range(len(m["skin_table"]) + 1):
        for mat in range(len(m["names"]) + 1):
            want = int(m["skin_table"][skin_id, mat]) if skin_id < len(m["skin_table"]) and mat < len(m["names"]) else 0  # Model.cpp:349-357
            assert files.material_index(skin_id, mat) == want
            if ref is not None:
                assert ref.material_index(skin_id, mat) == want
    for mat, name in enumerate(m["names"]):
        for d, directory in enumerate(m["dirs"]):
            assert files.material_path(mat, d) == directory + name
            if ref is not None:
                assert ref.material_path(mat, d) == directory + name


def test_mdl_rejects_what_the_reference_rejects_and_never_reads_out_of_bounds(built, oracle_mod):
    import vistrace_b200 as vt

    m = make_model(seed=5)
    good = (m["mdl"], m["vvd"], m["vtx"])

    def both_reject(mdl, vvd, vtx):
        with pytest.raises(RuntimeError):
            vt.MdlFiles(mdl, vvd, vtx).info()
        if oracle_mod.available("reference"):
            assert not oracle_mod.RefModel(mdl, vvd, vtx).valid

    both_reject(b"IDSX" + good[0][4:], good[1], good[2])                                     # wrong id (MDLParser.cpp:50)
    both_reject(make_model(seed=5, version=49)["mdl"], good[1], good[2])                     # version > 48
    both_reject(good[0], good[1][:4] + (5).to_bytes(4, "little") + good[1][8:], good[2])     # vvd version != 4 (VVDParser.cpp:41-45)
    other = make_model(seed=5, checksum=0x11112222)
    both_reject(good[0], other["vvd"], good[2])                                              # vvd checksum mismatch
    both_reject(good[0], good[1], other["vtx"])                                              # vtx checksum mismatch (VTXParser.cpp:44-47)
    both_reject(good[0], good[1], (6).to_bytes(4, "little") + good[2][4:])                   # vtx version != 7
    both_reject(good[0], good[1][:80], good[2])                                              # vvd truncated below its vertex count (VVDParser.cpp:49)
    # truncations and bit-flips the reference would walk off the end of: an error here, never a crash
    rng = np.random.default_rng(0)
    for which in range(3):
        for cut in (0.3, 0.6, 0.9):
            files = list(good)
            files[which] = files[which][: int(len(files[which]) * cut)]
            f = vt.MdlFiles(*files)
            try:
                f.info()
                for bg in range(len(m["body"])):
                    for val in range(f.bodygroup_values(bg)):
                        f.mesh_triangles(bg, val)
            except RuntimeError:
                pass
    for _ in range(200):
        files = [bytearray(b) for b in good]
        which = int(rng.integers(0, 3))
        for _ in range(3):
            pos = int(rng.integers(12, len(files[which]) - 4))
            files[which][pos:pos + 4] = rng.integers(0, 256, 4, dtype=np.uint8).tobytes()
        f = vt.MdlFiles(*[bytes(b) for b in files])
        try:
            info = f.info()
            for bg in range(min(4, int(info["n_bodygroups"]))):
                for val in range(min(4, f.bodygroup_values(bg))):
                    f.mesh_triangles(bg, val)
            f.bind_matrices() if int(info["n_bones"]) < 100000 else None
        except RuntimeError:
            pass


def test_mdl_triangles_through_skin_triangles_into_a_scene(built):
    """The ingestion chain of PopulateAccel for an entity (source/objects/AccelStruct.cpp:716-749): Model::GetMesh -> material through
    the skin table -> SkinTriangle with the entity's bone matrices -> the flat triangle list a scene is built from (host only)."""
    import vistrace_b200 as vt
    from vistrace_b200 import abi

    m = make_model(seed=3)
    files = vt.MdlFiles(m["mdl"], m["vvd"], m["vtx"])
    binds = files.bind_matrices()
    bones = np.tile(np.eye(4, dtype=f4).reshape(1, 16), (len(binds), 1))
    bones[:, 12:15] += np.arange(len(binds), dtype=f4)[:, None]  # a translation per bone (column-major glm::mat4)
    parts = []
    for bg in range(len(m["body"])):
        tris, skin = files.mesh_triangles(bg, 0)  # bodygroup value 0 of every body group
        tris["material"] = [files.material_index(1, int(x)) for x in tris["material"]]
        tris["ent_idx"] = 1
        ok = np.isfinite(tris["normals"]).all((1, 2))
        parts.append(vt.skin_triangles(tris[ok], skin[ok], bones, binds))
    tris = np.concatenate(parts)
    assert len(tris) > 5 and np.isfinite(tris["p"]).all()
    scene = abi.SceneData(tris, np.zeros(len(m["names"]), abi.MATERIAL), np.zeros(2, abi.ENTITY))
    nodes, prims = vt.build_bvh(scene)
    assert len(prims) == len(tris) and len(nodes) >= 1
