"""CPU tests of the oracle: the plain-C restatement (oracle/vt_oracle.c) against
  (a) golden vectors generated from the UNMODIFIED reference (tests/golden/*.npz),
  (b) the known-answer tests the vendored bvh library ships, and
  (c) the reference itself (oracle/_ref/libvt_ref.so) when it is present.
Everything is compared bit for bit: the restatement follows the reference's arithmetic order."""
import numpy as np
import pytest

from conftest import load_golden

FLT_MAX = np.finfo(np.float32).max


def _port_scene(oracle_mod, scene, z):
    cpu = oracle_mod.CpuScene(scene, "port", build_bvh=False)
    cpu.set_bvh(z["nodes"], z["prim_indices"])
    return cpu


@pytest.mark.parametrize("name", ["foliage_small", "props_small"])
def test_port_matches_golden_hits_and_attrs(oracle_mod, name):
    scene, z = load_golden(name)
    cpu = _port_scene(oracle_mod, scene, z)
    np.testing.assert_array_equal(cpu.tri_derived().view(np.uint32), z["tri_derived"].view(np.uint32))
    for rays_k, hits_k, attrs_k in (("rays", "hits", "attrs"), ("bounce_rays", "bounce_hits", "bounce_attrs"), ("extra_rays", "extra_hits", "extra_attrs")):
        if rays_k not in z:
            continue
        got = cpu.traverse(z[rays_k], want_attrs=True, want_stats=True)
        assert got["hits"].tobytes() == z[hits_k].tobytes(), (name, rays_k)
        assert got["attrs"].tobytes() == z[attrs_k].tobytes(), (name, rays_k)
        if rays_k == "rays":  # SingleRayTraverser::Statistics of the reference run
            assert [got["steps"], got["isects"]] == [int(v) for v in z["stats"]]
        # the attribute stage on its own reproduces the same records from the golden hits
        assert cpu.trace_result(z[rays_k], z[hits_k]).tobytes() == z[attrs_k].tobytes()


def test_port_texture_sampler_matches_golden(oracle_mod):
    scene, z = load_golden("foliage_small")
    cpu = _port_scene(oracle_mod, scene, z)
    for i in range(int(z["n_textures"])):  # texture 0 wraps, texture 1 clamps; mips 0..5 incl. fractional levels
        got = cpu.sample(i, z["tex_uvm"])
        np.testing.assert_array_equal(got.view(np.uint32), z[f"tex{i}_samples"].view(np.uint32))


def test_kat_node_intersector_flat_box_negative_zero(oracle_mod):
    """libs/bvh/test/node_intersectors.cpp:18-36 — box flat in z, ray direction (0, -0, 1): must be hit."""
    z = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "kat_node_intersect.npz"))
    entry, exit_ = oracle_mod.node_intersect(z["node"], z["ray"], "port")
    assert entry <= exit_
    assert [np.float32(entry), np.float32(exit_)] == list(z["entry_exit"])


def test_kat_simple_example_quad(oracle_mod):
    """libs/bvh/test/simple_example.cpp:63-83 — ray (0,0,0)->+Z, t in [0,100] must hit the 2-triangle quad at z = 1.
    (VisTrace's primitive is two-sided here: oneSided = false, so winding does not matter.)"""
    from vistrace_b200 import abi

    tris = np.zeros(2, abi.TRI_IN)
    tris["p"][0] = [[1, -1, 1], [1, 1, 1], [-1, 1, 1]]
    tris["p"][1] = [[1, -1, 1], [-1, -1, 1], [-1, 1, 1]]
    tris["normals"], tris["tangents"], tris["alphas"] = (0, 0, 1), (1, 0, 0), 1.0
    tris["uvs"] = [[0, 0], [1, 0], [1, 1]]
    scene = abi.SceneData(tris)
    cpu = oracle_mod.CpuScene(scene, "port", build_bvh=False)
    nodes = np.zeros(1, abi.NODE)  # a single leaf root (single_ray_traverser.hpp:72-73)
    nodes["bounds"], nodes["prim_count"], nodes["first"] = (-1, 1, -1, 1, 1, 1), 2, 0
    cpu.set_bvh(nodes, np.arange(2, dtype=np.uint64))
    ray = np.zeros(1, abi.RAY)
    ray["d"], ray["tmax"] = (0, 0, 1), 100.0
    hit = cpu.traverse(ray)["hits"][0]
    assert hit["prim"] != abi.VT_MISS and hit["t"] == np.float32(1.0)
    ray["tmax"] = 0.5  # closed interval [tmin, tmax] (Primitives.h:189)
    assert cpu.traverse(ray)["hits"][0]["prim"] == abi.VT_MISS
    ray["tmax"] = 1.0
    assert cpu.traverse(ray)["hits"][0]["prim"] != abi.VT_MISS


def test_exact_tie_is_won_by_the_later_candidate(oracle_mod):
    """`t <= tmax` + tmax = t (single_ray_traverser.hpp:55-60): of two coplanar duplicates the one tested later wins."""
    from vistrace_b200 import abi

    tris = np.zeros(2, abi.TRI_IN)
    tris["p"][:] = [[-1, -1, 2], [1, -1, 2], [0, 1, 2]]
    tris["normals"], tris["tangents"], tris["alphas"] = (0, 0, 1), (1, 0, 0), 1.0
    tris["uvs"] = [[0, 0], [1, 0], [0, 1]]
    cpu = oracle_mod.CpuScene(abi.SceneData(tris), "port", build_bvh=False)
    nodes = np.zeros(1, abi.NODE)
    nodes["bounds"], nodes["prim_count"] = (-1, 1, -1, 1, 2, 2), 2
    ray = np.zeros(1, abi.RAY)
    ray["d"], ray["tmax"] = (0, 0, 1), FLT_MAX
    for order in ([0, 1], [1, 0]):
        cpu.set_bvh(nodes, np.array(order, np.uint64))
        assert cpu.traverse(ray)["hits"][0]["prim"] == order[1]


def test_backface_cull_and_nocull(oracle_mod):
    """oneSided && !nocull && dot(n, d) > 0 rejects (Primitives.h:173-174)."""
    from vistrace_b200 import abi

    tris = np.zeros(1, abi.TRI_IN)
    tris["p"][0] = [[-1, -1, 0], [1, -1, 0], [0, 1, 0]]
    tris["normals"], tris["tangents"], tris["alphas"], tris["one_sided"] = (0, 0, 1), (1, 0, 0), 1.0, 1
    tris["uvs"] = [[0, 0], [1, 0], [0, 1]]
    nodes = np.zeros(1, abi.NODE)
    nodes["bounds"], nodes["prim_count"] = (-1, 1, -1, 1, 0, 0), 1
    rays = np.zeros(2, abi.RAY)
    rays["o"], rays["d"], rays["tmax"] = [(0, 0, 1), (0, 0, -1)], [(0, 0, -1), (0, 0, 1)], FLT_MAX
    mats = abi.default_materials(1)
    cpu = oracle_mod.CpuScene(abi.SceneData(tris, mats), "port", build_bvh=False)
    cpu.set_bvh(nodes, np.zeros(1, np.uint64))
    hit = cpu.traverse(rays)["hits"]["prim"] != abi.VT_MISS
    assert hit.sum() == 1  # exactly one side is culled
    mats["flags"] = abi.VT_MATFLAG_NOCULL
    cpu = oracle_mod.CpuScene(abi.SceneData(tris, mats), "port", build_bvh=False)
    cpu.set_bvh(nodes, np.zeros(1, np.uint64))
    assert (cpu.traverse(rays)["hits"]["prim"] != abi.VT_MISS).all()


def test_port_matches_live_reference(oracle_mod):
    """With /root/reference compiled (oracle/_ref): restatement == reference on a fresh seeded scene, both hierarchies."""
    import vistrace_b200 as vt
    from vistrace_b200 import scenes

    if not oracle_mod.available("reference"):
        pytest.skip("oracle/_ref/libvt_ref.so not built (needs /root/reference)")
    scene = scenes.scene_foliage(n_cards=1500, tex_size=64, seed=31)
    rays = scenes.pinhole_rays(200, 120, (0, -48, 20), (0, 0, 8))
    ref = oracle_mod.CpuScene(scene, "reference")
    port = oracle_mod.CpuScene(scene, "port", build_bvh=False)
    np.testing.assert_array_equal(ref.tri_derived().view(np.uint32), port.tri_derived().view(np.uint32))
    for bvh in (ref.get_bvh(), vt.build_bvh(scene)):  # the reference's PLOC tree, then the product builder's tree
        ref.set_bvh(*bvh)
        port.set_bvh(*bvh)
        a = ref.traverse(rays, want_attrs=True, want_stats=True)
        b = port.traverse(rays, want_attrs=True, want_stats=True)
        assert a["hits"].tobytes() == b["hits"].tobytes() and a["attrs"].tobytes() == b["attrs"].tobytes()
        assert (a["steps"], a["isects"]) == (b["steps"], b["isects"])
    uvm = np.random.default_rng(1).uniform(-3, 3, (5000, 3)).astype(np.float32)
    for i in range(2):
        np.testing.assert_array_equal(ref.sample(i, uvm).view(np.uint32), port.sample(i, uvm).view(np.uint32))


def _skinned27(out):
    """{p0, e1, e2, normals, tangents} from skinned vt_tri_in records, with the constructor's e1 = p0 - p1, e2 = p2 - p0."""
    p = out["p"]
    return np.concatenate([p[:, 0], p[:, 0] - p[:, 1], p[:, 2] - p[:, 0], out["normals"].reshape(-1, 9), out["tangents"].reshape(-1, 9)], 1).astype(np.float32)


def test_skin_triangles_port_and_product_match_reference_golden(oracle_mod):
    """SkinTriangle (source/objects/AccelStruct.cpp:66-108): the committed outputs of the reference's own function vs the
    C restatement and the product's host code (vt_skin_triangles), bit for bit; live reference too when present."""
    import os

    import vistrace_b200 as vt
    from conftest import GOLDEN

    z = np.load(os.path.join(GOLDEN, "skin_small.npz"))
    for skin, bones, binds, key in ((z["skin"], z["bones"], z["binds"], "skinned"), (None, z["bones"][:1], z["binds"][:1], "skinned_one_bone")):
        want = z[key]
        port = oracle_mod.skin_triangles(z["tris"], skin, bones, binds, "port")
        np.testing.assert_array_equal(port.view(np.uint32), want.view(np.uint32))
        got = _skinned27(vt.skin_triangles(z["tris"], skin, bones, binds))
        np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
        if oracle_mod.available("reference"):
            live = oracle_mod.skin_triangles(z["tris"], skin, bones, binds, "reference")
            np.testing.assert_array_equal(live.view(np.uint32), want.view(np.uint32))
    # the transform really moved things, and a bad bone id is an error, not a crash
    assert np.abs(z["skinned"][:, :3] - z["tris"]["p"][:, 0]).max() > 1.0
    bad = z["skin"].copy()
    bad["bone_ids"][0, 0, 0] = 99
    with pytest.raises(RuntimeError, match="bone id"):
        vt.skin_triangles(z["tris"], bad, z["bones"], z["binds"])


def _material_case():
    from vistrace_b200 import scenes

    scene = scenes.scene_materials()
    rays = np.concatenate([scenes.pinhole_rays(160, 90, (0, -70, 30), (0, 0, 4)), scenes.random_rays(4000, (-40, -40, 0), (40, 40, 30), seed=8)])
    rng = np.random.default_rng(5)
    cones = np.stack([rng.uniform(0.0, 0.05, len(rays)), rng.uniform(1e-4, 0.02, len(rays))], -1).astype(np.float32)
    cones[::4] = -1.0  # every fourth ray without a cone: mip 0 (TraceResult.cpp:53)
    return scene, rays, cones


def test_port_matches_reference_on_every_shading_branch(oracle_mod):
    """Normal maps, vertex-transition blending, all twelve detail blend modes, MRAO, UV transforms and the cone-footprint
    texture LOD (source/objects/TraceResult.cpp:11-43, 89-253): restatement vs the reference's own TraceResult, bit for bit."""
    if not oracle_mod.available("reference"):
        pytest.skip("needs oracle/_ref (the compiled reference)")
    scene, rays, cones = _material_case()
    ref = oracle_mod.CpuScene(scene, "reference")
    port = oracle_mod.CpuScene(scene, "port", build_bvh=False)
    port.set_bvh(*ref.get_bvh())
    want = ref.traverse(rays, want_attrs=True)
    got = port.traverse(rays, want_attrs=True)
    assert got["hits"].tobytes() == want["hits"].tobytes() and got["attrs"].tobytes() == want["attrs"].tobytes()
    hit = want["hits"]["prim"] != 0xFFFFFFFF
    mats = scene.tris["material"][want["hits"]["prim"][hit]]
    assert set(np.unique(mats)) >= set([0, 2, 3, 4, 5] + list(range(6, 18)))  # every material is actually hit
    a = ref.trace_result(rays, want["hits"], cones=cones)
    b = port.trace_result(rays, want["hits"], cones=cones)
    assert a.tobytes() == b.tobytes()
    lod = a["base_mip"][hit & (cones[:, 0] >= 0)]
    assert (lod > 0).any() and not (a["albedo"] == want["attrs"]["albedo"]).all()  # the cones change the sampled mip
    assert len(np.unique(a["metalness"][hit])) > 10 and (a["flags"][hit] & 4).any()  # MRAO sampled, water hit


@pytest.mark.parametrize("kind", ["reference", "port"])
def test_cornell_box_golden_image_of_the_bvh_library(built, oracle_mod, kind):
    """libs/bvh/test/CMakeLists.txt:57-82: every builder of the library must reproduce scene/cornell_box_reference.png with the
    benchmark's camera (shading = |normalised triangle normal|).  Both checkers do, over the reference's PLOC + LeafCollapser tree
    (reference kind: built by the library itself; port: the product's bit-identical rebuild of it) and over the product's SAH tree."""
    import vistrace_b200 as vt
    from conftest import cornell_box

    if not oracle_mod.available(kind):
        pytest.skip(f"oracle kind {kind} not built")
    scene, rays, to_image, want = cornell_box()
    trees = [vt.build_bvh_ploc(scene), vt.build_bvh(scene)]
    if kind == "reference":
        trees.insert(0, None)  # the library's own build
    for tree in trees:
        cpu = oracle_mod.CpuScene(scene, kind, build_bvh=tree is None)
        if tree is not None:
            cpu.set_bvh(*tree)
        img = to_image(cpu.traverse(rays)["hits"])
        differing = int((img != want).any(-1).sum())
        assert differing <= 8, f"{differing} of {want.shape[0] * want.shape[1]} pixels differ from the golden image"


def test_port_equals_reference_on_random_hostile_inputs(built, oracle_mod):
    """Seeded fuzz of the C restatement against the compiled reference: zero-area, grid-aligned, duplicated and coplanar triangles,
    one- and two-sided; rays with zero / negative-zero / tiny / huge direction components, through vertices and along edges,
    finite and infinite intervals.  Hit records, TraceResult records (NaNs included) and both statistics counters: byte for byte."""
    import vistrace_b200 as vt

    if not (oracle_mod.available("reference") and oracle_mod.available("port")):
        pytest.skip("needs both checkers")
    from conftest import hostile_case

    rng = np.random.default_rng(99)
    for it in range(18):
        scene, rays, kind = hostile_case(it, rng)
        tree = vt.build_bvh_ploc(scene) if it % 2 else vt.build_bvh(scene)
        out = {}
        for k in ("reference", "port"):
            cpu = oracle_mod.CpuScene(scene, k, build_bvh=False)
            cpu.set_bvh(*tree)
            out[k] = cpu.traverse(rays, want_attrs=True, want_stats=True)
        a, b = out["reference"], out["port"]
        assert a["hits"].tobytes() == b["hits"].tobytes(), (it, kind)
        assert (a["steps"], a["isects"]) == (b["steps"], b["isects"]), (it, kind)
        assert a["attrs"].tobytes() == b["attrs"].tobytes(), (it, kind)


def _bsdf_case(n=20000, seed=3):
    """TraceResult-like records with every kind of frame: front / back facing, metallic 0 .. 1 (incl. exactly 1), rough 0 .. 1."""
    from vistrace_b200 import abi

    rng = np.random.default_rng(seed)
    f4 = np.float32
    a = np.zeros(n, abi.ATTR)
    nrm = rng.normal(size=(n, 3)).astype(f4)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True).astype(f4)
    t = np.cross(nrm, rng.normal(size=(n, 3))).astype(f4)
    t /= np.linalg.norm(t, axis=1, keepdims=True).astype(f4)
    a["normal"], a["tangent"], a["binormal"] = nrm, t, np.cross(t, nrm).astype(f4)
    a["albedo"] = rng.uniform(0, 1.2, (n, 3)).astype(f4)  # > 1 exercises the clamp of PrepShadingData
    a["metalness"] = rng.uniform(-0.1, 1.0, n).astype(f4)
    a["metalness"][::17] = 1.0  # pDiffuse = 0: SampleBSDF falls through every branch
    a["roughness"] = rng.uniform(-0.1, 1.1, n).astype(f4)
    wo = rng.normal(size=(n, 3)).astype(f4)
    wo /= np.linalg.norm(wo, axis=1, keepdims=True).astype(f4)
    rnd = rng.uniform(0, 1, (n, 3)).astype(f4)
    rnd[::13, 1] = 0.0  # r1 = 0: a sample in the tangent plane, pdf 0
    return a, wo, rnd


def test_bsdf_diffuse_port_equals_reference(oracle_mod):
    """The C port's restatement of SampleBSDF (diffuse lobe) against the reference's own function driven by a scripted ISampler."""
    if not oracle_mod.available("reference"):
        pytest.skip("needs oracle/_ref")
    a, wo, rnd = _bsdf_case()
    want, wret = oracle_mod.sample_bsdf_diffuse(a, wo, rnd, "reference")
    got, gret = oracle_mod.sample_bsdf_diffuse(a, wo, rnd, "port")
    assert (wret == 1).all() and (gret == 1).all()
    np.testing.assert_array_equal(got["lobe"], want["lobe"])
    assert (want["lobe"][::17] == 0).all() and (want["lobe"] == 1).sum() > 0.9 * len(a)
    for f in ("scattered", "pdf", "weight"):
        np.testing.assert_array_equal(got[f].view(np.uint32), want[f].view(np.uint32), err_msg=f)  # same libm, same order: bit-identical
