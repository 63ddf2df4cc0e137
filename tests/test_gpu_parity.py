"""-m gpu parity tests: the CUDA path, called through the C ABI, against the CPU oracle.

Bars (BASELINE.json north_star): hit/miss, primitive id and (t, u, v) BIT-EXACT; TraceResult
attributes within 1e-5 relative (they are bit-exact in practice; the tolerance covers sqrt/div
differences the spec allows).  Every parity test runs on both node layouts:
  * "exact":   the kernel walks the same nodes in the reference's order with the reference's
               arithmetic, so even exact ties must agree — whole hit buffers compare byte for byte;
  * "compact" / "quad": conservatively quantised binary / 4-wide nodes; a record may differ from the oracle's only as a TIE — both
               sides hit and |dt| <= 1e-6 |t| (north_star: "near-ties counted and reported") — and
               every other record must be bit-identical.  The synthetic scenes have no duplicate
               geometry, so in practice the tie count is 0 and the buffers are identical too.
"""
import os

import numpy as np
import pytest

from conftest import attr_max_rel_err, compare_hits, oracle_kinds

pytestmark = pytest.mark.gpu

FLT_MAX = np.finfo(np.float32).max
ATTR_FLOAT_FIELDS = ("pos", "distance", "normal", "alpha", "tangent", "metalness", "binormal", "roughness",
                     "geometric_normal", "base_mip", "albedo", "uvw", "tex_uv")
ATTR_INT_FIELDS = ("ent_id", "submat_idx", "flags", "prim")


@pytest.fixture(scope="module")
def vt(built):
    import vistrace_b200

    assert vistrace_b200.lib().vt_device_count() >= 1, "no CUDA device"
    return vistrace_b200


@pytest.fixture(params=["quad", "compact", "exact"])
def layout(request):
    return request.param


TIES = {"quad": 0, "compact": 0, "exact": 0}
LEAKS = {"quad": 0, "compact": 0, "exact": 0}


def same_hits(got, want, layout, rays=None, cpu=None, max_leaks=None):
    """Hit buffers agree: byte for byte ("exact"), or — quantised layouts — up to COUNTED exceptions of the two kinds
    vistrace_b200/report.py defines: exact ties (t bit-identical, the twin primitive) and verified reference leaks (the engine
    reports a CLOSER hit, or a hit where the reference traversal misses, and the checker's own triangle test accepts exactly
    that (t, u, v) for that primitive; only checked when the caller passes `rays` and a CPU checker).  Seen once in 5.6 M
    bounce rays of the closed 5 M-triangle scene (a ray through the shared edge of two wall triangles: the reference reports
    a MISS inside a closed room).

    Everything else fails: a FARTHER hit than the reference's (near-tie or not — a lost candidate, what conservative boxes
    rule out), a miss where the reference hits, t/u/v bits differing on the same primitive, an unverified closer hit."""
    if got.tobytes() == want.tobytes():
        return True
    if layout == "exact":
        return False
    from vistrace_b200.report import classify_hits

    rep = classify_hits(got, want, rays, cpu.tri_intersect if cpu is not None else None)
    TIES[layout] += rep["exact_tie"]
    LEAKS[layout] += rep["leak"]
    print(f"[parity] {rep['differing']} of {len(got)} records differ: {rep['exact_tie']} exact ties, {rep['leak']} verified reference leaks "
          f"({rep['near_leak']} near-ties, {rep['leak_vs_miss']} hit-vs-miss), {rep['lost']} lost, {rep['unverified']} unverified, {rep['tuv_bits']} t/u/v bits")
    if max_leaks is None:
        max_leaks = max(1, len(got) // 1000000)
    return rep["ok"] and rep["leak"] <= max_leaks


def _check_against(vt, oracle_mod, scene, rays, kind, bvh_from, layout, any_hit=False):
    """Run GPU and oracle over the SAME hierarchy and compare everything."""
    from vistrace_b200 import abi

    accel = vt.Accel(0, layout=layout)
    cpu = oracle_mod.CpuScene(scene, kind, build_bvh=(bvh_from == "reference"))
    if bvh_from == "reference":
        accel.populate(scene, bvh=cpu.get_bvh())  # the reference's own PLOC + LeafCollapser tree
    else:
        accel.populate(scene)  # product builder; hand the same tree to the oracle
        cpu.set_bvh(*accel.get_bvh())
    np.testing.assert_array_equal(accel.tri_derived().view(np.uint32), cpu.tri_derived().view(np.uint32))
    want = cpu.traverse(rays, want_attrs=True)
    if any_hit:
        hits = accel.traverse(rays, any_hit=True)
        np.testing.assert_array_equal(hits["prim"] == abi.VT_MISS, want["hits"]["prim"] == abi.VT_MISS)
        return accel, cpu, hits, None, want
    assert accel.layout == layout
    hits, attrs = accel.traverse(rays, want_attrs=True)
    rep = compare_hits(hits, want["hits"])
    assert rep["hit_miss_mismatch"] == 0 and rep["tuv_bit_mismatch"] == 0, rep
    assert rep["prim_mismatch"] == 0 or layout != "exact", rep
    assert same_hits(hits, want["hits"], layout, rays, cpu), rep
    err = attr_max_rel_err(attrs, want["attrs"])
    for f in ATTR_FLOAT_FIELDS:
        assert err[f] <= 1e-5, (f, err[f])  # tolerance from BASELINE.json north_star
    for f in ATTR_INT_FIELDS:
        assert err[f] == 0, (f, err[f])
    miss = hits["prim"] == abi.VT_MISS
    if miss.any():  # miss records are zero apart from prim = VT_MISS
        assert not attrs[miss].view(np.uint32).reshape(int(miss.sum()), 32)[:, :-1].any()
    return accel, cpu, hits, attrs, want


@pytest.mark.parametrize("kind", oracle_kinds())
@pytest.mark.parametrize("bvh_from", ["product", "reference"])
def test_config1_heightfield_primary_and_bounce(vt, oracle_mod, kind, bvh_from, layout):
    from vistrace_b200 import scenes

    if bvh_from == "reference" and kind != "reference":
        pytest.skip("the reference-built tree needs oracle/_ref")
    scene = scenes.scene_heightfield(96)
    rays = scenes.pinhole_rays(480, 270, (0, -80, 60), (0, 0, 5))
    accel, cpu, hits, attrs, want = _check_against(vt, oracle_mod, scene, rays, kind, bvh_from, layout)
    bounce, _ = scenes.bounce_rays(attrs, spp=2)
    assert len(bounce) > 1000
    got = accel.traverse(bounce)
    assert same_hits(got, cpu.traverse(bounce)["hits"], layout, bounce, cpu)


@pytest.mark.parametrize("kind", oracle_kinds())
def test_config2_props_primary_and_shadow(vt, oracle_mod, kind, layout):
    from vistrace_b200 import abi, scenes

    scene = scenes.scene_props_skinned(24, 31, 15, 24)  # props baked to world space by SkinTriangle (vt_skin_triangles)
    rays = scenes.pinhole_rays(480, 270, (0, -95, 40), (0, 0, 10))
    accel, cpu, hits, attrs, want = _check_against(vt, oracle_mod, scene, rays, kind, "product", layout)
    assert len(np.unique(attrs["ent_id"][hits["prim"] != abi.VT_MISS])) > 5  # several entities visible
    shadow, parent = scenes.shadow_rays(attrs)
    # the device-side generator (K3b) writes the same rays into their parents' slots, masked slots elsewhere
    sun = np.array((0.3, 0.2, 0.93), np.float32)
    sun = (sun / np.sqrt((sun * sun).sum(dtype=np.float32))).astype(np.float32)
    dev_rays, live = accel.shadow_rays(attrs, sun)
    assert live == len(shadow) and (dev_rays["tmax"] < 0).sum() == len(attrs) - live
    assert dev_rays[parent].tobytes() == shadow.tobytes()
    light = np.array((10.0, -20.0, 60.0), np.float32)
    pt_rays, _ = accel.shadow_rays(attrs, light, point_light=True)
    np.testing.assert_array_equal(pt_rays["d"][parent], (light[None, :] - shadow["o"]).astype(np.float32))
    assert (pt_rays["tmax"][parent] == 1.0).all()
    got = accel.traverse(shadow)
    assert same_hits(got, cpu.traverse(shadow)["hits"], layout, shadow, cpu)
    occl = accel.traverse(shadow, any_hit=True)  # early-out variant: only hit / no-hit is defined
    np.testing.assert_array_equal(occl["prim"] == abi.VT_MISS, got["prim"] == abi.VT_MISS)


@pytest.mark.parametrize("kind", oracle_kinds())
@pytest.mark.parametrize("bvh_from", ["product", "reference"])
def test_config4_foliage_alpha_test_and_attrs(vt, oracle_mod, kind, bvh_from, layout):
    from vistrace_b200 import abi, scenes

    if bvh_from == "reference" and kind != "reference":
        pytest.skip("the reference-built tree needs oracle/_ref")
    scene = scenes.scene_foliage(n_cards=6000, tex_size=128)
    rays = scenes.pinhole_rays(480, 270, (0, -48, 20), (0, 0, 8))
    accel, cpu, hits, attrs, want = _check_against(vt, oracle_mod, scene, rays, kind, bvh_from, layout)
    mats = scene.tris["material"][hits["prim"][hits["prim"] != abi.VT_MISS]]
    assert (mats >= 2).sum() > 1000  # alpha-tested cards are actually being hit ...
    assert (attrs["alpha"][hits["prim"] != abi.VT_MISS][mats >= 2] >= 0.5 - 1e-6).all()  # ... and only where opaque
    bounce, _ = scenes.bounce_rays(attrs, spp=1)
    assert same_hits(accel.traverse(bounce), cpu.traverse(bounce)["hits"], layout, bounce, cpu)


@pytest.mark.parametrize("kind", oracle_kinds())
def test_incoherent_rays_finite_tmax(vt, oracle_mod, kind, layout):
    from vistrace_b200 import scenes

    scene = scenes.scene_props(8, 31, 15, 16)
    rays = scenes.random_rays(60000, (-90, -90, -5), (90, 90, 70), seed=11)
    rays["tmax"][::3] = 25.0  # shadow-style finite interval
    rays["tmin"][::5] = 3.0
    rays["d"][::7] *= 3.5  # un-normalised directions: t is parametric (AccelStruct.cpp:810-815)
    _check_against(vt, oracle_mod, scene, rays, kind, "product", layout)


def test_edge_cases(vt, oracle_mod, layout):
    from vistrace_b200 import abi, scenes

    # (a) a scene small enough for the root to be a leaf (single_ray_traverser.hpp:72-73)
    tiny = abi.SceneData(scenes.box((-1, -1, -1), (1, 1, 1), inward=False)[:2])
    rays = scenes.random_rays(4096, (-3, -3, -3), (3, 3, 3), seed=3)
    accel, cpu, hits, attrs, want = _check_against(vt, oracle_mod, tiny, rays, "port", "product", layout)
    assert (hits["prim"] != abi.VT_MISS).any()
    # (b) empty batch
    assert len(accel.traverse(rays[:0])) == 0
    # (c) ragged batch sizes around the warp / CTA granularity
    for n in (1, 31, 32, 33, 127, 129, 1000):
        assert accel.traverse(rays[:n]).tobytes() == want["hits"][:n].tobytes()
    # (d) argument rules (AccelStruct.cpp:805-806): invalid rays become counted misses
    bad = rays[:64].copy()
    bad["tmin"][:10] = -1.0
    bad["tmax"][10:20] = 0.0
    bad["tmax"][20:30] = np.nan
    got = accel.traverse(bad)
    assert (got["prim"][:30] == abi.VT_MISS).all() and accel.invalid_rays == 30
    assert got[30:].tobytes() == want["hits"][30:64].tobytes()
    # (e) axis-parallel rays and signed zeros (the octant logic; libs/bvh/test/node_intersectors.cpp)
    scene = scenes.scene_heightfield(32)
    ax = np.zeros(6 * 500, abi.RAY)
    rng = np.random.default_rng(0)
    ax["o"] = rng.uniform(-40, 40, (len(ax), 3)).astype(np.float32)
    ax["o"][:, 2] = rng.uniform(12, 40, len(ax))
    dirs = np.array([[0, 0, -1], [0, -0.0, -1], [-0.0, 0, -1], [1, 0, 0], [0, 1, -0.0], [-1, -0.0, 0]], np.float32)
    ax["d"] = np.tile(dirs, (500, 1))
    ax["tmax"] = FLT_MAX
    _check_against(vt, oracle_mod, scene, ax, "port", "product", layout)
    # (f) un-normalised directions of extreme magnitude (t is parametric, AccelStruct.cpp:810-815): |d| ~ 2^70 and
    # 2^100 send a warp of the quad kernel down its two-fma plane form (slab_quad: |inv| < 2^-60), |d| ~ 2^-30 hits
    # safe_inverse's clamp (vector.hpp:69-74); mixed into ordinary rays so both forms run side by side
    wild = scenes.pinhole_rays(96, 54, (0, -80, 60), (0, 0, 5))
    scale = np.ones(len(wild), np.float32)
    scale[::7] = 2.0**70
    scale[3::11] = 2.0**100
    scale[5::13] = 2.0**-30
    wild["d"] *= scale[:, None]
    _check_against(vt, oracle_mod, scene, wild, "port", "product", layout)
    wild["d"][::5, 0] = np.float32(2.0**90)  # one huge component only
    _check_against(vt, oracle_mod, scene, wild, "port", "product", layout)


def test_exact_ties_duplicate_geometry(vt, oracle_mod, layout):
    """Every triangle stored twice: each hit is an exact tie between two candidates.  The reference lets the
    LATER tested candidate win (`t <= tmax`, single_ray_traverser.hpp:55-60).  The exact layout reproduces
    the winner; the compact layout reproduces t/u/v bit for bit and may report the twin."""
    from vistrace_b200 import abi, scenes

    base = scenes.scene_heightfield(40)
    n = base.n_tris
    scene = abi.SceneData(np.concatenate([base.tris, base.tris]), base.materials, base.entities)
    rays = scenes.pinhole_rays(320, 180, (0, -80, 60), (0, 0, 5))
    accel = vt.Accel(0, layout=layout).populate(scene)
    cpu = oracle_mod.CpuScene(scene, "port", build_bvh=False)
    cpu.set_bvh(*accel.get_bvh())
    want = cpu.traverse(rays)["hits"]
    got = accel.traverse(rays)
    assert (got["prim"] != abi.VT_MISS).all()
    for f in ("t", "u", "v"):
        np.testing.assert_array_equal(got[f].view(np.uint32), want[f].view(np.uint32))
    np.testing.assert_array_equal(got["prim"] % n, want["prim"] % n)  # the same triangle or its twin
    if layout == "exact":
        assert got.tobytes() == want.tobytes()
    else:
        print(f"[parity] {layout} layout: {int((got['prim'] != want['prim']).sum())} of {len(rays)} exact ties resolved to the twin")


def test_traversal_statistics_match_the_reference_counters(vt, oracle_mod):
    """SingleRayTraverser::Statistics (single_ray_traverser.hpp:132-135): on the exact layout the kernel takes
    exactly the reference's traversal steps and runs exactly its primitive intersections; the compact layout
    runs a superset of the steps (conservative boxes), only slightly larger."""
    from vistrace_b200 import scenes

    scene = scenes.scene_props(12, 31, 15, 16)
    rays = np.concatenate([scenes.pinhole_rays(320, 180, (0, -95, 40), (0, 0, 10)), scenes.random_rays(30000, (-90, -90, -5), (90, 90, 70), seed=5)])
    exact = vt.Accel(0, layout="exact").populate(scene)
    cpu = oracle_mod.CpuScene(scene, "port", build_bvh=False)
    cpu.set_bvh(*exact.get_bvh())
    want = cpu.traverse(rays, want_stats=True)
    assert exact.traverse_stats(rays) == (want["steps"], want["isects"])
    compact = vt.Accel(0, layout="compact").populate(scene, bvh=exact.get_bvh())
    steps, tests = compact.traverse_stats(rays)
    assert want["steps"] <= steps <= 1.10 * want["steps"] and 0.95 * want["isects"] <= tests <= 1.10 * want["isects"]
    quad = vt.Accel(0, layout="quad").populate(scene, bvh=exact.get_bvh())
    qsteps, qtests = quad.traverse_stats(rays)
    assert 0.3 * want["steps"] <= qsteps <= 0.75 * want["steps"] and 0.9 * want["isects"] <= qtests <= 1.3 * want["isects"]
    print(f"[stats] exact {want['steps']} steps / {want['isects']} tests; compact {steps} / {tests}; quad {qsteps} / {qtests}")


@pytest.mark.parametrize("kind", oracle_kinds())
def test_every_shading_branch_and_cone_lod(vt, oracle_mod, kind):
    """K2 on a scene that reaches every TraceResult branch (source/objects/TraceResult.cpp:11-43, 89-253): one and two
    normal maps, vertex-transition blending (smoothstep and masked), second base texture, the twelve detail blend
    modes, MRAO, UV transforms / texScale, water — without cones (mip 0) and with per-ray cones (trilinear LOD from
    CalcFootprint + TriUVInfoToTexLOD, source/Utils.h:75-78)."""
    from test_oracle import _material_case
    from vistrace_b200 import abi

    scene, rays, cones = _material_case()
    accel = vt.Accel(0).populate(scene)
    cpu = oracle_mod.CpuScene(scene, kind, build_bvh=False)
    cpu.set_bvh(*accel.get_bvh())
    want = cpu.traverse(rays, want_attrs=True)
    hits, attrs = accel.traverse(rays, want_attrs=True)
    assert same_hits(hits, want["hits"], accel.layout)
    for got_attrs, want_attrs in ((attrs, want["attrs"]), (accel.traverse(rays, want_attrs=True, cones=cones)[1], cpu.trace_result(rays, hits, cones=cones))):
        err = attr_max_rel_err(got_attrs, want_attrs)
        for f in ATTR_FLOAT_FIELDS:
            assert err[f] <= 1e-5, (f, err[f])
        for f in ATTR_INT_FIELDS:
            assert err[f] == 0, (f, err[f])
    hit = hits["prim"] != abi.VT_MISS
    assert len(np.unique(scene.tris["material"][hits["prim"][hit]])) >= 17


def test_trace_result_stage_alone(vt, oracle_mod):
    """K2 on its own: attributes for hits produced elsewhere (here: by the oracle)."""
    from vistrace_b200 import scenes

    scene = scenes.scene_foliage(n_cards=3000, tex_size=64)
    rays = scenes.pinhole_rays(320, 180, (0, -48, 20), (0, 0, 8))
    accel = vt.Accel(0).populate(scene)
    cpu = oracle_mod.CpuScene(scene, "port", build_bvh=False)
    cpu.set_bvh(*accel.get_bvh())
    want = cpu.traverse(rays, want_attrs=True)
    attrs = accel.trace_result(rays, want["hits"])
    err = attr_max_rel_err(attrs, want["attrs"])
    for f in ATTR_FLOAT_FIELDS:
        assert err[f] <= 1e-5, (f, err[f])


def test_single_ray_traverse_surface(vt, oracle_mod):
    """Host buffers, n = 1: what the Lua-facing accel:Traverse does per call."""
    from vistrace_b200 import abi, scenes

    scene = scenes.scene_heightfield(32)
    accel = vt.Accel(0).populate(scene)
    ray = np.zeros(1, abi.RAY)
    ray["o"], ray["d"], ray["tmax"] = (0, 0, 50), (0, 0, -1), FLT_MAX
    hit, attr = accel.traverse(ray, want_attrs=True)
    assert hit["prim"][0] != abi.VT_MISS and attr["flags"][0] & abi.VT_ATTR_FRONT_FACING
    ray["d"] = (0, 0, 1)
    hit, attr = accel.traverse(ray, want_attrs=True)
    assert attr["flags"][0] & abi.VT_ATTR_HIT_SKY  # the room's ceiling is a sky brush: a hit, not a miss


def test_large_batch_properties(vt):
    """BASELINE-size batch (1920x1080) checked through size-independent properties: every ray of a
    closed room hits, t reproduces the hit position, and splitting the batch changes nothing."""
    from vistrace_b200 import abi, scenes

    scene = scenes.scene_heightfield(224)
    rays = scenes.pinhole_rays(1920, 1080, (0, -80, 60), (0, 0, 5))
    accel = vt.Accel(0).populate(scene)
    hits, attrs = accel.traverse(rays, want_attrs=True)
    assert (hits["prim"] != abi.VT_MISS).all()
    pos_from_t = rays["o"] + rays["d"] * hits["t"][:, None]
    assert np.abs(pos_from_t - attrs["pos"]).max() < 2e-2
    assert np.abs(attrs["uvw"].sum(-1) - 1).max() < 1e-5
    half = len(rays) // 2
    again = np.concatenate([accel.traverse(rays[:half]), accel.traverse(rays[half:])])
    assert again.tobytes() == hits.tobytes()


def test_full_size_scene_layouts_agree(vt, oracle_mod):
    """BASELINE config 3 at full size (5 005 460 triangles, 1920x1080 primary + 4 spp bounce rays, ~7.7 M rays): the
    quantised layouts against the exact layout (which the tests above pin to the oracle bit for bit).  Records may
    differ only as ties — the axis-aligned room has edges where two triangles give the same t — or as verified
    reference leaks (same_hits); both are counted."""
    from vistrace_b200 import abi, scenes

    scene = scenes.scene_terrain_closed(1582)
    rays = scenes.pinhole_rays(1920, 1080, (0.0, -330.0, 200.0), (0.0, 0.0, 10.0))
    exact = vt.Accel(0, layout="exact").populate(scene)
    bvh = exact.get_bvh()
    want = exact.trace_diffuse_wave(rays, 4, seed=11, want_bounce_rays=True)
    assert (want["hits"]["prim"] != abi.VT_MISS).all()  # closed scene
    live = want["bounce_rays"]["tmax"] >= 0
    exact.close()
    cpu = oracle_mod.CpuScene(scene, "reference" if oracle_mod.available("reference") else "port", build_bvh=False)  # triangle test only
    bounce = want["bounce_rays"][live]
    for layout in ("quad", "compact"):
        accel = vt.Accel(0, layout=layout).populate(scene, bvh=bvh)
        assert accel.layout == layout
        assert same_hits(accel.traverse(rays), want["hits"], layout, rays, cpu)
        assert same_hits(accel.traverse(bounce), want["bounce_hits"][live], layout, bounce, cpu)
        accel.close()


@pytest.mark.parametrize("name", ["foliage_small", "props_small"])
def test_cuda_matches_reference_golden_vectors(vt, name, layout):
    """The committed fixtures are outputs of the UNMODIFIED reference (tests/golden/make_golden.py): its own
    PLOC + LeafCollapser tree, its hits, its TraceResult values.  The CUDA path must reproduce them."""
    from conftest import load_golden

    scene, z = load_golden(name)
    accel = vt.Accel(0, layout=layout).populate(scene, bvh=(z["nodes"], z["prim_indices"]))
    assert accel.layout == layout
    np.testing.assert_array_equal(accel.tri_derived().view(np.uint32), z["tri_derived"].view(np.uint32))
    for rays_k, hits_k, attrs_k in (("rays", "hits", "attrs"), ("bounce_rays", "bounce_hits", "bounce_attrs"), ("extra_rays", "extra_hits", "extra_attrs")):
        if rays_k not in z:
            continue
        hits, attrs = accel.traverse(z[rays_k], want_attrs=True)
        assert same_hits(hits, z[hits_k], layout), (name, rays_k)
        err = attr_max_rel_err(attrs, z[attrs_k])
        for f in ATTR_FLOAT_FIELDS:
            assert err[f] <= 1e-5, (name, rays_k, f, err[f])
        for f in ATTR_INT_FIELDS:
            assert err[f] == 0, (name, rays_k, f)


def test_bounce_ray_generation_and_diffuse_wave(vt, oracle_mod, layout):
    """K3 + the one-call wave: the generated rays follow hemisphere_cos / CalcRayOrigin, the wave equals the
    piecewise calls, tiling does not change a bit, and the oracle agrees on every generated ray."""
    import os

    from vistrace_b200 import abi, scenes

    scene = scenes.scene_heightfield(64)
    rays = scenes.pinhole_rays(320, 180, (0, -80, 60), (0, 0, 5))
    accel = vt.Accel(0, layout=layout).populate(scene)
    hits, attrs = accel.traverse(rays, want_attrs=True)
    spp = 3
    brays, live = accel.bounce_rays(attrs, spp, seed=42)
    masked = brays["tmax"] < 0
    spawn = (attrs["prim"] != abi.VT_MISS) & ((attrs["flags"] & abi.VT_ATTR_HIT_SKY) == 0)
    np.testing.assert_array_equal(~masked, np.repeat(spawn, spp))
    assert live == int((~masked).sum()) and 0 < live < len(brays)
    a = np.repeat(attrs, spp)[~masked]
    b = brays[~masked]
    sgn = np.where((a["flags"] & abi.VT_ATTR_FRONT_FACING) != 0, 1.0, -1.0)[:, None].astype(np.float32)
    norm = np.linalg.norm(b["d"], axis=1)  # interpolated T/B/N are unit but not exactly orthogonal: |d| is only ~1
    assert 0.6 < norm.min() and norm.max() < 1.4
    assert ((b["d"] * a["normal"] * sgn).sum(-1) >= -0.35 * norm).all()               # viewer-side hemisphere (up to the TBN skew)
    np.testing.assert_array_equal(b["o"], scenes.calc_ray_origin(a["pos"], a["geometric_normal"] * sgn))  # CalcRayOrigin, bit-exact
    cosines = (b["d"] * a["normal"] * sgn).sum(-1) / norm
    assert abs(cosines.mean() - 2.0 / 3.0) < 0.03                                     # cosine-weighted: E[cos] = 2/3
    # traversal of the generated rays is bit-exact against the oracle
    cpu = oracle_mod.CpuScene(scene, "port", build_bvh=False)
    cpu.set_bvh(*accel.get_bvh())
    bhits = accel.traverse(brays)
    want = cpu.traverse(b)["hits"]
    assert same_hits(bhits[~masked], want, layout)
    assert (bhits["prim"][masked] == abi.VT_MISS).all() and accel.invalid_rays == 0   # masked slots: silent misses
    # the one-call wave, whole and tiled
    whole = accel.trace_diffuse_wave(rays, spp, seed=42, want_attrs=True, want_bounce_rays=True)
    os.environ["VT_WAVE_TILE"] = "5000"
    tiled = accel.trace_diffuse_wave(rays, spp, seed=42, want_attrs=True, want_bounce_rays=True)
    del os.environ["VT_WAVE_TILE"]
    for k, ref in (("hits", hits), ("attrs", attrs), ("bounce_rays", brays), ("bounce_hits", bhits)):
        assert whole[k].tobytes() == ref.tobytes(), k
        assert tiled[k].tobytes() == ref.tobytes(), k
    assert whole["live_bounce"] == live == tiled["live_bounce"]


def test_accumulate_sky_framebuffer(vt):
    import torch

    from vistrace_b200 import abi, scenes

    scene = scenes.scene_heightfield(48)
    rays = scenes.pinhole_rays(160, 90, (0, -80, 60), (0, 0, 5))
    accel = vt.Accel(0).populate(scene)
    spp = 4
    w = accel.trace_diffuse_wave(rays, spp, seed=3, want_attrs=True)
    to_dev = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).cuda()
    d_attrs, d_bh = to_dev(w["attrs"]), to_dev(w["bounce_hits"])
    fb = torch.zeros(len(rays) * 3, dtype=torch.float32, device="cuda")
    accel.accumulate_sky_device(d_attrs.data_ptr(), d_bh.data_ptr(), len(rays), spp, 0.5, fb.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = fb.cpu().numpy().reshape(-1, 3)
    a, bh = w["attrs"], w["bounce_hits"].reshape(-1, spp)
    sky_mat = (scene.materials["surf_flags"] & abi.VT_SURF_SKY) != 0
    esc = (bh["prim"] == abi.VT_MISS) | sky_mat[scene.tris["material"][np.minimum(bh["prim"], scene.n_tris - 1)]]
    vis = esc.sum(1).astype(np.float32) / np.float32(spp)
    is_sky = (a["flags"] & abi.VT_ATTR_HIT_SKY) != 0
    want = np.where(is_sky[:, None], a["albedo"], a["albedo"] * vis[:, None]) * np.float32(0.5)
    want[a["prim"] == abi.VT_MISS] = 0
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-7)
    # the one-call render (host rays in, host RGBFFF image out), whole and tiled, is the same image bit for bit
    import os

    img, live = accel.render_diffuse_wave(rays, spp, seed=3, weight=0.5)
    assert live == spp * int((~is_sky & (a["prim"] != abi.VT_MISS)).sum())
    np.testing.assert_array_equal(img, got)
    os.environ["VT_WAVE_TILE"] = "3000"
    img2, _ = accel.render_diffuse_wave(rays, spp, seed=3, weight=0.5)
    del os.environ["VT_WAVE_TILE"]
    np.testing.assert_array_equal(img2, got)


def test_ray_queue_generators_and_queued_traversal(vt, layout):
    """Ray queue: the queued generators list exactly the slots they filled, write the miss record of every masked slot, and
    the queued traversal (closest hit after bounce rays, any hit after shadow rays) leaves the same hit buffer, byte for
    byte, as tracing every slot — including when the hit buffer starts out as garbage."""
    import torch

    from vistrace_b200 import abi, scenes

    scene = scenes.scene_heightfield(64)
    rays = scenes.pinhole_rays(320, 180, (0, -80, 60), (0, 0, 5))
    accel = vt.Accel(0, layout=layout).populate(scene)
    hits, attrs = accel.traverse(rays, want_attrs=True)
    n, spp = len(rays), 3
    spawn = (attrs["prim"] != abi.VT_MISS) & ((attrs["flags"] & abi.VT_ATTR_HIT_SKY) == 0)
    assert 0 < spawn.sum() < n  # the open height field leaves sky pixels: masked slots exist
    to_dev = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).cuda()
    sh = torch.cuda.current_stream().cuda_stream
    d_attrs = to_dev(attrs)
    # --- bounce rays, closest hit
    brays, live = accel.bounce_rays(attrs, spp, seed=42)
    want_hits = accel.traverse(brays)
    d_brays = torch.zeros(n * spp * 32, dtype=torch.uint8, device="cuda")
    d_bhits = torch.full((n * spp * 16,), 0xAB, dtype=torch.uint8, device="cuda")  # garbage: every slot must be written
    d_queue = torch.full((n * spp,), -1, dtype=torch.int32, device="cuda")
    d_count = torch.full((1,), 12345, dtype=torch.int64, device="cuda")            # the call zeroes it
    accel.bounce_rays_queued_device(d_attrs.data_ptr(), n, spp, 42, d_brays.data_ptr(), d_queue.data_ptr(), d_count.data_ptr(),
                                    d_bhits.data_ptr(), stream=sh)
    accel.traverse_queued_device(d_brays.data_ptr(), d_queue.data_ptr(), d_count.data_ptr(), n * spp, d_bhits.data_ptr(), stream=sh)
    torch.cuda.synchronize()
    assert int(d_count.item()) == live
    queue = d_queue.cpu().numpy().view(np.uint32)
    np.testing.assert_array_equal(np.sort(queue[:live]), np.flatnonzero(np.repeat(spawn, spp)))
    assert (queue[live:] == 0xFFFFFFFF).all()
    assert d_brays.cpu().numpy().tobytes() == brays.tobytes()
    assert d_bhits.cpu().numpy().tobytes() == want_hits.tobytes()
    # --- shadow rays, any hit (only hit / miss is defined for an occlusion query)
    sun = (0.3, -0.2, 0.9)
    srays, slive = accel.shadow_rays(attrs, sun)
    want_occ = accel.traverse(srays, any_hit=True)["prim"] != abi.VT_MISS
    d_srays = torch.zeros(n * 32, dtype=torch.uint8, device="cuda")
    d_shits = torch.full((n * 16,), 0xCD, dtype=torch.uint8, device="cuda")
    accel.shadow_rays_queued_device(d_attrs.data_ptr(), n, sun, d_srays.data_ptr(), d_queue.data_ptr(), d_count.data_ptr(),
                                    d_shits.data_ptr(), stream=sh)
    accel.traverse_queued_device(d_srays.data_ptr(), d_queue.data_ptr(), d_count.data_ptr(), n, d_shits.data_ptr(), any_hit=True, stream=sh)
    torch.cuda.synchronize()
    assert int(d_count.item()) == slive == int(spawn.sum())
    assert d_srays.cpu().numpy().tobytes() == srays.tobytes()
    got_occ = np.frombuffer(d_shits.cpu().numpy().tobytes(), abi.HIT)["prim"] != abi.VT_MISS
    np.testing.assert_array_equal(got_occ, want_occ)
    assert accel.invalid_rays == 0


@pytest.mark.parametrize("kind", oracle_kinds())
@pytest.mark.parametrize("path", ["device", "host"])
def test_refit_moved_props(vt, oracle_mod, kind, layout, path, monkeypatch):
    """accel:Rebuild through vt_accel_refit: props move, the hierarchy keeps its structure, boxes are refitted — by K5 on
    the device for the quad layout (path "device"), on the host otherwise.  The GPU result on the refitted tree equals the
    checker's on the SAME refitted tree (its own refit: bvh::HierarchyRefitter for the reference kind) and — up to counted
    ties — a from-scratch build of the moved scene; the device-refitted quads prune like the host-derived ones."""
    from test_host import _moved_props
    from vistrace_b200 import abi, scenes

    if path == "device" and layout != "quad":
        pytest.skip("K5 refits the quad layout; the other layouts take the host path")
    monkeypatch.setenv("VT_REFIT_DEVICE", "1" if path == "device" else "0")
    scene = scenes.scene_props(8, 21, 11, 12)
    moved = _moved_props(scene)
    rays = np.concatenate([scenes.pinhole_rays(320, 180, (0, -95, 40), (0, 0, 10)), scenes.random_rays(20000, (-90, -90, -5), (90, 90, 60), seed=12)])
    accel = vt.Accel(0, layout=layout).populate(scene)
    before = accel.traverse(rays)
    nodes0, prims0 = accel.get_bvh()
    accel.refit(moved)
    assert accel.layout == layout
    nodes1, prims1 = accel.get_bvh()
    assert (nodes1["first"] == nodes0["first"]).all() and (prims1 == prims0).all() and (nodes1["bounds"] != nodes0["bounds"]).any()
    cpu = oracle_mod.CpuScene(scene, kind, build_bvh=False)
    cpu.set_bvh(nodes0, prims0)
    cpu.refit(moved)
    assert np.array_equal(cpu.get_bvh()[0]["bounds"], nodes1["bounds"])
    np.testing.assert_array_equal(accel.tri_derived().view(np.uint32), cpu.tri_derived().view(np.uint32))
    hits, attrs = accel.traverse(rays, want_attrs=True)
    want = cpu.traverse(rays, want_attrs=True)
    assert same_hits(hits, want["hits"], layout, rays, cpu)
    err = attr_max_rel_err(attrs, want["attrs"])
    assert max(err[f] for f in ATTR_FLOAT_FIELDS) <= 1e-5
    assert (hits["prim"] != before["prim"]).sum() + (hits["t"] != before["t"]).sum() > 100  # the props really moved
    if path == "device":  # second refit on the prepared state (back to the start), then forward again: same records
        accel.refit(scene)
        assert same_hits(accel.traverse(rays), before, layout)
        accel.refit(moved)
        assert accel.traverse(rays).tobytes() == hits.tobytes()
        monkeypatch.setenv("VT_REFIT_DEVICE", "0")
        host = vt.Accel(0, layout=layout).populate(scene).refit(moved)
        assert same_hits(host.traverse(rays), hits, layout)
        s_dev, s_host = accel.traverse_stats(rays), host.traverse_stats(rays)
        assert abs(s_dev[0] - s_host[0]) <= 0.03 * s_host[0], (s_dev, s_host)  # same pruning power: node visits within 3 %
    fresh = vt.Accel(0, layout=layout).populate(moved).traverse(rays)
    rep = compare_hits(hits, fresh)
    # two different trees over the same triangles: same answer but for exact ties / a verified reference leak (same_hits)
    assert rep["hit_miss_mismatch"] <= 1 and rep["tuv_bit_mismatch"] == 0 and rep["prim_mismatch"] <= 3, rep
    with pytest.raises(RuntimeError):
        accel.refit(abi.SceneData(moved.tris[:-1], moved.materials, moved.entities))


def test_hostile_random_inputs(vt, oracle_mod, layout):
    """The seeded hostile scenes of tests/test_oracle.py through the CUDA path: zero-area, grid-aligned, duplicated and coplanar
    triangles; rays with zero / -0.0 / tiny / huge direction components, through vertices and along edges.  Exact layout: the
    checker's hit buffer byte for byte; quantised layouts: up to counted ties / verified reference leaks (same_hits)."""
    from conftest import hostile_case
    from vistrace_b200 import abi

    kind_name = "reference" if oracle_mod.available("reference") else "port"
    rng = np.random.default_rng(99)
    for it in range(12):
        scene, rays, kind = hostile_case(it, rng)
        tree = vt.build_bvh_ploc(scene) if it % 2 else vt.build_bvh(scene)
        accel = vt.Accel(0, layout=layout).populate(scene, bvh=tree)
        cpu = oracle_mod.CpuScene(scene, kind_name, build_bvh=False)
        cpu.set_bvh(*tree)
        got, want = accel.traverse(rays), cpu.traverse(rays)["hits"]
        if accel.layout == "exact":  # also where a quantised layout fell back to exact (e.g. a root leaf it cannot hold)
            assert got.tobytes() == want.tobytes(), (it, kind)
        else:
            rep = compare_hits(got, want)
            assert rep["tuv_bit_mismatch"] == 0, (it, kind, rep)
            # every differing record is an exact tie, or a candidate the checker's own triangle test accepts with exactly these
            # (t, u, v) but its traverser never reached (a box test rounded the ray out: see same_hits) — any number of them here,
            # the inputs are built to sit on edges and vertices; never a farther hit, never a lost one
            assert same_hits(got, want, accel.layout, rays, cpu, max_leaks=len(rays)), (it, kind, rep)
        accel.close()


@pytest.mark.parametrize("builder", ["product", "ploc"])
def test_cornell_box_golden_image_of_the_bvh_library(vt, layout, builder, monkeypatch):
    """The bvh library's golden-image test (libs/bvh/test/CMakeLists.txt:57-82: every builder must reproduce
    scene/cornell_box_reference.png with the benchmark's camera) through the CUDA path, on every node layout, over the product's
    tree and over the rebuilt PLOC + LeafCollapser tree.  tests/test_oracle.py holds the same check for the CPU checkers."""
    from conftest import cornell_box

    scene, rays, to_image, want = cornell_box()
    if builder == "ploc":
        monkeypatch.setenv("VT_BUILDER", "ploc")
    accel = vt.Accel(0, layout=layout).populate(scene)
    assert accel.layout == layout
    img = to_image(accel.traverse(rays))
    differing = int((img != want).any(-1).sum())
    assert differing == 0, f"{differing} of {want.shape[0] * want.shape[1]} pixels differ from the golden image"


@pytest.mark.parametrize("scene_name", ["props", "duplicates"])
def test_ploc_builder_gives_the_reference_answers_ties_included(vt, oracle_mod, scene_name, monkeypatch):
    """VT_BUILDER=ploc + exact layout: the engine builds the reference's own hierarchy (vt_bvh_ploc.cpp) and, with the reference's
    visit order, returns its hit buffer byte for byte — exact ties between duplicated triangles included — while the
    reference runs on the tree IT built; nothing is handed over."""
    from vistrace_b200 import abi, scenes

    if not oracle_mod.available("reference"):
        pytest.skip("needs oracle/_ref")
    base = scenes.scene_props(6, 15, 9, 8)
    scene = base if scene_name == "props" else abi.SceneData(np.concatenate([base.tris, base.tris]), base.materials, base.entities)
    rays = np.concatenate([scenes.pinhole_rays(320, 180, (0, -95, 40), (0, 0, 10)), scenes.random_rays(20000, (-90, -90, -5), (90, 90, 60), seed=3)])
    monkeypatch.setenv("VT_BUILDER", "ploc")
    accel = vt.Accel(0, layout="exact").populate(scene)
    cpu = oracle_mod.CpuScene(scene, "reference", build_bvh=True)
    want_nodes, want_prims = cpu.get_bvh()
    nodes, prims = accel.get_bvh()
    assert nodes.tobytes() == want_nodes.tobytes() and prims.tobytes() == want_prims.tobytes()
    hits, attrs = accel.traverse(rays, want_attrs=True)
    want = cpu.traverse(rays, want_attrs=True)
    assert hits.tobytes() == want["hits"].tobytes()
    err = attr_max_rel_err(attrs, want["attrs"])
    assert max(err[f] for f in ATTR_FLOAT_FIELDS) <= 1e-5
    quad = vt.Accel(0, layout="quad").populate(scene)  # the quantised layout over the same tree: same records up to counted ties
    assert same_hits(quad.traverse(rays), want["hits"], "quad", rays, cpu)


@pytest.mark.parametrize("scene_name", ["foliage", "props", "duplicates"])
def test_child_order_and_collapse_rule_do_not_change_the_records(vt, oracle_mod, scene_name, monkeypatch):
    """The quad kernel orders a node's children by entry distance or by entry + exit (VT_KEY_ORDER, chosen per scene from the
    sibling overlap of its hierarchy) and the wide nodes come from the SAH-optimal or the largest-child collapse (VT_COLLAPSE):
    order and grouping only prune — with the canonical tie rule every combination returns the same hit buffer byte for byte,
    duplicated geometry included, and that buffer agrees with the checker (which order prunes better depends on the scene:
    profiles/r2_child_order.md)."""
    from vistrace_b200 import abi, scenes

    base = scenes.scene_foliage(n_cards=6000, tex_size=64, ground_quads=16) if scene_name == "foliage" else scenes.scene_props(6, 15, 9, 8)
    scene = abi.SceneData(np.concatenate([base.tris, base.tris]), base.materials, base.entities) if scene_name == "duplicates" else base
    eye = (0, -48, 20) if scene_name == "foliage" else (0, -95, 40)
    rays = np.concatenate([scenes.pinhole_rays(320, 180, eye, (0, 0, 8)), scenes.random_rays(20000, (-45, -45, 0), (45, 45, 30), seed=6)])
    got, visits = {}, {}
    for collapse in ("dp", "greedy"):
        for order in ("entry", "mid"):
            monkeypatch.setenv("VT_COLLAPSE", collapse)
            monkeypatch.setenv("VT_KEY_ORDER", order)
            accel = vt.Accel(0, layout="quad").populate(scene)
            assert accel.layout == "quad"
            got[collapse, order] = accel.traverse(rays)
            visits[collapse, order] = accel.traverse_stats(rays)
            any_hit = accel.traverse(rays, any_hit=True)
            np.testing.assert_array_equal(any_hit["prim"] == abi.VT_MISS, got[collapse, order]["prim"] == abi.VT_MISS)
    first = got["dp", "entry"]
    for key, hits in got.items():
        assert hits.tobytes() == first.tobytes(), key
    kind = "reference" if oracle_mod.available("reference") else "port"
    cpu = oracle_mod.CpuScene(scene, kind, build_bvh=False)
    cpu.set_bvh(*accel.get_bvh())
    assert same_hits(first, cpu.traverse(rays)["hits"], "quad", rays, cpu)
    assert len({v for v in visits.values()}) > 1  # the knobs did change how the rays walk, not what they find
    monkeypatch.delenv("VT_COLLAPSE")
    monkeypatch.delenv("VT_KEY_ORDER")
    auto = vt.Accel(0, layout="quad").populate(scene)  # auto: whatever it picks, the same records
    assert auto.traverse(rays).tobytes() == first.tobytes()


@pytest.mark.parametrize("scene_name", ["props", "foliage"])
def test_reinsertion_optimised_tree_gives_the_checkers_answers(vt, oracle_mod, scene_name, monkeypatch):
    """VT_REINSERT (builder-quality option, vt_bvh_reinsert.cpp): vt_accel_populate optimises the product builder's tree before it is
    flattened; the engine over that tree equals the checker over the SAME tree — byte for byte on the exact layout, up to counted
    ties on the quantised ones — and the plain build's hit records up to exact ties; on separate objects it takes fewer node visits."""
    from vistrace_b200 import scenes

    scene = scenes.scene_props(8, 21, 11, 12) if scene_name == "props" else scenes.scene_foliage(n_cards=3000, tex_size=64, ground_quads=16)
    eye = (0, -95, 40) if scene_name == "props" else (0, -48, 20)
    rays = np.concatenate([scenes.pinhole_rays(320, 180, eye, (0, 0, 10)), scenes.random_rays(20000, (-90, -90, -5), (90, 90, 60), seed=4)])
    plain = vt.Accel(0, layout="quad").populate(scene)
    monkeypatch.setenv("VT_REINSERT", "3")
    monkeypatch.setenv("VT_REINSERT_FRACTION", "0.3")
    kind = "reference" if oracle_mod.available("reference") else "port"
    cpu = oracle_mod.CpuScene(scene, kind, build_bvh=False)
    want = None
    for layout in ("exact", "compact", "quad"):
        accel = vt.Accel(0, layout=layout).populate(scene)
        assert accel.layout == layout
        nodes, prims = accel.get_bvh()
        if want is None:
            assert nodes.tobytes() != plain.get_bvh()[0].tobytes()  # the pass did move nodes
            cpu.set_bvh(nodes, prims)
            want = cpu.traverse(rays, want_attrs=True)
        hits, attrs = accel.traverse(rays, want_attrs=True)
        if layout == "exact":
            assert hits.tobytes() == want["hits"].tobytes()
            err = attr_max_rel_err(attrs, want["attrs"])
            assert max(err[f] for f in ATTR_FLOAT_FIELDS) <= 1e-5
        else:
            assert same_hits(hits, want["hits"], layout, rays, cpu)
    assert same_hits(plain.traverse(rays), want["hits"], "quad", rays, cpu)
    if scene_name == "props":  # separate objects: fewer steps for the reference's own traverser over the optimised tree
        steps_opt = cpu.traverse(rays, want_stats=True)["steps"]
        cpu.set_bvh(*plain.get_bvh())
        assert steps_opt < cpu.traverse(rays, want_stats=True)["steps"]


def test_refit_range_one_moved_entity(vt, oracle_mod):
    """vt_accel_refit_range: only the moved entities' triangles go up; the resident scene ends up byte-identical in effect to a
    whole-scene refit — same hit records, same derived triangles, same refitted host boxes — and equal to the checker."""
    from test_host import _moved_props
    from vistrace_b200 import abi, scenes

    scene = scenes.scene_props(8, 21, 11, 12)
    moved = _moved_props(scene)
    rays = np.concatenate([scenes.pinhole_rays(320, 180, (0, -95, 40), (0, 0, 10)), scenes.random_rays(20000, (-90, -90, -5), (90, 90, 60), seed=12)])
    whole = vt.Accel(0, layout="quad").populate(scene).refit(moved)
    # its own copy of the scene: refit_range keeps the Python-side triangle array of the handle in step
    part = vt.Accel(0, layout="quad").populate(abi.SceneData(scene.tris.copy(), scene.materials, scene.entities))
    for ent in range(1, len(scene.entities)):  # entity by entity, each a contiguous run of the triangle array
        idx = np.nonzero(scene.tris["ent_idx"] == ent)[0]
        assert idx[-1] - idx[0] + 1 == len(idx)
        part.refit_range(moved.tris[idx[0]: idx[-1] + 1], int(idx[0]))
    assert part.layout == "quad"
    h_whole, a_whole = whole.traverse(rays, want_attrs=True)
    h_part, a_part = part.traverse(rays, want_attrs=True)
    assert h_part.tobytes() == h_whole.tobytes() and a_part.tobytes() == a_whole.tobytes()
    np.testing.assert_array_equal(part.tri_derived().view(np.uint32), whole.tri_derived().view(np.uint32))
    assert np.array_equal(part.get_bvh()[0]["bounds"], whole.get_bvh()[0]["bounds"])
    # the ranged walk (touched quads and their ancestors only) leaves the very quads of the whole-tree pass: the same node visits
    # and triangle tests ray by ray (a stale ancestor box would prune differently), the same node-area sum
    s_whole, s_part = whole.traverse_ray_stats(rays), part.traverse_ray_stats(rays)
    assert np.array_equal(s_part[0], s_whole[0]) and np.array_equal(s_part[1], s_whole[1])
    assert part.refit_quality()[0] == pytest.approx(whole.refit_quality()[0], rel=1e-9) and whole.refit_quality()[0] != 1.0  # atomic double sums: order varies
    kind = "reference" if oracle_mod.available("reference") else "port"
    cpu = oracle_mod.CpuScene(moved, kind, build_bvh=False)
    cpu.set_bvh(*part.get_bvh())
    assert same_hits(h_part, cpu.traverse(rays)["hits"], "quad", rays, cpu)
    # back to the start through the whole-tree walk of the same entry point (VT_REFIT_RANGE_WALK=0): the as-built quads again
    os.environ["VT_REFIT_RANGE_WALK"] = "0"
    try:
        for ent in range(1, len(scene.entities)):
            idx = np.nonzero(scene.tris["ent_idx"] == ent)[0]
            part.refit_range(scene.tris[idx[0]: idx[-1] + 1], int(idx[0]))
    finally:
        del os.environ["VT_REFIT_RANGE_WALK"]
    fresh = vt.Accel(0, layout="quad").populate(scene)
    assert part.traverse(rays).tobytes() == fresh.traverse(rays).tobytes()
    s_fresh, s_back = fresh.traverse_ray_stats(rays), part.traverse_ray_stats(rays)
    assert np.array_equal(s_back[0], s_fresh[0]) and np.array_equal(s_back[1], s_fresh[1])
    for ent in range(1, len(scene.entities)):  # and forward again on the ranged walk: counters were left clean
        idx = np.nonzero(scene.tris["ent_idx"] == ent)[0]
        part.refit_range(moved.tris[idx[0]: idx[-1] + 1], int(idx[0]))
    assert part.traverse(rays).tobytes() == h_whole.tobytes()
    s_part = part.traverse_ray_stats(rays)
    assert np.array_equal(s_part[0], s_whole[0]) and np.array_equal(s_part[1], s_whole[1])
    with pytest.raises(RuntimeError):
        part.refit_range(moved.tris[:4], scene.n_tris - 2)  # past the end
    with pytest.raises(RuntimeError):
        vt.Accel(0, layout="exact").populate(scene).refit_range(moved.tris[:4], 0)  # quad layout only


def test_refit_falls_back_when_a_box_leaves_the_float_grid(vt, oracle_mod):
    """K5 cannot requantise a box whose coordinates exceed the grid range (|E| <= 60): it reports that, vt_accel_refit
    re-derives the layout on the host, which drops to the exact layout — results still equal the checker's."""
    from test_host import _moved_props
    from vistrace_b200 import abi, scenes

    scene = scenes.scene_props(4, 13, 9, 8)
    moved = _moved_props(scene)
    far = np.nonzero(moved.tris["ent_idx"] == 2)[0][0]
    moved.tris["p"][far, 1] = (1.0e30, -1.0e30, 1.0e30)  # one vertex far outside any representable grid
    rays = scenes.pinhole_rays(200, 120, (0, -95, 40), (0, 0, 10))
    accel = vt.Accel(0, layout="quad").populate(scene)
    assert accel.layout == "quad"
    accel.refit(moved)
    assert accel.layout == "exact"
    kind = "reference" if oracle_mod.available("reference") else "port"
    cpu = oracle_mod.CpuScene(moved, kind, build_bvh=False)
    cpu.set_bvh(*accel.get_bvh())
    assert accel.traverse(rays).tobytes() == cpu.traverse(rays)["hits"].tobytes()
    accel.refit(scene)  # and back: the exact layout is refitted on the host
    cpu = oracle_mod.CpuScene(scene, kind, build_bvh=False)
    cpu.set_bvh(*accel.get_bvh())
    assert accel.traverse(rays).tobytes() == cpu.traverse(rays)["hits"].tobytes()


@pytest.mark.parametrize("kind", oracle_kinds())
def test_alpha_test_through_dxt_compressed_vtf_files(vt, oracle_mod, kind, layout):
    """The ingestion row end to end: the foliage scene's base textures arrive as DXT5 / DXT1-one-bit-alpha VTF FILES,
    go through vt_vtf_decode (pinned to the reference's parser in tests/test_vtf.py) and drive the alpha test and the
    TraceResult albedo/alpha of both the GPU path and the checker."""
    from test_vtf import make_vtf
    from vistrace_b200 import abi, scenes

    base = scenes.scene_foliage(n_cards=1500, tex_size=32, ground_quads=8)
    files = [make_vtf("DXT5", 64, 64, 7, low=(16, 16), seed=21), make_vtf("DXT1_ONEBITALPHA", 32, 32, 6, flags=0x4 | 0x8, seed=22)]
    texs = [vt.vtf_decode(f) for f in files]
    assert len(base.textures) == len(texs)
    scene = abi.SceneData(base.tris, base.materials, base.entities, textures=texs)
    rays = np.concatenate([scenes.pinhole_rays(320, 180, (0, -48, 20), (0, 0, 8)), scenes.random_rays(15000, (-45, -45, -3), (45, 45, 50), seed=5)])
    accel, cpu, hits, attrs, want = _check_against(vt, oracle_mod, scene, rays, kind, "product", layout)
    hit = hits["prim"] != abi.VT_MISS
    alpha_tested = (scene.materials["flags"][scene.tris["material"][hits["prim"][hit]]] & abi.VT_MATFLAG_ALPHATEST) != 0
    assert alpha_tested.sum() > 1000 and len(np.unique(attrs["alpha"][hit][alpha_tested])) > 20  # texels really sampled


@pytest.mark.parametrize("kind", oracle_kinds())
def test_alpha_test_through_16_bit_vtf_files(vt, oracle_mod, kind, layout):
    """The 16-bit VTF formats end to end: files -> vt_vtf_decode -> WIDE texels (uint16 numerators + divisor codes) -> the alpha
    test inside K1 and the albedo / alpha of K2.  Against the unmodified reference with RGBA16161616 files (its own parser reads
    them); against the C port additionally with BGRA4444 and RGB565, whose ParsePixel values leave [0, 1] (alpha up to 16,
    green up to 32: exactly what the reference samples)."""
    from test_vtf import make_vtf
    from vistrace_b200 import abi, scenes

    base = scenes.scene_foliage(n_cards=1500, tex_size=32, ground_quads=8)
    fmts = ("RGBA16161616", "RGBA16161616F") if kind == "reference" else ("BGRA4444", "RGB565")
    files = [make_vtf(fmts[0], 64, 64, 7, seed=31), make_vtf(fmts[1], 32, 32, 6, flags=0x4 | 0x8, seed=32)]
    texs = [vt.vtf_decode(f) for f in files]
    assert all(t[5] & abi.VT_TEXEL_WIDE for t in texs)
    scene = abi.SceneData(base.tris, base.materials, base.entities, textures=texs)
    rays = np.concatenate([scenes.pinhole_rays(320, 180, (0, -48, 20), (0, 0, 8)), scenes.random_rays(15000, (-45, -45, -3), (45, 45, 50), seed=5)])
    accel, cpu, hits, attrs, want = _check_against(vt, oracle_mod, scene, rays, kind, "product", layout)
    hit = hits["prim"] != abi.VT_MISS
    alpha_tested = (scene.materials["flags"][scene.tris["material"][hits["prim"][hit]]] & abi.VT_MATFLAG_ALPHATEST) != 0
    assert alpha_tested.sum() > 1000 and len(np.unique(attrs["alpha"][hit][alpha_tested])) > 20  # texels really sampled
    if kind != "reference":
        assert attrs["alpha"][hit].max() > 1.0  # BGRA4444's unmasked alpha shift: the reference's value, not a clamped one


def test_two_gpu_sharded_trace_nccl(vt):
    """Real multi-GPU plumbing when the box has >= 2 GPUs (gpurun --gpus 2): replicated hierarchy, ray shards, NCCL gather."""
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from conftest import ROOT

    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29571", f"{ROOT}/tests/multi_gpu_worker.py"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MULTI_GPU_OK" in r.stdout


def test_refit_range_onto_an_alphatest_material_enables_the_alpha_test(vt, oracle_mod):
    """a scene populated WITHOUT any alpha-tested triangle selects K1's no-alpha template; a refit that moves triangles
    onto an alpha-tested material must switch the template (has_alphatest) — on the device path as well as on the host path."""
    from vistrace_b200 import abi, scenes

    base = scenes.scene_foliage(n_cards=800, tex_size=32, ground_quads=8)
    alpha_mats = np.nonzero((base.materials["flags"] & abi.VT_MATFLAG_ALPHATEST) != 0)[0]
    plain = int(np.nonzero((base.materials["flags"] & abi.VT_MATFLAG_ALPHATEST) == 0)[0][0])
    assert len(alpha_mats) > 0
    cards = np.nonzero(np.isin(base.tris["material"], alpha_mats))[0]
    rays = np.concatenate([scenes.pinhole_rays(240, 135, (0, -48, 20), (0, 0, 8)), scenes.random_rays(10000, (-45, -45, -3), (45, 45, 50), seed=5)])
    kind = "reference" if oracle_mod.available("reference") else "port"
    cpu = oracle_mod.CpuScene(base, kind, build_bvh=False)
    for path in ("range", "device", "host"):
        start = base.tris.copy()
        start["material"][cards] = plain  # nothing alpha-tested at populate time (a fresh copy: refit_range updates the scene it was given)
        scene0 = abi.SceneData(start, base.materials, base.entities, base.textures)
        accel = vt.Accel(0, layout="quad").populate(scene0)
        opaque = accel.traverse(rays)
        if path == "range":  # contiguous runs of card triangles, one refit_range each
            runs = np.split(cards, np.nonzero(np.diff(cards) != 1)[0] + 1)
            for r in runs:
                accel.refit_range(base.tris[r[0]: r[-1] + 1], int(r[0]))
        else:
            import os

            os.environ["VT_REFIT_DEVICE"] = "1" if path == "device" else "0"
            try:
                accel.refit(base)
            finally:
                del os.environ["VT_REFIT_DEVICE"]
        cpu.set_bvh(*accel.get_bvh())
        want = cpu.traverse(rays)["hits"]
        got = accel.traverse(rays)
        assert same_hits(got, want, "quad", rays, cpu), path
        assert (got["prim"] != opaque["prim"]).sum() > 100, path  # rays now pass through the transparent texels
        accel.close()


def test_refit_quality_and_rebuild_trigger(vt):
    """vt_accel_refit_quality: 1.0 as built, grows when moved geometry loosens the refitted boxes, returns to 1.0 when the geometry
    moves back; with a rebuild ratio set, vt_accel_refit rebuilds from scratch once the refitted tree exceeds it."""
    from test_host import _moved_props
    from vistrace_b200 import scenes

    scene = scenes.scene_props(8, 21, 11, 12)
    moved = _moved_props(scene)
    rays = scenes.pinhole_rays(200, 120, (0, -95, 40), (0, 0, 10))
    accel = vt.Accel(0, layout="quad").populate(scene)
    assert accel.refit_quality() == (1.0, 0)
    accel.refit(moved)
    ratio, rebuilds = accel.refit_quality()
    assert ratio > 1.002 and rebuilds == 0, ratio  # the props left the places their subtrees were built for (a small move: +0.55 % here)
    refit_hits = accel.traverse(rays)
    accel.refit(scene)
    back, _ = accel.refit_quality()
    assert abs(back - 1.0) < 1e-9, back
    accel.set_refit_rebuild_ratio(1.0 + (ratio - 1.0) / 2)
    accel.refit(moved)  # exceeds the tolerance: rebuilt
    ratio2, rebuilds = accel.refit_quality()
    assert rebuilds == 1 and ratio2 == 1.0
    rep = compare_hits(accel.traverse(rays), refit_hits)  # another tree over the same triangles: same answer up to exact ties
    assert rep["hit_miss_mismatch"] == 0 and rep["tuv_bit_mismatch"] == 0 and rep["prim_mismatch"] <= 3, rep
    fresh = vt.Accel(0, layout="quad").populate(moved)
    assert accel.traverse_stats(rays) == fresh.traverse_stats(rays)  # the rebuilt tree IS the tree a fresh populate builds


@pytest.mark.parametrize("layout_id,layout_name", [(0, "exact"), (2, "quad")])
def test_compiled_reference_side_binding(built, oracle_mod, layout_id, layout_name):
    """INTEGRATION.md sections 2-4 COMPILED (oracle/ref_binding.cpp -> oracle/_ref/libvt_ref_binding.so): the real reference
    AccelStruct runs its own PopulateAccel build sequence (source/objects/AccelStruct.cpp:762-775), hands its own containers and
    its own collapsed bvh::Bvh to vt_accel_populate_with_bvh — textures through the public IVTFTexture interface only (GetPixel) —
    and its own per-ray Traverse statements (:810-831) are compared with the batched GPU entry on the same rays."""
    import ctypes as C
    import os

    from conftest import ROOT
    from vistrace_b200 import abi, scenes

    so = os.path.join(ROOT, "oracle", "_ref", "libvt_ref_binding.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libvt_ref_binding.so is built where /root/reference exists (make -C oracle binding)")
    lib = C.CDLL(so)
    lib.vtbind_selfcheck.restype = C.c_int
    lib.vtbind_selfcheck.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_char_p, C.c_uint64]
    for scene, rays in ((scenes.scene_foliage(n_cards=3000, tex_size=64, ground_quads=16),
                         np.concatenate([scenes.pinhole_rays(320, 180, (0, -48, 20), (0, 0, 8)), scenes.random_rays(20000, (-45, -45, -3), (45, 45, 50), seed=5)])),
                        (scenes.scene_props(8, 21, 11, 12),
                         np.concatenate([scenes.pinhole_rays(320, 180, (0, -95, 40), (0, 0, 10)), scenes.random_rays(20000, (-90, -90, -5), (90, 90, 60), seed=3)]))):
        rays = np.ascontiguousarray(rays, abi.RAY)
        report = np.zeros(8, np.uint64)
        worst = C.c_double(0)
        err = C.create_string_buffer(512)
        rc = lib.vtbind_selfcheck(C.cast(scene.ptr(), C.c_void_p), rays.ctypes.data, len(rays), layout_id, report.ctypes.data, C.addressof(worst), err, 512)
        assert rc == 0, err.value.decode()
        n, bytes_diff, hit_miss, tuv, prim, compared, n_tex, n_tris = (int(v) for v in report)
        print(f"[binding] {layout_name}: {n} rays, {bytes_diff} records differ, {hit_miss} hit/miss, {tuv} t/u/v, {prim} prim; {n_tex} textures via GetPixel; attr err {worst.value:.3g}")
        assert n == len(rays) and n_tris == scene.n_tris and compared > 0.3 * n
        assert hit_miss == 0 and tuv == 0
        assert worst.value <= 1e-5  # tolerance from BASELINE.json north_star
        if layout_name == "exact":
            assert bytes_diff == 0  # the reference's tree, the reference's visit order: its hit buffer byte for byte
        else:
            assert prim <= 3  # quantised layout: exact ties only
        if len(scene.textures):
            assert n_tex >= len(scene.textures)
