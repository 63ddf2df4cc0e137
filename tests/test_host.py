"""CPU tests of the host side of the product: the C-ABI library loads and exports every symbol the
header declares, the hierarchy builder and the flattening step are correct (checked by running the
ORACLE's traverser over the product's tree), and nothing computes on a machine without a GPU."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, compare_hits


def test_library_exports_every_declared_symbol(built):
    import vistrace_b200 as vt
    from vistrace_b200 import binding

    header = open(os.path.join(ROOT, "include", "vistrace_b200.h")).read()
    declared = set(re.findall(r"\b(vt_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 18
    L = vt.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/vistrace_b200.h but not exported"
    assert declared == set(binding.SYMBOLS), declared ^ set(binding.SYMBOLS)


def test_header_is_plain_c_and_links_against_the_library(built, tmp_path):
    """The boundary is a C ABI: the header compiles as C99 and as C++11 with pedantic warnings as errors, and a C program that
    calls into the library links and runs (no compute: vt_device_count, vt_last_error, a VTF header probe)."""
    import shutil
    import subprocess

    import vistrace_b200 as vt

    if not shutil.which("gcc"):
        pytest.skip("no C compiler")
    src = tmp_path / "abi.c"
    src.write_text(
        '#include <stdio.h>\n#include <string.h>\n#include "vistrace_b200.h"\n'
        "int main(void) {\n"
        "    vt_ray r; vt_hit h; vt_vtf_info info;\n"
        "    unsigned char junk[16];\n"
        "    memset(&r, 0, sizeof r); memset(&h, 0, sizeof h); memset(junk, 0, sizeof junk);\n"
        '    if (vt_vtf_read_info(junk, sizeof junk, &info) == 0) return 2; /* bad signature must fail ... */\n'
        '    if (vt_last_error()[0] == 0) return 3;                        /* ... and say why */\n'
        '    printf("%d %u %u %u\\n", vt_device_count() >= -1, (unsigned)sizeof(vt_ray), (unsigned)sizeof(vt_hit), (unsigned)sizeof(vt_tri_in));\n'
        "    return 0;\n}\n")
    inc = os.path.join(ROOT, "include")
    lib = vt.library_path()
    for cc, std in (("gcc", "-std=c99"), ("g++", "-std=c++11")):
        if not shutil.which(cc):
            continue
        subprocess.check_call([cc, std, "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c" if cc == "gcc" else "c++", "-I", inc, str(src)])
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-std=c99", "-I", inc, str(src), "-o", str(exe), lib, "-Wl,-rpath," + os.path.dirname(lib)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out
    assert out.stdout.split() == ["1", "32", "16", "152"]


def test_record_sizes_match_header(built):
    from vistrace_b200 import abi

    assert (abi.RAY.itemsize, abi.HIT.itemsize, abi.NODE.itemsize, abi.TRI_IN.itemsize, abi.ATTR.itemsize) == (32, 16, 32, 152, 128)


def test_no_cpu_fallback(built):
    """Without a CUDA device the engine refuses to work instead of detouring through a CPU path."""
    import vistrace_b200 as vt

    if vt.lib().vt_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        vt.Accel(0)


def test_ingestion_survives_mutation_fuzz_under_sanitizers(built):
    """tools/fuzz/run.sh: the host-only ingestion code (VTF / MDL / BSP) compiled with AddressSanitizer + UBSan reads a few hundred
    mutated synthetic files from exact-size heap buffers, and the hierarchy code (builders, flatten, layouts, refit, reinsertion) is
    run over hostile geometry and mutated node arrays — everything is either processed or refused, nothing reads out of bounds."""
    import shutil
    import subprocess

    if not shutil.which("g++"):
        pytest.skip("no C++ compiler")
    probe = subprocess.run("echo 'int main(){}' | g++ -x c++ - -fsanitize=address,undefined -o /dev/null", shell=True, capture_output=True)
    if probe.returncode != 0:
        pytest.skip("g++ cannot link the sanitizer runtimes here")
    out = subprocess.run([os.path.join(ROOT, "tools", "fuzz", "run.sh"), "7", "300"], capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-3000:])
    assert "no sanitizer finding" in out.stdout and out.stdout.count("accepted") >= 3 and "hierarchy: seed 7" in out.stdout


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "vistrace_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "oracle/" not in src and "vt_oracle" not in src and "vtref_" not in src, f


@pytest.mark.parametrize("scene_name", ["heightfield", "foliage", "props"])
def test_builder_tree_is_valid_and_equivalent(oracle_mod, scene_name):
    """The product builder emits a well-formed bvh::Bvh<float>-form tree; the oracle traversing it finds the same
    closest t as over the brute-force single-leaf tree for every ray (primitive may differ only on exact ties)."""
    import vistrace_b200 as vt
    from vistrace_b200 import abi, scenes

    scene = {"heightfield": lambda: scenes.scene_heightfield(48), "foliage": lambda: scenes.scene_foliage(1200, tex_size=32),
             "props": lambda: scenes.scene_props(5, 15, 9, 12)}[scene_name]()
    nodes, prims = vt.build_bvh(scene)
    n = scene.n_tris
    assert len(nodes) % 2 == 1 and sorted(prims.tolist()) == list(range(n))
    leaves = nodes[nodes["prim_count"] > 0]
    assert leaves["prim_count"].sum() == n and leaves["prim_count"].max() <= 4
    inner = nodes[nodes["prim_count"] == 0]
    assert (inner["first"] % 2 == 1).all() and len(set(inner["first"].tolist())) == len(inner)
    # children are inside their parent
    for i in np.random.default_rng(0).choice(len(inner), min(300, len(inner)), replace=False):
        p = inner[i]
        for c in (nodes[p["first"]], nodes[p["first"] + 1]):
            assert (c["bounds"][0::2] >= p["bounds"][0::2]).all() and (c["bounds"][1::2] <= p["bounds"][1::2]).all()
    rays = scenes.random_rays(4000, (-45, -45, -2), (45, 45, 40), seed=5)
    cpu = oracle_mod.CpuScene(scene, "port", build_bvh=False)
    cpu.set_bvh(nodes, prims)
    got = cpu.traverse(rays)["hits"]
    brute = np.zeros(1, abi.NODE)
    brute["bounds"], brute["prim_count"] = (-1e30, 1e30, -1e30, 1e30, -1e30, 1e30), n
    cpu.set_bvh(brute, np.arange(n, dtype=np.uint64))
    want = cpu.traverse(rays)["hits"]
    rep = compare_hits(got, want)
    assert rep["hit_miss_mismatch"] == 0 and rep["prim_mismatch"] == rep["prim_mismatch_exact_tie"], rep
    np.testing.assert_array_equal(got["t"], want["t"])
    # deterministic: same input -> same tree
    nodes2, prims2 = vt.build_bvh(scene)
    assert nodes.tobytes() == nodes2.tobytes() and prims.tobytes() == prims2.tobytes()


@pytest.mark.parametrize("scene_name", ["heightfield", "foliage", "props"])
def test_reinsertion_optimisation_keeps_the_tree_valid_and_the_answers(oracle_mod, scene_name):
    """vt_optimize_bvh (builder-quality option; the reference library's counterpart is
    bvh::ParallelReinsertionOptimizer): the optimised array is a well-formed depth-first bvh::Bvh-form tree over the SAME leaves,
    every inner box is exactly the union of its children, the sum of inner-node areas went down as reported, the oracle's
    traverser finds the same hits over it (exact ties aside) and the result is deterministic."""
    import vistrace_b200 as vt
    from vistrace_b200 import binding, scenes

    scene = {"heightfield": lambda: scenes.scene_heightfield(48), "foliage": lambda: scenes.scene_foliage(1200, tex_size=32),
             "props": lambda: scenes.scene_props(5, 15, 9, 12)}[scene_name]()
    nodes, prims = vt.build_bvh(scene)
    opt, before, after, moves = binding.optimize_bvh(nodes, iterations=6, fraction=0.3)
    assert len(opt) == len(nodes) and moves > 0 and 0 < after < before
    # the same leaves (primitive ranges untouched), a tree again: every pair referenced once, children behind their parents
    leaf = lambda a: sorted(zip(a["first"][a["prim_count"] > 0].tolist(), a["prim_count"][a["prim_count"] > 0].tolist()))
    assert leaf(opt) == leaf(nodes)
    inner = np.nonzero(opt["prim_count"] == 0)[0]
    first = opt["first"][inner]
    assert (first % 2 == 1).all() and len(set(first.tolist())) == len(inner) == (len(opt) - 1) // 2 and (first > inner).all()
    lo, hi = opt["bounds"][:, 0::2], opt["bounds"][:, 1::2]
    np.testing.assert_array_equal(lo[inner], np.minimum(lo[first], lo[first + 1]))
    np.testing.assert_array_equal(hi[inner], np.maximum(hi[first], hi[first + 1]))

    def inner_area(a):
        e = (a["bounds"][:, 1::2] - a["bounds"][:, 0::2]).astype(np.float64)[a["prim_count"] == 0]
        return float((e[:, 0] * e[:, 1] + e[:, 1] * e[:, 2] + e[:, 2] * e[:, 0]).sum())

    assert inner_area(nodes) == pytest.approx(before, rel=1e-5) and inner_area(opt) == pytest.approx(after, rel=1e-5)
    flat = vt.flatten_bvh(opt, prims)  # the product's own validation: adjacent odd pairs, every primitive once, depth <= 64
    assert sorted(flat["leaf_order"].tolist()) == list(range(scene.n_tris))
    quads = binding.build_quads(opt, prims)
    assert sorted(quads["leaf_order"].tolist()) == list(range(scene.n_tris)) and quads["max_stack"] <= 64
    rays = np.concatenate([scenes.pinhole_rays(96, 54, (0, -48, 20), (0, 0, 8)), scenes.random_rays(4000, (-45, -45, -2), (45, 45, 40), seed=5)])
    cpu = oracle_mod.CpuScene(scene, "port", build_bvh=False)
    cpu.set_bvh(nodes, prims)
    want = cpu.traverse(rays, want_stats=True)
    cpu.set_bvh(opt, prims)
    got = cpu.traverse(rays, want_stats=True)
    rep = compare_hits(got["hits"], want["hits"])
    assert rep["hit_miss_mismatch"] == 0 and rep["prim_mismatch"] == rep["prim_mismatch_exact_tie"], rep
    np.testing.assert_array_equal(got["hits"]["t"], want["hits"]["t"])
    if scene_name == "props":  # separate objects on a ground plane: the case the pass is made for
        assert got["steps"] < want["steps"]
    again = binding.optimize_bvh(nodes, iterations=6, fraction=0.3)
    assert again[0].tobytes() == opt.tobytes() and again[1:] == (before, after, moves)
    # nothing to do / nothing done leaves the array as it was
    assert binding.optimize_bvh(nodes, iterations=0)[0].tobytes() == nodes.tobytes()
    tiny = np.zeros(3, nodes.dtype)  # a root over two leaves: nothing can move
    tiny["bounds"] = [(0, 2, 0, 1, 0, 1), (0, 1, 0, 1, 0, 1), (1, 2, 0, 1, 0, 1)]
    tiny["prim_count"], tiny["first"] = [0, 1, 1], [1, 0, 1]
    out = binding.optimize_bvh(tiny, iterations=4)
    assert out[0].tobytes() == tiny.tobytes() and out[3] == 0


def test_collapse_rule_is_chosen_by_sibling_overlap(built, monkeypatch):
    """VT_COLLAPSE=auto (the default): a dense volume of overlapping cards keeps the largest-child rule, surfaces and separate objects
    get the SAH-optimal plan (vt_bvh_collapse.cpp: sibling_overlap against VT_COLLAPSE_OVERLAP, profiles/r2_child_order.md); either
    way the quads hold every triangle once and fewer wide nodes than binary inner nodes."""
    from vistrace_b200 import binding, scenes

    cases = {"dense foliage": (scenes.scene_foliage(n_cards=12000, extent=16.0, tex_size=32, ground_quads=8), "greedy"),
             "height field": (scenes.scene_heightfield(64), "dp"), "props": (scenes.scene_props(6, 15, 9, 8), "dp")}
    for name, (scene, expect) in cases.items():
        nodes, prims = binding.build_bvh(scene)
        built_by = {}
        for mode in ("auto", "dp", "greedy"):
            monkeypatch.setenv("VT_COLLAPSE", mode)
            built_by[mode] = binding.build_quads(nodes, prims)
        monkeypatch.delenv("VT_COLLAPSE")
        default = binding.build_quads(nodes, prims)
        assert default["quads"].tobytes() == built_by["auto"]["quads"].tobytes()
        assert built_by["auto"]["quads"].tobytes() == built_by[expect]["quads"].tobytes(), name
        assert built_by["dp"]["quads"].tobytes() != built_by["greedy"]["quads"].tobytes()
        assert len(built_by["dp"]["quads"]) < len(built_by["greedy"]["quads"]) < (len(nodes) - 1) // 2 + 1
        for q in built_by.values():
            assert sorted(q["leaf_order"].tolist()) == list(range(scene.n_tris))
    monkeypatch.setenv("VT_COLLAPSE_OVERLAP", "0.0")  # threshold override: everything counts as overlapping
    nodes, prims = binding.build_bvh(cases["props"][0])
    assert binding.build_quads(nodes, prims)["quads"].tobytes() == built_by["greedy"]["quads"].tobytes()


def test_reinsertion_on_the_reference_objects_own_hierarchy(built):
    """INTEGRATION.md "Builder options" compiled (oracle/ref_binding.cpp, vtbind_optimize_check; host only): vt_optimize_bvh runs in
    place on the REAL bvh::Bvh<float> of the reference's AccelStruct — built by its own PLOC + LeafCollapser sequence
    (source/objects/AccelStruct.cpp:762-770) — and the reference's own traverser (`:810-831`) then walks its own, optimised,
    containers: same hits (exact ties aside), fewer traversal steps on a scene of separate objects."""
    import ctypes as C

    from vistrace_b200 import scenes

    so = os.path.join(ROOT, "oracle", "_ref", "libvt_ref_binding.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libvt_ref_binding.so is built where /root/reference exists (make -C oracle binding)")
    lib = C.CDLL(so)
    lib.vtbind_optimize_check.restype = C.c_int
    lib.vtbind_optimize_check.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_char_p, C.c_uint64]
    scene = scenes.scene_props(8, 21, 11, 12)
    rays = np.ascontiguousarray(np.concatenate([scenes.pinhole_rays(160, 90, (0, -95, 40), (0, 0, 10)), scenes.random_rays(8000, (-90, -90, -5), (90, 90, 60), seed=8)]))
    report, areas, err = np.zeros(8, np.uint64), np.zeros(2, np.float64), C.create_string_buffer(512)
    rc = lib.vtbind_optimize_check(C.cast(scene.ptr(), C.c_void_p), rays.ctypes.data, len(rays), 6, 0.3, report.ctypes.data, areas.ctypes.data, err, 512)
    assert rc == 0, err.value.decode()
    n, hit_miss, tuv, prim, moves, steps_before, steps_after, nodes = (int(v) for v in report)
    assert n == len(rays) and hit_miss == 0 and tuv == 0 and prim <= 2, report
    assert moves > 0 and 0 < areas[1] < areas[0] and steps_after < steps_before, (report, areas)


def test_reinsertion_on_random_hostile_scenes(oracle_mod):
    """Seeded fuzz: triangle soups, zero-area and grid-aligned triangles, duplicated geometry, coplanar layers (conftest.hostile_case),
    product tree and the reference's own PLOC tree, several batch fractions: the optimised array is always a valid tree over the same
    leaves, no worse in inner-node area, and the checker finds the same hits over it — t bit-identical, the primitive too unless two
    candidates tie exactly."""
    import vistrace_b200 as vt
    from conftest import hostile_case
    from vistrace_b200 import binding

    rng = np.random.default_rng(20261018)
    moved_any = 0
    for it in range(36):
        scene, rays, kind = hostile_case(it, rng, n_rays=1500)
        nodes, prims = vt.build_bvh_ploc(scene) if it % 3 == 2 else vt.build_bvh(scene)
        opt, before, after, moves = binding.optimize_bvh(nodes, iterations=1 + it % 5, fraction=(0.05, 0.2, 0.5)[it % 3])
        assert len(opt) == len(nodes) and after <= before
        moved_any += moves > 0
        if moves == 0:
            assert opt.tobytes() == nodes.tobytes()
            continue
        leaf = lambda a: sorted(zip(a["first"][a["prim_count"] > 0].tolist(), a["prim_count"][a["prim_count"] > 0].tolist()))
        assert leaf(opt) == leaf(nodes)
        flat = vt.flatten_bvh(opt, prims)  # validates pairs, coverage and depth
        assert sorted(flat["leaf_order"].tolist()) == list(range(scene.n_tris))
        cpu = oracle_mod.CpuScene(scene, "port", build_bvh=False)
        cpu.set_bvh(nodes, prims)
        want = cpu.traverse(rays)["hits"]
        cpu.set_bvh(opt, prims)
        got = cpu.traverse(rays)["hits"]
        rep = compare_hits(got, want)
        assert rep["hit_miss_mismatch"] == 0 and rep["prim_mismatch"] == rep["prim_mismatch_exact_tie"], (it, kind, rep)
        np.testing.assert_array_equal(got["t"], want["t"])
    assert moved_any >= 12


def test_reinsertion_rejects_malformed_arrays(built):
    from vistrace_b200 import binding, scenes

    nodes, _ = binding.build_bvh(scenes.scene_heightfield(16))
    bad = nodes.copy()
    k = int(np.nonzero(bad["prim_count"] == 0)[0][3])
    bad["first"][k] = len(bad) + 5  # child pair outside the array
    with pytest.raises(RuntimeError):
        binding.optimize_bvh(bad)
    bad = nodes.copy()
    bad["first"][k] = bad["first"][k] + 1  # even index: not a sibling pair
    with pytest.raises(RuntimeError):
        binding.optimize_bvh(bad)
    bad = nodes.copy()
    inner = np.nonzero(bad["prim_count"] == 0)[0]
    bad["first"][inner[5]] = bad["first"][inner[4]]  # two parents for one pair: not a tree
    with pytest.raises(RuntimeError):
        binding.optimize_bvh(bad)
    with pytest.raises(RuntimeError):
        binding.optimize_bvh(nodes[:-1])  # even node count


def _emulate_pairs(flat, tris_derived, rays):
    """Tiny numpy traverser over the FLATTENED layout (pairs + leaf order) — checks the structure the GPU walks."""
    pairs, order = flat["pairs"], flat["leaf_order"]
    out = np.full(len(rays), 0xFFFFFFFF, np.uint32)
    for ri, r in enumerate(rays):
        o, d = r["o"].astype(np.float64), r["d"].astype(np.float64)
        best_t, best = float(r["tmax"]), 0xFFFFFFFF
        stack = [0] if len(pairs) else []
        runs = [(0, flat["root_leaf_count"])] if flat["root_leaf_count"] else []
        while stack or runs:
            for first, cnt in runs:
                for s in range(first, first + cnt):
                    t = tris_derived[order[s]].astype(np.float64)
                    p0, e1, e2, n = t[0:3], t[3:6], t[6:9], t[9:12]
                    nd = n @ d
                    if nd == 0:
                        continue
                    c = p0 - o
                    rr = np.cross(d, c)
                    u, v = (rr @ e2) / nd, (rr @ e1) / nd
                    tt = (n @ c) / nd
                    if u >= 0 and v >= 0 and 1 - u - v >= 0 and r["tmin"] <= tt <= best_t:
                        best_t, best = tt, order[s]
            runs = []
            if not stack:
                break
            p = pairs[stack.pop()]
            for b, cnt, first in ((p["l_bounds"], p["l_count"], p["l_first"]), (p["r_bounds"], p["r_count"], p["r_first"])):
                with np.errstate(divide="ignore", invalid="ignore"):
                    t0 = (b[0::2] - o) / d
                    t1 = (b[1::2] - o) / d
                lo, hi = np.nanmax(np.minimum(t0, t1)), np.nanmin(np.maximum(t0, t1))
                if max(lo, r["tmin"]) <= min(hi, best_t) * (1 + 1e-6) + 1e-6:
                    if cnt:
                        runs.append((int(first), int(cnt)))
                    else:
                        stack.append(int(first))
        out[ri] = best
    return out


@pytest.mark.parametrize("bfs_pairs", [0, 7, 100000])
def test_flatten_preserves_the_tree(oracle_mod, bfs_pairs):
    import vistrace_b200 as vt
    from vistrace_b200 import abi, scenes

    scene = scenes.scene_props(3, 11, 7, 6)
    nodes, prims = vt.build_bvh(scene)
    flat = vt.flatten_bvh(nodes, prims, bfs_pairs)
    n_pairs = (len(nodes) - 1) // 2
    assert len(flat["pairs"]) == n_pairs and sorted(flat["leaf_order"].tolist()) == list(range(scene.n_tris))
    assert 1 <= flat["max_depth"] <= 60
    p = flat["pairs"]
    inner_refs = np.concatenate([p["l_first"][p["l_count"] == 0], p["r_first"][p["r_count"] == 0]])
    assert sorted(inner_refs.tolist()) == list(range(1, n_pairs))  # every pair but the root pair is referenced exactly once
    # pair 0 holds the children of the root, bounds copied verbatim
    assert p[0]["l_bounds"].tobytes() == nodes[nodes[0]["first"]]["bounds"].tobytes()
    cpu = oracle_mod.CpuScene(scene, "port", build_bvh=False)
    cpu.set_bvh(nodes, prims)
    rays = scenes.random_rays(150, (-80, -80, 0), (80, 80, 60), seed=9)
    want = cpu.traverse(rays)["hits"]
    got = _emulate_pairs(flat, cpu.tri_derived(), rays)
    assert (got == want["prim"]).mean() > 0.98  # float64 emulation vs float32 oracle: only grazing rays may differ
    assert ((got == abi.VT_MISS) == (want["prim"] == abi.VT_MISS)).mean() > 0.98


def test_flatten_rejects_malformed_trees(built):
    import vistrace_b200 as vt
    from vistrace_b200 import abi

    nodes = np.zeros(3, abi.NODE)
    nodes[0]["first"] = 2  # children must sit at an odd index
    nodes[1]["prim_count"] = nodes[2]["prim_count"] = 1
    with pytest.raises(RuntimeError, match="child index"):
        vt.flatten_bvh(nodes, np.arange(2, dtype=np.uint64))
    nodes[0]["first"] = 1
    nodes[2]["first"] = 1
    with pytest.raises(RuntimeError, match="cover|past the end"):
        vt.flatten_bvh(nodes, np.arange(3, dtype=np.uint64))
    chain = np.zeros(2 * 70 + 1, abi.NODE)  # a 70-deep degenerate chain overflows the 64-entry stack
    for d in range(70):
        parent = 0 if d == 0 else 2 * d - 1
        chain[parent]["first"] = 2 * d + 1
        chain[2 * d + 2]["prim_count"], chain[2 * d + 2]["first"] = 1, d
    chain[2 * 69 + 1]["prim_count"], chain[2 * 69 + 1]["first"] = 1, 70
    with pytest.raises(RuntimeError, match="deeper"):
        vt.flatten_bvh(chain, np.arange(71, dtype=np.uint64))


@pytest.mark.parametrize("collapse", ["auto", "dp", "greedy"])
def test_quad_builder_rejects_malformed_trees_under_every_collapse_rule(built, collapse, monkeypatch):
    """Found by tools/fuzz/hierarchy_driver.cpp: the largest-child collapse followed grandchild references before anything had checked
    them (a hostile `first` read past the node array).  Every rule now refuses a child reference that is even, zero, out of range or
    shared by two parents — vt_build_quads is what vt_accel_populate_with_bvh runs on a caller's tree."""
    from vistrace_b200 import binding, scenes

    monkeypatch.setenv("VT_COLLAPSE", collapse)
    scene = scenes.scene_foliage(n_cards=3000, extent=8.0, tex_size=32, ground_quads=4)  # dense: auto picks the largest-child rule
    nodes, prims = binding.build_bvh(scene)
    assert len(binding.build_quads(nodes, prims)["quads"]) > 100
    inner = np.nonzero(nodes["prim_count"] == 0)[0]
    deep = [int(i) for i in inner if i > 40][:50]
    for k, value in enumerate([len(nodes) + 1000, 0xFFFFFFF1, 0, 2, len(nodes) - 1, int(nodes["first"][inner[3]])]):
        bad = nodes.copy()
        bad["first"][deep[3 * k]] = value
        with pytest.raises(RuntimeError):
            binding.build_quads(bad, prims)
    bad = nodes.copy()
    bad["prim_count"][deep[20]] = 0xFFFFFFFF  # an inner node turned into an impossible leaf
    with pytest.raises(RuntimeError):
        binding.build_quads(bad, prims)
    with pytest.raises(RuntimeError):
        binding.build_quads(nodes[:-2], prims)  # truncated array: some pair is gone


def test_ploc_builder_terminates_on_non_finite_geometry(built):
    """Found by tools/fuzz/hierarchy_driver.cpp: with NaN planes the cluster distance is not symmetric, a level of the clustering can
    be left without any mutual nearest-neighbour pair, and the loop repeated that level for ever.  The build now always shrinks a
    level; finite scenes are unaffected (the bit-identity tests above), non-finite triangles end up in a tree like any other."""
    import vistrace_b200 as vt
    from vistrace_b200 import abi

    rng = np.random.default_rng(1)
    for value, vertex in ((np.nan, 0), (np.inf, 0), (-np.inf, 2), (np.nan, 1)):
        tris = np.zeros(40, abi.TRI_IN)
        tris["p"] = rng.uniform(-20, 20, (40, 3, 3))
        tris["p"][[0, 7, 14, 21], vertex, 0] = value
        for collapse in (False, True):
            nodes, prims = vt.build_bvh_ploc(abi.SceneData(tris), collapse=collapse)
            assert sorted(prims.tolist()) == list(range(40)) and len(nodes) % 2 == 1
            assert nodes["prim_count"][nodes["prim_count"] > 0].sum() == 40
        nodes, prims = vt.build_bvh(abi.SceneData(tris))  # the product builder on the same triangles
        assert sorted(prims.tolist()) == list(range(40))


def test_scene_generators_are_deterministic_and_oriented():
    from vistrace_b200 import scenes

    a, b = scenes.scene_heightfield(16), scenes.scene_heightfield(16)
    assert a.tris.tobytes() == b.tris.tobytes()
    n = scenes.reference_normal(a.tris[:-12])
    assert (n[:, 2] > 0).all()  # terrain front faces look up (n = cross(p0-p1, p2-p0), Primitives.h:82,93)
    rays = scenes.pinhole_rays(8, 4, (0, -80, 60), (0, 0, 5))
    np.testing.assert_allclose(np.linalg.norm(rays["d"], axis=1), 1.0, atol=1e-6)
    pos = np.array([[10.0, -0.01, 300.0]], np.float32)
    nrm = np.array([[0.0, 0.0, 1.0]], np.float32)
    o = scenes.calc_ray_origin(pos, nrm)  # source/VisTrace.cpp:1495-1517
    assert o[0, 2] > pos[0, 2] and o[0, 0] == pos[0, 0] and o[0, 1] == pos[0, 1]


def _decode_cpairs(c):
    """numpy restatement of slab_cpair's decode (vt_traverse.cu): float32 fma-free because every step is exact."""
    scale = (c["exp"].astype(np.uint32) << 23).view(np.float32)                      # 2^E
    magic = (np.uint32(0x4B000000) | c["q"].astype(np.uint32)).view(np.float32)      # 2^23 + q
    plane = magic.astype(np.float64) * scale.astype(np.float64)[:, :, None] + c["origin_adj"].astype(np.float64)[:, :, None]
    assert (plane.astype(np.float32).astype(np.float64) == plane).all()              # (k + q) * 2^E is an exact float
    return plane.astype(np.float32)  # [pair, axis, {l.lo, l.hi, r.lo, r.hi}]


@pytest.mark.parametrize("scene_name", ["heightfield", "foliage", "props", "far_from_origin"])
def test_compact_pairs_are_conservative_and_address_the_same_children(built, scene_name):
    import vistrace_b200 as vt
    from vistrace_b200 import scenes

    if scene_name == "heightfield":
        scene = scenes.scene_heightfield(48)
    elif scene_name == "foliage":
        scene = scenes.scene_foliage(n_cards=1500, tex_size=16, ground_quads=8)
    elif scene_name == "props":
        scene = scenes.scene_props(5, 15, 9, 8)
    else:  # Source-engine map scale: coordinates of +-16384 with sub-unit triangles
        scene = scenes.scene_props(5, 15, 9, 8)
        scene.tris["p"] = scene.tris["p"] * np.float32(0.02) + np.array([16000.0, -15900.0, 8000.0], np.float32)
    nodes, prims = vt.build_bvh(scene)
    flat = vt.flatten_bvh(nodes, prims, 0)
    p = flat["pairs"]
    c = vt.compact_pairs(p)
    plane = _decode_cpairs(c)
    for a in range(3):
        assert (plane[:, a, 0] <= p["l_bounds"][:, 2 * a]).all() and (plane[:, a, 1] >= p["l_bounds"][:, 2 * a + 1]).all()
        assert (plane[:, a, 2] <= p["r_bounds"][:, 2 * a]).all() and (plane[:, a, 3] >= p["r_bounds"][:, 2 * a + 1]).all()
        # ... and tight: never more than one grid cell of slack per plane
        cell = (c["exp"][:, a].astype(np.uint32) << 23).view(np.float32)
        assert (p["l_bounds"][:, 2 * a] - plane[:, a, 0] < cell).all() and (plane[:, a, 3] - p["r_bounds"][:, 2 * a + 1] < cell).all()
        # the grid is as fine as 8 bits allow: the pair spans more than 127 cells unless float resolution binds
        lo = np.minimum(p["l_bounds"][:, 2 * a], p["r_bounds"][:, 2 * a]).astype(np.float64)
        hi = np.maximum(p["l_bounds"][:, 2 * a + 1], p["r_bounds"][:, 2 * a + 1]).astype(np.float64)
        ulp_bound = np.maximum(np.abs(lo), np.abs(hi)) / 2.0**22 >= cell
        assert (((hi - lo) / cell > 127) | ulp_bound | (hi == lo)).all()
    lcount, rcount = c["counts"] & 15, c["counts"] >> 4
    nxt = np.arange(1, len(c) + 1, dtype=np.uint32)
    lfirst = np.where(lcount == 0, nxt, c["ref"])
    rfirst = np.where(lcount == 0, c["ref"], np.where(rcount == 0, nxt, c["ref"] + lcount))
    assert (lcount == p["l_count"]).all() and (rcount == p["r_count"]).all()
    assert (lfirst == p["l_first"]).all() and (rfirst == p["r_first"]).all()


def test_compact_pairs_reject_what_they_cannot_hold(built):
    import vistrace_b200 as vt
    from vistrace_b200.binding import PAIR

    p = np.zeros(1, PAIR)
    p["l_count"], p["r_count"], p["l_first"], p["r_first"] = 16, 1, 0, 16
    with pytest.raises(RuntimeError, match="more than 15"):
        vt.compact_pairs(p)
    p["l_count"] = 1
    p["r_first"] = 1
    p["l_bounds"][0, 1] = np.inf
    with pytest.raises(RuntimeError, match="non-finite"):
        vt.compact_pairs(p)
    p["l_bounds"][0, 1] = 0
    p["r_first"] = 5  # both leaves: the right run must follow the left one
    with pytest.raises(RuntimeError, match="depth-first"):
        vt.compact_pairs(p)


def _tri_hit(t, o, d, tmin, best_t):
    p0, e1, e2, n = t[0:3], t[3:6], t[6:9], t[9:12]
    nd = n @ d
    if nd == 0:
        return None
    c = p0 - o
    rr = np.cross(d, c)
    u, v = (rr @ e2) / nd, (rr @ e1) / nd
    tt = (n @ c) / nd
    return tt if (u >= 0 and v >= 0 and 1 - u - v >= 0 and tmin <= tt <= best_t) else None


@pytest.mark.parametrize("scene_name", ["heightfield", "props"])
def test_quad_layout_structure_and_walk(oracle_mod, scene_name):
    """The 4-wide layout: every triangle in exactly one leaf run, every quad referenced once, decoded boxes contain
    the boxes of the binary nodes they replace, and a numpy walk over the quads finds the oracle's hits."""
    import vistrace_b200 as vt
    from vistrace_b200 import abi, scenes

    scene = scenes.scene_heightfield(40) if scene_name == "heightfield" else scenes.scene_props(4, 13, 9, 8)
    nodes, prims = vt.build_bvh(scene)
    qb = vt.build_quads(nodes, prims)
    quads, order = qb["quads"], qb["leaf_order"]
    assert sorted(order.tolist()) == list(range(scene.n_tris)) and qb["root_leaf_count"] == 0
    assert 0 < qb["max_stack"] <= 64 and len(quads) < (len(nodes) - 1) // 2  # fewer, wider nodes
    refs = quads["ref"]
    valid = (refs != 0xFFFFFFFF).astype(np.uint8)
    count, idx = refs >> 28, refs & 0x0FFFFFFF
    assert (refs[valid == 0] == 0xFFFFFFFF).all() and valid.sum(1).min() >= 2
    inner = (valid == 1) & (count == 0)
    assert sorted(idx[inner].tolist()) == list(range(1, len(quads)))  # a tree: each quad but the root referenced once
    leaf = (valid == 1) & (count > 0)
    runs = sorted(zip(idx[leaf].tolist(), count[leaf].tolist()))
    assert runs[0][0] == 0 and all(a + n == b for (a, n), (b, _) in zip(runs, runs[1:])) and runs[-1][0] + runs[-1][1] == scene.n_tris
    # decoded planes: exact floats on the power-of-two grid
    scale = quads["scale"].astype(np.float64)                                                            # [quad, axis], powers of two
    assert ((quads["scale"].view(np.uint32) & 0x007FFFFF) == 0).all()
    magic = float(vt.quad_plane_offset()) + quads["q"].astype(np.float64)                                 # [quad, axis, lo/hi, child]
    plane = magic * scale[:, :, None, None] + quads["origin_adj"].astype(np.float64)[:, :, None, None]
    assert (plane.astype(np.float32).astype(np.float64) == plane).all()
    # leaf children: the decoded box contains every vertex of the run's triangles
    cpu = oracle_mod.CpuScene(scene, "port", build_bvh=False)
    cpu.set_bvh(nodes, prims)
    td = cpu.tri_derived().astype(np.float64)
    verts = np.stack([td[:, 0:3], td[:, 0:3] - td[:, 3:6], td[:, 0:3] + td[:, 6:9]], 1)                    # p0, p0 - e1, p0 + e2
    for qi, ci in zip(*np.nonzero(leaf)):
        v = verts[order[idx[qi, ci]: idx[qi, ci] + count[qi, ci]]].reshape(-1, 3)
        assert (plane[qi, :, 0, ci] <= v.min(0)).all() and (plane[qi, :, 1, ci] >= v.max(0)).all()
    # inner children: the decoded box contains the child quad's own decoded children (nesting)
    for qi, ci in zip(*np.nonzero(inner)):
        c = idx[qi, ci]
        cv = valid[c] == 1
        assert (plane[qi, :, 0, ci] <= plane[c, :, 0][:, cv].min(1)).all() and (plane[qi, :, 1, ci] >= plane[c, :, 1][:, cv].max(1)).all()
    # walk
    rays = scenes.random_rays(120, (-45, -45, 0), (45, 45, 40), seed=4)
    want = cpu.traverse(rays)["hits"]
    got = np.full(len(rays), 0xFFFFFFFF, np.uint32)
    for ri, r in enumerate(rays):
        o, d = r["o"].astype(np.float64), r["d"].astype(np.float64)
        best_t, stack = float(r["tmax"]), [0]
        while stack:
            ref = stack.pop()
            if ref >> 28:
                for s in range(ref & 0x0FFFFFFF, (ref & 0x0FFFFFFF) + (ref >> 28)):
                    tt = _tri_hit(td[order[s]], o, d, r["tmin"], best_t)
                    if tt is not None:
                        best_t, got[ri] = tt, order[s]
                continue
            for ci in range(4):
                if not valid[ref, ci]:
                    continue
                with np.errstate(divide="ignore", invalid="ignore"):
                    t0, t1 = (plane[ref, :, 0, ci] - o) / d, (plane[ref, :, 1, ci] - o) / d
                lo, hi = np.nanmax(np.minimum(t0, t1)), np.nanmin(np.maximum(t0, t1))
                if max(lo, r["tmin"]) <= min(hi, best_t) * (1 + 1e-6) + 1e-6:
                    stack.append(int(refs[ref, ci]))
    assert (got == want["prim"]).mean() > 0.98 and ((got == abi.VT_MISS) == (want["prim"] == abi.VT_MISS)).mean() > 0.98


def _moved_props(scene, seed=9):
    """Same triangles, same order, new vertices: every prop (entity >= 1) gets its own rigid shift + a little jitter."""
    from vistrace_b200 import abi

    rng = np.random.default_rng(seed)
    tris = scene.tris.copy()
    n_ent = len(scene.entities)
    shift = rng.uniform(-6.0, 6.0, (n_ent, 3)).astype(np.float32)
    shift[0] = 0.0  # the world stays where it is
    tris["p"] = tris["p"] + shift[tris["ent_idx"]][:, None, :] + rng.uniform(-0.05, 0.05, tris["p"].shape).astype(np.float32) * (tris["ent_idx"] > 0)[:, None, None]
    return abi.SceneData(tris, scene.materials, scene.entities)


@pytest.mark.parametrize("tree", ["product", "reference"])
def test_refit_matches_the_reference_hierarchy_refitter(built, oracle_mod, tree):
    """vt_refit_bvh against bvh::HierarchyRefitter itself (oracle/_ref, leaf update of libs/bvh/test/refit_bvh.cpp:79-89)
    and against the C restatement: same boxes for every node, structure untouched, boxes nested and tight."""
    import vistrace_b200 as vt
    from vistrace_b200 import scenes

    scene = scenes.scene_props(6, 15, 9, 8)
    moved = _moved_props(scene)
    kinds = [k for k in ("reference", "port") if oracle_mod.available(k)]
    if tree == "reference" and "reference" not in kinds:
        pytest.skip("the reference-built tree needs oracle/_ref")
    if tree == "reference":
        nodes, prims = oracle_mod.CpuScene(scene, "reference", build_bvh=True).get_bvh()
    else:
        nodes, prims = vt.build_bvh(scene)
    got = vt.refit_bvh(moved, nodes, prims)
    assert (got["first"] == nodes["first"]).all() and (got["prim_count"] == nodes["prim_count"]).all()
    assert (got["bounds"] != nodes["bounds"]).any()
    for kind in kinds:
        cpu = oracle_mod.CpuScene(scene, kind, build_bvh=False)
        cpu.set_bvh(nodes, prims)
        cpu.refit(moved)
        want, _ = cpu.get_bvh()
        assert np.array_equal(got["bounds"], want["bounds"]), kind  # numeric equality: +0 == -0
        assert (want["first"] == nodes["first"]).all()
    # leaves hold their triangles exactly; parents are the union of their children
    p = moved.tris["p"]
    leaf = np.nonzero(got["prim_count"] > 0)[0]
    for i in leaf[:: max(1, len(leaf) // 200)]:
        v = p[prims[got["first"][i]: got["first"][i] + got["prim_count"][i]].astype(np.int64)].reshape(-1, 3)
        b = got["bounds"][i]
        assert np.allclose(b[0::2], v.min(0), rtol=0, atol=1e-4) and np.allclose(b[1::2], v.max(0), rtol=0, atol=1e-4)
        assert (b[0::2] <= v.min(0) + 1e-4).all() and (b[1::2] >= v.max(0) - 1e-4).all()
    inner = np.nonzero(got["prim_count"] == 0)[0]
    l, r = got["bounds"][got["first"][inner]], got["bounds"][got["first"][inner] + 1]
    assert np.array_equal(got["bounds"][inner][:, 0::2], np.minimum(l[:, 0::2], r[:, 0::2]))
    assert np.array_equal(got["bounds"][inner][:, 1::2], np.maximum(l[:, 1::2], r[:, 1::2]))


def test_refit_rejects_a_changed_topology(built):
    import vistrace_b200 as vt
    from vistrace_b200 import abi, scenes

    scene = scenes.scene_heightfield(12)
    nodes, prims = vt.build_bvh(scene)
    fewer = abi.SceneData(scene.tris[:-2], scene.materials, scene.entities)
    with pytest.raises(RuntimeError):
        vt.refit_bvh(fewer, nodes, prims[:-2])  # leaves address primitives past the end


@pytest.mark.parametrize("name", ["props_small", "foliage_small"])
def test_ploc_builder_reproduces_the_golden_reference_hierarchy(built, name):
    """vt_build_bvh_ploc against the hierarchy the UNMODIFIED reference built for the golden scenes (PLOC + LeafCollapser,
    stored by tests/golden/make_golden.py): node array and primitive indices bit for bit."""
    import vistrace_b200 as vt
    from conftest import load_golden

    scene, z = load_golden(name)
    nodes, prims = vt.build_bvh_ploc(scene)
    assert nodes.tobytes() == np.ascontiguousarray(z["nodes"]).tobytes()
    assert prims.tobytes() == np.ascontiguousarray(z["prim_indices"], np.uint64).tobytes()


@pytest.mark.parametrize("scene_name", ["heightfield", "props", "far_from_origin", "two_triangles", "one_triangle", "duplicates"])
def test_ploc_builder_reproduces_the_reference_hierarchy(built, oracle_mod, scene_name):
    """The same against the reference run here (oracle/_ref): ordinary scenes, map-scale coordinates, the degenerate sizes where
    the collapser turns the root into a leaf, and duplicated geometry (equal Morton codes and equal merge distances: every
    tie-break of the stable sort and of the neighbour search is exercised)."""
    import vistrace_b200 as vt
    from vistrace_b200 import abi, scenes

    if not oracle_mod.available("reference"):
        pytest.skip("needs oracle/_ref")
    if scene_name == "heightfield":
        scene = scenes.scene_heightfield(64)
    elif scene_name == "props":
        scene = scenes.scene_props(9, 21, 11, 12)
    elif scene_name == "far_from_origin":
        scene = scenes.scene_props(5, 15, 9, 8)
        scene.tris["p"] = scene.tris["p"] * np.float32(0.02) + np.array([16000.0, -15900.0, 8000.0], np.float32)
    elif scene_name == "duplicates":
        base = scenes.scene_props(3, 9, 7, 6)
        scene = abi.SceneData(np.concatenate([base.tris, base.tris, base.tris[::3]]), base.materials, base.entities)
    else:
        base = scenes.scene_heightfield(4)
        scene = abi.SceneData(base.tris[: 2 if scene_name == "two_triangles" else 1], base.materials, base.entities)
    ref = oracle_mod.CpuScene(scene, "reference", build_bvh=True)
    want_nodes, want_prims = ref.get_bvh()
    nodes, prims = vt.build_bvh_ploc(scene)
    assert len(nodes) == len(want_nodes) and nodes.tobytes() == want_nodes.tobytes()
    assert prims.tobytes() == want_prims.tobytes()
    if scene.n_tris > 2:  # and it is a hierarchy the engine accepts: every triangle once, depth within the stack
        flat = vt.flatten_bvh(nodes, prims, 0)
        assert sorted(flat["leaf_order"].tolist()) == list(range(scene.n_tris))


def test_ploc_builder_random_scenes(built, oracle_mod):
    """Seeded fuzz against the reference's build: triangle soups, duplicated and degenerate triangles, far-apart clusters, huge
    coordinates — identical arrays.  A scene with a zero-extent axis makes the REFERENCE convert NaN to unsigned (undefined
    behaviour, morton.hpp:52-57), so there only validity and determinism of our tree are checked."""
    import vistrace_b200 as vt
    from vistrace_b200 import abi

    rng = np.random.default_rng(123)
    kinds = ["soup", "duplicates", "degenerate", "clustered", "huge", "flat"]
    for it in range(90):
        kind = kinds[it % len(kinds)]
        n = int(rng.integers(1, 300))
        p = rng.uniform(-50, 50, (n, 3, 3)).astype(np.float32)
        if kind == "duplicates":
            p = np.concatenate([p, p[: n // 2], p[: n // 3]])
        elif kind == "degenerate":
            p[::3, 1] = p[::3, 0]
            p[::5, 2] = p[::5, 0]
        elif kind == "clustered":
            p = (p * 0.001 + rng.integers(0, 3, (len(p), 1, 1)) * 1000).astype(np.float32)
        elif kind == "huge":
            p = (p * 1e6).astype(np.float32)
        elif kind == "flat":
            p[:, :, 2] = 3.0
        tris = np.zeros(len(p), abi.TRI_IN)
        tris["p"] = p
        scene = abi.SceneData(tris)
        nodes, prims = vt.build_bvh_ploc(scene)
        if kind == "flat":
            again = vt.build_bvh_ploc(scene)
            assert nodes.tobytes() == again[0].tobytes() and prims.tobytes() == again[1].tobytes()
            assert sorted(prims.tolist()) == list(range(len(p)))
            leaves = nodes[nodes["prim_count"] > 0]
            assert int(leaves["prim_count"].sum()) == len(p)
            continue
        if not oracle_mod.available("reference"):
            continue
        want_nodes, want_prims = oracle_mod.CpuScene(scene, "reference", build_bvh=True).get_bvh()
        assert nodes.tobytes() == want_nodes.tobytes() and prims.tobytes() == want_prims.tobytes(), (it, kind, len(p))
