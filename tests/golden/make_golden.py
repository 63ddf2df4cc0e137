"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container (needs /root/reference -> oracle/_ref/libvt_ref.so):

    python tests/golden/make_golden.py

Every array in the .npz files is either an input (scene, hierarchy, rays) or an output of the
reference's own code run through oracle/ref_harness.cpp.  The fixtures travel to the GPU box, where
/root/reference does not exist; tests compare the C restatement (oracle/vt_oracle.c) and the CUDA
path against them bit for bit.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402
from vistrace_b200 import abi, scenes  # noqa: E402


def pack_scene(sc):
    d = {"tris": sc.tris, "materials": sc.materials, "entities": sc.entities, "n_textures": np.array(len(sc.textures))}
    for i, (w, h, m, fl, px) in enumerate(sc.textures):
        d[f"tex{i}_hdr"] = np.array([w, h, m, fl], np.int64)
        d[f"tex{i}_px"] = px
    return d


def golden_scene(name, sc, rays, extra_rays=None):
    ref = oracle.CpuScene(sc, "reference")
    nodes, prims = ref.get_bvh()
    out = pack_scene(sc)
    out.update(nodes=nodes, prim_indices=prims, tri_derived=ref.tri_derived())
    r = ref.traverse(rays, want_attrs=True, want_stats=True)
    out.update(rays=rays, hits=r["hits"], attrs=r["attrs"], stats=np.array([r["steps"], r["isects"]], np.uint64))
    bounce, _ = scenes.bounce_rays(r["attrs"], spp=2, key=11)
    rb = ref.traverse(bounce, want_attrs=True)
    out.update(bounce_rays=bounce, bounce_hits=rb["hits"], bounce_attrs=rb["attrs"])
    if extra_rays is not None:
        re = ref.traverse(extra_rays, want_attrs=True)
        out.update(extra_rays=extra_rays, extra_hits=re["hits"], extra_attrs=re["attrs"])
    if sc.textures:
        rng = np.random.default_rng(5)
        uvm = np.concatenate([rng.uniform(-2.5, 2.5, (4000, 2)), rng.uniform(-1.0, 9.0, (4000, 1))], 1).astype(np.float32)
        uvm[:500, 2] = 0.0
        uvm[500:600, :2] = np.array([[0.0, 0.0], [1.0, 1.0], [-1e-9, 0.5], [0.5, -1e-9], [0.9999, 0.9999]] * 20, np.float32)
        out["tex_uvm"] = uvm
        for i in range(len(sc.textures)):
            out[f"tex{i}_samples"] = ref.sample(i, uvm)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: (v.shape, str(v.dtype)[:12]) for k, v in out.items() if hasattr(v, "shape")})


def golden_skin():
    """SkinTriangle (source/objects/AccelStruct.cpp:66-108): weighted multi-bone case + the one-bone overload."""
    tris, skin, bones, binds = scenes.skin_case()
    out = {"tris": tris, "skin": skin, "bones": bones, "binds": binds,
           "skinned": oracle.skin_triangles(tris, skin, bones, binds, "reference"),
           "skinned_one_bone": oracle.skin_triangles(tris, None, bones[:1], binds[:1], "reference")}
    np.savez_compressed(os.path.join(HERE, "skin_small.npz"), **out)
    print("skin_small", {k: v.shape for k, v in out.items()})


def golden_vtf():
    """Synthetic VTF files (tests/test_vtf.py: make_vtf) and every texel of them as the reference's VTFTexture reports it."""
    sys.path.insert(0, os.path.dirname(HERE))
    from test_vtf import make_vtf

    cases = {"dxt1_low": ("DXT1", 32, 32, 6, {"low": (16, 16)}), "dxt3_odd": ("DXT3", 20, 12, 3, {"seed": 4}),
             "dxt5_res": ("DXT5", 16, 16, 5, {"minor": 3, "resources": True, "low": (16, 16), "seed": 9}),
             "dxt1a_tiny": ("DXT1_ONEBITALPHA", 4, 4, 3, {"seed": 6}), "bgr888": ("BGR888", 16, 8, 5, {}),
             "ia88_v71": ("IA88", 8, 4, 4, {"minor": 1}), "argb8888": ("ARGB8888", 8, 8, 4, {}), "a8": ("A8", 8, 8, 4, {})}
    out = {}
    for name, (fmt, w, h, mips, kw) in cases.items():
        data = make_vtf(fmt, w, h, mips, **kw)
        n_tex = sum(max(1, w >> m) * max(1, h >> m) for m in range(mips))
        want = oracle.vtf_pixels(data, n_tex)
        assert want is not None, name
        out["file_" + name] = np.frombuffer(data, np.uint8)
        out["want_" + name] = want
    np.savez_compressed(os.path.join(HERE, "vtf_small.npz"), **out)
    print("vtf_small", {k: v.shape for k, v in out.items() if k.startswith("want_")})


def golden_cornell():
    """The bvh library's own golden-image test (libs/bvh/test/CMakeLists.txt:57-82): the Cornell box of libs/bvh/test/scene,
    triangulated as libs/bvh/test/obj.hpp:57-95 does (fans), and the reference picture every builder must reproduce."""
    from PIL import Image

    root = "/root/reference/libs/bvh/test/scene"
    verts, tris = [], []
    for line in open(os.path.join(root, "cornell_box.obj")):
        tok = line.split()
        if not tok or tok[0].startswith("#"):
            continue
        if tok[0] == "v":
            verts.append([np.float32(x) for x in tok[1:4]])
        elif tok[0] == "f":
            idx = [int(t.split("/")[0]) for t in tok[1:]]
            pts = [verts[len(verts) + i if i < 0 else i - 1] for i in idx]
            p0, p1 = pts[0], pts[1]
            for v in pts[2:]:
                tris.append([p0, p1, v])
                p1 = v
    image = np.asarray(Image.open(os.path.join(root, "cornell_box_reference.png")).convert("RGB"), np.uint8)
    np.savez_compressed(os.path.join(HERE, "cornell_box.npz"), p=np.asarray(tris, np.float32), image=image)
    print("cornell_box", len(tris), "triangles, image", image.shape)


def main():
    assert oracle.available("reference"), "build oracle/_ref first: make -C oracle ref"
    if sys.argv[1:] == ["cornell"]:
        return golden_cornell()
    if sys.argv[1:] == ["skin"]:
        return golden_skin()
    if sys.argv[1:] == ["vtf"]:
        return golden_vtf()
    # foliage: alpha test (wrap + clamp textures), two-sided cards, one-sided ground, sky room
    sc = scenes.scene_foliage(n_cards=400, tex_size=32, ground_quads=8, seed=7)
    rays = scenes.pinhole_rays(96, 54, (0, -48, 20), (0, 0, 8))
    rnd = scenes.random_rays(3000, (-45, -45, -3), (45, 45, 50), seed=2)
    rnd["tmax"][::3] = 30.0
    rnd["tmin"][::4] = 2.0
    rnd["d"][::5] *= 2.5
    golden_scene("foliage_small", sc, rays, rnd)
    # props: multi-entity, per-entity colours and ids, shadow-style rays
    sc = scenes.scene_props(6, 15, 9, 8)
    rays = scenes.pinhole_rays(96, 54, (0, -95, 40), (0, 0, 10))
    golden_scene("props_small", sc, rays)
    # known answers of the vendored bvh tests
    # libs/bvh/test/node_intersectors.cpp:18-36 — flat box z in [2.1, 2.1], direction (0, -0, 1): must intersect
    node = np.zeros(1, abi.NODE)
    node["bounds"] = (-1, 1, -1, 1, 2.1, 2.1)
    ray = np.zeros(1, abi.RAY)
    ray["o"], ray["d"], ray["tmin"], ray["tmax"] = (0.25, 0.25, 0.0), (0.0, -0.0, 1.0), 0.0, 100.0
    entry, exit_ = oracle.node_intersect(node, ray, "reference")
    np.savez(os.path.join(HERE, "kat_node_intersect.npz"), node=node, ray=ray, entry_exit=np.array([entry, exit_], np.float32))
    print("kat node", entry, exit_)
    golden_skin()
    golden_vtf()
    golden_cornell()


if __name__ == "__main__":
    main()
