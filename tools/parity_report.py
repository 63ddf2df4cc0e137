"""Correctness report at FULL BASELINE sizes: the shipped configuration (quad layout, product tree)
against the UNMODIFIED reference traversing the same tree on the host, every ray of the wave, per config.

usage (under gpurun): python tools/parity_report.py [--configs 1,2,4,3,5] [--out gpurun_out/parity.jsonl]
       python tools/parity_report.py --selftest     (no GPU: the C port stands in for the engine on a small scene)

Per config: rays, hit/miss mismatches, primitive mismatches split into exact ties (dt = 0), near ties (|dt| / t < 1e-6) and
other, bit mismatches of t/u/v where the primitive agrees, largest relative error of every vt_attr float field (primary rays).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
import vistrace_b200 as vt  # noqa: E402
from vistrace_b200 import abi, scenes  # noqa: E402

FLOAT_FIELDS = [f for f in abi.ATTR.names if abi.ATTR[f].base == np.float32 or abi.ATTR[f].subdtype and abi.ATTR[f].subdtype[0] == np.float32]


def compare(got, want, rays=None, cpu=None):
    """classify_hits (vistrace_b200/report.py) + the column names of SURVEY.md section 8d."""
    from vistrace_b200.report import classify_hits

    rep = classify_hits(got, want, rays, cpu.tri_intersect if cpu is not None else None)
    miss_g, miss_w = got["prim"] == abi.VT_MISS, want["prim"] == abi.VT_MISS
    both = ~miss_g & ~miss_w
    rep["hit_miss_mismatch"] = int((miss_g != miss_w).sum())
    rep["prim_mismatch"] = int((both & (got["prim"] != want["prim"])).sum())
    # "other" = a primitive / hit-miss difference that is neither an exact tie nor a checker-verified leak: a bug if non-zero
    rep["other"] = rep["lost"] + rep["unverified"]
    rep["tuv_bit_mismatch_same_prim"] = rep["tuv_bits"]
    return rep


def attr_err(got, want):
    ok = (got["prim"] != abi.VT_MISS) & (want["prim"] != abi.VT_MISS) & (got["prim"] == want["prim"])
    out = {}
    for f in FLOAT_FIELDS:
        x, y = got[f][ok].astype(np.float64), want[f][ok].astype(np.float64)
        out[f] = float(np.max(np.abs(x - y) / np.maximum(np.abs(y), 1e-6))) if x.size else 0.0
    return {"max_rel_err": round(max(out.values()), 9) if out else 0.0, "worst_field": max(out, key=out.get) if out else "", "records": int(ok.sum())}


def config(cfg):
    if cfg == 1:
        return "config1 100k height field", scenes.scene_heightfield(224), (1920, 1080), ((0, -80, 60), (0, 0, 5)), None
    if cfg == 2:
        return "config2 1M skinned props + shadow", scenes.scene_props_skinned(256, 63, 31, 64), (1920, 1080), ((0, -95, 40), (0, 0, 10)), "shadow"
    if cfg == 3:
        return "config3 5M terrain + 4 spp bounce", scenes.scene_terrain_closed(1582), (1920, 1080), ((0.0, -330.0, 200.0), (0.0, 0.0, 10.0)), "bounce4"
    if cfg == 4:
        return "config4 1M alpha-tested foliage + bounce", scenes.scene_foliage(n_cards=500000, tex_size=256, ground_quads=64), (1920, 1080), ((0, -48, 20), (0, 0, 8)), "bounce1"
    return "config5 20M terrain + props, 4K: primary + shadow + bounce", scenes.scene_terrain_closed(2980, n_props=143), (3840, 2160), ((0.0, -330.0, 200.0), (0.0, 0.0, 10.0)), "shadow+bounce1"


def run(cfg, engine_factory, kind, small=False, builder="product"):
    """builder = "product": the engine's own tree, handed to the checker as well (same hierarchy on both sides);
    builder = "ploc": the engine rebuilds the reference's PLOC + LeafCollapser tree (vt_build_bvh_ploc) while the reference
    checker builds ITS OWN — nothing is handed over."""
    t0 = time.time()
    name, scene, (W, H), cam, secondary = config(cfg)
    if small:
        W, H = W // 8, H // 8
    if builder == "ploc":
        bvh = vt.build_bvh_ploc(scene)
        cpu = oracle.CpuScene(scene, kind, build_bvh=(kind == "reference"))
        if kind != "reference":
            cpu.set_bvh(*bvh)
    else:
        bvh = vt.build_bvh(scene)
        cpu = oracle.CpuScene(scene, kind, build_bvh=False)
        cpu.set_bvh(*bvh)
    engine = engine_factory(scene, bvh)
    rays = scenes.pinhole_rays(W, H, *cam)
    hits, attrs = engine(rays, True)
    want = cpu.traverse(rays, want_attrs=True)
    res = {"config": cfg, "name": name, "n_tris": int(scene.n_tris), "checker": kind, "builder": builder, "primary": compare(hits, want["hits"], rays, cpu),
           "primary_attrs": attr_err(attrs, want["attrs"]), "byte_identical": hits.tobytes() == want["hits"].tobytes()}
    waves = []
    if secondary:
        for part in secondary.split("+"):
            if part == "shadow":
                sec = scenes.shadow_rays(want["attrs"])
                sec = sec[0] if isinstance(sec, tuple) else sec
            else:
                sec, _ = scenes.bounce_rays(want["attrs"], spp=int(part[6:]))
            sec = np.ascontiguousarray(sec[sec["tmax"] >= 0])
            g = engine(sec, False)
            w = cpu.traverse(sec)["hits"]
            waves.append({part: compare(g, w, sec, cpu)})  # shadow rays too are traced closest-hit here: every field is defined
            res["byte_identical"] = res["byte_identical"] and g.tobytes() == w.tobytes()
    res["secondary"] = waves
    res["seconds"] = round(time.time() - t0, 1)
    cpu.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,4,3,5")
    ap.add_argument("--out", default="")
    ap.add_argument("--selftest", action="store_true")
    args = ap.parse_args()
    if args.selftest:  # the C port plays the engine, the reference (or the port again) checks it: exercises the report code
        def factory(scene, bvh):
            port = oracle.CpuScene(scene, "port", build_bvh=False)
            port.set_bvh(*bvh)

            def engine(rays, want_attrs):
                r = port.traverse(rays, want_attrs=want_attrs)
                return (r["hits"], r["attrs"]) if want_attrs else r["hits"]
            return engine
        kind = "reference" if oracle.available("reference") else "port"
        print(json.dumps(run(1, factory, kind, small=True)))
        return
    kind = "reference" if oracle.available("reference") else "port"

    def factory(scene, bvh):
        accel = vt.Accel(0).populate(scene, bvh=bvh)

        def engine(rays, want_attrs):
            return accel.traverse(rays, want_attrs=True) if want_attrs else accel.traverse(rays)
        return engine

    for cfg in [int(c) for c in args.configs.split(",")]:
        res = run(cfg, factory, kind)
        line = json.dumps(res)
        print(line, flush=True)
        if args.out:
            with open(args.out, "a") as f:
                f.write(line + "\n")


if __name__ == "__main__":
    main()
