"""Reinsertion optimisation (vt_optimize_bvh / VT_REINSERT) on the CPU: SAH inner-node area and the reference traverser's own step / test
counters per ray on the product builder's tree before and after, per scene kind; hit buffers must be identical up to exact ties.
usage: python tools/reinsert_probe.py [--quads 400] [--iterations 1 2 4 8]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
import vistrace_b200 as vt  # noqa: E402
from vistrace_b200 import binding, scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quads", type=int, default=400)
    ap.add_argument("--iterations", type=int, nargs="+", default=[1, 2, 4, 8])
    ap.add_argument("--fraction", type=float, default=0.05)
    args = ap.parse_args()
    kind = "reference" if oracle.available("reference") else "port"
    cases = {
        "terrain": (scenes.scene_terrain_closed(args.quads), (0, -330 * args.quads / 1582, 200 * args.quads / 1582), (0, 0, 10)),
        "props": (scenes.scene_props(16, 31, 15, 12), (0, -95, 40), (0, 0, 10)),
        "foliage": (scenes.scene_foliage(n_cards=20000, extent=20.0, tex_size=64, ground_quads=8), (0, -9.6, 20), (0, 0, 8)),
    }
    for name, (scene, eye, look) in cases.items():
        primary = scenes.pinhole_rays(320, 180, eye, look)
        nodes, prims = vt.build_bvh(scene)
        cpu = oracle.CpuScene(scene, kind, build_bvh=False)
        cpu.set_bvh(nodes, prims)
        base = cpu.traverse(primary, want_attrs=True, want_stats=True)
        ok = base["hits"]["prim"] != 0xFFFFFFFF
        # incoherent rays: cosine bounce off the primary hits (as the bench's second wave)
        accel_rays = scenes.bounce_rays(base["attrs"], spp=2, key=3)[0]
        base_b = cpu.traverse(accel_rays, want_stats=True)
        row = {"scene": name, "tris": int(scene.n_tris), "nodes": int(len(nodes)),
               "steps_per_primary_ray": round(base["steps"] / len(primary), 2), "tests_per_primary_ray": round(base["isects"] / len(primary), 2),
               "steps_per_incoherent_ray": round(base_b["steps"] / len(accel_rays), 2), "tests_per_incoherent_ray": round(base_b["isects"] / len(accel_rays), 2)}
        print(json.dumps(row), flush=True)
        for it in args.iterations:
            t0 = time.time()
            opt, before, after, moves = binding.optimize_bvh(nodes, it, args.fraction)
            dt = time.time() - t0
            cpu.set_bvh(opt, prims)
            a = cpu.traverse(primary, want_stats=True)
            b = cpu.traverse(accel_rays, want_stats=True)
            differ = int(((a["hits"]["prim"] != base["hits"]["prim"]) | (a["hits"]["t"] != base["hits"]["t"])).sum())
            ties = int(((a["hits"]["prim"] != base["hits"]["prim"]) & (a["hits"]["t"] == base["hits"]["t"])).sum())
            print(json.dumps({"scene": name, "iterations": it, "seconds": round(dt, 2), "moves": int(moves), "area_ratio": round(after / before, 4),
                              "steps_per_primary_ray": round(a["steps"] / len(primary), 2), "tests_per_primary_ray": round(a["isects"] / len(primary), 2),
                              "steps_per_incoherent_ray": round(b["steps"] / len(accel_rays), 2), "tests_per_incoherent_ray": round(b["isects"] / len(accel_rays), 2),
                              "primary_records_differing": differ, "of_which_exact_ties": ties}), flush=True)


if __name__ == "__main__":
    main()
