"""Experiment (CPU only): would a reinsertion pass (Meister & Bittner; the library's ParallelReinsertionOptimizer) on top of the
product's binned-SAH tree cut traversal work?  Counts the reference traverser's steps and triangle tests per ray before and after.
usage: python tools/reinsertion_probe.py [--quads 700] [--props 0]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
import vistrace_b200 as vt  # noqa: E402
from vistrace_b200 import scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--quads", type=int, default=700)
ap.add_argument("--props", type=int, default=0)
ap.add_argument("--foliage", type=int, default=0)
args = ap.parse_args()
if args.foliage:
    scene, cam = scenes.scene_foliage(n_cards=args.foliage, tex_size=64, ground_quads=32), ((0, -48, 20), (0, 0, 8))
else:
    scene, cam = scenes.scene_terrain_closed(args.quads, n_props=args.props), ((0, -330, 200), (0, 0, 10))
rays = scenes.pinhole_rays(480, 270, *cam)
for name, tree in (("product binned SAH", vt.build_bvh(scene)), ("PLOC + LeafCollapser", vt.build_bvh_ploc(scene))):
    cpu = oracle.CpuScene(scene, "reference", build_bvh=False)
    cpu.set_bvh(*tree)
    r = cpu.traverse(rays, want_attrs=True, want_stats=True)
    bounce, _ = scenes.bounce_rays(r["attrs"], spp=2)
    bounce = bounce[bounce["tmax"] >= 0]
    b = cpu.traverse(bounce, want_stats=True)
    print(f"{name}: {scene.n_tris} tris, {len(tree[0])} nodes | primary steps/tests per ray {r['steps'] / len(rays):.2f} / {r['isects'] / len(rays):.2f} | "
          f"bounce {b['steps'] / len(bounce):.2f} / {b['isects'] / len(bounce):.2f}", flush=True)
    t0 = time.time()
    c0, c1 = cpu.reinsertion_optimize()
    dt = time.time() - t0
    r2 = cpu.traverse(rays, want_stats=True)
    b2 = cpu.traverse(bounce, want_stats=True)
    assert r2["hits"]["t"].tobytes() == r["hits"]["t"].tobytes()
    print(f"   + reinsertion ({dt:.1f} s, SAH cost {c0:.1f} -> {c1:.1f}): primary {r2['steps'] / len(rays):.2f} / {r2['isects'] / len(rays):.2f} | "
          f"bounce {b2['steps'] / len(bounce):.2f} / {b2['isects'] / len(bounce):.2f}", flush=True)
