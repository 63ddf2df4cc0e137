"""CPU emulation of the quad kernel's walk (development tool, not product code): the decoded planes of vt_build_quads, children ordered by
entry distance or by entry + exit, nearest first, the others pushed farthest first, the canonical tie rule — and the triangle test
(alpha test included) through the checker's tri_intersect.  Counts quad visits and triangle tests per ray for the SAH-optimal and the
greedy collapse; it reproduces the GPU's vt_accel_traverse_stats counters to about 1 % and is how profiles/r2_child_order.md found
why config 4's camera rays got slower under the SAH-optimal collapse (--trace prints one ray's walk).
usage: python tools/quad_walk_emulator.py {foliage N_CARDS | terrain N_QUADS | props N_PROPS} [--trace]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
import vistrace_b200 as vt  # noqa: E402
from vistrace_b200 import binding, scenes  # noqa: E402

KIND = sys.argv[1] if len(sys.argv) > 1 else "foliage"
NC = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
TRACE = "--trace" in sys.argv
if KIND == "foliage":
    scene, eye, look = scenes.scene_foliage(n_cards=NC, tex_size=256, ground_quads=64), (0, -48, 20), (0, 0, 8)
elif KIND == "terrain":
    scene, eye, look = scenes.scene_terrain_closed(NC), (0, -330 * NC / 1582, 200 * NC / 1582), (0, 0, 10)
else:
    scene, eye, look = scenes.scene_props(NC, 31, 15, 16), (0, -95, 40), (0, 0, 10)
primary = scenes.pinhole_rays(48, 27, eye, look)
nodes, prims = vt.build_bvh(scene)
cpu = oracle.CpuScene(scene, "reference" if oracle.available("reference") else "port", build_bvh=False)
cpu.set_bvh(nodes, prims)
want = cpu.traverse(primary, want_stats=True)
print("binary tree, reference order: steps %.1f tests %.1f per primary ray" % (want["steps"] / len(primary), want["isects"] / len(primary)), flush=True)
full = cpu.traverse(scenes.pinhole_rays(192, 108, eye, look), want_attrs=True)
bounce = scenes.bounce_rays(full["attrs"], spp=1, key=3)[0]
bounce = bounce[np.random.default_rng(1).choice(len(bounce), min(1300, len(bounce)), replace=False)]
want_b = cpu.traverse(bounce)
OFF = float(vt.quad_plane_offset())


def prepare(q):
    quads, order = q["quads"], q["leaf_order"]
    plane = (OFF + quads["q"].astype(np.float64)) * quads["scale"].astype(np.float64)[:, :, None, None] + quads["origin_adj"].astype(np.float64)[:, :, None, None]
    lo, hi = plane[:, :, 0, :].transpose(0, 2, 1).copy(), plane[:, :, 1, :].transpose(0, 2, 1).copy()  # [quad, child, axis]
    refs = quads["ref"]
    valid = (refs != 0xFFFFFFFF) & ((refs >> 28 == 0) | ((refs & 0x0FFFFFFF) < scene.n_tris))  # not the sentinel leaf of an empty slot
    return lo, hi, refs, valid, order


def walk(Q, rays, mode, trace=None):
    lo, hi, refs, valid, order = Q
    visits = tests = 0
    out = []
    for ri, r in enumerate(rays):
        o, d = r["o"].astype(np.float64), r["d"].astype(np.float64)
        with np.errstate(divide="ignore"):
            inv = 1.0 / d
        tmin, best, bp = float(r["tmin"]), float(r["tmax"]), -1
        rr = r.copy()
        stack = [0]
        while stack:
            ref = stack.pop()
            cnt = ref >> 28
            if cnt:
                s0 = ref & 0x0FFFFFFF
                for s in range(s0, s0 + cnt):
                    tests += 1
                    rr["tmax"] = best
                    ok, tuv = cpu.tri_intersect(int(order[s]), rr)
                    if ok and (tuv[0] < best or bp < 0 or order[s] > bp):  # equal t: the larger original index wins
                        best, bp = float(tuv[0]), int(order[s])
                continue
            visits += 1
            with np.errstate(invalid="ignore"):
                t0, t1 = (lo[ref] - o) * inv, (hi[ref] - o) * inv
            near, far = np.minimum(t0, t1).max(axis=1), np.maximum(t0, t1).min(axis=1)
            entry, exit_ = np.maximum(near, tmin), np.minimum(far, best)
            hit = valid[ref] & (entry <= exit_)
            key = entry.astype(np.float32) if mode == "entry" else (entry + exit_).astype(np.float32)
            idx = [c for c in np.lexsort((np.arange(4), key)) if hit[c]]
            if trace is not None and ri == trace and visits < 40:
                print("   quad", ref, "hit", hit.tolist(), "entry", np.round(entry, 2).tolist(), "far", np.round(far, 1).tolist(), "order", [int(c) for c in idx])
            for c in reversed(idx):
                stack.append(int(refs[ref, c]))
        out.append(bp)
    return round(visits / len(rays), 2), round(tests / len(rays), 2), out


def agree(p, w):
    return round(float(np.mean(np.array(p, np.int64) == np.where(w["hits"]["prim"] == 0xFFFFFFFF, -1, w["hits"]["prim"].astype(np.int64)))), 4)


for collapse in ("greedy", "dp"):
    os.environ["VT_COLLAPSE"] = collapse
    Q = prepare(binding.build_quads(nodes, prims))
    for mode in ("entry", "mid"):
        v, t, p = walk(Q, primary, mode)
        vb, tb, pb = walk(Q, bounce, mode)
        print(KIND, NC, "collapse", collapse, "order", mode, "| primary: visits", v, "tests", t, "same prim as the reference", agree(p, want),
              "| bounce: visits", vb, "tests", tb, "same prim", agree(pb, want_b), flush=True)
    if TRACE:
        print("trace of primary ray 141 //", collapse)
        walk(Q, primary[141:142], "entry", trace=0)
