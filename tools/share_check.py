"""Tail work-sharing on / off must give byte-identical hit buffers (the canonical tie rule makes the answer independent of the
visit order).  Two engines over the same scene, primary + bounce waves of the bench frame and of small shards.
usage: python tools/share_check.py [--quads 1582] [--seeds 3]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vistrace_b200 as vt  # noqa: E402
from vistrace_b200 import scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--quads", type=int, default=1582)
ap.add_argument("--seeds", type=int, default=3)
args = ap.parse_args()
scene = scenes.scene_terrain_closed(args.quads)
rays = scenes.pinhole_rays(1920, 1080, (0.0, -330.0, 200.0), (0.0, 0.0, 10.0))
os.environ["VT_TAIL_SHARE"] = "0"
off = vt.Accel(0).populate(scene)
os.environ["VT_TAIL_SHARE"] = "1"
on = vt.Accel(0).populate(scene, bvh=off.get_bvh())
out = {"waves": []}


def check(name, r):
    a, b = off.traverse(r), on.traverse(r)
    bad = np.nonzero(a.view(np.uint8).reshape(len(a), -1).any(1) & (a.tobytes() != b.tobytes()) & (a != b))[0] if a.tobytes() != b.tobytes() else []
    rec = {"wave": name, "rays": int(len(r)), "differing": int(len(bad)), "invalid_off": int(off.invalid_rays), "invalid_on": int(on.invalid_rays)}
    if len(bad):
        rec["first"] = [{"i": int(i), "off": [float(a["t"][i]), int(a["prim"][i])], "on": [float(b["t"][i]), int(b["prim"][i])]} for i in bad[:5]]
    out["waves"].append(rec)
    return a


hits, attrs = off.traverse(rays, want_attrs=True)
check("primary", rays)
for cnt in (1, 7, 33, 100, 1000, 4097, 33333):  # launches that run dry at once, lane counts that are not a multiple of 32
    check(f"primary, first {cnt} horizon rays", np.ascontiguousarray(rays[1920 * 840:1920 * 840 + cnt]))
for w in (8, 64):
    idx = vt.shard_indices(len(rays), w, 0, 8192)
    check(f"primary shard 1/{w}", np.ascontiguousarray(rays[idx]))
for seed in range(args.seeds):
    brays, _ = off.bounce_rays(attrs, 4, seed=1003 + seed)
    live = np.ascontiguousarray(brays[brays["tmax"] >= 0])
    check(f"bounce seed {seed}", live)
    check(f"bounce seed {seed} first 700k", np.ascontiguousarray(live[:700000]))
print(json.dumps(out))
