"""Rebuild latency: vt_accel_populate phase by phase (VT_TIMING=1 lines on stderr) next to the reference's own
build sequence (oracle/_ref: PLOC + LeafCollapser, source/objects/AccelStruct.cpp:762-770) on the same host.
usage: python tools/build_time.py [--quads 1582]   (under gpurun: populate needs the GPU)"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["VT_TIMING"] = "1"
import oracle  # noqa: E402
import vistrace_b200 as vt  # noqa: E402
from vistrace_b200 import scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--quads", type=int, default=1582)
ap.add_argument("--props", type=int, default=0)
args = ap.parse_args()
scene = scenes.scene_terrain_closed(args.quads, n_props=args.props)
out = {"n_tris": int(scene.n_tris), "host_threads": os.cpu_count()}
for name in ("product", "ploc"):
    if name == "ploc":
        os.environ["VT_BUILDER"] = "ploc"
    best = 1e9
    for _ in range(2):
        a = vt.Accel(0)
        t0 = time.time()
        a.populate(scene)
        best = min(best, time.time() - t0)
        a.close()
    out[f"populate_{name}_s"] = round(best, 3)
os.environ.pop("VT_BUILDER", None)
if oracle.available("reference"):
    best = 1e9
    for _ in range(2):
        t0 = time.time()
        cpu = oracle.CpuScene(scene, "reference", build_bvh=True)  # Triangle ctor + PLOC + LeafCollapser + traverser set-up
        best = min(best, time.time() - t0)
        out["reference_threads"] = cpu.max_threads
        cpu.close()
    out["reference_populate_s"] = round(best, 3)
print(json.dumps(out))
