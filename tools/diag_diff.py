"""Diagnostic (GPU): config 3 at full size, quantised layout vs exact layout; dumps every differing record
(ray, both hits) to gpurun_out/diag_diff.npz and checks both primitives with the checker's triangle test."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402  (diagnostic tool, not product)
import vistrace_b200 as vt  # noqa: E402
from vistrace_b200 import scenes  # noqa: E402

layout = sys.argv[1] if len(sys.argv) > 1 else "quad"
scene = scenes.scene_terrain_closed(1582)
rays = scenes.pinhole_rays(1920, 1080, (0.0, -330.0, 200.0), (0.0, 0.0, 10.0))
exact = vt.Accel(0, layout="exact").populate(scene)
bvh = exact.get_bvh()
want = exact.trace_diffuse_wave(rays, 4, seed=11, want_bounce_rays=True)
live = want["bounce_rays"]["tmax"] >= 0
exact.close()
accel = vt.Accel(0, layout=layout).populate(scene, bvh=bvh)
out = {}
cpu = oracle.CpuScene(scene, "reference" if oracle.available("reference") else "port", build_bvh=False)
for name, rr, ww in (("primary", rays, want["hits"]), ("bounce", want["bounce_rays"][live], want["bounce_hits"][live])):
    got = accel.traverse(rr)
    diff = np.nonzero((got.view(np.uint32).reshape(-1, 4) != ww.view(np.uint32).reshape(-1, 4)).any(1))[0]
    print(name, "differ:", len(diff))
    out[name + "_rays"], out[name + "_exact"], out[name + "_" + layout] = rr[diff], ww[diff], got[diff]
    for i in diff[:16]:
        print(" ray", rr[i], "\n   exact", ww[i], "\n   ", layout, got[i])
        for p in {int(ww[i]["prim"]), int(got[i]["prim"])}:
            if p != 0xFFFFFFFF:
                print("    checker tri_intersect prim", p, cpu.tri_intersect(p, rr[i]))
os.makedirs("gpurun_out", exist_ok=True)
np.savez("gpurun_out/diag_diff.npz", **out)
