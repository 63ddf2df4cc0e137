"""Rebuild vs refit on a large scene (run under gpurun): python tools/refit_time.py [--quads 1582] [--props 64]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import vistrace_b200 as vt  # noqa: E402
from vistrace_b200 import abi, scenes  # noqa: E402
from test_host import _moved_props  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quads", type=int, default=1582)
    ap.add_argument("--props", type=int, default=64)
    args = ap.parse_args()
    scene = scenes.scene_terrain_closed(args.quads, n_props=args.props)
    moved = _moved_props(scene)
    rays = scenes.pinhole_rays(1920, 1080, (0, -330, 200), (0, 0, 10))
    accel = vt.Accel(0)
    t0 = time.time(); accel.populate(scene); t_build = time.time() - t0
    h0 = accel.traverse(rays)
    os.environ["VT_REFIT_DEVICE"] = "0"
    t0 = time.time(); accel.refit(moved); t_refit_host = time.time() - t0
    h_host = accel.traverse(rays)
    os.environ["VT_REFIT_DEVICE"] = "1"
    accel.refit(scene)  # back (first device refit: prepares parent / slot tables)
    t0 = time.time(); accel.refit(moved); t_refit = time.time() - t0
    t0 = time.time(); accel.refit(moved); t_refit2 = time.time() - t0
    h1 = accel.traverse(rays)
    assert h1.tobytes() == h_host.tobytes() or int(((h1["prim"] != h_host["prim"]) | (h1["t"] != h_host["t"])).sum()) < 10
    st1 = accel.traverse_stats(rays)
    t0 = time.time(); fresh = vt.Accel(0).populate(moved); t_rebuild = time.time() - t0
    h2 = fresh.traverse(rays)
    st2 = fresh.traverse_stats(rays)
    differ = int(((h1["prim"] != h2["prim"]) | (h1["t"] != h2["t"])).sum())
    # one moved entity through vt_accel_refit_range: the ranged walk (touched quads + ancestors) against the whole-tree walk
    ent = len(scene.entities) - 1
    idx = np.nonzero(scene.tris["ent_idx"] == ent)[0]
    run_a, run_b = scene.tris[idx[0]: idx[-1] + 1].copy(), moved.tris[idx[0]: idx[-1] + 1].copy()  # refit_range writes into the handle's scene
    ranged = {}
    wall = {"1": [], "0": []}
    os.environ["VT_TIMING"] = "1"  # the library prints the device time of the walk itself to stderr
    for k in range(12):  # the two walks alternate; the last call (k = 11) leaves the moved run in place
        walk = "1" if k % 4 < 2 else "0"
        os.environ["VT_REFIT_RANGE_WALK"] = walk
        t0 = time.time(); accel.refit_range(run_b if k % 2 else run_a, int(idx[0])); wall[walk].append(time.time() - t0)
    del os.environ["VT_REFIT_RANGE_WALK"], os.environ["VT_TIMING"]
    ranged["ranged_walk_call_ms"] = round(float(np.median(wall["1"])) * 1e3, 3)
    ranged["whole_tree_walk_call_ms"] = round(float(np.median(wall["0"])) * 1e3, 3)
    ranged["triangles"] = int(len(idx))
    assert accel.traverse(rays).tobytes() == h1.tobytes()  # k = 5 put the moved run back
    print(json.dumps({"tris": scene.n_tris, "refit_range_one_entity": ranged, "populate_s": round(t_build, 3), "refit_host_s": round(t_refit_host, 3), "refit_device_s": round(t_refit, 3), "refit_device_again_s": round(t_refit2, 3), "rebuild_moved_s": round(t_rebuild, 3),
                      "rays_changed_by_the_move": int(((h0["prim"] != h1["prim"]) | (h0["t"] != h1["t"])).sum()),
                      "refit_vs_rebuild_records_differing": differ,
                      "node_visits_per_ray_refit": round(st1[0] / len(rays), 2), "node_visits_per_ray_rebuild": round(st2[0] / len(rays), 2)}))


if __name__ == "__main__":
    main()
