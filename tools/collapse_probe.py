"""Which wide-node collapse for which scene (run under gpurun): primary and bounce K1 time, quad visits and triangle tests per ray with the
SAH-optimal plan (VT_COLLAPSE=dp) and the round-1 rule (greedy), next to the sibling-overlap figure VT_COLLAPSE=auto decides on.
usage: python tools/collapse_probe.py [--out file.jsonl]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vistrace_b200 as vt  # noqa: E402
from vistrace_b200 import scenes  # noqa: E402
from gpu_explore import time_traverse, to_dev  # noqa: E402


def sibling_overlap(nodes):
    inner = np.nonzero(nodes["prim_count"] == 0)[0]
    f = nodes["first"][inner]
    b = nodes["bounds"].astype(np.float64)
    e = np.minimum(b[f][:, 1::2], b[f + 1][:, 1::2]) - np.maximum(b[f][:, 0::2], b[f + 1][:, 0::2])
    ov = np.where((e >= 0).all(axis=1), e[:, 0] * e[:, 1] + e[:, 1] * e[:, 2] + e[:, 2] * e[:, 0], 0.0)
    ee = b[inner][:, 1::2] - b[inner][:, 0::2]
    return float(ov.sum() / (ee[:, 0] * ee[:, 1] + ee[:, 1] * ee[:, 2] + ee[:, 2] * ee[:, 0]).sum())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    cases = [
        ("foliage 500k cards (config 4)", lambda: scenes.scene_foliage(n_cards=500000, tex_size=256, ground_quads=64), ((0, -48, 20), (0, 0, 8))),
        ("foliage 100k cards", lambda: scenes.scene_foliage(n_cards=100000, tex_size=256, ground_quads=64), ((0, -48, 20), (0, 0, 8))),
        ("foliage 20k cards", lambda: scenes.scene_foliage(n_cards=20000, tex_size=256, ground_quads=64), ((0, -48, 20), (0, 0, 8))),
        ("props 256 (config 2)", lambda: scenes.scene_props(), ((0, -95, 40), (0, 0, 10))),
        ("terrain 1.3M", lambda: scenes.scene_terrain_closed(800), ((0.0, -330.0 * 800 / 1582, 200.0 * 800 / 1582), (0.0, 0.0, 10.0))),
    ]
    out = open(args.out, "w") if args.out else None
    for name, mk, cam in cases:
        scene = mk()
        nodes, prims = vt.build_bvh(scene)
        rays = scenes.pinhole_rays(1920, 1080, *cam)
        row = {"scene": name, "tris": int(scene.n_tris), "sibling_overlap": round(sibling_overlap(nodes), 4)}
        for mode in ("dp", "greedy"):
            os.environ["VT_COLLAPSE"] = mode
            accel = vt.Accel(0, layout="quad").populate(scene, bvh=(nodes, prims))
            d_rays = to_dev(rays)
            d_hits = torch.empty(len(rays) * 16, dtype=torch.uint8, device="cuda")
            d_attrs = torch.empty(len(rays) * 128, dtype=torch.uint8, device="cuda")
            accel.traverse_device(d_rays.data_ptr(), len(rays), d_hits.data_ptr(), d_attrs.data_ptr())
            d_b = torch.empty(len(rays) * 32, dtype=torch.uint8, device="cuda")
            accel.bounce_rays_device(d_attrs.data_ptr(), len(rays), 1, 7, d_b.data_ptr())
            torch.cuda.synchronize()
            sp, sb = accel.traverse_stats(d_rays.data_ptr(), len(rays)), accel.traverse_stats(d_b.data_ptr(), len(rays))
            row[mode] = {"primary_ms": round(time_traverse(accel, d_rays, len(rays), d_hits), 3), "bounce_ms": round(time_traverse(accel, d_b, len(rays), d_hits), 3),
                         "primary_visits_tests": [round(sp[0] / len(rays), 2), round(sp[1] / len(rays), 2)],
                         "bounce_visits_tests": [round(sb[0] / len(rays), 2), round(sb[1] / len(rays), 2)]}
            accel.close()
        del os.environ["VT_COLLAPSE"]
        print(json.dumps(row), flush=True)
        if out:
            out.write(json.dumps(row) + "\n")
            out.flush()


if __name__ == "__main__":
    main()
