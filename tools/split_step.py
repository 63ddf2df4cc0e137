"""Experiment: the resident bench step as ONE stream vs P pixel slices on P streams (do tails / memory-bound stages overlap?).
usage (under gpurun): python tools/split_step.py [--parts 1,2,3,4]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vistrace_b200 as vt  # noqa: E402
from vistrace_b200 import scenes  # noqa: E402

SPP = 4


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--parts", default="1,2,3,4")
    ap.add_argument("--quads", type=int, default=1582)
    args = ap.parse_args()
    scene = scenes.scene_terrain_closed(args.quads)
    rays = scenes.pinhole_rays(1920, 1080, (0, -330, 200), (0, 0, 10))
    n = len(rays)
    accel = vt.Accel(0).populate(scene)
    dev = lambda nb: torch.empty(int(nb), dtype=torch.uint8, device="cuda")
    d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1)).cuda()
    d_hits, d_attrs, d_brays, d_bhits = dev(n * 16), dev(n * 128), dev(n * SPP * 32), dev(n * SPP * 16)
    d_queue, d_fb = dev(n * SPP * 4), torch.zeros(n * 3, dtype=torch.float32, device="cuda")
    main_stream = torch.cuda.current_stream()
    for parts in [int(x) for x in args.parts.split(",")]:
        streams = [torch.cuda.Stream() for _ in range(parts)]
        qcounts = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(parts)]
        cuts = [n * i // parts for i in range(parts + 1)]

        def step(seed):
            fork = torch.cuda.Event()
            fork.record(main_stream)
            for p, s in enumerate(streams):
                b, m = cuts[p], cuts[p + 1] - cuts[p]
                s.wait_event(fork)
                sh = s.cuda_stream
                accel.traverse_device(d_rays.data_ptr() + b * 32, m, d_hits.data_ptr() + b * 16, d_attrs.data_ptr() + b * 128, stream=sh)
                accel.bounce_rays_queued_device(d_attrs.data_ptr() + b * 128, m, SPP, seed, d_brays.data_ptr() + b * SPP * 32,
                                                d_queue.data_ptr() + b * SPP * 4, qcounts[p].data_ptr(), d_bhits.data_ptr() + b * SPP * 16, stream=sh)
                accel.traverse_queued_device(d_brays.data_ptr() + b * SPP * 32, d_queue.data_ptr() + b * SPP * 4, qcounts[p].data_ptr(), m * SPP,
                                             d_bhits.data_ptr() + b * SPP * 16, stream=sh)
                accel.accumulate_sky_device(d_attrs.data_ptr() + b * 128, d_bhits.data_ptr() + b * SPP * 16, m, SPP, 1.0, d_fb.data_ptr() + b * 12, stream=sh)
                join = torch.cuda.Event()
                join.record(s)
                main_stream.wait_event(join)

        for it in range(3):
            step(it)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main_stream)
        for it in range(20):
            step(100 + it)
        e1.record(main_stream)
        torch.cuda.synchronize()
        print(f"parts={parts}: {e0.elapsed_time(e1) / 20:.3f} ms/step", flush=True)


if __name__ == "__main__":
    main()
