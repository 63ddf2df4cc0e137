"""Timeline of one host-pointer render wave (VT_WAVE_TRACE=1): when each tile's stages finish on the device."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vistrace_b200 as vt  # noqa: E402
from vistrace_b200 import abi, scenes  # noqa: E402

scene = scenes.scene_terrain_closed(1582)
rays = scenes.pinhole_rays(1920, 1080, (0, -330, 200), (0, 0, 10))
n = len(rays)
accel = vt.Accel(0).populate(scene)
h_rays_t, h_fb_t = torch.empty(n * 32, dtype=torch.uint8).pin_memory(), torch.empty(n * 12, dtype=torch.uint8).pin_memory()
h_rays = h_rays_t.numpy().view(abi.RAY)
h_rays[:] = rays
h_fb = h_fb_t.numpy().view(np.float32).reshape(n, 3)
for it in range(4):
    accel.render_diffuse_wave(h_rays, 4, seed=it, weight=1.0, out=h_fb)
os.environ["VT_WAVE_TRACE"] = "1"
import time
t0 = time.perf_counter()
accel.render_diffuse_wave(h_rays, 4, seed=9, weight=1.0, out=h_fb)
print(f"wall {1e3 * (time.perf_counter() - t0):.3f} ms (with event overhead)")
