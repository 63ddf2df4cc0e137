"""PCIe probe: pinned H2D / D2H bandwidth and the e2e wave at several tile sizes."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vistrace_b200 as vt
from vistrace_b200 import abi, scenes
n = 166 * 1000 * 1000
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    print(name, f"{n/dt/1e9:.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
print("bidirectional", f"{n/dt/1e9:.1f} GB/s each way")
scene = scenes.scene_terrain_closed(int(sys.argv[1]) if len(sys.argv) > 1 else 1582)
rays = scenes.pinhole_rays(1920, 1080, (0, -330, 200), (0, 0, 10))
accel = vt.Accel(0).populate(scene)
SPP = 4; nr = len(rays)
pin = lambda nb: torch.empty(nb, dtype=torch.uint8).pin_memory()
hr, hh, hb = pin(nr * 32), pin(nr * 16), pin(nr * SPP * 16)
hrays = hr.numpy().view(abi.RAY); hrays[:] = rays
out = {"hits": hh.numpy().view(abi.HIT), "bounce_hits": hb.numpy().view(abi.HIT)}
import ctypes as C
live = C.c_uint64(0)
def raw():
    rc = accel.L.vt_accel_trace_diffuse_wave(accel.h, hrays.ctypes.data, nr, SPP, 1, out["hits"].ctypes.data, None, None, out["bounce_hits"].ctypes.data, C.addressof(live), 0, None)
    assert rc == 0
for tile in (1 << 18, 1 << 21):
    os.environ["VT_WAVE_TILE"] = str(tile)
    raw(); raw()
    t = time.perf_counter()
    for i in range(10): raw()
    dt = (time.perf_counter() - t) / 10
    print(f"raw ctypes tile {tile}: {dt*1e3:.2f} ms/step")
# device-only wave for comparison
dr = torch.from_numpy(rays.view(np.uint8).reshape(-1)).cuda()
dh, da, dbr, dbh = (torch.empty(nr * k, dtype=torch.uint8, device="cuda") for k in (16, 128, SPP * 32, SPP * 16))
st = torch.cuda.current_stream().cuda_stream
accel.trace_diffuse_wave_device(dr.data_ptr(), nr, SPP, 1, dh.data_ptr(), da.data_ptr(), dbr.data_ptr(), dbh.data_ptr(), stream=st)
torch.cuda.synchronize(); t = time.perf_counter()
for i in range(10): accel.trace_diffuse_wave_device(dr.data_ptr(), nr, SPP, i, dh.data_ptr(), da.data_ptr(), dbr.data_ptr(), dbh.data_ptr(), stream=st)
torch.cuda.synchronize(); print(f"device-only wave: {(time.perf_counter()-t)/10*1e3:.2f} ms/step")
for tile in (1 << 17, 1 << 18, 1 << 19):
    os.environ["VT_WAVE_TILE"] = str(tile)
    for _ in range(2): accel.trace_diffuse_wave(hrays, SPP, seed=1, out=out)
    t = time.perf_counter()
    for i in range(10): r = accel.trace_diffuse_wave(hrays, SPP, seed=i, out=out)
    dt = (time.perf_counter() - t) / 10
    print(f"tile {tile}: {dt*1e3:.2f} ms/step, {(nr + r['live_bounce'])/dt/1e6:.0f} Mrays/s")
