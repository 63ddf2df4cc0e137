import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bsp_files, mdl_files, test_vtf
seed = int(sys.argv[1]); N = int(sys.argv[2]); out = sys.argv[3]
rng = np.random.default_rng(seed)
def mutate(b):
    b = bytearray(b)
    for _ in range(int(rng.integers(0, 8))):
        if len(b) == 0: break
        mode = rng.integers(0, 5); pos = int(rng.integers(0, len(b)))
        if mode == 0: b[pos] = int(rng.integers(0, 256))
        elif mode == 1 and pos + 4 <= len(b): b[pos:pos+4] = int(rng.choice([0, 1, 0x7FFFFFFF, 0xFFFFFFFF, 0x80000000, len(b), len(b)-1, len(b)+1, 65535, 65536, 0x01000000])).to_bytes(4, "little")
        elif mode == 2: del b[pos:]
        elif mode == 3 and pos + 2 <= len(b): b[pos:pos+2] = int(rng.choice([0, 0xFFFF, 0x7FFF, 0x8000, 256])).to_bytes(2, "little")
        elif mode == 4 and pos + 4 <= len(b):
            v = int.from_bytes(b[pos:pos+4], "little"); v = (v + int(rng.choice([-1, 1, -4, 4, 16, -16, 1 << 16]))) & 0xFFFFFFFF; b[pos:pos+4] = v.to_bytes(4, "little")
    return bytes(b)
for k in ("bsp", "vtf", "mdl"): os.makedirs(os.path.join(out, k), exist_ok=True)
maps = [bsp_files.make_map(seed=s, layout=l, version=v, sprp_version=sp)["data"] for s, l, v, sp in ((0, "grid", 20, 6), (1, "tjunc", 19, 4), (2, "single", 21, 5), (3, None, 20, 6))]
models = [mdl_files.make_model(seed=s, fixups=f) for s, f in ((0, False), (1, True))]
fmts = ['RGBA8888', 'RGB888', 'BGR888', 'RGB565', 'DXT1', 'DXT3', 'DXT5', 'BGRA8888', 'BGR565', 'BGRA4444', 'BGRA5551', 'RGBA16161616F', 'I8', 'P8', 'DXT1_ONEBITALPHA', 'UV88']
vtfs = [test_vtf.make_vtf(fm, 16, 8, 4, frames=2, resources=bool(len(fm) % 2), minor=2 + (len(fm) % 4), low=(16, 16) if len(fm) % 3 == 0 else None) for fm in fmts]
vtfs += [test_vtf.make_vtf('RGBA8888', 8, 8, 3, flags=0x4000, minor=4), test_vtf.make_vtf('DXT5', 8, 8, 1, depth=4, minor=2)]
for i in range(N):
    open(os.path.join(out, "bsp", f"{i:05d}"), "wb").write(mutate(maps[i % len(maps)]))
    open(os.path.join(out, "vtf", f"{i:05d}"), "wb").write(mutate(vtfs[i % len(vtfs)]))
    m = models[i % 2]; parts = [m["mdl"], m["vvd"], m["vtx"]]; j = int(rng.integers(0, 3)); parts[j] = mutate(parts[j])
    for name, p in zip(("mdl", "vvd", "vtx"), parts): open(os.path.join(out, "mdl", f"{i:05d}.{name}"), "wb").write(p)
