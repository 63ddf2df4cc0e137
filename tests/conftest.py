import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Make sure every native piece exists (the driver runs build() first; this is a safety net)."""
    import __graft_entry__ as g

    g.build()
    return True


@pytest.fixture(scope="session")
def oracle_mod(built):
    import oracle

    return oracle


def oracle_kinds():
    import oracle

    kinds = ["port"]
    if oracle.available("reference"):
        kinds.append("reference")
    return kinds


def compare_hits(a, b):
    """Parity report between two hit arrays (SURVEY.md §8d 'Correctness report')."""
    from vistrace_b200 import abi

    miss_a, miss_b = a["prim"] == abi.VT_MISS, b["prim"] == abi.VT_MISS
    rep = {"n": len(a), "hit_miss_mismatch": int((miss_a != miss_b).sum())}
    both = ~miss_a & ~miss_b
    diff = both & (a["prim"] != b["prim"])
    rep["prim_mismatch"] = int(diff.sum())
    rep["prim_mismatch_exact_tie"] = int((diff & (a["t"] == b["t"])).sum())
    same = both & ~diff
    rep["tuv_bit_mismatch"] = int(
        (same & ((a["t"].view(np.uint32) != b["t"].view(np.uint32)) | (a["u"].view(np.uint32) != b["u"].view(np.uint32)) | (a["v"].view(np.uint32) != b["v"].view(np.uint32)))).sum()
    )
    return rep


def attr_max_rel_err(a, b):
    """Largest relative error per float field of vt_attr over the records both sides hit."""
    from vistrace_b200 import abi

    ok = (a["prim"] != abi.VT_MISS) & (b["prim"] != abi.VT_MISS) & (a["prim"] == b["prim"])
    out = {}
    for f in abi.ATTR.names:
        x, y = a[f][ok], b[f][ok]
        if x.dtype == np.float32:
            x64, y64 = x.astype(np.float64), y.astype(np.float64)
            den = np.maximum(np.abs(x64), 1e-6) if x.ndim == 1 else np.maximum(np.sqrt((x64 * x64).sum(-1, keepdims=True)), 1e-6)
            with np.errstate(invalid="ignore"):
                e = np.abs(x64 - y64) / den
            e = np.where(np.isfinite(x64) & np.isfinite(y64), e, np.where((x64 == y64) | (np.isnan(x64) & np.isnan(y64)), 0.0, np.inf))
            out[f] = float(e.max()) if e.size else 0.0
        else:
            out[f] = int((x != y).sum())
    return out


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    """tests/golden/<name>.npz -> (SceneData, dict of arrays); fixtures come from the unmodified
    reference (tests/golden/make_golden.py)."""
    from vistrace_b200 import abi

    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    texs = []
    for i in range(int(z["n_textures"])):
        w, h, m, fl = (int(v) for v in z[f"tex{i}_hdr"])
        texs.append((w, h, m, fl, z[f"tex{i}_px"]))
    scene = abi.SceneData(z["tris"], z["materials"], z["entities"], texs)
    return scene, z


def cornell_box():
    """The bvh library's golden-image known answer (libs/bvh/test/CMakeLists.txt:57-82, benchmark.cpp:121-176,451-463):
    (SceneData of the 36 Cornell-box triangles, the 1080 x 720 camera rays of `--eye 0 0.9 2.5 --dir 0 0.001 -1 --up 0 1 0 --fov 60`
    in the benchmark's float arithmetic, a function hits -> uint8 image, the reference picture)."""
    from vistrace_b200 import abi

    f4 = np.float32
    g = np.load(os.path.join(GOLDEN, "cornell_box.npz"))
    tris = np.zeros(len(g["p"]), abi.TRI_IN)
    tris["p"] = g["p"]
    tris["normals"][:, :, 2] = 1.0
    tris["tangents"][:, :, 0] = 1.0
    scene = abi.SceneData(tris)  # two-sided, one flag-free material: TriangleBackfaceCull degenerates to bvh::Triangle::intersect

    def dot(a, b):
        d = a[..., 0] * b[..., 0]
        d = d + a[..., 1] * b[..., 1]
        return d + a[..., 2] * b[..., 2]

    def normalize(v):  # vector.hpp:151-154: v * (1 / sqrt(dot))
        return (v * (f4(1) / np.sqrt(dot(v, v)))[..., None]).astype(f4)

    def cross(a, b):  # vector.hpp:159-167
        return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1], a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                         a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], -1).astype(f4)

    W, H = 1080, 720
    eye, cdir, up, fov = np.array([0, 0.9, 2.5], f4), np.array([0, 0.001, -1], f4), np.array([0, 1, 0], f4), f4(60)
    d = normalize(cdir)
    iu = normalize(cross(d, up))
    iv = normalize(cross(iu, d))
    iw = f4(np.tan(np.float64(fov * f4(3.14159265 * (1.0 / 180.0) * 0.5))))
    ratio = f4(H) / f4(W)
    iu = (iu * iw).astype(f4)
    iv = ((iv * iw).astype(f4) * ratio).astype(f4)
    i = np.arange(W, dtype=f4)[None, :]
    j = np.arange(H, dtype=f4)[:, None]
    u = (f4(2) * (i + f4(0.5)) / f4(W) - f4(1)).astype(f4) + np.zeros((H, 1), f4)
    v = (f4(2) * (j + f4(0.5)) / f4(H) - f4(1)).astype(f4) + np.zeros((1, W), f4)
    dirs = normalize(((iu[None, None, :] * u[..., None]).astype(f4) + (iv[None, None, :] * v[..., None]).astype(f4)).astype(f4) + d[None, None, :])
    rays = np.zeros(W * H, abi.RAY)  # index = W * j + i
    rays["o"] = eye
    rays["d"] = dirs.reshape(-1, 3)
    rays["tmin"] = 0.0
    rays["tmax"] = np.finfo(f4).max

    p = g["p"]
    n = cross(p[:, 0] - p[:, 1], p[:, 2] - p[:, 0])  # Triangle: e1 = p0 - p1, e2 = p2 - p0, n = cross(e1, e2)
    shade = np.abs(normalize(n))

    def to_image(hits):
        px = np.zeros((W * H, 3), f4)
        hit = hits["prim"] != abi.VT_MISS
        px[hit] = shade[hits["prim"][hit]]
        img = np.maximum(np.minimum(px * f4(255), f4(255)), f4(0)).astype(np.uint8).reshape(H, W, 3)
        return img[::-1]  # the benchmark writes row j = height - 1 first

    return scene, rays, to_image, g["image"]


def hostile_case(it, rng, n_rays=3000):
    """One seeded hostile scene + ray set (kind = it % 6: soup, zero-area, grid-aligned, duplicated, coplanar layers, soup again):
    rays with zero / negative-zero / tiny / huge direction components, grid-aligned origins and directions for the grid kind,
    finite and infinite intervals.  Shared by the oracle fuzz (tests/test_oracle.py) and the GPU fuzz (tests/test_gpu_parity.py)."""
    from vistrace_b200 import abi

    n = int(rng.integers(1, 300))
    p = rng.uniform(-20, 20, (n, 3, 3)).astype(np.float32)
    kind = it % 6
    if kind == 1:
        p[::3, 1] = p[::3, 0]
    elif kind == 2:
        p = np.round(p)
    elif kind == 3:
        p = np.concatenate([p, p])
    elif kind == 4:
        p[:, :, 2] = np.round(p[:, :, 2] / 10) * 10
    tris = np.zeros(len(p), abi.TRI_IN)
    tris["p"] = p
    tris["normals"] = rng.normal(size=(len(p), 3, 3))
    tris["tangents"] = rng.normal(size=(len(p), 3, 3))
    tris["uvs"] = rng.uniform(-2, 2, (len(p), 3, 2))
    tris["alphas"] = rng.uniform(0, 1, (len(p), 3))
    tris["one_sided"] = rng.integers(0, 2, len(p))
    scene = abi.SceneData(tris)
    m = n_rays
    rays = np.zeros(m, abi.RAY)
    rays["o"] = rng.uniform(-25, 25, (m, 3))
    rays["d"] = rng.normal(size=(m, 3))
    if kind == 2:
        rays["o"] = np.round(rays["o"])
        rays["d"] = np.round(rays["d"] * 2) / 2
    rays["d"][::7, 0] = 0.0
    rays["d"][::11, 1] = -0.0
    rays["d"][::13] *= 1e-3
    rays["d"][::17] *= 1e4
    rays["tmin"] = np.where(rng.random(m) < 0.3, rng.uniform(0, 5, m), 0)
    rays["tmax"] = np.where(rng.random(m) < 0.3, rng.uniform(5, 60, m), np.finfo(np.float32).max)
    rays["d"][(rays["d"] == 0).all(1), 2] = 1.0
    return scene, rays, kind
