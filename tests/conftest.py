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
