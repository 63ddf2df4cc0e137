"""Correctness report between two hit buffers (SURVEY.md section 8d "Correctness report", BASELINE.json north_star:
"hit/miss and primitive ID bit-exact wherever the closest hit is unique ... near-ties counted and reported").

Pure numpy, no checker inside: the caller passes `tri_check(prim, ray) -> (accepted, [t, u, v] float32)` — the CHECKER's own
TriangleBackfaceCull::intersect (source/objects/Primitives.h:168-215) — when records that are closer than the reference's
are to be verified.  tests/ and bench.py's cpu_baseline leg pass oracle.CpuScene.tri_intersect; the product never does.

Classes of a record where `got` (the engine) and `want` (the reference traversal) differ in any bit:

  exact_tie   both hit, t bit-identical, another primitive: two triangles at the same distance; the reference lets the
              candidate it tests LAST win (`t <= tmax`, libs/bvh/include/bvh/single_ray_traverser.hpp:55-60), a layout
              that visits equal-distance subtrees in another order reports the twin.
  leak        the engine reports a hit the reference traversal did not consider — closer than the reference's hit
              (`near_leak` when within 1e-6 relative: a near-tie) or a hit where the reference misses (`leak_vs_miss`) —
              AND the checker's own triangle test accepts exactly that (t, u, v) for that primitive and ray.  The triangle
              is then a genuine hit under the reference's arithmetic which its traverser never tested because a
              FastNodeIntersector box test (node_intersectors.hpp:35-47) rounded the ray out of an ancestor box; the
              quantised layouts' boxes CONTAIN the reference's, so they keep such a candidate.
  unverified  closer / hit-vs-miss records the checker does not confirm (or no checker was passed): a failure.
  lost        the reference has a hit and the engine misses, or reports a FARTHER hit (any amount, near-tie or not): a
              lost candidate — exactly what conservative boxes rule out; always a failure.
  tuv_bits    same primitive, t, u or v differ in any bit: a failure (the triangle test is the exact one).
"""
import numpy as np

from . import abi


def classify_hits(got, want, rays=None, tri_check=None):
    got = np.ascontiguousarray(got, abi.HIT)
    want = np.ascontiguousarray(want, abi.HIT)
    assert len(got) == len(want)
    rep = {"rays": int(len(got)), "hits": int((want["prim"] != abi.VT_MISS).sum()), "differing": 0, "exact_tie": 0, "leak": 0, "near_leak": 0,
           "leak_vs_miss": 0, "unverified": 0, "lost": 0, "tuv_bits": 0}
    diff = np.nonzero((got.view(np.uint32).reshape(-1, 4) != want.view(np.uint32).reshape(-1, 4)).any(1))[0]
    rep["differing"] = int(len(diff))
    if len(diff) == 0:
        rep["ok"] = True
        return rep
    g, w = got[diff], want[diff]
    g_hit, w_hit = g["prim"] != abi.VT_MISS, w["prim"] != abi.VT_MISS
    both = g_hit & w_hit
    same_prim = both & (g["prim"] == w["prim"])
    rep["tuv_bits"] = int(same_prim.sum())
    t_equal = g["t"].view(np.uint32) == w["t"].view(np.uint32)
    tie = both & ~same_prim & t_equal
    rep["exact_tie"] = int(tie.sum())
    closer = (g_hit & ~w_hit) | (both & ~same_prim & ~t_equal & (g["t"] < w["t"]))
    lost = (~g_hit & w_hit) | (both & ~same_prim & ~t_equal & ~(g["t"] < w["t"]))
    rep["lost"] = int(lost.sum())
    for j in np.nonzero(closer)[0]:
        ok = False
        if rays is not None and tri_check is not None:
            accepted, tuv = tri_check(int(g["prim"][j]), rays[diff[j]])
            ok = bool(accepted) and np.asarray(tuv, np.float32).tobytes() == np.array([g["t"][j], g["u"][j], g["v"][j]], np.float32).tobytes()
        if not ok:
            rep["unverified"] += 1
            continue
        rep["leak"] += 1
        if not w_hit[j]:
            rep["leak_vs_miss"] += 1
        elif abs(float(g["t"][j]) - float(w["t"][j])) <= 1e-6 * abs(float(w["t"][j])):
            rep["near_leak"] += 1
    rep["ok"] = rep["unverified"] == 0 and rep["lost"] == 0 and rep["tuv_bits"] == 0
    return rep


def merge_reports(reports):
    """Sum of several classify_hits reports (one per wave)."""
    out = {}
    for r in reports:
        for k, v in r.items():
            if k == "ok":
                out[k] = out.get(k, True) and bool(v)
            else:
                out[k] = out.get(k, 0) + int(v)
    return out
