"""Small-shape tour of every kernel family for compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_case.py
    compute-sanitizer --tool racecheck python tools/sanitize_case.py     (shared-memory hazards)
    compute-sanitizer --tool synccheck python tools/sanitize_case.py

Each step is also checked against the oracle, so a run that passes is a parity run too.  Shapes are kept tiny: the tools slow
kernels down by two orders of magnitude."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from pyarrowspace_b200.api import ArrowSpaceBuilder  # noqa: E402

QUICK = bool(os.environ.get("SANITIZE_QUICK"))
rng = np.random.default_rng(3)


def clustered(n, f, ncl=8):
    c = rng.normal(size=(ncl, f))
    return np.abs(c[rng.integers(0, ncl, n)] + 0.2 * rng.normal(size=(n, f))) + 0.05


def same(a, b, what):
    ok = np.array_equal(a, b)
    print("%-58s %s" % (what, "ok" if ok else "MISMATCH"), flush=True)
    return ok


ok = True
gp = {"eps": 0.6, "k": 5, "topk": 6, "p": 2.0, "sigma": 0.3}
# feature graph + lambdas (Gram DMMA + TMA, selection, CSR, taumode) and the three search routes
for n, f in ((700, 48), (300, 130)) if QUICK else ((1500, 48), (600, 130), (400, 1600)):
    x = clustered(n, f)
    a, g = ArrowSpaceBuilder.build(gp, x)
    s, og = oracle.build(gp, x)
    ok &= same(g.edges(), og.edges(), "feature graph edges %dx%d" % (n, f))
    ok &= bool(np.allclose(a.lambdas(), s.lambdas(), rtol=1e-9, atol=0))
    q = x[:: max(1, n // 300)] * 1.01
    for route in ("tc", "fp64"):
        os.environ["ASP_SEARCH_STAGE1"] = route
        idx, sc = a.search_batch(q, g, 0.7)
        oidx, osc, _ = s.search_batch(q, og, 0.7)
        ok &= same(idx, oidx, "search (%s candidates) %d queries" % (route, len(q)))
    os.environ.pop("ASP_SEARCH_STAGE1")
    idx, sc = a.search_batch(q[:3], g, 0.7)                           # GEMV / small-batch route
    ok &= same(idx, s.search_batch(q[:3], og, 0.7)[0], "search, 3 queries")
    for pool in (None, 40):                                           # hybrid search: shortlist + re-ranking kernels (csrc/hybrid.cuh)
        idx, sc = a.search_hybrid_batch(q, g, 0.4, pool=pool)
        oidx, osc, _ = s.search_hybrid_batch(q, og, 0.4, pool or 0)
        ok &= same(idx, oidx, "hybrid search, shortlist %s" % (pool or "default")) and same(sc, osc, "hybrid scores bit for bit")
    del a, g
# unpinned switches (symmetrise / laplacian variants, synthetic lambda, l2 distance)
x = clustered(500, 40)
for sw in ({"profile": "kat12"}, {"lambda_form": "synthetic", "tau_mode": "median_abs"}, {"distance": "l2", "symmetrise": "min", "laplacian": "rw"}):
    a, g = ArrowSpaceBuilder.build({"eps": 5.0, "k": 4, "topk": 3, "p": 2.0, "sigma": None} if "distance" in sw else gp, x, **sw)
    s, og = oracle.build({"eps": 5.0, "k": 4, "topk": 3, "p": 2.0, "sigma": None} if "distance" in sw else gp, x, **sw)
    ok &= same(g.csr()[1], og.csr()[1], "switches %s" % sw)
    ok &= bool(np.allclose(a.lambdas(), s.lambdas(), rtol=1e-9, atol=0))
    del a, g
# item graph (tcgen05 candidates need n >= 8192; the FP64 pass below that)
for n, f, mode in ((900, 32, "fp64"),) if QUICK else ((1200, 32, "fp64"), (8448, 64, "tc")):
    x = clustered(n, f, 24)
    os.environ["ASP_KNN_STAGE1"] = mode
    a, g = ArrowSpaceBuilder.build_item_graph({"eps": 0.3, "k": 6, "topk": 3, "p": 2.0, "sigma": None}, x)
    _, og = oracle.build({"eps": 0.3, "k": 6, "topk": 3, "p": 2.0, "sigma": None}, x, nodes="items")
    ok &= same(g.edges(), og.edges(), "item graph (%s) %dx%d" % (mode, n, f))
    os.environ.pop("ASP_KNN_STAGE1")
    del a, g
# pre-graph reduction
x = clustered(900 if QUICK else 2500, 50)
a, g = ArrowSpaceBuilder.build(gp, x, reduction={"n_clusters": 40})
s, og, cent, info = oracle.build_reduced(gp, x, reduction={"n_clusters": 40})
ok &= same(g.centroids(), cent, "reduction centroids")
ok &= same(g.edges(), og.edges(), "reduction graph")
print("SANITIZE_CASE", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
