"""search_hybrid at the C4 shape: throughput of the shortlist + re-ranking path for the shortlist lengths
that select each route (tensor-core candidates <= 31, batched exact scan beyond, no shortlist), with an oracle check of a
sample.  Not yet run: written after the round's GPU budget was spent.  Usage: python tools/hybrid_time.py [queries_per_call]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import oracle
from pyarrowspace_b200 import api, synth
from pyarrowspace_b200.api import ArrowSpaceBuilder

cfg = synth.config("C4")
n, f, tau = cfg["n"], cfg["f"], cfg["tau"]
x = synth.make_items(n, f, cfg["seed"], cfg["scale"])
xd = torch.from_numpy(x).cuda()
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
q, _ = synth.make_queries(x[:65536], nq, cfg["seed"], cfg["scale"])
out = {"n": n, "f": f, "queries_per_call": nq, "cases": []}
s = g = None
for topk, pool, nqc in [(10, None, nq), (15, None, nq), (15, 31, nq), (10, 40, 1024), (10, n, nq)]:
    gp = dict(cfg["graph_params"], topk=topk)
    aspace, gl = ArrowSpaceBuilder.build(gp, xd)
    qq = q[:nqc]
    aspace.search_hybrid_batch(qq, gl, tau, pool=pool)
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); idx, sc = aspace.search_hybrid_batch(qq, gl, tau, pool=pool)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    so, go = oracle.build(gp, x)
    pick = np.unique(np.linspace(0, nqc - 1, 16).astype(np.int64))
    oidx, osc, _ = so.search_hybrid_batch(qq[pick], go, tau, pool or 0)
    case = {"topk": topk, "pool": pool, "shortlist": api.stat("hybrid_pool"), "queries": nqc, "ms": min(ts),
            "queries_per_s": nqc / (min(ts) * 1e-3), "tensor_core_candidates": api.stat("search_stage1_is_tc"),
            "exact_scan_queries": api.stat("search_slow_queries"),
            "oracle_idx_equal": bool(np.array_equal(idx[pick], oidx)),
            "oracle_score_max_rel_err": float(np.max(np.abs(sc[pick] - osc) / np.abs(osc)))}
    out["cases"].append(case)
    print(json.dumps(case), flush=True)
    del aspace, gl, so, go
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "hybrid_time.json"), "w"), indent=1)
