"""The hybrid-search leg of bench.py, run by bench.py in a SEPARATE process with a timeout so that nothing it
does can cost the bench its JSON line: ArrowSpace.search_hybrid_batch on the C4 workload (host queries in, host results out --
the public call, so the figure is end to end), default shortlist, then an untimed oracle check of a sample of the queries.
Prints one line "HYBRID_LEG {json}".  Usage: python tools/hybrid_leg.py [--items N] [--queries Q] [--device D]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", type=int, default=None)
    ap.add_argument("--queries", type=int, default=16384)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--parity-queries", type=int, default=16)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    import numpy as np
    import torch
    from pyarrowspace_b200 import _lib, api, synth
    from pyarrowspace_b200.api import ArrowSpaceBuilder
    cfg = synth.config("C4")
    n, f, gp, tau = a.items or cfg["n"], cfg["f"], cfg["graph_params"], cfg["tau"]
    torch.cuda.set_device(a.device)
    x = synth.make_items(n, f, cfg["seed"], cfg["scale"])
    q, _ = synth.make_queries(x[:min(n, 65536)], a.queries, cfg["seed"] + 7, cfg["scale"])
    aspace, gl = ArrowSpaceBuilder.build(gp, x, device=a.device)
    lib, ctx = _lib.load(), _lib.context(a.device)
    for _ in range(2):
        aspace.search_hybrid_batch(q, gl, tau)
    ts = []
    l0 = lib.asp_ctx_launch_count(ctx)
    for _ in range(a.reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        idx, sc = aspace.search_hybrid_batch(q, gl, tau)
        ts.append(time.perf_counter() - t0)
    launches = (lib.asp_ctx_launch_count(ctx) - l0) // a.reps
    out = {"metric": "hybrid search queries/s (top-%d, shortlist %d, %d x %d f64)" % (gp["topk"], int(api.stat("hybrid_pool", a.device)), n, f),
           "value": a.queries * len(ts) / sum(ts), "unit": "queries/s", "ms_per_call": sum(ts) / len(ts) * 1e3,
           "ms_per_call_min": min(ts) * 1e3, "queries_per_call": a.queries, "calls": a.reps, "warmup_calls": 2, "shortlist": int(api.stat("hybrid_pool", a.device)), "tau": tau, "gpu_launches": int(launches),
           "h2d_bytes_per_call": a.queries * f * 8, "d2h_bytes_per_call": a.queries * gp["topk"] * 16 + a.queries * 8,
           "tensor_core_candidates": api.stat("search_stage1_is_tc", a.device) == 1.0,
           "exact_scan_queries": api.stat("search_slow_queries", a.device),
           "note": "wall clock of the public call with host buffers (upload, shortlist at tau = 1 through the search path, "
                   "re-ranking kernels, read-back); restated semantics, parity unpinned against the crate (DESIGN.md section 6)"}
    if a.parity_queries > 0:
        import oracle
        try:
            oracle.set_num_threads(len(os.sched_getaffinity(0)))
        except AttributeError:
            pass
        t0 = time.perf_counter()
        s, g = oracle.build(gp, x)
        pick = np.unique(np.linspace(0, a.queries - 1, a.parity_queries).astype(np.int64))
        oidx, osc, _ = s.search_hybrid_batch(q[pick], g, tau, 0)
        plain, _, _ = s.search_batch(q[pick], g, tau)
        m = oidx >= 0
        rel = float(np.max(np.abs(sc[pick][m] - osc[m]) / np.abs(osc[m]))) if m.any() else 0.0
        same = bool(np.array_equal(idx[pick], oidx))
        out["parity_check"] = {"queries": int(len(pick)), "items": n, "idx_equal": same, "score_max_rel_err": rel, "rtol": 1e-9,
                               "scores_bit_identical": bool(np.array_equal(sc[pick][m], osc[m])),
                               "queries_where_hybrid_differs_from_plain_search": int((oidx != plain).any(axis=1).sum()),
                               "ok": bool(same and rel <= 1e-9), "seconds": time.perf_counter() - t0}
    print("HYBRID_LEG " + json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
