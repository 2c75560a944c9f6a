"""Mean-zero regime (bench.py regimes.mean_zero): the three-term split against ONE fp16 term with its (wide) band."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from pyarrowspace_b200 import api, synth
from pyarrowspace_b200.api import ArrowSpaceBuilder

cfg = synth.config("C4")
n, f, gp = cfg["n"], cfg["f"], cfg["graph_params"]
x = torch.from_numpy(synth.make_items(n, f, cfg["seed"], cfg["scale"], shift=0.0)).cuda()
aspace, gl = ArrowSpaceBuilder.build(gp, x, tau_mode="median_abs")
keys = ("search_stage1_ms", "search_stage2_ms", "search_terms", "search_slow_queries", "search_rescored_per_query", "search_delta_cos_max")
ref = {}
for nq in (8192, 65536):
    q = torch.from_numpy(synth.make_fresh_queries(f, nq, cfg["seed"] + 100)[0]).cuda()
    for terms in ("3", "1"):
        os.environ["ASP_TC_TERMS"] = terms
        aspace.search_batch(q[:1024], gl, 0.62)
        ts = []
        for _ in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter(); idx, sc = aspace.search_batch(q, gl, 0.62); torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        st = {k: api.stat(k) for k in keys}
        same = None
        if terms == "3":
            ref[nq] = (idx.clone(), sc.clone())
        else:
            same = bool(torch.equal(idx, ref[nq][0]) and torch.equal(sc, ref[nq][1]))
        print("queries", nq, "terms", terms, "wall %.2f ms" % min(ts), "%.0f q/s" % (nq / min(ts) * 1e3), st, "equal to 3-term:", same, flush=True)
        if st["search_slow_queries"] > 64 and nq == 8192:
            print("one-term mode overflows here; skipping the 64k batch"); sys.exit(0)
os.environ.pop("ASP_TC_TERMS")
