"""Where the time of ONE search call goes at small per-rank batches (the N = 8 situation: 8192 queries per rank and step):
device stages (CUDA events) against the host wall clock between them (stats search_host_*_us)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from pyarrowspace_b200 import api, synth
from pyarrowspace_b200.api import ArrowSpaceBuilder

cfg = synth.config("C4")
n, f, gp = cfg["n"], cfg["f"], cfg["graph_params"]
x = synth.make_items(n, f, cfg["seed"], cfg["scale"])
xd = torch.from_numpy(x).cuda()
aspace, gl = ArrowSpaceBuilder.build(gp, xd)
keys = ["search_host_call_us", "search_host_lambda_us", "search_host_device_batch_us", "search_host_prep_us",
        "search_host_stage1_launched_us", "search_stage1_ms", "search_stage2_ms"]
out = {}
for nq in [int(v) for v in (sys.argv[1:] or ["8192", "65536"])]:
    q, _ = synth.make_queries(x[:65536], nq, cfg["seed"], cfg["scale"])
    qd = torch.from_numpy(q).cuda()
    for _ in range(5):
        aspace.search_batch(qd, gl, 0.62)
    rows = []
    for _ in range(20):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        aspace.search_batch(qd, gl, 0.62)
        torch.cuda.synchronize(); wall = (time.perf_counter() - t0) * 1e6
        rows.append([wall] + [api.stat(k) for k in keys])
    med = np.median(np.array(rows), axis=0)
    out[nq] = dict(zip(["python_wall_us"] + keys, [float(v) for v in med]))
    print(nq, json.dumps(out[nq]), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "call_breakdown.json"), "w"), indent=1)
