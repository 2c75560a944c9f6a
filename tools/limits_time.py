"""What the documented limits of the search path cost: top-k beyond the 16-entry kernel variant, beyond the
tensor-core pass, tau at the edge of / outside the tensor-core pass.  C4 items, 8192 device-resident queries per call."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from pyarrowspace_b200 import api, synth
from pyarrowspace_b200.api import ArrowSpaceBuilder

cfg = synth.config("C4")
n, f = cfg["n"], cfg["f"]
x = synth.make_items(n, f, cfg["seed"], cfg["scale"])
xd = torch.from_numpy(x).cuda()
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
q, _ = synth.make_queries(x[:65536], nq, cfg["seed"], cfg["scale"])
qd = torch.from_numpy(q).cuda()
out = {"n": n, "f": f, "queries_per_call": nq, "cases": []}
for topk, tau, nqc in [(10, 0.62, nq), (16, 0.62, nq), (17, 0.62, nq), (26, 0.62, nq), (32, 0.62, nq), (33, 0.62, 256),
                       (10, 1.0, nq), (10, 0.05, nq), (10, 0.0005, 1024), (10, 0.0, 1024)]:
    gp = dict(cfg["graph_params"], topk=topk)
    aspace, gl = ArrowSpaceBuilder.build(gp, xd)
    qq = qd[:nqc]
    aspace.search_batch(qq, gl, tau)
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); aspace.search_batch(qq, gl, tau); torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    case = {"topk": topk, "tau": tau, "queries": nqc, "ms": min(ts), "queries_per_s": nqc / (min(ts) * 1e-3),
            "tensor_core_candidates": api.stat("search_stage1_is_tc"), "exact_scan_queries": api.stat("search_slow_queries"),
            "stage1_ms": api.stat("search_stage1_ms"), "stage2_ms": api.stat("search_stage2_ms")}
    out["cases"].append(case)
    print(json.dumps(case), flush=True)
    del aspace, gl
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "limits_time.json"), "w"), indent=1)
