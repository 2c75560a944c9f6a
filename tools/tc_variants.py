"""Profiling helper: time the tcgen05 candidate kernel in its diagnostic variants (results of variants 2/3 are wrong)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pyarrowspace_b200 import api, synth
from pyarrowspace_b200.api import ArrowSpaceBuilder
n, f, Q = int(os.environ.get("N", 1000000)), 384, int(os.environ.get("Q", 16384))
c = synth.config("C4")
x = torch.from_numpy(synth.make_items(n, f, c["seed"], c["scale"])).cuda()
q = torch.from_numpy(synth.make_queries(synth.make_items(n, f, c["seed"], c["scale"], rows=(0, 65536)), Q, 44, c["scale"])[0]).cuda()
aspace, gl = ArrowSpaceBuilder.build(c["graph_params"], x)
out = {}
for v in sys.argv[1:] or ["0", "2", "3"]:
    os.environ["ASP_TC_VARIANT"] = v
    ts = []
    for i in range(4):
        try:
            aspace.search_batch(q, gl, c["tau"])
        except Exception as e:
            pass
        ts.append(api.stat("search_stage1_ms"))
    out[v] = min(ts[1:])
    print("variant", v, "stage1 ms", ts, flush=True)
print(json.dumps(out))
