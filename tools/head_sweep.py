"""Head size of the two-piece host pipeline (asp_search_batch with host pointers): interleaved A/B at the C4 bench shape."""
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
aspace, gl = ArrowSpaceBuilder.build(gp, torch.from_numpy(x).cuda())
nq = 65536
q, _ = synth.make_queries(x[:65536], nq, cfg["seed"], cfg["scale"])
q_pin = torch.from_numpy(q).pin_memory()
q_np = q_pin.numpy()
out_idx = torch.empty((nq, gp["topk"]), dtype=torch.int64).pin_memory().numpy()
out_sc = torch.empty((nq, gp["topk"]), dtype=torch.float64).pin_memory().numpy()
heads = [int(v) for v in (sys.argv[1:] or ["0", "3072", "4096", "6144", "8192"])]
res = {h: [] for h in heads}
for rnd in range(9):
    for h in heads:
        if h:
            os.environ["ASP_PIPE_HEAD"] = str(h)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        aspace.search_batch(q_np, gl, 0.62, out=(out_idx, out_sc))
        dt = (time.perf_counter() - t0) * 1e3
        os.environ.pop("ASP_PIPE_HEAD", None)
        if rnd:
            res[h].append(dt)
qd = torch.from_numpy(q).cuda()
dev = []
for _ in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter(); aspace.search_batch(qd, gl, 0.62); torch.cuda.synchronize()
    dev.append((time.perf_counter() - t0) * 1e3)
for h in heads:
    print("head", h or "default", "median %.2f ms  min %.2f" % (float(np.median(res[h])), min(res[h])), flush=True)
print("device-resident median %.2f ms" % float(np.median(dev[1:])))
json.dump({"heads": {str(h): res[h] for h in heads}, "device_resident": dev}, open(os.path.join(ROOT, "gpurun_out", "head_sweep.json"), "w"))
