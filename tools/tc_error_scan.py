"""Measures the tcgen05 fp16-split cosine error against f64 at the C4 shape (dump mode, a slice of queries) and prints
the stage split of one search step.  Output: gpurun_out/tc_error_scan.json"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pyarrowspace_b200 import api, synth, _lib
from pyarrowspace_b200.api import ArrowSpaceBuilder

n, f = int(os.environ.get("N", 1000000)), int(os.environ.get("F", 384))
nq_err = int(os.environ.get("QE", 128))
c = synth.config("C4")
x = synth.make_items(n, f, c["seed"], c["scale"])
q, sel = synth.make_queries(x[:65536], 16384, 44, c["scale"])
aspace, gl = ArrowSpaceBuilder.build(c["graph_params"], x)
out = {}
dots = np.zeros((nq_err, n), dtype=np.float32)
qe = np.ascontiguousarray(q[:nq_err])
_lib.check(_lib.load().asp_debug_tc_dots(aspace._h, qe.ctypes.data, nq_err, dots.ctypes.data))
xt = torch.from_numpy(x).cuda()
qt = torch.from_numpy(qe).cuda()
exact = (qt @ xt.T) / (qt.norm(dim=1)[:, None] * xt.norm(dim=1)[None, :])
err = (torch.from_numpy(dots).cuda().double() - exact).abs()
band = api.stat("search_delta_cos_max")
out["terms"] = api.stat("search_terms"); out["rho_q_max"] = api.stat("search_rho_q_max"); out["rho_x_max"] = api.stat("search_rho_x_max")
out["err_max"] = float(err.max()); out["err_mean"] = float(err.mean()); out["band"] = band
out["band_over_max"] = band / out["err_max"]; out["dots_checked"] = int(err.numel())
print(out, flush=True)
qd = torch.from_numpy(q).cuda()
for rep in range(4):
    idx, sc = aspace.search_batch(qd, gl, c["tau"])
    st = {k: api.stat(k) for k in ("search_stage1_ms", "search_stage2_ms", "search_rescored_per_query", "search_exact_per_query", "search_slow_queries", "search_terms", "search_delta_cos_max")}
    print(rep, st, flush=True)
out["stats"] = st
out["top1_is_source"] = float((idx.cpu().numpy()[:, 0] == sel).mean())
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tc_error_scan.json"), "w"), indent=1)
print(json.dumps(out))
