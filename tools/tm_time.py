"""K3 (lambda pass) timing on the C4-sized workload, one GPU: the fused kernel with the ALU / histogram median, and with tau modes
that need no selection (mean, fixed) -- the difference is what the median costs next to the graph walk.
Writes gpurun_out/tm_time.json.   python tools/tm_time.py [N] [F] [k]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pyarrowspace_b200 import api  # noqa: E402
from pyarrowspace_b200.api import ArrowSpaceBuilder  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
f = int(sys.argv[2]) if len(sys.argv) > 2 else 384
k = int(sys.argv[3]) if len(sys.argv) > 3 else 25
g = torch.Generator(device="cuda").manual_seed(1)
centres = torch.randn(256, f, generator=g, device="cuda", dtype=torch.float64)
lab = torch.randint(0, 256, (n,), generator=g, device="cuda")
x = centres[lab] + 0.3 * torch.randn(n, f, generator=g, device="cuda", dtype=torch.float64)
x = x / x.norm(dim=1, keepdim=True) * 100.0 + 25.0
gp = {"eps": 10.0, "k": k, "topk": 10, "p": 2.0, "sigma": None}
cases = [("median_interp", {}, {}), ("median_alu", {"ASP_TM_MEDIAN": "alu"}, {}), ("median_abs_interp", {}, {"tau_mode": "median_abs"}),
         ("mean", {}, {"tau_mode": "mean"}), ("synthetic", {}, {"lambda_form": "synthetic", "tau_mode": "mean"})]
out = {"n": n, "f": f, "k": k}
ref = None
for rnd in range(4):
    for name, env, sw in cases:
        os.environ.update(env)
        aspace, gl = ArrowSpaceBuilder.build(gp, x, **sw)
        for key in env:
            os.environ.pop(key)
        out.setdefault(name, []).append(api.stat("lambda_ms"))
        if rnd == 0:
            lam = aspace.lambdas()
            if name == "median_interp":
                ref = lam
                out["upper_nnz"] = (gl.nnz - f) // 2
            elif name.startswith("median"):
                out[name + "_equals_interp_bitwise"] = bool(np.array_equal(lam, ref))
        del aspace, gl
for name, _, _ in cases:
    out[name + "_median_ms"] = float(np.median(out[name][1:]))
    print(name, out[name], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tm_time.json"), "w"), indent=1)
print(json.dumps({k_: v for k_, v in out.items() if not isinstance(v, list)}))
