"""Profiling driver for K3 (taumode_kernel) at the C4 shape: a few builds so that `ncu -k regex:taumode_kernel -s 1 -c 1`
captures a warm launch.    python tools/tm_prof.py [tau_mode]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from pyarrowspace_b200 import api  # noqa: E402
from pyarrowspace_b200.api import ArrowSpaceBuilder  # noqa: E402

n, f = int(os.environ.get("N", 1000000)), int(os.environ.get("F", 384))
tau_mode = sys.argv[1] if len(sys.argv) > 1 else "median"
g = torch.Generator(device="cuda").manual_seed(1)
centres = torch.randn(256, f, generator=g, device="cuda", dtype=torch.float64)
lab = torch.randint(0, 256, (n,), generator=g, device="cuda")
x = centres[lab] + 0.3 * torch.randn(n, f, generator=g, device="cuda", dtype=torch.float64)
x = x / x.norm(dim=1, keepdim=True) * 100.0 + 25.0
gp = {"eps": 10.0, "k": 25, "topk": 10, "p": 2.0, "sigma": None}
for rep in range(3):
    aspace, gl = ArrowSpaceBuilder.build(gp, x, tau_mode=tau_mode)
    print(tau_mode, rep, api.stat("lambda_ms"), flush=True)
    del aspace, gl
