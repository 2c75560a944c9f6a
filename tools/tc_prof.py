"""Profiling driver for the tcgen05 candidate kernel at the C4 shape (16k-query batches so that ncu replays stay short):
    [ASP_TC_PAIR=1] python tools/tc_prof.py      then   ncu -k regex:tc_gemm_kernel -s 2 -c 1 ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from pyarrowspace_b200 import api, synth  # noqa: E402
from pyarrowspace_b200.api import ArrowSpaceBuilder  # noqa: E402

n, Q = int(os.environ.get("N", 1000000)), int(os.environ.get("Q", 16384))
c = synth.config("C4")
g = torch.Generator(device="cuda").manual_seed(1)
f = c["f"]
centres = torch.randn(256, f, generator=g, device="cuda", dtype=torch.float64)
lab = torch.randint(0, 256, (n,), generator=g, device="cuda")
x = centres[lab] + 0.3 * torch.randn(n, f, generator=g, device="cuda", dtype=torch.float64)
x = x / x.norm(dim=1, keepdim=True) * 100.0 + 25.0
sel = torch.randint(0, n, (Q,), generator=g, device="cuda")
q = x[sel] / 100.0 + 0.01 * torch.randn(Q, f, generator=g, device="cuda", dtype=torch.float64)
aspace, gl = ArrowSpaceBuilder.build(c["graph_params"], x)
for rep in range(4):
    aspace.search_batch(q, gl, c["tau"])
    print(rep, {k: api.stat(k) for k in ("search_stage1_ms", "search_stage2_ms", "search_cta_pair", "search_rescored_per_query", "search_slow_queries")}, flush=True)
