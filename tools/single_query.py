"""A handful of single-query searches (the reference's call shape) on a C4-sized space, for an ncu launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pyarrowspace_b200.api import ArrowSpaceBuilder
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
f = 384
g = torch.Generator(device="cuda").manual_seed(1)
centres = torch.randn(256, f, generator=g, device="cuda", dtype=torch.float64)
lab = torch.randint(0, 256, (n,), generator=g, device="cuda")
x = centres[lab] + 0.3 * torch.randn(n, f, generator=g, device="cuda", dtype=torch.float64)
x = x / x.norm(dim=1, keepdim=True) * 100.0 + 25.0
gp = {"eps": 10.0, "k": 25, "topk": 10, "p": 2.0, "sigma": None}
aspace, gl = ArrowSpaceBuilder.build(gp, x)
q = (x[:8] / 100.0).cpu().numpy()
for mode in ("fp64", "tc"):
    os.environ["ASP_SEARCH_STAGE1"] = mode
    for i in range(3):
        r = aspace.search(q[i], gl, 0.62)
    print(mode, r[:3])
