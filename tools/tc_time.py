"""Quick timing of the tcgen05 candidate kernel on device-generated data (no host data generation):
    python tools/tc_time.py [N] [Q]   with env ASP_TC_VARIANT / ASP_TC_ARES / ASP_TC_TERMS sweeps inside."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pyarrowspace_b200 import api
from pyarrowspace_b200.api import ArrowSpaceBuilder
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
f = 384
g = torch.Generator(device="cuda").manual_seed(1)
x = (torch.randn(n, f, generator=g, device="cuda", dtype=torch.float64) * 5.0 + 30.0)
q = x[torch.randint(0, n, (Q,), generator=g, device="cuda")] / 100.0 + 0.0005 * torch.randn(Q, f, generator=g, device="cuda", dtype=torch.float64)
aspace, gl = ArrowSpaceBuilder.build({"eps": 10.0, "k": 25, "topk": 10, "p": 2.0, "sigma": None}, x)
out = {}
configs = [("base", {}), ("pair", {"ASP_TC_PAIR": "1"}), ("pair_var2", {"ASP_TC_PAIR": "1", "ASP_TC_VARIANT": "2"}), ("pair_var4", {"ASP_TC_PAIR": "1", "ASP_TC_VARIANT": "4"}), ("var4", {"ASP_TC_VARIANT": "4"}), ("var2", {"ASP_TC_VARIANT": "2"}), ("var3", {"ASP_TC_VARIANT": "3"}),
                  ("nores", {"ASP_TC_ARES": "0"}), ("nores_var2", {"ASP_TC_ARES": "0", "ASP_TC_VARIANT": "2"}),
                  ("terms3", {"ASP_TC_TERMS": "3"}), ("terms3_var2", {"ASP_TC_TERMS": "3", "ASP_TC_VARIANT": "2"})]
if os.environ.get("QUICK"):
    configs = configs[:int(os.environ['QUICK'])]
for name, env in configs:
    for k in ("ASP_TC_VARIANT", "ASP_TC_ARES", "ASP_TC_TERMS", "ASP_TC_PAIR"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ts = []
    for i in range(4):
        aspace.search_batch(q, gl, 0.62)
        ts.append(api.stat("search_stage1_ms"))
    out[name] = {"stage1_ms": min(ts[1:]), "terms": api.stat("search_terms"), "ares": api.stat("search_a_resident"), "pair": api.stat("search_cta_pair"),
                 "rescored": api.stat("search_rescored_per_query"), "stage2_ms": api.stat("search_stage2_ms")}
    print(name, out[name], flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tc_time.json"), "w"), indent=1)
