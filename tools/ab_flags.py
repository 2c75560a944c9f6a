"""Interleaved A/B of environment-variable variants on the C4-sized workload (1M x 384 f64, device-generated), one GPU.

    python tools/ab_flags.py "ASP_TM_PREFETCH=1" "ASP_MEDIAN_PREFETCH=1" "ASP_TM_PREFETCH=1,ASP_MEDIAN_PREFETCH=1" "ASP_TC_PAIR=1"

Every variant (plus the baseline with none of the variables) is measured in the same process, round-robin, so that box-to-box and
power-cap drift hits all of them alike.  Per variant: build stage times (gram / graph / lambda, device-resident items), one
64k-query search step (device-resident queries) with its stage-1 / stage-2 split, and whether lambdas and result lists are bitwise
equal to the baseline's.  Writes gpurun_out/ab_flags.json."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pyarrowspace_b200 import api  # noqa: E402
from pyarrowspace_b200.api import ArrowSpaceBuilder  # noqa: E402


def parse(spec):
    return dict(kv.split("=", 1) for kv in spec.split(",") if kv)


def main():
    variants = [("baseline", {})] + [(s, parse(s)) for s in sys.argv[1:]]
    n, f, nq, rounds = int(os.environ.get("AB_N", 1000000)), int(os.environ.get("AB_F", 384)), 65536, int(os.environ.get("AB_ROUNDS", 5))
    g = torch.Generator(device="cuda").manual_seed(1)
    centres = torch.randn(256, f, generator=g, device="cuda", dtype=torch.float64)
    lab = torch.randint(0, 256, (n,), generator=g, device="cuda")
    x = centres[lab] + 0.3 * torch.randn(n, f, generator=g, device="cuda", dtype=torch.float64)
    x = x / x.norm(dim=1, keepdim=True) * 100.0 + 25.0
    sel = torch.randint(0, n, (nq,), generator=g, device="cuda")
    q = x[sel] / 100.0 + 0.01 * torch.randn(nq, f, generator=g, device="cuda", dtype=torch.float64)
    gp = {"eps": 10.0, "k": 25, "topk": 10, "p": 2.0, "sigma": None}
    out = {name: {"build_ms": [], "gram_ms": [], "lambda_ms": [], "step_ms": [], "stage1_ms": [], "stage2_ms": []} for name, _ in variants}
    ref = {}
    for rnd in range(rounds + 1):                                      # round 0 warms every variant up (caches, pools)
        for name, env in variants:
            os.environ.update(env)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            aspace, gl = ArrowSpaceBuilder.build(gp, x)
            torch.cuda.synchronize(); t_build = (time.perf_counter() - t0) * 1e3
            st = {k: api.stat(k) for k in ("gram_ms", "lambda_ms")}
            aspace.search_batch(q[:4096], gl, 0.62)                    # builds the fp16 cache outside the timed step
            torch.cuda.synchronize(); t0 = time.perf_counter()
            idx, sc = aspace.search_batch(q, gl, 0.62)
            torch.cuda.synchronize(); t_step = (time.perf_counter() - t0) * 1e3
            s1, s2 = api.stat("search_stage1_ms"), api.stat("search_stage2_ms")
            lam = aspace.lambdas()
            for k in env:
                os.environ.pop(k)
            if rnd == 0:
                if name == "baseline":
                    ref = {"lam": lam, "idx": idx.cpu().numpy(), "sc": sc.cpu().numpy()}
                out[name]["bitwise_equal_to_baseline"] = bool(np.array_equal(lam, ref["lam"]) and np.array_equal(idx.cpu().numpy(), ref["idx"])
                                                               and np.array_equal(sc.cpu().numpy(), ref["sc"]))
            else:
                d = out[name]
                d["build_ms"].append(t_build); d["gram_ms"].append(st["gram_ms"]); d["lambda_ms"].append(st["lambda_ms"])
                d["step_ms"].append(t_step); d["stage1_ms"].append(s1); d["stage2_ms"].append(s2)
            del aspace, gl
    for name, d in out.items():
        med = {k: float(np.median(v)) for k, v in d.items() if isinstance(v, list) and v}
        d["median"] = med
        print("%-60s equal=%s  build %.2f (gram %.2f, lambda %.2f)  step %.2f (stage1 %.2f, stage2 %.2f)"
              % (name, d.get("bitwise_equal_to_baseline"), med["build_ms"], med["gram_ms"], med["lambda_ms"], med["step_ms"],
                 med["stage1_ms"], med["stage2_ms"]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ab_flags.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
