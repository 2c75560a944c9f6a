"""Pre-graph reduction at the C4 shape on one GPU: stage times, the reported statistics, and what the
centroid graph does to lambdas / results compared with the plain build (behavioural, not a parity statement)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from pyarrowspace_b200 import synth  # noqa: E402
from pyarrowspace_b200.api import ArrowSpaceBuilder, stat  # noqa: E402


def main():
    cfg = synth.config(os.environ.get("CFG", "C4"))
    n, f, gp = int(os.environ.get("N", cfg["n"])), cfg["f"], cfg["graph_params"]
    x = synth.make_items(n, f, 44, 100.0)
    xd = torch.from_numpy(x).cuda()
    out = {"n": n, "f": f}
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        aspace, gl = ArrowSpaceBuilder.build(gp, xd, reduction=True)
        torch.cuda.synchronize(); dt = time.time() - t0
        out["reduced_build_s"] = dt
        out["info"] = gl.reduction
        out["stages_ms"] = {k: stat(k) for k in ("reduce_ms", "reduce_two_nn_ms", "reduce_kmeans_ms", "reduce_assign_ms",
                                                   "reduce_assign_passes", "gram_ms", "graph_ms", "lambda_ms")}
        print("rep", rep, json.dumps(out), flush=True)
    ns, K = out["info"]["n_sampled"], out["info"]["n_clusters"]
    passes = out["stages_ms"]["reduce_assign_passes"]
    dp = 3.0 * ns * K * f * passes
    out["assign_dp_ops_per_s"] = dp / (out["stages_ms"]["reduce_assign_ms"] * 1e-3)
    t0 = time.time()
    a0, g0 = ArrowSpaceBuilder.build(gp, xd)
    torch.cuda.synchronize(); out["plain_build_s"] = time.time() - t0
    lam_r, lam_p = aspace.lambdas(), a0.lambdas()
    out["lambda_corr_with_plain_build"] = float(np.corrcoef(lam_r, lam_p)[0, 1])
    q, _ = synth.make_queries(x[:65536], 2048, 44, 100.0)
    ir, _ = aspace.search_batch(q, gl, 0.62)
    ip, _ = a0.search_batch(q, g0, 0.62)
    out["top10_overlap_with_plain_build"] = float(np.mean([len(set(a) & set(b)) / 10.0 for a, b in zip(ir, ip)]))
    out["graph_nnz"] = {"reduced": gl.nnz, "plain": g0.nnz}
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "reduce_time.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
