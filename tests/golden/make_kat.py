"""Regenerates tests/golden/oracle_vectors.npz: oracle outputs on the reference's two KAT inputs and on
two small seeded synthetic cases (+ the item graph of a third, + the pre-graph reduction of a fourth, + the hybrid search on the first).  kat.json itself is transcribed from the reference's README.md:37-69
and tests/test_0.py:4-61 (the reference engine -- crate arrowspace 0.18.0 -- is not vendored and cannot
be imported here, so there is no reference run to record).

    python tests/golden/make_kat.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import oracle  # noqa: E402
from pyarrowspace_b200 import synth  # noqa: E402


def run_case(items, gp, queries, tau):
    s, g = oracle.build(gp, items)
    indptr, indices, data = g.csr()
    idx, sc, lq = s.search_batch(queries, g, tau)
    return dict(indptr=indptr, indices=indices, data=data, lambdas=s.lambdas(), idx=idx, score=sc, lambda_q=lq)


def main():
    kat = json.load(open(os.path.join(HERE, "kat.json")))
    out = {}
    r = kat["readme"]
    for k, v in run_case(np.array(r["items"]), r["graph_params"], np.array([r["query"]]), r["tau"]).items():
        out["readme_" + k] = v
    t = kat["test_0"]
    it = np.array(t["items"])
    q = (it[t["query_item"]] * t["query_scale"]).reshape(1, -1)
    for tau in t["expected_top3"]:
        for k, v in run_case(it, t["graph_params"], q, float(tau)).items():
            out["test0_%s_%s" % (tau, k)] = v
    for name, (n, f, seed, gp, nq) in {"synthA": (600, 48, 5, {"eps": 0.6, "k": 5, "topk": 10, "p": 2.0, "sigma": 0.3}, 16),
                                       "synthB": (1500, 100, 6, {"eps": 1.0, "k": 12, "topk": 5, "p": 2.0, "sigma": None}, 16)}.items():
        x = synth.make_items(n, f, seed, scale=100.0, n_clusters=16)
        qs, _ = synth.make_queries(x, nq, seed)
        for k, v in run_case(x, gp, qs, 0.62).items():
            out["%s_%s" % (name, k)] = v
    # item graph (nodes = items): Laplacian CSR only
    x = synth.make_items(2500, 48, 7, scale=100.0, n_clusters=12)
    s, g = oracle.build({"eps": 0.5, "k": 8, "topk": 3, "p": 2.0, "sigma": 0.2}, x, nodes="items")
    indptr, indices, data = g.csr()
    out.update(itemsC_indptr=indptr, itemsC_indices=indices.astype(np.int32), itemsC_data=data)
    # pre-graph reduction: centroids, statistics, the centroid graph and the lambdas it gives every item
    x = synth.make_items(3000, 40, 8, scale=100.0, n_clusters=10)
    s, g, cent, info = oracle.build_reduced({"eps": 0.6, "k": 5, "topk": 5, "p": 2.0, "sigma": 0.3}, x, reduction={"max_iters": 6})
    indptr, indices, data = g.csr()
    out.update(reducedD_centroids=cent, reducedD_indptr=indptr, reducedD_indices=indices.astype(np.int32), reducedD_data=data,
               reducedD_lambdas=s.lambdas(),
               reducedD_info=np.array([info["n_sampled"], info["n_probes"], info["two_nn_mean_ratio"], info["intrinsic_dim"],
                                       info["n_clusters"], info["iters"], info["converged"]], dtype=np.float64))
    # hybrid search: default shortlist and an explicit one, on the synthA inputs
    x = synth.make_items(600, 48, 5, scale=100.0, n_clusters=16)
    qs, _ = synth.make_queries(x, 16, 5)
    s, g = oracle.build({"eps": 0.6, "k": 5, "topk": 10, "p": 2.0, "sigma": 0.3}, x)
    for tag, pool in (("hybridE", 0), ("hybridE_pool13", 13)):
        idx, sc, lq = s.search_hybrid_batch(qs, g, 0.3, pool)      # tau 0.3: the lambda term decides places
        out.update({tag + "_idx": idx, tag + "_score": sc, tag + "_lambda_q": lq})
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
