"""First-contact diagnostics for a GPU box: FP64 / HBM peaks, then the hot path stage by stage against
the oracle, printing errors instead of asserting.  Writes gpurun_out/diag.json."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import oracle  # noqa: E402
from pyarrowspace_b200 import _lib, api, synth  # noqa: E402
from arrowspace import ArrowSpaceBuilder  # noqa: E402

out = {}


def ev_time(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts) // 2]


def peaks():
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_properties(0).multi_processor_count, "SMs")
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    best, med = ev_time(lambda: torch.matmul(a, b), reps=5, warm=2)
    out["fp64_dgemm_tflops_best"] = 2 * n ** 3 / best / 1e9
    out["fp64_dgemm_tflops_median"] = 2 * n ** 3 / med / 1e9
    t0 = time.time()
    cnt = 0
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    while time.time() - t0 < 3.0:
        torch.matmul(a, b)
        cnt += 1
        torch.cuda.synchronize()
    s1.record()
    torch.cuda.synchronize()
    out["fp64_dgemm_tflops_sustained"] = cnt * 2 * n ** 3 / s0.elapsed_time(s1) / 1e9
    del a, b
    x = torch.empty(1 << 28, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    best, med = ev_time(lambda: y.copy_(x), reps=8, warm=2)
    out["hbm_copy_gbs_best"] = 2 * x.numel() * 8 / best / 1e6
    del x, y
    print("peaks", {k: round(v, 1) for k, v in out.items()})


def stage_checks(n, f, gp, nq, tag):
    res = {}
    x = synth.make_items(n, f, 17, n_clusters=32)
    q, sel = synth.make_queries(x, nq, 17)
    t = time.time()
    s, g = oracle.build(gp, x)
    res["oracle_build_s"] = time.time() - t
    lib, ctx = _lib.load(), _lib.context()
    # Gram
    hs = C.c_void_p()
    _lib.check(lib.asp_space_create(ctx, x.ctypes.data, n, f, n, 1, 0, C.byref(hs)))
    segs = torch.zeros((8, f, f), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    _lib.check(lib.asp_space_gram_partials(hs, segs.data_ptr()))
    _lib.check(lib.asp_ctx_synchronize(ctx))
    gram = segs.cpu().numpy().sum(0)
    ref = oracle.gram_columns(x)
    scale = np.sqrt(np.outer(np.diag(ref), np.diag(ref)))
    res["gram_max_rel_err"] = float((np.abs(gram - ref) / scale).max())
    res["gram_symmetric"] = bool(np.array_equal(gram, gram.T))
    lib.asp_free_space(hs)
    # full build
    t = time.time()
    aspace, gl = ArrowSpaceBuilder.build(gp, x)
    res["gpu_build_wall_s"] = time.time() - t
    for k in ("upload_ms", "gram_ms", "graph_ms", "lambda_ms", "need_exact_pairs"):
        res[k] = api.stat(k)
    ip, ix, dt = gl.csr()
    oip, oix, odt = g.csr()
    res["csr_indptr_equal"] = bool(np.array_equal(ip, oip))
    res["csr_indices_equal"] = bool(ip.shape == oip.shape and ix.shape == oix.shape and np.array_equal(ix, oix))
    res["nnz"] = [int(gl.nnz), int(g.nnz)]
    if res["csr_indices_equal"]:
        res["csr_data_max_rel"] = float((np.abs(dt - odt) / np.maximum(np.abs(odt), 1e-300)).max())
    lam, olam = aspace.lambdas(), s.lambdas()
    res["lambda_max_rel"] = float((np.abs(lam - olam) / np.abs(olam)).max())
    res["norms_equal"] = bool(np.array_equal(aspace.norms(), s.norms()))
    # search
    for tau in (0.62, 1.0):
        t = time.time()
        idx, sc = aspace.search_batch(q, gl, tau)
        res["gpu_search_wall_s_%s" % tau] = time.time() - t
        res["search_stage1_ms_%s" % tau] = api.stat("search_stage1_ms")
        res["slow_queries_%s" % tau] = api.stat("search_slow_queries")
        m = min(nq, 256)
        t = time.time()
        oidx, osc, olq = s.search_batch(q[:m], g, tau)
        res["oracle_search_s_per_query_%s" % tau] = (time.time() - t) / m
        res["idx_equal_%s" % tau] = bool(np.array_equal(idx[:m], oidx))
        res["idx_mismatch_rows_%s" % tau] = int((idx[:m] != oidx).any(axis=1).sum())
        ok = oidx >= 0
        res["score_max_rel_%s" % tau] = float((np.abs(sc[:m][ok] - osc[ok]) / np.abs(osc[ok])).max())
        res["top1_is_source_%s" % tau] = float((idx[:, 0] == sel).mean())
    one = aspace.search(q[1], gl, 0.62)
    res["gemv_equals_gemm"] = bool([i for i, _ in one] == list(aspace.search_batch(q[:16], gl, 0.62)[0][1]))
    out[tag] = res
    print(tag, json.dumps(res, indent=1))


if __name__ == "__main__":
    which = sys.argv[1:] or ["peaks", "small", "mid"]
    if "peaks" in which:
        peaks()
    if "small" in which:
        stage_checks(2000, 48, {"eps": 0.5, "k": 5, "topk": 10, "p": 2.0, "sigma": 0.25}, 300, "small")
    if "mid" in which:
        stage_checks(50000, 384, {"eps": 0.5, "k": 4, "topk": 10, "p": 2.0, "sigma": 0.25}, 2048, "mid")
    if "big" in which:
        stage_checks(200000, 384, {"eps": 10.0, "k": 25, "topk": 10, "p": 2.0, "sigma": None}, 8192, "big")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "diag_%s.json" % "_".join(which)), "w"), indent=1)
