"""Latency / throughput of the search across call shapes on the C4-sized space (1M x 384 by default):
the reference's own shape (ONE query per ArrowSpace.search call), small batches, and the 64k host batch with and
without the copy/compute pipeline.  Writes gpurun_out/latency.json."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from pyarrowspace_b200 import api
from pyarrowspace_b200.api import ArrowSpaceBuilder

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
f = int(sys.argv[2]) if len(sys.argv) > 2 else 384
g = torch.Generator(device="cuda").manual_seed(1)
centres = torch.randn(256, f, generator=g, device="cuda", dtype=torch.float64)
lab = torch.randint(0, 256, (n,), generator=g, device="cuda")
x = centres[lab] + 0.3 * torch.randn(n, f, generator=g, device="cuda", dtype=torch.float64)
x = x / x.norm(dim=1, keepdim=True) * 100.0 + 25.0
gp = {"eps": 10.0, "k": 25, "topk": 10, "p": 2.0, "sigma": None}
aspace, gl = ArrowSpaceBuilder.build(gp, x)
sel = torch.randint(0, n, (65536,), generator=g, device="cuda")
q_dev = x[sel] / 100.0 + 0.01 * torch.randn(65536, f, generator=g, device="cuda", dtype=torch.float64)
q_pin = q_dev.cpu().pin_memory()
q_np = q_pin.numpy()
out = {"n": n, "f": f, "single": {}, "batch": {}, "host64k": {}}


def force(mode):
    if mode is None:
        os.environ.pop("ASP_SEARCH_STAGE1", None)
    else:
        os.environ["ASP_SEARCH_STAGE1"] = mode


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, r


ref = None
for mode in ("fp64", "tc"):
    force(mode)
    it = iter(range(10 ** 9))
    dt, r = timed(lambda: aspace.search(q_np[next(it) % 512], gl, 0.62), 200)
    out["single"][mode] = {"ms_per_query": dt * 1e3, "queries_per_s": 1.0 / dt,
                           "hbm_gbs_if_f64_scan": 8.0 * n * f / dt / 1e9}
    print("single", mode, out["single"][mode], flush=True)
    for nq in (8, 64, 256, 1024, 4096, 16384):
        if mode == "fp64" and nq > 4096:
            continue
        dt, r = timed(lambda: aspace.search_batch(q_dev[:nq], gl, 0.62), 5 if nq >= 1024 else 20)
        out["batch"].setdefault(str(nq), {})[mode] = {"ms": dt * 1e3, "queries_per_s": nq / dt,
                                                      "is_tc": api.stat("search_stage1_is_tc"), "slow": api.stat("search_slow_queries")}
        print("batch", nq, mode, out["batch"][str(nq)][mode], flush=True)
        if nq == 1024:
            if ref is None:
                ref = (r[0].cpu().numpy(), r[1].cpu().numpy())
            else:
                assert np.array_equal(ref[0], r[0].cpu().numpy()) and np.array_equal(ref[1], r[1].cpu().numpy()), "paths differ"
force(None)
variants = [("pipelined", {}), ("head3k", {"ASP_PIPE_HEAD": "3072"}), ("head1k", {"ASP_PIPE_HEAD": "1024"}),
            ("single_shot", {"ASP_NO_PIPELINE": "1"})]
for label, env in variants:                            # warm every variant once
    os.environ.update(env); aspace.search_batch(q_np, gl, 0.62)
    for k in env:
        os.environ.pop(k)
for rnd in range(16):                                  # interleaved call by call: box drift hits every variant alike
    for label, env in variants:
        os.environ.update(env)
        t0 = time.perf_counter(); r2 = aspace.search_batch(q_np, gl, 0.62); dt = (time.perf_counter() - t0) * 1e3
        for k in env:
            os.environ.pop(k)
        d = out["host64k"].setdefault(label, {"ms_calls": []})
        d["ms_calls"].append(dt)
        d["chunks"] = api.stat("search_pipeline_chunks")
        if rnd >= 12 and d["chunks"] == 2.0:
            d.setdefault("timeline", []).append({k: round(api.stat("search_pipe_" + k + "_ms"), 2) for k in ("up0", "up1", "start0", "done0", "start1", "done1")} | {"wall": round(dt, 2)})
    torch.cuda.synchronize(); t0 = time.perf_counter(); r2 = aspace.search_batch(q_dev, gl, 0.62); torch.cuda.synchronize()
    out["host64k"].setdefault("device_resident_calls", {"ms_calls": []})["ms_calls"].append((time.perf_counter() - t0) * 1e3)
for label, d in out["host64k"].items():
    d["ms_median"] = float(np.median(d["ms_calls"]))
    d["queries_per_s"] = 65536 / d["ms_median"] * 1e3
    print("host64k", label, "median %.2f ms" % d["ms_median"], "min %.2f" % min(d["ms_calls"]), d.get("chunks"), d.get("timeline"), flush=True)
dt, r = timed(lambda: aspace.search_batch(q_dev, gl, 0.62), 5)
out["host64k"]["device_resident"] = {"ms": dt * 1e3, "queries_per_s": 65536 / dt}
print("dev64k", out["host64k"]["device_resident"], flush=True)
# raw PCIe figures for the same buffers
dst = torch.empty_like(q_dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(2):
    e0.record(); dst.copy_(q_pin, non_blocking=True); e1.record(); e1.synchronize()
out["h2d_201MB_ms"] = e0.elapsed_time(e1)
print("h2d", out["h2d_201MB_ms"], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "latency.json"), "w"), indent=1)
