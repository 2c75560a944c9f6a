#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

metric   : search queries/s (top-10) on the BEIR/MS MARCO-shaped synthetic 1M x 384 f64 workload (C4),
           with the build (feature graph + Laplacian + lambda) timed beside it as items/s.
step     : one pass of the search hot path over one batch of Q synthetic queries (default 16384)
           against all N items (row-sharded over the ranks when --gpus > 1, results merged).
value    : whole-job queries/s with items and queries resident in HBM.
e2e      : the same through the public API (ArrowSpace.search_batch) with HOST buffers: pinned
           host -> device copy of the batch and device -> host read of (idx, score) inside the timed region.
roofline : dominant kernel = search_gemm_kernel (FP64 DMMA): 2*Q*N_local*F algorithmic FLOP per launch /
           its CUDA-event duration; peak = cuBLAS DGEMM measured in this run (MEASURED_PEAKS.json has no
           FP64 figure).  The build's kernels are reported under "build".
cpu_baseline / --impl reference : the CPU oracle (oracle/, a restatement: the reference's Rust engine is not
           vendored and cannot be built here) timed on this box's host cores.

Launch: python bench.py --gpus N --steps K --warmup W   (N > 1 under torchrun, one rank per GPU).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLAGSHIP = "C4"


_REAL_STDOUT = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version banner on the
    first communicator), so file descriptor 1 is pointed at stderr for the whole run and the line goes to the saved one."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--items", type=int, default=None, help="override N (default: the C4 shape, 1,000,000)")
    ap.add_argument("--features", type=int, default=None)
    ap.add_argument("--queries", type=int, default=65536, help="queries per step (SURVEY.md 8(d): C4 is batched 64k per call)")
    ap.add_argument("--build-reps", type=int, default=3)
    ap.add_argument("--cpu-sample-items", type=int, default=100_000)
    ap.add_argument("--cpu-sample-queries", type=int, default=64)
    ap.add_argument("--ref-queries", type=int, default=256, help="--impl reference: queries per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-item-graph", action="store_true", help="skip the item-graph (nodes = items) build of the C4 matrix")
    return ap.parse_args()


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        # median over the samples taken under load (>= 60 % of the highest draw seen)
        load = [s for s, p in zip(sm, pw) if pw and p >= 0.6 * max(pw)] or sm
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- CPU oracle legs
def cpu_oracle_run(cfg, n_items, nq, reps, label):
    """Oracle build on n_items rows of the workload + `reps` search batches of nq queries, all host threads."""
    import oracle
    from pyarrowspace_b200 import synth
    # all the host threads this process may use: torchrun exports OMP_NUM_THREADS=1 to every rank, which would turn the
    # CPU arm into a single-thread run at N > 1 (rank 0 is the only rank that works here; the others exit)
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    oracle.set_num_threads(cores)
    x = synth.make_items(cfg["n"], cfg["f"], cfg["seed"], cfg["scale"], rows=(0, n_items))
    q, _ = synth.make_queries(x, nq, cfg["seed"], cfg["scale"])
    t0 = time.perf_counter()
    s, g = oracle.build(cfg["graph_params"], x)
    t_build = time.perf_counter() - t0
    s.search_batch(q[: max(1, nq // 8)], g, cfg["tau"])              # warm-up
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        s.search_batch(q, g, cfg["tau"])
        ts.append(time.perf_counter() - t0)
    return {"build_s": t_build, "search_s": ts, "threads": oracle.num_threads(), "n_items": n_items, "nq": nq}


def run_reference(args, cfg):
    """--impl reference: the CPU implementation of the path (the oracle port) on the host cores, same config/metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.items or cfg["n"]
    cfg = dict(cfg, n=n)
    nq = args.ref_queries
    r = cpu_oracle_run(cfg, n, nq, args.warmup + args.steps, "reference")
    ts = r["search_s"][args.warmup:]
    total = sum(ts)
    val = nq * len(ts) / total
    line = {
        "impl": "reference", "metric": "search queries/s (top-10, 1M x 384 f64)", "value": val, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(ts),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C4 BEIR/MS MARCO-shaped synthetic %d x %d f64, top-%d, tau %.2f" % (n, cfg["f"], cfg["graph_params"]["topk"], cfg["tau"]),
                   "queries_per_step": nq, "graph_params": cfg["graph_params"]},
        "cpu_baseline": {"value": val, "unit": "queries/s", "cores": r["threads"], "kind": "port",
                         "sample": "oracle (C + OpenMP) search of %d queries per step against all %d items; build %.1f s (%.0f items/s) untimed"
                                   % (nq, n, r["build_s"], n / r["build_s"])},
        "build": {"items_per_s": n / r["build_s"], "seconds": r["build_s"]},
        "e2e": {"value": val, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def gram_roofline(n_local, f, ms, peak):
    """K1: 2*N*F^2 algorithmic FLOP; the kernel executes only the upper-triangular 128x128 output tiles
    (symmetry), so `frac` is EXECUTED FLOP / peak and the symmetry gain is reported separately."""
    nt = (f + 3) // 4 * 4
    nt = (nt + 127) // 128
    executed = 2.0 * n_local * (nt * (nt + 1) // 2) * 128 * 128
    algorithmic = 2.0 * n_local * f * f
    return {"ms": ms, "bound": "tensor", "algorithmic_flop": algorithmic, "executed_flop": executed,
            "achieved_tflops": executed / (ms * 1e-3) / 1e12, "frac": executed / (ms * 1e-3) / 1e12 / peak,
            "symmetry_speedup_vs_algorithmic": algorithmic / executed,
            "note": "ms is the stage time (slice kernel + reduces + scratch allocation), not the kernel alone"}


# --------------------------------------------------------------------------- GPU arm
def main():
    args = parse_args()
    quiet_stdout()
    from pyarrowspace_b200 import synth
    cfg = synth.config(FLAGSHIP)
    if args.features:
        cfg["f"] = args.features
    if args.impl == "reference":
        return run_reference(args, cfg)

    import torch
    import torch.distributed as dist
    from pyarrowspace_b200 import _lib, api
    from pyarrowspace_b200.api import ArrowSpaceBuilder, shard_rows

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.items or cfg["n"]
    f, gp, tau, Q = cfg["f"], cfg["graph_params"], cfg["tau"], args.queries
    topk = gp["topk"]
    lib = _lib.load()
    ctx = _lib.context(local)
    stream = torch.cuda.Stream(device=dev)
    _lib.check(lib.asp_ctx_set_stream(ctx, stream.cuda_stream))       # the library's kernels go on this stream

    def barrier_sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- FP64 tensor peak on this GPU (cuBLAS DGEMM), the roofline denominator for the DMMA kernels
    with torch.cuda.stream(stream):
        a = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
        b = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
        best = 1e9
        for i in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); torch.matmul(a, b); e1.record(stream); e1.synchronize()
            if i:
                best = min(best, e0.elapsed_time(e1))
        fp64_peak_tflops = 2 * 4096 ** 3 / best / 1e9
        del a, b

    # ---- synthetic inputs: this rank's row shard (identical bytes to what the oracle sees)
    r0, r1 = shard_rows(n, world, rank)
    x_host = torch.from_numpy(synth.make_items(n, f, cfg["seed"], cfg["scale"], rows=(r0, r1))).pin_memory()
    nbatch = 2
    qrng_src = synth.make_items(n, f, cfg["seed"], cfg["scale"], rows=(0, min(n, 65536)))
    q_host = [torch.from_numpy(synth.make_queries(qrng_src, Q, cfg["seed"] + b_, cfg["scale"])[0]).pin_memory()
              for b_ in range(nbatch)]
    with torch.cuda.stream(stream):
        x_dev = x_host.to(dev, non_blocking=True)
        q_dev = [q.to(dev, non_blocking=True) for q in q_host]
    stream.synchronize()

    def build(items):
        if world == 1:
            return ArrowSpaceBuilder.build(gp, items, device=local)
        return ArrowSpaceBuilder.build_sharded(gp, items, n, device=local)

    # ---- build: items resident in HBM (value) and from pinned host memory (e2e)
    launches0 = lib.asp_ctx_launch_count(ctx)
    build_ms, build_e2e_ms, stages = [], [], {}
    with torch.cuda.stream(stream):
        for rep in range(1 + args.build_reps):
            barrier_sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            aspace, gl = build(x_dev)
            e1.record(stream)
            barrier_sync()
            if rep:
                build_ms.append(max_over_ranks(e0.elapsed_time(e1)))
                for k in ("gram_ms", "graph_ms", "lambda_ms"):
                    stages.setdefault(k, []).append(api.stat(k, local))
            del aspace, gl
        for rep in range(2):
            barrier_sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            aspace, gl = build(x_host.numpy())
            e1.record(stream)
            barrier_sync()
            build_e2e_ms.append(max_over_ranks(e0.elapsed_time(e1)))
            if rep == 0:
                del aspace, gl
    build_launches = (lib.asp_ctx_launch_count(ctx) - launches0) // (3 + args.build_reps)
    feature_graph_nnz = int(gl.nnz)

    # ---- search: W warm-up steps, then exactly K timed steps
    clocks = ClockSampler(local)
    with torch.cuda.stream(stream):
        for w in range(args.warmup):
            aspace.search_batch(q_dev[w % nbatch], gl, tau)
        barrier_sync()
        if rank == 0:
            clocks.start()
        l0 = lib.asp_ctx_launch_count(ctx)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stage1 = []
        e0.record(stream)
        for k in range(args.steps):
            idx, sc = aspace.search_batch(q_dev[k % nbatch], gl, tau)
            stage1.append(api.stat("search_stage1_ms", local))
        e1.record(stream)
        barrier_sync()
        step_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
        launches = lib.asp_ctx_launch_count(ctx) - l0
        slow = api.stat("search_slow_queries", local)
        stage1_is_tc = api.stat("search_stage1_is_tc", local) == 1.0
        rescored = api.stat("search_rescored_per_query", local) if stage1_is_tc else None
        sstat = {k: api.stat(k, local) for k in ("search_terms", "search_stage2_ms", "search_a_resident", "search_delta_cos_max",
                                                  "search_rho_q_max", "search_rho_x_max", "search_exact_per_query")}
        # end to end: pinned host queries in, host results out, every step
        for w in range(2):
            aspace.search_batch(q_host[w % nbatch].numpy(), gl, tau)
        barrier_sync()
        e0.record(stream)
        for k in range(args.steps):
            idx_h, sc_h = aspace.search_batch(q_host[k % nbatch].numpy(), gl, tau)
        e1.record(stream)
        barrier_sync()
        e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    clk = clocks.stop() if rank == 0 else None

    # ---- the reference's own call shape: ONE query per ArrowSpace.search call (src/lib.rs:132-174), host vector in,
    # Python list of (index, score) out; wall clock per call (includes the ctypes call, the upload and the read-back)
    single = None
    if world == 1:
        qs = q_host[0].numpy()
        for i in range(20):
            aspace.search(qs[i], gl, tau)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for i in range(200):
            hits = aspace.search(qs[20 + i], gl, tau)
        dt = (time.perf_counter() - t0) / 200
        single = {"ms_per_call": dt * 1e3, "queries_per_s": 1.0 / dt,
                  "f64_scan_equivalent_gbs": 8.0 * n * f / dt / 1e9,
                  "note": "candidate pass streams the fp16 operands (%.2f GB) instead of the f64 rows (%.2f GB); exact f64 stage 2"
                          % (2.0 * n * (((f + 3 + 63) // 64) * 64) / 1e9, 8.0 * n * f / 1e9)}

    # ---- item graph (nodes = items) of the same matrix: the graph-build workload of C4 (eps / k-NN lists of all 1M items
    # against all 1M items on the tensor cores, exact stage 2, Laplacian CSR); sharded over the ranks when N > 1
    item_graph = None
    if not args.no_item_graph:
        del aspace, gl
        ig_ms = []
        with torch.cuda.stream(stream):
            for rep in range(2):
                barrier_sync()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                if world == 1:
                    a_ig, g_ig = ArrowSpaceBuilder.build_item_graph(gp, x_dev, device=local)
                else:
                    a_ig, g_ig = ArrowSpaceBuilder.build_item_graph_sharded(gp, x_dev, n, r0, device=local)
                e1.record(stream)
                barrier_sync()
                ig_ms.append(max_over_ranks(e0.elapsed_time(e1)))
                ig_nnz = g_ig.nnz
                ig_stats = {k: api.stat(k, local) for k in ("knn_stage1_ms", "knn_stage2_ms", "knn_slow_rows", "knn_rescored_per_row")}
                del a_ig, g_ig
        ig_exec = 2.0 * n * (r1 - r0) * 16.0 * ((f + 3 + 15) // 16)
        item_graph = {"ms": min(ig_ms), "items_per_s": n / (min(ig_ms) * 1e-3), "nnz": int(ig_nnz),
                      "rows_per_rank": r1 - r0, "stage1_ms": ig_stats["knn_stage1_ms"], "stage2_ms": ig_stats["knn_stage2_ms"],
                      "exact_scan_rows": ig_stats["knn_slow_rows"], "rescored_per_row": ig_stats["knn_rescored_per_row"],
                      "algorithmic_flop": 2.0 * n * (r1 - r0) * f, "executed_fp16_flop": ig_exec,
                      "stage1_tflops": ig_exec / (ig_stats["knn_stage1_ms"] * 1e-3) / 1e12 if ig_stats["knn_stage1_ms"] > 0 else None,
                      "note": "rows of this rank against all %d items; N > 1: + all-gather of the item shards and of the lists" % n}

    # sanity: a perturbed copy must find its source item (size-independent property)
    stage1_ms = max_over_ranks(float(np.mean(stage1)))
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    n_local = r1 - r0
    gemm_flop = 2.0 * Q * n_local * f                       # algorithmic: one f64-accurate score per (query, item)
    gram_ms = float(np.mean(stages["gram_ms"]))
    lam_ms = float(np.mean(stages["lambda_ms"]))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    prof = {}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "roofline.json")))
    except Exception:
        pass
    if stage1_is_tc:
        terms = int(sstat["search_terms"])
        ksteps = (f + 3 + 15) // 16 + (2 * ((f + 15) // 16) if terms == 3 else 0)   # K=16 MMA steps per (query, item) tile pair
        executed = 2.0 * Q * n_local * 16.0 * ksteps
        sustained = peaks.get("bf16_tflops_sustained", 1400.0)
        burst = peaks.get("bf16_tflops", sustained)
        achieved = executed / (stage1_ms * 1e-3) / 1e12
        peak = sustained if achieved <= sustained else burst
        roofline = {"kernel": "tc_gemm_kernel (tcgen05.mma kind::f16, fp16 operands, f32 TMEM accumulators, TMA SWIZZLE_128B; items in "
                              "lambda order, rank-1 mean-direction term + %s of the residuals; single-compare epilogue, thresholds "
                              "shared across CTAs); exact f64 stage 2 follows"
                              % ("ONE fp16 term" if terms == 1 else "the two-term fp16 split (3 MMA terms)"),
                    "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "peak_nominal": 2250.0, "frac_of_nominal": achieved / 2250.0,     # dense 16-bit datasheet figure (report both)
                    "traffic": prof.get("tc_gemm_dram_bytes_per_launch"),
                    "algorithmic": "2*Q*N_local*F = %.3e FLOP per launch; EXECUTED 2*Q*N_local*16*%d = %.3e fp16 tensor FLOP "
                                   "(achieved/frac count the executed FLOP against the measured 16-bit dense tensor peak)"
                                   % (gemm_flop, ksteps, executed),
                    "algorithmic_tflops": gemm_flop / (stage1_ms * 1e-3) / 1e12,
                    "fp64_tensor_peak_tflops": fp64_peak_tflops,
                    "kernel_ms": stage1_ms, "share_of_step": stage1_ms / step_ms,
                    "stage2_ms": sstat["search_stage2_ms"],
                    "mma_terms": terms, "query_operand_resident": sstat["search_a_resident"] == 1.0,
                    "band_cos_max": sstat["search_delta_cos_max"],
                    "residual_norms": {"rho_q_max": sstat["search_rho_q_max"], "rho_x_max": sstat["search_rho_x_max"]},
                    "rescored_candidates_per_query": rescored,
                    "reference_order_rescored_per_query": sstat["search_exact_per_query"],
                    "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peak == sustained
                                    else "MEASURED_PEAKS.json bf16_tflops (burst): the kernel ran above the sustained figure %.1f" % sustained)
                                   if "bf16_tflops_sustained" in peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"}
    else:
        achieved = gemm_flop / (stage1_ms * 1e-3) / 1e12
        roofline = {"kernel": "search_gemm_kernel (FP64 DMMA.8x8x4, TMA fed, fused score/top-k epilogue)",
                    "bound": "tensor", "achieved": achieved, "peak": fp64_peak_tflops, "unit": "TFLOP/s",
                    "frac": achieved / fp64_peak_tflops, "traffic": prof.get("search_gemm_dram_bytes_per_launch"),
                    "algorithmic": "2*Q*N_local*F = %.3e FLOP per launch" % gemm_flop,
                    "kernel_ms": stage1_ms, "share_of_step": stage1_ms / step_ms,
                    "peak_source": "cuBLAS DGEMM 4096^3 measured in this run (MEASURED_PEAKS.json has no FP64 entry)"}
    line = {
        "metric": "search queries/s (top-10, 1M x 384 f64)",
        "value": Q / (step_ms * 1e-3), "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C4 BEIR/MS MARCO-shaped synthetic %d x %d f64 items (seed %d, x%g), %d queries per step, top-%d, tau %.2f"
                               % (n, f, cfg["seed"], cfg["scale"], Q, topk, tau),
                   "graph_params": gp, "queries_per_step": Q, "sharding": "rows over %d rank(s)" % world,
                   "l2": "inputs larger than L2 (item shard %.2f GB)" % (n_local * f * 8 / 1e9),
                   "stage1": "tcgen05 fp16 candidates + exact f64 rescoring" if stage1_is_tc else "FP64 DMMA",
                   "exact_rescan_queries_last_step": slow},
        "e2e": {"value": Q / (e2e_ms * 1e-3), "unit": "queries/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": Q * f * 8, "d2h_bytes_per_step": Q * topk * 16},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "build": {"items_per_s": n / (float(np.mean(build_ms)) * 1e-3), "ms": float(np.mean(build_ms)),
                  "e2e_items_per_s": n / (min(build_e2e_ms) * 1e-3), "e2e_ms": min(build_e2e_ms),
                  "h2d_bytes": n_local * f * 8, "gpu_launches": int(build_launches),
                  "gram": gram_roofline(n_local, f, gram_ms, fp64_peak_tflops),
                  "graph_ms": float(np.mean(stages["graph_ms"])),
                  "lambda": {"ms": lam_ms, "bound": "hbm", "achieved_gbs": (8.0 * n_local * f + 8.0 * n_local) / (lam_ms * 1e-3) / 1e9,
                             "frac": (8.0 * n_local * f + 8.0 * n_local) / (lam_ms * 1e-3) / 1e9 / hbm_peak, "peak_gbs": hbm_peak,
                             # the bound that actually binds at this graph density: one 8-byte shared-memory gather per
                             # strictly-upper non-zero of L per item, against 128 B/clk/SM
                             "gather": {"upper_nnz": (feature_graph_nnz - f) // 2,
                                        "smem_bytes": 8.0 * n_local * ((feature_graph_nnz - f) // 2),
                                        "smem_floor_ms": 8.0 * n_local * ((feature_graph_nnz - f) // 2) /
                                                         (148 * 128.0 * (clk or {}).get("sm_mhz", 1900.0) * 1e6) * 1e3,
                                        "note": "lambda_ms = median_kernel (radix selection, HBM pass 1) + taumode_kernel (HBM pass 2 + gathers)"}}},
        "item_graph": item_graph,
        "single_query": single,
        "clocks": clk,
    }
    if not args.no_cpu_baseline:
        ns = min(n, args.cpu_sample_items)
        r = cpu_oracle_run(dict(cfg, n=n), ns, args.cpu_sample_queries, 3, "sample")
        per_q = min(r["search_s"]) / r["nq"] * (n / ns)           # a scan is linear in the item count
        line["cpu_baseline"] = {
            "value": 1.0 / per_q, "unit": "queries/s", "cores": r["threads"], "kind": "port",
            "sample": "oracle (C + OpenMP) on the first %d items and %d queries of the workload; search time scaled x%.0f "
                      "to %d items (linear scan); oracle build of the sample %.2f s = %.0f items/s"
                      % (ns, r["nq"], n / ns, n, r["build_s"], ns / r["build_s"]),
            "build_items_per_s": ns / r["build_s"]}
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
