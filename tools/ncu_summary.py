"""Summarise ncu reports (gpurun_out/*.ncu-rep, gpurun_out/launches.csv) into profiles/ (tracked).

    python tools/ncu_summary.py <round-tag>
"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg.per_second", "launch__shared_mem_per_block_static"]


def short(name):
    for key in ("tc_gemm_kernel", "tc_rescore_kernel", "project_split_kernel", "median_kernel", "search_gemm_kernel", "gram_slice_kernel", "taumode_kernel", "rescore_kernel", "search_gemv_kernel",
                "feature_select_kernel", "gram_segment_reduce", "gram_final_reduce", "sort_rows", "fill_kernel",
                "count_mirror", "scan_", "weights_compact", "exact_", "topk_merge", "reciprocal", "zero_lambda",
                "knn_"):
        if key in name:
            return key
    return name[:60]


def launches(tag):
    path = os.path.join(OUT, "launches.csv")
    if not os.path.exists(path):
        return
    rows = [l for l in open(path) if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    per = defaultdict(list)
    order = []
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v_us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3 if unit in ("ms", "msecond") else v
        per[short(r["Kernel Name"])].append(v_us)
        order.append((r["ID"], short(r["Kernel Name"]), v_us))
    total = sum(sum(v) for v in per.values())
    with open(os.path.join(PROF, "launches_%s.md" % tag), "w") as fh:
        fh.write("# ncu launch list (%s): gpu__time_duration.sum, --clock-control none\n\n" % tag)
        fh.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES.\n\n")
        fh.write("| kernel | launches | total us | mean us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            fh.write("| %s | %d | %.1f | %.1f | %.1f%% |\n" % (k, len(v), sum(v), sum(v) / len(v), 100 * sum(v) / total))
        fh.write("\n## in launch order\n\n| id | kernel | us |\n|---|---|---:|\n")
        for i, k, v in order:
            fh.write("| %s | %s | %.1f |\n" % (i, k, v))
    print("wrote launches_%s.md: %d launches, %.1f ms total" % (tag, len(order), total / 1e3))


def full(tag, reports=None):
    summary = {}
    for fn in sorted(os.listdir(OUT)):
        if not fn.endswith(".ncu-rep") or (reports and fn not in reports):
            continue
        p = subprocess.run(["ncu", "-i", os.path.join(OUT, fn), "--page", "raw", "--csv"], capture_output=True, text=True)
        rows = list(csv.reader(io.StringIO(p.stdout)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = short(r[col["Kernel Name"]])
            ent = {"report": fn, "kernel": r[col["Kernel Name"]][:160]}
            for k in KEEP:
                if k in col:
                    try:
                        ent[k] = float(r[col[k]].replace(",", ""))
                        ent[k + "|unit"] = units[col[k]]
                    except ValueError:
                        pass
            summary.setdefault(name, []).append(ent)
    if summary:
        json.dump(summary, open(os.path.join(PROF, "ncu_full_%s.json" % tag), "w"), indent=1)
        print("wrote ncu_full_%s.json:" % tag, {k: len(v) for k, v in summary.items()})
    return summary


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    launches(tag)
    full(tag, set(sys.argv[2:]) or None)
