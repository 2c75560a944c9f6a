"""BASELINE.json config C5 (MS MARCO eps sweep): item graph (nodes = items) of N x 768 f64 synthetic embeddings across the
GPUs of one box -- all-gather of the item shards over NCCL (the halo rows), every rank resolves its rows against ALL
items on the tensor cores, all-gather of the neighbour lists, Laplacian CSR on every rank.

    torchrun --nproc-per-node 8 tools/c5_sweep.py [N=8800000] [F=768] [eps list=10,5,15]

Shards are generated on the device (seed 45, per 65536-row block, so any rank can regenerate any block).  There is no
oracle at this size; correctness is checked through a planted property: the first 1000 rows of every rank are perturbed
copies (1e-4 relative) of rows owned by the NEXT rank, so each must list its source as a neighbour (and the source it).
Writes gpurun_out/c5_sweep.json (rank 0)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from pyarrowspace_b200 import api  # noqa: E402
from pyarrowspace_b200.api import ArrowSpaceBuilder  # noqa: E402

SEED, BLOCK, NCL, PLANT, PLANT_OFF = 45, 65536, 256, 1000, 5000


def gen_block(b, n, f, centres, dev):
    g = torch.Generator(device=dev).manual_seed(SEED * 100003 + b)
    m = min(BLOCK, n - b * BLOCK)
    lab = torch.randint(0, NCL, (m,), generator=g, device=dev)
    x = centres[lab] + 0.3 * torch.randn(m, f, generator=g, device=dev, dtype=torch.float64)
    x /= x.norm(dim=1, keepdim=True)
    return x.mul_(100.0).add_(25.0)


def gen_rows(r0, r1, n, f, centres, dev):
    out = torch.empty((r1 - r0, f), dtype=torch.float64, device=dev)
    for b in range(r0 // BLOCK, (r1 - 1) // BLOCK + 1):
        blk = gen_block(b, n, f, centres, dev)
        lo, hi = max(b * BLOCK, r0), min(b * BLOCK + blk.shape[0], r1)
        out[lo - r0:hi - r0] = blk[lo - b * BLOCK:hi - b * BLOCK]
        del blk
    return out


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_800_000
    f = int(sys.argv[2]) if len(sys.argv) > 2 else 768
    eps_list = [float(e) for e in (sys.argv[3] if len(sys.argv) > 3 else "10,5,15").split(",")]
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert n % world == 0, "equal row blocks"
    per = n // world
    r0 = rank * per
    centres = torch.randn(NCL, f, generator=torch.Generator(device=dev).manual_seed(SEED), device=dev, dtype=torch.float64)
    t0 = time.time()
    shard = gen_rows(r0, r0 + per, n, f, centres, dev)
    # planted pairs: my rows [0, PLANT) <- rows [PLANT_OFF, PLANT_OFF + PLANT) of the next rank, perturbed
    src0 = ((rank + 1) % world) * per + PLANT_OFF
    src = gen_rows(src0, src0 + PLANT, n, f, centres, dev)
    g = torch.Generator(device=dev).manual_seed(SEED * 7 + rank)
    shard[:PLANT] = src * (1.0 + 1e-4 * torch.randn(PLANT, f, generator=g, device=dev, dtype=torch.float64))
    del src
    torch.cuda.synchronize()
    gen_s = time.time() - t0
    out = {"n": n, "f": f, "world": world, "rows_per_rank": per, "k": 25, "generate_s": gen_s, "runs": []}
    for eps in eps_list:
        gp = {"eps": eps, "k": 25, "topk": 10, "p": 2.0, "sigma": None}
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.time()
        if world > 1:
            aspace, gl = ArrowSpaceBuilder.build_item_graph_sharded(gp, shard, n, r0, device=local)
        else:
            aspace, gl = ArrowSpaceBuilder.build_item_graph(gp, shard, device=local)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        wall = time.time() - t0
        st = {k: api.stat(k, local) for k in ("knn_stage1_ms", "knn_stage2_ms", "knn_slow_rows", "knn_rows_two_term", "knn_rescored_per_row")}
        run = {"eps": eps, "wall_s": wall, "nnz": int(gl.nnz), "nnodes": int(gl.nnodes), **st,
               "stage1_executed_pflops": 2.0 * per * n * 16.0 * ((f + 3 + 15) // 16) / (st["knn_stage1_ms"] * 1e-3) / 1e15,
               "gpu_mem_gb": torch.cuda.mem_get_info(dev)[0] / 1e9}
        if rank == 0:
            t1 = time.time()
            indptr, indices, data = gl.csr()
            ok_fwd = ok_bwd = 0
            for j in range(PLANT):
                row = indices[indptr[j]:indptr[j + 1]]
                ok_fwd += int((src0 + j) in row)
                s_row = indices[indptr[src0 + j]:indptr[src0 + j + 1]]
                ok_bwd += int(j in s_row)
            deg = np.diff(indptr) - 1
            run.update({"planted_found_forward": ok_fwd, "planted_found_backward": ok_bwd, "planted": PLANT,
                        "degree_mean": float(deg.mean()), "degree_max": int(deg.max()), "isolated_nodes": int((deg == 0).sum()),
                        "row_sums_max_abs": float(np.abs(np.add.reduceat(data, indptr[:-1])).max()),
                        "check_s": time.time() - t1})
            print("C5", json.dumps(run), flush=True)
        out["runs"].append(run)
        del aspace, gl
        if world > 1:
            dist.barrier()
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", "c5_sweep_n%d_w%d.json" % (n, world)), "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
