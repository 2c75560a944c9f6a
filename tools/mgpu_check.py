"""torchrun --nproc-per-node N tools/mgpu_check.py : sharded build + search over NCCL against the CPU oracle
and against the single-GPU result (bit identical).  CHECK_PEER=1 also runs the peer-memory merge (csrc/peer.cu)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from pyarrowspace_b200 import shard_rows, synth  # noqa: E402
from pyarrowspace_b200.api import ArrowSpaceBuilder  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for n, f, gp in [(20000, 96, {"eps": 0.5, "k": 6, "topk": 10, "p": 2.0, "sigma": 0.25}),
                     (3001, 384, {"eps": 10.0, "k": 25, "topk": 10, "p": 2.0, "sigma": None})]:
        r0, r1 = shard_rows(n, world, rank)
        shard = synth.make_items(n, f, 9, n_clusters=16, rows=(r0, r1))
        full = synth.make_items(n, f, 9, n_clusters=16)
        q, _ = synth.make_queries(full, 500, 9)
        aspace, gl = ArrowSpaceBuilder.build_sharded(gp, shard, n, device=local)
        idx, sc = aspace.search_batch(q, gl, 0.62)
        lam = aspace.lambdas()
        if os.environ.get("CHECK_PEER"):                      # csrc/peer.cu: P2P exchange + flag-driven merge vs the NCCL route
            os.environ["ASP_PEER_MERGE"] = "1"
            same = True
            for rep in range(3):                              # three calls: both parities and a reused slot
                idx_p, sc_p = aspace.search_batch(q, gl, 0.62)
                same = same and np.array_equal(idx_p, idx) and np.array_equal(sc_p, sc)
            os.environ.pop("ASP_PEER_MERGE")
            flag = torch.tensor([1 if same else 0], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if rank == 0:
                print("world", world, "n", n, "peer-memory merge == NCCL merge on every rank (bitwise):", bool(flag.item()))
            ok = ok and bool(flag.item())
        if rank == 0:
            import oracle
            s, g = oracle.build(gp, full)
            oidx, osc, _ = s.search_batch(q, g, 0.62)
            e = [np.array_equal(gl.edges(), g.edges()), np.array_equal(idx, oidx),
                 bool(np.allclose(sc, osc, rtol=1e-9, atol=0)),
                 bool(np.allclose(lam, s.lambdas()[r0:r1], rtol=1e-9, atol=0))]
            a1, g1 = ArrowSpaceBuilder.build(gp, full, device=local)
            idx1, sc1 = a1.search_batch(q, g1, 0.62)
            e += [all(np.array_equal(a, b) for a, b in zip(gl.csr(), g1.csr())), np.array_equal(idx, idx1),
                  np.array_equal(sc, sc1), np.array_equal(lam, a1.lambdas()[r0:r1])]
            print("world", world, "n", n, "f", f, "edges/idx/score/lambda vs oracle:", e[:4], " vs single GPU (bitwise):", e[4:])
            ok = ok and all(e)
        dist.barrier()
    if os.environ.get("QUICK"):                               # build / search / merge checks only
        if rank == 0:
            print("MGPU_CHECK", "OK" if ok else "FAILED")
        dist.destroy_process_group()
        sys.exit(0 if ok else 1)
    # item graph across the ranks: halo all-gather of the item shards, rows resolved per rank on the tensor cores,
    # all-gather of the neighbour lists; every rank must end with the single-GPU graph (bitwise) == the oracle's edges
    import time
    for n, f, gp in [(24000, 96, {"eps": 0.3, "k": 12, "topk": 3, "p": 2.0, "sigma": None}),
                     (int(os.environ.get("IG_N", 60000)), 384, {"eps": 10.0, "k": 25, "topk": 3, "p": 2.0, "sigma": None})]:
        per = (n + world - 1) // world
        r0, r1 = min(rank * per, n), min((rank + 1) * per, n)
        full = synth.make_items(n, f, 11, n_clusters=16)
        shard = torch.from_numpy(full[r0:r1].copy()).cuda()
        torch.cuda.synchronize(); dist.barrier(); t0 = time.time()
        aspace, gl = ArrowSpaceBuilder.build_item_graph_sharded(gp, shard, n, r0, device=local)
        torch.cuda.synchronize(); dist.barrier(); dt = time.time() - t0
        os.environ["ASP_KNN_STAGE1"] = "tc"
        torch.cuda.synchronize(); t1 = time.time()
        a1, g1 = ArrowSpaceBuilder.build_item_graph(gp, torch.from_numpy(full).cuda(), device=local)
        torch.cuda.synchronize(); d1 = time.time() - t1
        os.environ.pop("ASP_KNN_STAGE1", None)
        same = all(np.array_equal(a, b) for a, b in zip(gl.csr(), g1.csr()))
        flags = torch.tensor([1 if same else 0], device="cuda")
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        if rank == 0:
            e = [bool(flags.item())]
            if n <= 30000:
                import oracle
                s, g = oracle.build(gp, full, nodes="items")
                e.append(np.array_equal(gl.edges(), g.edges()))
            print("world", world, "item graph n", n, "f", f, "sharded %.3f s, single GPU %.3f s" % (dt, d1),
                  "all ranks == single GPU (bitwise):", e[0], "" if len(e) < 2 else "== oracle edges: %s" % e[1])
            ok = ok and all(e)
        del aspace, gl, a1, g1
        dist.barrier()
    if rank == 0:
        print("MGPU_CHECK", "OK" if ok else "FAILED")
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
