"""world_size-2 (and 4) gloo runs of the multi-GPU host logic on CPU.

`pyarrowspace_b200.distributed.sharded_build` is the orchestration the GPU ranks run (segment
all-gather, rank-ordered continuation of exact column sums, per-shard lambdas).  Here it is driven with
an oracle-backed engine (test infrastructure, lives in this file) over gloo, and must reproduce the
single-process oracle: same edges, same lambdas, and left-to-right sums that continue across ranks
bit-exactly."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class OracleEngine:
    """Stands in for CudaEngine: same methods, numpy/oracle arithmetic, CPU tensors for the collectives."""

    def __init__(self, force_need):
        import torch
        self.torch = torch
        self.force_need = force_need       # ask for these pairs once, to exercise the rank-ordered chain
        self.asked = False
        self.exact = None

    def space_create(self, shard, n_total, world, rank):
        from pyarrowspace_b200 import shard_rows
        r0, r1 = shard_rows(n_total, world, rank)
        assert shard.shape[0] == r1 - r0
        return {"x": np.ascontiguousarray(shard), "world": world, "rank": rank, "n_total": n_total}

    def gram_partials(self, space, f):
        from pyarrowspace_b200 import shard_rows
        x, world, rank, n = space["x"], space["world"], space["rank"], space["n_total"]
        segs = np.zeros((8, f, f))
        r0, _ = shard_rows(n, world, rank)
        for e in range(rank * 8 // world, (rank + 1) * 8 // world):
            e0, e1 = shard_rows(n, 8, e)
            blk = x[e0 - r0:e1 - r0]
            segs[e] = blk.T @ blk if len(blk) else 0.0
        return self.torch.from_numpy(segs)

    def graph_from_gram(self, segs, f, n_total, cgp, sw, pairs, sums):
        if self.force_need is not None and not self.asked:
            self.asked = True
            return None, np.asarray(self.force_need, dtype=np.int32)
        self.exact = (pairs, sums)
        return {"gram": segs.numpy().sum(0)}, np.empty((0, 2), dtype=np.int32)

    def exact_pairs(self, space, pairs, sums):
        x = space["x"]
        for i, (a, b) in enumerate(pairs):
            s = sums[i].copy()
            for r in range(x.shape[0]):
                s[0] = s[0] + x[r, a] * x[r, b]
                s[1] = s[1] + x[r, a] * x[r, a]
                s[2] = s[2] + x[r, b] * x[r, b]
            sums[i] = s
        return sums

    def compute_lambdas(self, space, graph):
        space["gram"] = graph["gram"]

    def to_comm(self, arr):
        return self.torch.from_numpy(np.ascontiguousarray(arr))

    def from_comm(self, t):
        return t.numpy()


def _worker(rank, world, port, n, f, q):
    import torch.distributed as dist
    try:
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
        from pyarrowspace_b200 import _lib, shard_rows, synth
        from pyarrowspace_b200.distributed import sharded_build
        r0, r1 = shard_rows(n, world, rank)
        shard = synth.make_items(n, f, 3, n_clusters=4, rows=(r0, r1))
        eng = OracleEngine(force_need=[(0, 1), (2, 5)])
        cgp = _lib.make_params(0.5, 3, 4, 2.0, None)
        space, graph = sharded_build(eng, shard, n, cgp, _lib.make_switches(), None)
        q.put((rank, r0, r1, graph["gram"], eng.exact[0], eng.exact[1]))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:                                              # pragma: no cover
        q.put((rank, "error", repr(e)))
        raise


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_build_over_gloo(world, oracle_mod):
    import torch.multiprocessing as mp
    from pyarrowspace_b200 import synth
    n, f = 1000, 12
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, f, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda t: t[0])
    assert all(r[1] != "error" for r in res), res
    x = synth.make_items(n, f, 3, n_clusters=4)
    # shards tile the matrix and every rank ends with the same, complete Gram
    assert res[0][1] == 0 and res[-1][2] == n and all(res[i][2] == res[i + 1][1] for i in range(world - 1))
    ref = oracle_mod.gram_columns(x)
    for r in res:
        assert np.array_equal(r[3], res[0][3])
        np.testing.assert_allclose(r[3], ref, rtol=1e-12)
        # the exact sums were continued rank after rank: identical to ONE left-to-right pass over all rows
        pairs, sums = r[4], r[5]
        assert [tuple(p) for p in pairs] == [(0, 1), (2, 5)]
        assert sums[0][0] == ref[0, 1] and sums[0][1] == ref[0, 0] and sums[0][2] == ref[1, 1]
        assert sums[1][0] == ref[2, 5] and sums[1][1] == ref[2, 2] and sums[1][2] == ref[5, 5]


def test_synthetic_shards_are_consistent():
    """Rows generated per shard equal the rows of the whole matrix (ranks generate their own shard)."""
    from pyarrowspace_b200 import shard_rows, synth
    n, f = 200_000, 8
    full = synth.make_items(n, f, 44)
    for world in (2, 8):
        for r in (0, world - 1):
            r0, r1 = shard_rows(n, world, r)
            assert np.array_equal(synth.make_items(n, f, 44, rows=(r0, r1)), full[r0:r1])


# ----------------------------------------------------------------------------- item graph across ranks (steps 4-5)

class NumpyKnnEngine:
    """Stands in for CudaEngine in `sharded_item_graph`: brute-force neighbour lists in numpy (the oracle's distance and
    selection rule) and a numpy symmetrise + Laplacian."""

    def __init__(self):
        import torch
        self.torch = torch

    def full_space(self, x_full):
        return {"x": x_full.numpy().copy()}

    def knn_rows(self, space, cgp, r0, r1, sw=None):
        x = space["x"]
        n = x.shape[0]
        kk = min(int(cgp.k), n - 1)
        nrm = np.sqrt((x * x).sum(1))
        idx = np.full((r1 - r0, kk), -1, dtype=np.int32)
        dst = np.zeros((r1 - r0, kk))
        cnt = np.zeros(r1 - r0, dtype=np.int32)
        for i in range(r0, r1):
            c = (x @ x[i]) / (nrm * nrm[i])
            d = 1.0 - np.maximum(c, 0.0)
            cand = [(d[j], j) for j in range(n) if j != i and d[j] <= cgp.eps]
            cand.sort()
            for e, (dj, j) in enumerate(cand[:kk]):
                idx[i - r0, e], dst[i - r0, e] = j, dj
            cnt[i - r0] = min(len(cand), kk)
        t = self.torch
        return t.from_numpy(idx), t.from_numpy(dst), t.from_numpy(cnt)

    def graph_from_knn(self, n, idx, dist, cnt, cgp, sw):
        idx, dist, cnt = idx.numpy(), dist.numpy(), cnt.numpy()
        sigma = cgp.sigma if cgp.has_sigma else cgp.eps * 0.5
        w = np.zeros((n, n))
        for i in range(n):
            for e in range(cnt[i]):
                j = idx[i, e]
                wij = 1.0 / (1.0 + (dist[i, e] / sigma) ** cgp.p)
                w[i, j] = max(w[i, j], wij)
                w[j, i] = max(w[j, i], wij)
        return {"edges": sorted((i, j) for i in range(n) for j in range(i + 1, n) if w[i, j] > 0)}


def _item_graph_worker(rank, world, port, n, f, q):
    import torch
    import torch.distributed as dist
    try:
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
        from pyarrowspace_b200 import _lib, synth
        from pyarrowspace_b200.distributed import sharded_item_graph
        # deliberately ragged contiguous blocks
        cuts = [0] + [int(n * (r + 1) / world) + (3 if r % 2 == 0 and r + 1 < world else 0) for r in range(world)]
        cuts[-1] = n
        r0, r1 = cuts[rank], cuts[rank + 1]
        shard = torch.from_numpy(synth.make_items(n, f, 5, n_clusters=4)[r0:r1].copy())
        cgp = _lib.make_params(0.4, 4, 3, 2.0, None)
        space, graph = sharded_item_graph(NumpyKnnEngine(), shard, n, r0, cgp, _lib.make_switches(), None)
        q.put((rank, graph["edges"], space["x"].shape))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:                                              # pragma: no cover
        q.put((rank, "error", repr(e)))
        raise


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_item_graph_over_gloo(world, oracle_mod):
    """Halo all-gather of ragged item shards + all-gather of the per-rank neighbour lists reproduce the single-process
    oracle's item graph on every rank."""
    import torch.multiprocessing as mp
    from pyarrowspace_b200 import synth
    n, f = 240, 10
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_item_graph_worker, args=(r, world, port, n, f, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] != "error" for r in res), res
    x = synth.make_items(n, f, 5, n_clusters=4)
    s, g = oracle_mod.build({"eps": 0.4, "k": 4, "topk": 3, "p": 2.0, "sigma": None}, x, nodes="items")
    want = [tuple(int(v) for v in e) for e in g.edges()]
    for r in res:
        assert r[2] == (n, f)
        assert [tuple(e) for e in r[1]] == want


# ----------------------------------------------------------------------------- search grid (R item shards x C query slots)

def test_grid_arithmetic():
    from pyarrowspace_b200.distributed import auto_item_shards, grid_layout, query_slice
    assert grid_layout(8, 5, 2) == (2, 4, 1, 1)            # ranks 4..7 hold item shard 1; rank 5 answers query slot 1
    assert grid_layout(8, 5, 8) == (8, 1, 5, 0) and grid_layout(8, 5, 1) == (1, 8, 0, 5)
    with pytest.raises(ValueError):
        grid_layout(8, 0, 3)
    # the slices of a batch tile it exactly, in slot order, whatever the remainder
    for nq, slots in ((65536, 8), (10, 4), (3, 8), (0, 2), (1000, 3)):
        cuts = [query_slice(nq, slots, c) for c in range(slots)]
        assert cuts[0][0] == 0 and cuts[-1][1] == nq and all(cuts[i][1] == cuts[i + 1][0] for i in range(slots - 1))
        assert all(b - a <= per for a, b, per in cuts) and len({per for _, _, per in cuts}) == 1
    assert auto_item_shards(8, 1_000_000, 384, 170e9) == 1          # C4: replicate
    assert auto_item_shards(8, 8_800_000, 768, 170e9) == 2          # C5: two shards
    assert auto_item_shards(8, 40_000_000, 768, 170e9) == 8
    assert auto_item_shards(2, 40_000_000, 768, 170e9) == 2


class _GridEngine:
    """The regrouping half of CudaEngine on CPU tensors: a space is a dict of numpy arrays."""

    def __init__(self):
        import torch
        self.torch = torch

    def lambdas_norms(self, space, n_local):
        return self.torch.from_numpy(space["lam"].copy()), self.torch.from_numpy(space["nrm"].copy())

    def space_from_gathered(self, x, lam, nrm, n_total, shards, shard):
        return {"x": x.numpy(), "lam": lam.numpy(), "nrm": nrm.numpy(), "shards": shards, "shard": shard}

    def free_space(self, space):
        space["freed"] = True


def _regroup_worker(rank, world, port, n, f, item_shards, q):
    import torch
    import torch.distributed as dist
    try:
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
        from pyarrowspace_b200 import shard_rows
        from pyarrowspace_b200.distributed import regroup
        r0, r1 = shard_rows(n, world, rank)
        x = np.arange(n * f, dtype=np.float64).reshape(n, f)
        space = {"lam": x[r0:r1, 0] * 0.5, "nrm": x[r0:r1, 1] + 1.0}
        new_space, grid = regroup(_GridEngine(), space, torch.from_numpy(x[r0:r1].copy()), n, item_shards, None)
        q.put((rank, grid["R"], grid["C"], grid["r"], grid["c"], new_space.get("x"), new_space["lam"], new_space["nrm"],
               space.get("freed", False)))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:                                              # pragma: no cover
        q.put((rank, "error", repr(e)))
        raise


@pytest.mark.parametrize("world,item_shards", [(4, 1), (4, 2), (2, 2)])
def test_regroup_over_gloo(world, item_shards):
    """distributed.regroup: the C ranks of an item shard end up with the shard's rows, lambdas and norms in row order (the
    all-gathers of the R x C search grid), the build-time space is released; item_shards == world is a no-op."""
    import torch.multiprocessing as mp
    from pyarrowspace_b200 import shard_rows
    n, f = 1000, 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_regroup_worker, args=(r, world, port, n, f, item_shards, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] != "error" for r in res), res
    x = np.arange(n * f, dtype=np.float64).reshape(n, f)
    c_ = world // item_shards
    for rank, R, C_, r, c, gx, lam, nrm, freed in res:
        assert (R, C_, r, c) == (item_shards, c_, rank // c_, rank % c_)
        if c_ == 1:
            assert gx is None and not freed                              # untouched
            continue
        a, b = shard_rows(n, item_shards, r)
        assert freed and np.array_equal(gx, x[a:b]) and np.array_equal(lam, x[a:b, 0] * 0.5) and np.array_equal(nrm, x[a:b, 1] + 1.0)


# ----------------------------------------------------------------------------- pre-graph reduction across ranks

class _ReduceEngine:
    """Stands in for CudaEngine in sharded_reduction: the C ABI's host-side sampler, the oracle's reduction."""

    def __init__(self):
        import torch
        self.torch = torch

    def sampled_rows(self, shard, red, row0):
        import ctypes as C
        from pyarrowspace_b200 import _lib
        rows = np.empty(max(len(shard), 1), dtype=np.int32)
        cnt = C.c_int64(0)
        _lib.check(_lib.load().asp_reduction_sample(C.byref(red), int(row0), len(shard), rows.ctypes.data, C.byref(cnt)))
        return self.torch.from_numpy(np.ascontiguousarray(shard[rows[:cnt.value]]))

    def reduce_rows(self, rows, red, n_total):
        import oracle
        opts = {k: getattr(red, k) for k in ("sample_rate", "seed", "n_clusters", "max_iters", "probes")}
        cent, info = oracle.reduce(rows.numpy(), opts, n_total_for_k=n_total)
        return cent, info

    def feature_graph(self, cspace, cgp, sw):
        return {"centroids": cspace}


def _reduce_worker(rank, world, port, n, f, opts, q):
    import torch.distributed as dist
    try:
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
        from pyarrowspace_b200 import _lib, shard_rows, synth
        from pyarrowspace_b200.distributed import sharded_reduction
        r0, r1 = shard_rows(n, world, rank)
        shard = synth.make_items(n, f, 5, n_clusters=6, rows=(r0, r1))
        red = _lib.make_reduction(opts)
        graph, cent, info = sharded_reduction(_ReduceEngine(), shard, n, r0, _lib.make_params(0.5, 3, 4, 2.0, None),
                                              _lib.make_switches(), red, None)
        q.put((rank, cent, info, red.sample_rate))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:                                              # pragma: no cover
        q.put((rank, "error", repr(e)))
        raise


@pytest.mark.parametrize("world,opts", [(2, {"max_iters": 5}), (4, {"sample_rate": 0.35, "seed": 9, "n_clusters": 11}),
                                        (2, {"sample_rate": 1e-9, "n_clusters": 4, "probes": 64})])
def test_sharded_reduction_over_gloo(world, opts, oracle_mod):
    """Every rank samples its own rows (hash on GLOBAL row numbers), the kept rows are all-gathered in rank order, and the
    replicated reduction on them equals the single-process reduction of the whole matrix bit for bit (the third case keeps
    no row at all: the empty sample falls back to every row)."""
    import torch.multiprocessing as mp
    from pyarrowspace_b200 import synth
    n, f = 1500, 10
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_reduce_worker, args=(r, world, port, n, f, opts, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] is not None and not isinstance(r[1], str) for r in res), res
    x = synth.make_items(n, f, 5, n_clusters=6)
    want, info = oracle_mod.reduce(x, opts)
    for r in res:
        assert np.array_equal(r[1], want)
        assert {k: v for k, v in r[2].items() if k != "two_nn_mean_ratio"} == {k: v for k, v in info.items() if k != "two_nn_mean_ratio"}
        assert r[2]["two_nn_mean_ratio"] == info["two_nn_mean_ratio"] or np.isnan(info["two_nn_mean_ratio"])
        assert r[3] == opts.get("sample_rate", 0.6)        # the caller's options are not modified
