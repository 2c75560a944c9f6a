"""search_hybrid on the GPU (src/lib.rs:182-219) against the oracle's restatement (orc_search_hybrid).
Run with -m gpu.  Collected last on purpose: the path was written after the round's GPU budget was spent (its re-ranking
kernels are checked on the CPU, tests/test_hybrid_host.py; its shortlist is the validated search at tau = 1)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def _data(n, f, seed, nq, dup=0):
    rng = np.random.default_rng(seed)
    cent = rng.normal(size=(8, f))
    x = np.abs(cent[rng.integers(0, 8, n)] + 0.2 * rng.normal(size=(n, f))) + 0.05
    if dup:
        x[n // 2:n // 2 + dup] = x[:dup]
    q = x[rng.integers(0, n, nq)] * 1.02 + 0.01 * rng.normal(size=(nq, f))
    return x, q


def _assert_hits_equal(idx, sc, oidx, osc):
    assert np.array_equal(idx, oidx), "hybrid index lists differ at rows %s" % np.where((idx != oidx).any(axis=1))[0][:10]
    m = oidx >= 0
    np.testing.assert_allclose(sc[m], osc[m], rtol=RTOL, atol=0)
    assert np.isnan(sc[~m]).all()


@pytest.mark.parametrize("n,f,topk,pool,nq,dup", [
    (3000, 96, 10, None, 300, 0),       # shortlist 20: tensor-core candidates (batch >= 256)
    (3000, 96, 10, None, 7, 0),         # the same through the FP64 candidate pass (small batch)
    (5000, 64, 3, 24, 257, 12),         # 32-entry lists of the tensor-core pass, duplicated rows
    (2000, 130, 20, 40, 64, 0),         # shortlist 40: the batched exact scan
    (10, 16, 6, None, 9, 0),            # shortlist (12) cut to n
    (1500, 48, 4, 5000, 33, 0),         # pool >= n: search without the assertion
])
def test_hybrid_search_gpu_equals_oracle(oracle_mod, n, f, topk, pool, nq, dup):
    from arrowspace import ArrowSpaceBuilder
    x, q = _data(n, f, n + f + topk, nq, dup)
    gp = {"eps": 0.6, "k": 5, "topk": topk, "p": 2.0, "sigma": 0.3}
    aspace, gl = ArrowSpaceBuilder.build(gp, x)
    s, g = oracle_mod.build(gp, x)
    shortlist = min(n, max(topk, pool or min(2 * topk, 31)))
    # below n the shortlist pass always runs at tau = 1 and tau only enters the re-ranking kernel: any tau; a shortlist of
    # every item is the plain search at the caller's tau
    for tau in ((0.62, 1.0, 0.0) if shortlist < n else (0.62, 1.0)):
        idx, sc, lq = aspace.search_hybrid_batch(q, gl, tau, pool=pool, want_lambda=True)
        oidx, osc, olq = s.search_hybrid_batch(q, g, tau, pool or 0)
        np.testing.assert_allclose(lq, olq, rtol=RTOL, atol=0)
        _assert_hits_equal(idx, sc, oidx, osc)
    if pool is not None and pool >= n:
        i2, s2 = aspace.search_batch(q, gl, 0.62)
        i1, s1 = aspace.search_hybrid_batch(q, gl, 0.62, pool=pool)
        assert np.array_equal(i1, i2)
        np.testing.assert_allclose(s1, s2, rtol=RTOL, atol=0)


def test_search_hybrid_surface_and_errors(oracle_mod):
    """The reference's surface (src/lib.rs:182-219): same argument checks as search, list of (index, score), and no
    lambda_q != 0 assertion."""
    from arrowspace import ArrowSpaceBuilder, PanicException
    x = np.abs(np.random.default_rng(2).normal(size=(50, 6))) + 0.1
    gp = {"eps": 1.0, "k": 3, "topk": 3, "p": 2.0}
    aspace, gl = ArrowSpaceBuilder.build(gp, x)
    s, g = oracle_mod.build(gp, x)
    hits = aspace.search_hybrid(x[7].copy(), gl, 0.8)
    want = s.search_hybrid(x[7], g, 0.8)
    assert [i for i, _ in hits] == [i for i, _ in want]
    np.testing.assert_allclose([v for _, v in hits], [v for _, v in want], rtol=RTOL, atol=0)
    with pytest.raises(ValueError, match="query length 5 must match nfeatures 6"):
        aspace.search_hybrid(np.ones(5), gl, 0.5)
    with pytest.raises(ValueError, match="not contiguous"):
        aspace.search_hybrid(np.ones(12)[::2], gl, 0.5)
    with pytest.raises(TypeError):
        aspace.search_hybrid(np.ones(6, dtype=np.float32), gl, 0.5)
    a0, g0 = ArrowSpaceBuilder.build({"eps": 1.0, "k": 0, "topk": 2, "p": 2.0}, x)        # no edges: every lambda is 0
    with pytest.raises(PanicException, match="The lambdas are zero"):
        a0.search(x[0].copy(), g0, 0.5)
    hits = a0.search_hybrid(x[0].copy(), g0, 0.5)                                         # ... and search_hybrid answers
    assert hits[0][0] == 0 and len(hits) == 2


def test_golden_hybrid_search_gpu(golden):
    """The committed fixture (tests/golden/make_kat.py, case hybridE) through the CUDA path: no oracle call at run time."""
    from arrowspace import ArrowSpaceBuilder
    from pyarrowspace_b200 import synth
    x = synth.make_items(600, 48, 5, scale=100.0, n_clusters=16)
    q, _ = synth.make_queries(x, 16, 5)
    aspace, gl = ArrowSpaceBuilder.build({"eps": 0.6, "k": 5, "topk": 10, "p": 2.0, "sigma": 0.3}, x)
    for tag, pool in (("hybridE", None), ("hybridE_pool13", 13)):
        idx, sc, lq = aspace.search_hybrid_batch(q, gl, 0.3, pool=pool, want_lambda=True)
        assert np.array_equal(idx, golden[tag + "_idx"])
        np.testing.assert_allclose(sc, golden[tag + "_score"], rtol=RTOL, atol=0)
        np.testing.assert_allclose(lq, golden[tag + "_lambda_q"], rtol=RTOL, atol=0)
