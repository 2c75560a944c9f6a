"""Property tests of the path's invariants, on the oracle (CPU, hypothesis-generated inputs) and
on the CUDA path (-m gpu, the same properties on seeded inputs).  None of them needs a reference value: they hold for any
input, so they also hold at BASELINE.json's full sizes (tests/test_gpu_parity.py uses some of them there)."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st
from hypothesis.extra import numpy as hnp

FAST = settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])

shapes = st.tuples(st.integers(4, 40), st.integers(3, 12))
seeds = st.integers(0, 2 ** 31 - 1)
params = st.fixed_dictionaries({"eps": st.floats(0.05, 1.0), "k": st.integers(1, 6), "topk": st.integers(1, 5),
                                "p": st.sampled_from([1.0, 2.0, 3.0]), "sigma": st.one_of(st.none(), st.floats(0.05, 1.0))})


def _items(shape, seed):
    rng = np.random.default_rng(seed)
    return np.abs(rng.normal(size=shape)) + 0.05          # positive entries: the median tau stays off its floor


def _dense(graph):
    ip, ix, dt = graph.csr()
    m = graph.nnodes
    L = np.zeros((m, m))
    for a in range(m):
        L[a, ix[ip[a]:ip[a + 1]]] = dt[ip[a]:ip[a + 1]]
    return L


def _check_graph(L, gp, sw):
    m = L.shape[0]
    W = -(L - np.diag(np.diag(L)))
    assert (W >= 0).all(), "off-diagonal entries of a Laplacian are <= 0"
    if sw.get("laplacian", "combinatorial") == "combinatorial":
        np.testing.assert_allclose(L.sum(axis=1), 0.0, atol=1e-12)           # rows sum to zero: L = D - W
        if sw.get("symmetrise", "max") != "none":
            assert np.array_equal(L, L.T), "symmetrised graph"
        if sw.get("symmetrise", "max") == "min":                                # mutual edges only: the k cap holds per row
            assert ((W > 0).sum(axis=1) <= gp["k"]).all()
        assert ((W > 0).sum(axis=1) <= 2 * m).all()
    if sw.get("laplacian") == "rw":
        deg = (W > 0).sum(axis=1)
        np.testing.assert_allclose(L.sum(axis=1)[deg > 0], 0.0, atol=1e-12)  # I - D^-1 W


def _check_search(space, graph, x, gp, run):
    n = x.shape[0]
    q = x[n // 2] * 1.03 + 0.001
    hits1 = run(q, 1.0)
    cos = x @ q / (np.linalg.norm(x, axis=1) * np.linalg.norm(q))
    order = sorted(range(n), key=lambda i: (-cos[i], i))[:len(hits1)]
    gaps = np.diff(np.sort(cos)[::-1][:len(hits1) + 1])
    if len(gaps) == 0 or np.abs(gaps).min() > 1e-12:                           # tau = 1: pure cosine order
        assert [i for i, _ in hits1] == order
    assert len(hits1) == min(gp["topk"], n)
    hits = run(q, 0.6)
    sc = [s for _, s in hits]
    assert sc == sorted(sc, reverse=True), "scores descend"
    assert all(0.0 <= s <= 1.0 + 1e-12 for s in sc), "tau cos + (1 - tau) / (1 + |dl|) lies in [0, 1] for non-negative data"
    # scaling the QUERY changes lambda_q (the median tau is not scale invariant, tests/test_0.py) but not the cosine term
    h2 = run(q * 7.0, 1.0)
    np.testing.assert_allclose([s for _, s in h2], [s for _, s in hits1], rtol=1e-12)


@FAST
@given(shapes, seeds, params, st.sampled_from([{}, {"symmetrise": "min"}, {"symmetrise": "none"}, {"laplacian": "rw"},
                                               {"kernel": "gaussian"}, {"k_counts_self": True}]))
def test_oracle_invariants(oracle_mod, shape, seed, gp, sw):
    x = _items(shape, seed)
    s, g = oracle_mod.build(gp, x, **sw)
    _check_graph(_dense(g), gp, sw)
    lam = s.lambdas()
    assert lam.shape == (shape[0],) and np.isfinite(lam).all()
    if sw.get("laplacian", "combinatorial") == "combinatorial" and sw.get("symmetrise", "max") != "none":
        assert ((lam >= 0.0) & (lam < 1.0)).all(), "E >= 0 for a PSD Laplacian, so E / (E + tau) lies in [0, 1)"
    if g.nnz > g.nnodes and not np.any(lam == 0.0):
        try:
            _check_search(s, g, x, gp, lambda q, tau: s.search(q, g, tau))
        except oracle_mod.OracleError as e:                                    # lambda_q == 0 is the reference's panic, not a bug
            assert e.code == 3


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(12))
def test_gpu_invariants(oracle_mod, seed):
    """The same invariants on the CUDA path, plus agreement with the oracle on the drawn case."""
    from arrowspace import ArrowSpaceBuilder
    from pyarrowspace_b200._lib import LibraryError
    rng = np.random.default_rng(1000 + seed)
    shape = (int(rng.integers(4, 400)), int(rng.integers(3, 70)))
    gp = {"eps": float(rng.uniform(0.05, 1.0)), "k": int(rng.integers(1, 9)), "topk": int(rng.integers(1, 8)),
          "p": float(rng.choice([1.0, 2.0, 3.0])), "sigma": None if seed % 3 == 0 else float(rng.uniform(0.05, 1.0))}
    sw = [{}, {"symmetrise": "min"}, {"symmetrise": "none"}, {"laplacian": "rw"}, {"kernel": "gaussian"}, {"k_counts_self": True}][seed % 6]
    x = _items(shape, seed)
    aspace, gl = ArrowSpaceBuilder.build(gp, x, **sw)
    s, g = oracle_mod.build(gp, x, **sw)
    assert all(np.array_equal(a, b) for a, b in zip(gl.csr()[:2], g.csr()[:2]))
    _check_graph(_dense(gl), gp, sw)
    np.testing.assert_allclose(aspace.lambdas(), s.lambdas(), rtol=1e-9, atol=0)
    if gl.nnz > gl.nnodes and not np.any(s.lambdas() == 0.0):
        try:
            _check_search(aspace, gl, x, gp, lambda q, tau: aspace.search(q, gl, tau))
        except BaseException as e:                                             # the reference's lambda_q == 0 panic
            assert "lambdas are zero" in str(e)


@FAST
@given(st.integers(1, 32), st.integers(0, 400), seeds)
def test_shared_row_threshold_is_a_lower_bound_of_the_kth_best(topk, n_scores, seed):
    """The candidate kernel's running threshold (csrc/search_tc.cu epilogue): the four threads of a query row each keep the
    best rr = ceil(topk / 4) scores of THEIR quarter of the columns and the row's threshold is the minimum over the quarters of the
    rr-th best (minus infinity while a quarter has seen fewer).  Whatever the split, it never exceeds the row's topk-th best
    score, so no true top-k item is ever filtered."""
    rng = np.random.default_rng(seed)
    scores = np.round(rng.normal(size=n_scores), 1)                  # rounding makes ties
    quarter = rng.integers(0, 4, size=n_scores)                      # any assignment of columns to the four threads
    rr = (topk + 3) // 4
    published = []
    for t in range(4):
        mine = np.sort(scores[quarter == t])[::-1]
        published.append(mine[rr - 1] if len(mine) >= rr else -np.inf)
    threshold = min(published)
    kth = np.sort(scores)[::-1][topk - 1] if n_scores >= topk else -np.inf
    assert threshold <= kth
    if threshold > -np.inf:
        assert (scores >= threshold).sum() >= topk


@FAST
@given(shapes, seeds, params, st.integers(0, 50), st.sampled_from([0.0, 0.4, 0.62, 1.0]))
def test_hybrid_search_invariants(oracle_mod, shape, seed, gp, pool, tau):
    """Properties of the restated hybrid search (orc_search_hybrid) that need no reference value: every hit comes from the
    cosine shortlist; tau = 1 returns the head of the shortlist; a shortlist of every item is the plain search; a longer
    shortlist never lowers the k-th score; the score of a hit is the plain search's score of that item."""
    x = _items(shape, seed)
    n = shape[0]
    s, g = oracle_mod.build(gp, x)
    if np.any(s.lambdas() == 0.0) or g.nnz <= g.nnodes:
        return
    q = (x[n // 3] * 1.07 + 0.002).reshape(1, -1)
    topk = gp["topk"]
    m = min(n, max(topk, pool if pool > 0 else min(2 * topk, 31)))
    cos = s.scores(q[0], 0.5, 1.0)                                             # tau = 1: the cosine of every item
    short = sorted(range(n), key=lambda i: (-cos[i], i))[:m]
    idx, sc, lq = s.search_hybrid_batch(q, g, tau, pool)
    hits = [int(i) for i in idx[0] if i >= 0]
    assert len(hits) == min(topk, n) and set(hits) <= set(short)
    full = s.scores(q[0], lq[0], tau)
    assert [full[i] for i in hits] == [v for v in sc[0][:len(hits)]]            # same expression as the plain search
    assert list(sc[0][:len(hits)]) == sorted(sc[0][:len(hits)], reverse=True)
    one, _, _ = s.search_hybrid_batch(q, g, 1.0, pool)
    assert [int(i) for i in one[0] if i >= 0] == short[:min(topk, n)]
    if lq[0] != 0.0:
        whole, wsc, _ = s.search_hybrid_batch(q, g, tau, n)
        plain, psc, _ = s.search_batch(q, g, tau)
        assert np.array_equal(whole, plain) and np.array_equal(wsc[plain >= 0], psc[plain >= 0])
        assert wsc[0][len(hits) - 1] >= sc[0][len(hits) - 1]                     # more candidates cannot lower the k-th score


@FAST
@given(st.integers(2, 60), st.integers(2, 10), st.integers(1, 70), seeds, st.sampled_from([0.0, 0.5, 0.62, 1.0]))
def test_oracle_topk_selection_with_ties(oracle_mod, n, f, topk, seed, tau):
    """The arbiter's own top-k (a bounded heap + final sort, oracle.c orc_search) against a full sort of orc_scores on inputs
    full of exact ties (duplicated rows, values on a coarse grid): indices by (score desc, index asc), padded with -1 / NaN."""
    rng = np.random.default_rng(seed)
    x = np.round(np.abs(rng.normal(size=(n, f))), 1) + 0.1
    x[rng.integers(0, n, n // 2)] = x[rng.integers(0, n, n // 2)]                # duplicated rows: tied scores
    gp = {"eps": 0.8, "k": 3, "topk": topk, "p": 2.0, "sigma": 0.4}
    s, g = oracle_mod.build(gp, x)
    q = np.stack([x[rng.integers(0, n)] * 1.5, np.round(np.abs(rng.normal(size=f)), 1) + 0.1])
    try:
        idx, sc, lq = s.search_batch(q, g, tau)
    except oracle_mod.OracleError as e:
        assert e.code == 3                                                      # lambda_q == 0: the reference's panic
        return
    for qi in range(2):
        full = s.scores(q[qi], lq[qi], tau)
        order = np.lexsort((np.arange(n), -full))[:min(topk, n)]
        assert list(idx[qi][:len(order)]) == list(order) and (idx[qi][len(order):] == -1).all()
        assert np.array_equal(sc[qi][:len(order)], full[order]) and np.isnan(sc[qi][len(order):]).all()
