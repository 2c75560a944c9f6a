"""The CPU oracle against the reference's own known-answer material and
against its independent numpy mirror.  CPU only."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_np  # noqa: E402

from pyarrowspace_b200 import synth  # noqa: E402


def test_readme_kat_bit_exact(oracle_mod, kat):
    """README.md:69 -- the three published doubles, bit for bit."""
    r = kat["readme"]
    s, g = oracle_mod.build(r["graph_params"], np.array(r["items"]))
    hits = s.search(np.array(r["query"]), g, r["tau"])
    assert [(i, sc) for i, sc in hits] == [(i, sc) for i, sc in r["hits"]]
    assert g.nnodes == 3 and s.nitems == 3 and s.nfeatures == 3


@pytest.mark.parametrize("tau", ["1.0", "0.6", "0.55"])
def test_test0_orderings(oracle_mod, kat, tau):
    """tests/test_0.py:29-32,49-52,58-61 -- top-3 index order."""
    t = kat["test_0"]
    items = np.array(t["items"])
    s, g = oracle_mod.build(t["graph_params"], items)
    hits = s.search(items[t["query_item"]] * t["query_scale"], g, float(tau))
    assert len(hits) == 3
    assert [i for i, _ in hits] == t["expected_top3"][tau]


def test_test0_tau09_first_two_and_known_deviation(oracle_mod, kat):
    """tests/test_0.py:39-42.  Places 1-2 match; place 3 is the documented deviation of the default
    (unpinned) graph spec: item 3 instead of item 0, margin 8.7e-4."""
    t = kat["test_0"]
    items = np.array(t["items"])
    s, g = oracle_mod.build(t["graph_params"], items)
    hits = s.search(items[t["query_item"]] * t["query_scale"], g, 0.9)
    idx = [i for i, _ in hits]
    assert idx[:2] == t["expected_top3"]["0.9"][:2]
    assert idx[2] == t["known_deviation"]["oracle"]
    sc = s.scores(items[2] * 1.05, g.taumode(items[2] * 1.05)[2], 0.9)
    assert 0 < sc[3] - sc[0] < 2e-3


@pytest.mark.xfail(strict=True, reason="known deviation of the unpinned default graph spec (SURVEY.md Appendix A)")
def test_test0_tau09_third_place_reference(oracle_mod, kat):
    t = kat["test_0"]
    items = np.array(t["items"])
    s, g = oracle_mod.build(t["graph_params"], items)
    hits = s.search(items[2] * 1.05, g, 0.9)
    assert hits[2][0] == 0


@pytest.mark.parametrize("tau", ["1.0", "0.9", "0.6", "0.55"])
def test_test0_all_twelve_indices_under_profile_kat12(oracle_mod, kat, tau):
    """The passing sibling of the xfail above: the named switch set `kat12` (symmetrise none, laplacian sym, k counts
    self, topk prunes -- found by tools/fit_switches.py) reproduces every index tests/test_0.py:29-61 asserts."""
    t = kat["test_0"]
    items = np.array(t["items"])
    s, g = oracle_mod.build(t["graph_params"], items, profile="kat12")
    hits = s.search(items[t["query_item"]] * t["query_scale"], g, float(tau))
    assert [i for i, _ in hits] == t["expected_top3"][tau]


def test_readme_kat_bit_exact_under_profile_kat12(oracle_mod, kat):
    """README.md:69 is tau = 1 (pure cosine): bit exact whatever the graph switches are."""
    r = kat["readme"]
    s, g = oracle_mod.build(r["graph_params"], np.array(r["items"]), profile="kat12")
    hits = s.search(np.array(r["query"]), g, r["tau"])
    assert [(i, sc) for i, sc in hits] == [(i, sc) for i, sc in r["hits"]]


def test_appendix_a_self_check(oracle_mod, kat):
    """SURVEY.md Appendix A expected intermediate values on test_0."""
    t = kat["test_0"]
    items = np.array(t["items"])
    s, g = oracle_mod.build(t["graph_params"], items)
    e, tau, lam = g.taumode(items)
    np.testing.assert_allclose(e, [0.545592191302216, 0.4954836835763489, 0.502006439385439,
                                   0.5279264460367588, 0.5449817322807591], rtol=1e-12)
    np.testing.assert_allclose(tau, [0.38, 0.38, 0.365, 0.39, 0.375], rtol=1e-15)
    np.testing.assert_allclose(s.lambdas(), lam, rtol=0, atol=0)
    assert 2 * len(g.edges()) == 178          # directed non-zeros of W
    _, tq, lq = g.taumode(items[2] * 1.05)
    assert abs(tq - 0.38325) < 1e-12 and abs(lq - 0.5670745978802942) < 1e-12


def test_golden_vectors(oracle_mod, kat, golden):
    """The committed fixtures were produced by this oracle: guard against silent drift."""
    r = kat["readme"]
    s, g = oracle_mod.build(r["graph_params"], np.array(r["items"]))
    ip, ix, dt = g.csr()
    assert (ip == golden["readme_indptr"]).all() and (ix == golden["readme_indices"]).all()
    assert (dt == golden["readme_data"]).all() and (s.lambdas() == golden["readme_lambdas"]).all()
    x = synth.make_items(600, 48, 5, scale=100.0, n_clusters=16)
    q, _ = synth.make_queries(x, 16, 5)
    s, g = oracle_mod.build({"eps": 0.6, "k": 5, "topk": 10, "p": 2.0, "sigma": 0.3}, x)
    idx, sc, lq = s.search_batch(q, g, 0.62)
    assert (idx == golden["synthA_idx"]).all() and (sc == golden["synthA_score"]).all()
    assert (s.lambdas() == golden["synthA_lambdas"]).all() and (lq == golden["synthA_lambda_q"]).all()
    x = synth.make_items(2500, 48, 7, scale=100.0, n_clusters=12)                 # item graph (nodes = items)
    s, g = oracle_mod.build({"eps": 0.5, "k": 8, "topk": 3, "p": 2.0, "sigma": 0.2}, x, nodes="items")
    ip, ix, dt = g.csr()
    assert (ip == golden["itemsC_indptr"]).all() and (ix == golden["itemsC_indices"]).all() and (dt == golden["itemsC_data"]).all()


def test_golden_hybrid_search(oracle_mod, golden):
    """Drift guard of the hybrid search (tests/golden/make_kat.py, case hybridE: the synthA inputs, tau 0.3)."""
    x = synth.make_items(600, 48, 5, scale=100.0, n_clusters=16)
    q, _ = synth.make_queries(x, 16, 5)
    s, g = oracle_mod.build({"eps": 0.6, "k": 5, "topk": 10, "p": 2.0, "sigma": 0.3}, x)
    plain, _, _ = s.search_batch(q, g, 0.3)
    for tag, pool in (("hybridE", 0), ("hybridE_pool13", 13)):
        idx, sc, lq = s.search_hybrid_batch(q, g, 0.3, pool)
        assert (idx == golden[tag + "_idx"]).all() and (sc == golden[tag + "_score"]).all() and (lq == golden[tag + "_lambda_q"]).all()
        assert (idx != plain).any()                     # the shortlist matters on this case: not the plain search again


@pytest.mark.parametrize("nodes", ["feature_columns", "items"])
@pytest.mark.parametrize("kernel", ["inv_power", "gaussian"])
def test_c_oracle_equals_numpy_mirror(oracle_mod, nodes, kernel):
    rng = np.random.default_rng(3)
    x = np.abs(rng.normal(size=(37, 19))) + 0.05
    gp = {"eps": 0.45, "k": 4, "topk": 6, "p": 2.0, "sigma": None}
    s, g = oracle_mod.build(gp, x, nodes=nodes, kernel=kernel)
    m = oracle_np.build(x, gp["eps"], gp["k"], gp["topk"], gp["p"], None, nodes=nodes, kernel=kernel)
    import scipy.sparse as sp
    ip, ix, dt = g.csr()
    L = sp.csr_matrix((dt, ix, ip), shape=(g.nnodes, g.nnodes)).toarray()
    assert np.array_equal(L, m["L"])
    assert np.array_equal(g.edges(), np.array(m["edges"]).reshape(-1, 2))
    if nodes == "feature_columns":
        assert np.array_equal(s.lambdas(), m["lambdas"])
        q = x[5] * 0.9 + 0.01
        lq = oracle_np.taumode_lambda(q, m["L"])
        ref = oracle_np.search(x, m["lambdas"], q, lq, gp["topk"], 0.7)
        got = s.search(q, g, 0.7)
        assert got == [(i, float(v)) for i, v in ref]


@pytest.mark.parametrize("tau_mode", ["median", "median_abs", "mean", "fixed"])
@pytest.mark.parametrize("lambda_form", ["bounded", "synthetic"])
def test_switches_match_mirror(oracle_mod, tau_mode, lambda_form):
    rng = np.random.default_rng(11)
    x = rng.normal(size=(20, 12)) + 0.3
    gp = {"eps": 1.0, "k": 3, "topk": 4, "p": 2.0, "sigma": 0.5}
    s, g = oracle_mod.build(gp, x, tau_mode=tau_mode, lambda_form=lambda_form, tau_fixed=0.2)
    m = oracle_np.build(x, 1.0, 3, 4, 2.0, 0.5, tau_mode=tau_mode, lambda_form=lambda_form, tau_fixed=0.2)
    np.testing.assert_allclose(s.lambdas(), m["lambdas"], rtol=1e-13, atol=0)


GRAPH_SWITCH_CASES = [
    dict(symmetrise="avg"), dict(symmetrise="min"), dict(symmetrise="none"),
    dict(laplacian="sym"), dict(laplacian="rw"), dict(symmetrise="none", laplacian="sym"), dict(symmetrise="none", laplacian="rw"),
    dict(k_counts_self=True), dict(topk_prunes=True), dict(k_counts_self=True, topk_prunes=True),
    dict(distance="l2"), dict(distance="l2sq"), dict(profile="kat12"),
    dict(symmetrise="avg", laplacian="sym", lambda_form="synthetic"), dict(symmetrise="none", lambda_form="synthetic"),
]


def test_golden_reduction(oracle_mod, golden):
    """Drift guard of the pre-graph reduction (tests/golden/make_kat.py, case reducedD)."""
    from pyarrowspace_b200 import synth
    x = synth.make_items(3000, 40, 8, scale=100.0, n_clusters=10)
    s, g, cent, info = oracle_mod.build_reduced({"eps": 0.6, "k": 5, "topk": 5, "p": 2.0, "sigma": 0.3}, x, reduction={"max_iters": 6})
    assert np.array_equal(cent, golden["reducedD_centroids"])
    ip, ix, dt = g.csr()
    assert (ip == golden["reducedD_indptr"]).all() and (ix == golden["reducedD_indices"]).all() and (dt == golden["reducedD_data"]).all()
    assert (s.lambdas() == golden["reducedD_lambdas"]).all()
    want = golden["reducedD_info"]
    got = [info[k] for k in ("n_sampled", "n_probes", "two_nn_mean_ratio", "intrinsic_dim", "n_clusters", "iters", "converged")]
    assert np.array_equal(np.array(got, dtype=np.float64), want)


@pytest.mark.parametrize("case", GRAPH_SWITCH_CASES, ids=lambda c: ",".join("%s=%s" % kv for kv in c.items()))
def test_graph_switches_match_mirror(oracle_mod, case):
    """Every unpinned graph switch: C oracle == independent numpy restatement (L entry by entry, lambdas 1e-12)."""
    rng = np.random.default_rng(5)
    x = np.abs(rng.normal(size=(31, 17))) + 0.05
    gp = {"eps": 0.6, "k": 5, "topk": 3, "p": 2.0, "sigma": 0.3}
    if "distance" in case:
        gp["eps"], gp["sigma"] = (4.0, 2.0) if case["distance"] == "l2" else (16.0, 8.0)
    sw = dict(oracle_mod.PROFILES[case["profile"]]) if "profile" in case else dict(case)
    s, g = oracle_mod.build(gp, x, **case)
    m = oracle_np.build(x, gp["eps"], gp["k"], gp["topk"], gp["p"], gp["sigma"], **sw)
    import scipy.sparse as sp
    ip, ix, dt = g.csr()
    L = sp.csr_matrix((dt, ix, ip), shape=(g.nnodes, g.nnodes)).toarray()
    assert np.array_equal(L, m["L"])
    assert len(g.edges()) > 0
    np.testing.assert_allclose(s.lambdas(), m["lambdas"], rtol=1e-12, atol=0)


def test_graph_properties(oracle_mod):
    """Symmetric W, zero row sums, non-positive off-diagonals, sorted columns, <= k own neighbours."""
    x = synth.make_items(400, 40, 9, n_clusters=8)
    s, g = oracle_mod.build({"eps": 0.3, "k": 6, "topk": 5, "p": 2.0, "sigma": 0.1}, x)
    import scipy.sparse as sp
    ip, ix, dt = g.csr()
    L = sp.csr_matrix((dt, ix, ip), shape=(40, 40)).toarray()
    assert np.allclose(L, L.T, rtol=0, atol=0)
    assert np.abs(L.sum(axis=1)).max() < 1e-12
    off = L - np.diag(np.diag(L))
    assert (off <= 0).all()
    for a in range(40):
        assert (np.diff(ix[ip[a]:ip[a + 1]]) > 0).all()
    lam = s.lambdas()
    assert ((lam > 0) & (lam < 1)).all()


def test_tau_one_is_pure_cosine_and_scale_invariance(oracle_mod):
    x = synth.make_items(300, 24, 2, n_clusters=8)
    gp = {"eps": 0.5, "k": 4, "topk": 7, "p": 2.0, "sigma": 0.2}
    s, g = oracle_mod.build(gp, x)
    q = x[17] * 0.5 + 0.01
    h1 = s.search(q, g, 1.0)
    h2 = s.search(q * 3.0, g, 1.0)            # cosine term is scale invariant; lambda term is off at tau=1
    assert [i for i, _ in h1] == [i for i, _ in h2]
    cos = (x @ q) / (np.linalg.norm(x, axis=1) * np.linalg.norm(q))
    assert [i for i, _ in h1] == list(np.argsort(-cos, kind="stable")[:7])


def test_edge_cases(oracle_mod):
    with pytest.raises(oracle_mod.OracleError):
        oracle_mod.build({"eps": 1, "k": 1, "topk": 1, "p": 2.0}, np.zeros((0, 3)))
    # k = 0 -> no edges -> every lambda is 0 -> search refuses (src/lib.rs:156-159)
    x = np.abs(np.random.default_rng(0).normal(size=(6, 5))) + 0.1
    s, g = oracle_mod.build({"eps": 1.0, "k": 0, "topk": 2, "p": 2.0}, x)
    assert (s.lambdas() == 0).all() and len(g.edges()) == 0
    with pytest.raises(oracle_mod.OracleError) as ei:
        s.search(x[0], g, 0.5)
    assert ei.value.code == 3
    # topk larger than nitems -> nitems results; ties -> smaller index first
    x = np.array([[1.0, 2.0, 3.0], [1.0, 2.0, 3.0], [3.0, 2.0, 1.0]])
    s, g = oracle_mod.build({"eps": 1.0, "k": 2, "topk": 5, "p": 2.0, "sigma": 1.0}, x)
    hits = s.search(np.array([1.0, 2.0, 3.0]), g, 0.8)
    assert [i for i, _ in hits] == [0, 1, 2] and hits[0][1] == hits[1][1]
    # an all-zero item makes its Rayleigh quotient undefined (TAUMODE.md:13)
    with pytest.raises(oracle_mod.OracleError) as ei:
        oracle_mod.build({"eps": 1.0, "k": 2, "topk": 5, "p": 2.0}, np.array([[0.0, 0.0], [1.0, 2.0]]))
    assert ei.value.code == 2


# ----------------------------------------------------------------------------- pre-graph reduction

def _numpy_kmeans(S, K, iters):
    """Independent restatement of R4 (strided start, left-to-right squared distances, in-order member sums)."""
    ns, f = S.shape
    cent = np.stack([S[(j * ns) // K] for j in range(K)])
    assign = -np.ones(ns, dtype=np.int64)
    done, conv = 0, 0
    for _ in range(iters):
        d = np.zeros((ns, K))
        for t in range(f):
            diff = S[:, t][:, None] - cent[:, t][None, :]
            d += diff * diff
        a = d.argmin(axis=1)                     # first minimum: ties -> smaller centroid
        if np.array_equal(a, assign):
            conv = 1
            break
        assign = a
        for j in range(K):
            m = S[assign == j]
            if len(m):
                acc = np.zeros(f)
                for row in m:
                    acc = acc + row
                cent[j] = acc / len(m)
        done += 1
    return cent, done, conv


@pytest.mark.parametrize("n,f,red", [(900, 12, {}), (700, 31, {"n_clusters": 9, "sample_rate": 1.0, "max_iters": 30}),
                                     (500, 8, {"seed": 3, "sample_rate": 0.4, "max_iters": 3})])
def test_reduction_matches_an_independent_restatement(oracle_mod, n, f, red):
    rng = np.random.default_rng(n)
    cent = rng.normal(size=(6, f))
    x = cent[rng.integers(0, 6, n)] + 0.2 * rng.normal(size=(n, f))
    got, info = oracle_mod.reduce(x, red)
    rows = oracle_mod.reduction_sample(n, red)
    assert info["n_sampled"] == len(rows)
    K = red.get("n_clusters") or int(np.ceil(np.sqrt(n / 10.0)))
    assert info["n_clusters"] == K
    want, done, conv = _numpy_kmeans(x[rows], K, red.get("max_iters", 10))
    assert np.array_equal(got, want)
    assert (info["iters"], info["converged"]) == (done, conv)
    # two-NN statistic, brute force
    S = x[rows]
    P = min(2048, len(S))
    ratios = []
    for j in range(P):
        pos = (j * len(S)) // P
        d = np.zeros(len(S))
        for t in range(f):
            diff = S[pos, t] - S[:, t]
            d += diff * diff
        d[pos] = np.inf
        r = np.sqrt(np.sort(d)[:2])
        if r[0] > 0:
            ratios.append(r[1] / r[0])
    acc = 0.0
    for v in ratios:
        acc += v
    assert info["n_probes"] == len(ratios)
    assert info["two_nn_mean_ratio"] == acc / len(ratios)
    m = acc / len(ratios)
    assert info["intrinsic_dim"] == max(1, int(min(m / (m - 1.0), f)))


def test_reduction_rules_meet_the_published_data_point(oracle_mod):
    """tests/output/1760705545_v0_16/suggested_eps.md:8-11: N = 313841 -> K tested in [178, 179]; mean two-NN ratio 1.3560
    -> intrinsic dimension 3.  K rule on a small stand-in with n_total_for_k = 313841; the dimension rule on data of known
    dimension (a 2-D sheet and a 5-D ball embedded in 20 dimensions)."""
    rng = np.random.default_rng(0)
    x = rng.normal(size=(400, 6))
    _, info = oracle_mod.reduce(x, {"sample_rate": 1.0, "max_iters": 0, "probes": 0}, n_total_for_k=313841)
    assert info["n_clusters"] == 178
    assert int(1.3560 / (1.3560 - 1.0)) == 3
    basis = np.linalg.qr(rng.normal(size=(20, 20)))[0]
    for dim in (2, 5):
        pts = rng.uniform(size=(4000, dim)) @ basis[:dim]
        _, info = oracle_mod.reduce(pts, {"sample_rate": 1.0, "max_iters": 0, "probes": 1000})
        assert abs(info["intrinsic_dim"] - dim) <= 1, info


def test_reduced_build_uses_the_centroid_graph(oracle_mod):
    rng = np.random.default_rng(2)
    x = np.abs(rng.normal(size=(600, 16))) + 0.1
    gp = {"eps": 0.8, "k": 4, "topk": 3, "p": 2.0, "sigma": 0.4}
    s, g, cent, info = oracle_mod.build_reduced(gp, x, {"n_clusters": 20})
    g2 = oracle_mod.graph_from_nodes(np.ascontiguousarray(cent.T), gp)          # nodes = columns of the centroid matrix
    assert all(np.array_equal(a, b) for a, b in zip(g.csr(), g2.csr()))
    _, _, lam = g.taumode(x)
    assert np.array_equal(s.lambdas(), lam) and s.nitems == 600 and g.nnodes == 16
