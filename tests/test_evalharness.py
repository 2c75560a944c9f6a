"""Evaluation-output formats: the files pyarrowspace_b200.evalharness writes carry the columns / keys of
the reference's own evaluation scripts (tests/test_2_CVE_db.py:248-393, tests/test_3_beir.py:410-437).  CPU only: the
searches are stubbed with fixed result lists."""
import csv
import json

import numpy as np

from pyarrowspace_b200 import evalharness as eh


class _FakeSpace:
    """search_batch stub: deterministic result lists that depend on tau (so the three settings disagree a little)."""

    def search_batch(self, q, gl, tau):
        rng = np.random.default_rng(int(tau * 100))
        nq, k = q.shape[0], 20
        idx = np.stack([rng.permutation(60)[:k] for _ in range(nq)]).astype(np.int64)
        sc = np.sort(rng.uniform(0.2, 1.0, size=(nq, k)), axis=1)[:, ::-1].copy()
        return idx, sc


def _sweep(nq=4):
    return eh.run_tau_sweep(_FakeSpace(), None, np.zeros((nq, 8)))


def test_csv_columns_match_the_reference_scripts(tmp_path):
    sweep = _sweep()
    texts = ["q%d" % i for i in range(4)]
    ids = ["CVE-%04d" % i for i in range(60)]
    titles = ["title %d" % i for i in range(60)]
    rec = eh.compare(sweep, texts)
    eh.write_search_results(tmp_path / "s.csv", texts, sweep, ids, titles)
    eh.write_comparison(tmp_path / "c.csv", rec)
    eh.write_tail(tmp_path / "t.csv", rec)
    eh.write_summary(tmp_path / "u.csv", rec)
    want = {
        "s.csv": ["query_id", "query_text", "tau_method", "rank", "cve_id", "title", "score"],
        "c.csv": ["query_id", "query_text", "min_length", "spearman_cosine_hybrid", "spearman_cosine_taumode",
                  "spearman_hybrid_taumode", "kendall_cosine_hybrid", "kendall_cosine_taumode", "kendall_hybrid_taumode",
                  "ndcg_hybrid_vs_cosine", "ndcg_taumode_vs_cosine", "ndcg_taumode_vs_hybrid"],
        "t.csv": ["query_id", "query_text", "tau_method", "head_mean", "tail_mean", "tail_std", "tail_to_head_ratio", "tail_cv",
                  "tail_decay_rate", "n_tail_items", "total_items"],
        "u.csv": ["metric_type", "metric_name", "value", "std_dev"],
    }
    for name, cols in want.items():
        rows = list(csv.reader(open(tmp_path / name, encoding="utf-8")))
        assert rows[0] == cols, name
        assert len(rows) > 1
    rows = list(csv.DictReader(open(tmp_path / "s.csv", encoding="utf-8")))
    assert len(rows) == 4 * 3 * 20 and {r["tau_method"] for r in rows} == {"Cosine", "Hybrid", "Taumode"}
    assert all(len(r["score"].split(".")[1]) == 6 for r in rows)                       # the donor's "%.6f"
    urows = list(csv.DictReader(open(tmp_path / "u.csv", encoding="utf-8")))
    assert [r["metric_type"] for r in urows] == ["NDCG@10"] * 3 + ["Tail/Head Ratio"] * 3
    assert [r["metric_name"] for r in urows[3:]] == ["Cosine (τ=1.0)", "Hybrid (τ=0.8)", "Taumode (τ=0.62)"]


def test_metric_definitions():
    a = [(1, .9), (2, .8), (3, .7), (4, .6)]
    assert eh.rank_agreement(a, a) == (1.0, 1.0)
    rho, tau = eh.rank_agreement(a, a[::-1])
    assert abs(rho + 1.0) < 1e-12 and abs(tau + 1.0) < 1e-12
    assert eh.rank_agreement(a, [(9, .5)]) == (0.0, 0.0)
    assert abs(eh.ndcg_against(a, a, k=4) - 1.0) < 1e-12
    assert eh.ndcg_against([(7, .9)], a, k=4) == 0.0
    # against scikit-learn's ndcg_score, the function the donor script calls (test_2_CVE_db.py:197-203)
    from sklearn.metrics import ndcg_score
    pred = [(3, .95), (1, .9), (8, .7), (2, .4)]
    rel = {1: 4, 2: 3, 3: 2, 4: 1}
    want = ndcg_score(np.array([[rel.get(i, 0) for i, _ in pred]]), np.array([[s / .95 for _, s in pred]]), k=4)
    assert abs(eh.ndcg_against(pred, a, k=4) - want) < 1e-12
    st = eh.tail_statistics([(i, 1.0 - 0.01 * i) for i in range(20)])
    assert st["n_tail_items"] == 17 and st["total_items"] == 20 and abs(st["head_mean"] - 0.99) < 1e-12
    assert eh.tail_statistics(a[:3]) is None


def test_beir_json_keys(tmp_path):
    cos = [[(1, .9), (5, .8)], [(2, .9), (3, .8)]]
    lam = [[(5, .9), (1, .8)], [(3, .9), (2, .8)]]
    doc = eh.write_beir_json(tmp_path / "b.json", "MS MARCO (BeIR)", 1000, 0.62, cos, lam, [{5}, {3}])
    got = json.load(open(tmp_path / "b.json"))
    assert got == doc
    assert list(got) == ["dataset", "dataset_size", "num_queries", "tau", "metrics"]
    assert list(got["metrics"]) == ["cosine", "lambda_aware", "improvements"]
    assert list(got["metrics"]["cosine"]) == ["recall@10", "mrr", "ndcg@10"]
    assert list(got["metrics"]["improvements"]) == ["recall@10_pct", "mrr_pct", "ndcg@10_pct"]
    assert got["metrics"]["cosine"]["mrr"] == 0.5 and got["metrics"]["lambda_aware"]["mrr"] == 1.0
    assert abs(got["metrics"]["improvements"]["mrr_pct"] - 100.0) < 1e-12
