"""Evaluation outputs in the formats of the reference's own evaluation scripts -- SURVEY.md section 8(f) rank 4.

The reference publishes its quality tables as files written by tests/test_2_CVE_db.py (four CSVs) and tests/test_3_beir.py
(one JSON).  This module produces files with the SAME columns / keys from searches run on this backend, so the published
tables can be regenerated against it:

    search results CSV   query_id, query_text, tau_method, rank, cve_id, title, score   (test_2_CVE_db.py:248-272)
    comparison CSV       Spearman / Kendall / NDCG@10 between the three tau settings     (test_2_CVE_db.py:274-303)
    tail CSV             head / tail statistics of the score lists                        (test_2_CVE_db.py:305-340)
    summary CSV          metric_type, metric_name, value, std_dev                         (test_2_CVE_db.py:342-393)
    BEIR JSON            dataset, dataset_size, num_queries, tau, metrics{cosine, lambda_aware, improvements}
                                                                                          (test_3_beir.py:410-437)
The three tau settings and their labels are the donor script's (test_2_CVE_db.py:24-39: Cosine 1.0, Hybrid 0.8, Taumode 0.62).
All searches go through ArrowSpace.search_batch (one batched call per tau instead of one call per query and tau).
"""
import csv
import json

import numpy as np

TAU_SETTINGS = (("Cosine", 1.0), ("Hybrid", 0.8), ("Taumode", 0.62))
TAIL_LABELS = ("Cosine (τ=1.0)", "Hybrid (τ=0.8)", "Taumode (τ=0.62)")

SEARCH_FIELDS = ["query_id", "query_text", "tau_method", "rank", "cve_id", "title", "score"]
COMPARISON_FIELDS = ["query_id", "query_text", "min_length",
                     "spearman_cosine_hybrid", "spearman_cosine_taumode", "spearman_hybrid_taumode",
                     "kendall_cosine_hybrid", "kendall_cosine_taumode", "kendall_hybrid_taumode",
                     "ndcg_hybrid_vs_cosine", "ndcg_taumode_vs_cosine", "ndcg_taumode_vs_hybrid"]
TAIL_FIELDS = ["query_id", "query_text", "tau_method", "head_mean", "tail_mean", "tail_std", "tail_to_head_ratio", "tail_cv",
               "tail_decay_rate", "n_tail_items", "total_items"]
SUMMARY_FIELDS = ["metric_type", "metric_name", "value", "std_dev"]


def run_tau_sweep(aspace, gl, queries, settings=TAU_SETTINGS):
    """-> {label: list over queries of [(index, score), ...]} : one batched search per tau."""
    out = {}
    q = np.ascontiguousarray(queries, dtype=np.float64)
    for label, tau in settings:
        idx, sc = aspace.search_batch(q, gl, tau)
        out[label] = [[(int(i), float(s)) for i, s in zip(ri, rs) if i >= 0] for ri, rs in zip(np.asarray(idx), np.asarray(sc))]
    return out


def rank_agreement(a, b):
    """(Spearman rho, Kendall tau) of the positions of the items two result lists share; (0, 0) below two shared items."""
    from scipy.stats import kendalltau, spearmanr
    pos_a = {i: r for r, (i, _) in enumerate(a)}
    pos_b = {i: r for r, (i, _) in enumerate(b)}
    shared = [i for i in pos_a if i in pos_b]
    if len(shared) < 2:
        return 0.0, 0.0
    ra, rb = [pos_a[i] for i in shared], [pos_b[i] for i in shared]
    rho, tau = spearmanr(ra, rb)[0], kendalltau(ra, rb)[0]
    return (0.0 if np.isnan(rho) else float(rho)), (0.0 if np.isnan(tau) else float(tau))


def ndcg_against(pred, ref, k=10):
    """NDCG@k of `pred` when the first k entries of `ref` are the ground truth with graded relevance k, k-1, ..., 1 and the
    predicted items are ordered by their (max-normalised) scores."""
    rel = {i: k - r for r, (i, _) in enumerate(ref[:k])}
    gains = np.array([rel.get(i, 0) for i, _ in pred[:k]], dtype=np.float64)
    if gains.sum() == 0:
        return 0.0
    scores = np.array([s for _, s in pred[:k]], dtype=np.float64)
    order = np.argsort(-scores, kind="stable")
    disc = 1.0 / np.log2(np.arange(2, len(gains) + 2))
    dcg = float((gains[order] * disc).sum())
    ideal = float((np.sort(gains)[::-1] * disc).sum())
    return dcg / ideal if ideal > 0 else 0.0


def tail_statistics(results, k_head=3, k_tail=20):
    """Head / tail statistics of one score list (None when the list is not longer than the head)."""
    seg = [s for _, s in results[:k_tail]]
    if len(seg) <= k_head:
        return None
    head, tail = np.array(seg[:k_head]), np.array(seg[k_head:])
    hm, tm, ts = float(head.mean()), float(tail.mean()), float(tail.std())
    return {"head_mean": hm, "tail_mean": tm, "tail_std": ts,
            "tail_to_head_ratio": tm / hm if hm > 1e-10 else 0.0, "tail_cv": ts / tm if tm > 1e-10 else 0.0,
            "tail_decay_rate": float(tail[0] - tail[-1]) / len(tail) if len(tail) > 1 else 0.0,
            "n_tail_items": int(len(tail)), "total_items": int(len(seg))}


def compare(sweep, query_texts):
    """Per-query comparison records of the three tau settings (the donor's `comparison_metrics`)."""
    labels = [l for l, _ in TAU_SETTINGS]
    out = []
    for qi, text in enumerate(query_texts):
        c, h, t = (sweep[l][qi] for l in labels)
        m = min(len(c), len(h), len(t))
        c, h, t = c[:m], h[:m], t[:m]
        pairs = ((c, h), (c, t), (h, t))
        agree = [rank_agreement(a, b) for a, b in pairs]
        tails = {}
        for lab, res in zip(TAIL_LABELS, (c, h, t)):
            st = tail_statistics(res)
            if st is not None:
                tails[lab] = st
        out.append({"query": text, "min_length": m, "spearman": [a[0] for a in agree], "kendall": [a[1] for a in agree],
                    "ndcg": [ndcg_against(h, c), ndcg_against(t, c), ndcg_against(t, h)], "tail_metrics": tails})
    return out


def _f6(v):
    return "%.6f" % v


def write_search_results(path, query_texts, sweep, ids, titles, top=20):
    with open(path, "w", newline="", encoding="utf-8") as fh:
        w = csv.DictWriter(fh, fieldnames=SEARCH_FIELDS)
        w.writeheader()
        for qi, text in enumerate(query_texts):
            for label, _ in TAU_SETTINGS:
                for rank, (idx, score) in enumerate(sweep[label][qi][:top], 1):
                    w.writerow({"query_id": qi + 1, "query_text": text, "tau_method": label, "rank": rank, "cve_id": ids[idx],
                                "title": titles[idx], "score": _f6(score)})


def write_comparison(path, records):
    with open(path, "w", newline="", encoding="utf-8") as fh:
        w = csv.DictWriter(fh, fieldnames=COMPARISON_FIELDS)
        w.writeheader()
        for qi, m in enumerate(records):
            row = {"query_id": qi + 1, "query_text": m["query"], "min_length": m["min_length"]}
            for name, vals in (("spearman", m["spearman"]), ("kendall", m["kendall"])):
                for suffix, v in zip(("cosine_hybrid", "cosine_taumode", "hybrid_taumode"), vals):
                    row["%s_%s" % (name, suffix)] = _f6(v)
            for suffix, v in zip(("hybrid_vs_cosine", "taumode_vs_cosine", "taumode_vs_hybrid"), m["ndcg"]):
                row["ndcg_" + suffix] = _f6(v)
            w.writerow(row)


def write_tail(path, records):
    with open(path, "w", newline="", encoding="utf-8") as fh:
        w = csv.DictWriter(fh, fieldnames=TAIL_FIELDS)
        w.writeheader()
        for qi, m in enumerate(records):
            for label in TAIL_LABELS:
                st = m["tail_metrics"].get(label)
                if st is None:
                    continue
                row = {"query_id": qi + 1, "query_text": m["query"], "tau_method": label, "n_tail_items": st["n_tail_items"],
                       "total_items": st["total_items"]}
                for key in ("head_mean", "tail_mean", "tail_std", "tail_to_head_ratio", "tail_cv", "tail_decay_rate"):
                    row[key] = _f6(st[key])
                w.writerow(row)


def write_summary(path, records):
    with open(path, "w", newline="", encoding="utf-8") as fh:
        w = csv.DictWriter(fh, fieldnames=SUMMARY_FIELDS)
        w.writeheader()
        for j, name in enumerate(("Hybrid vs Cosine", "Taumode vs Cosine", "Taumode vs Hybrid")):
            vals = [m["ndcg"][j] for m in records]
            w.writerow({"metric_type": "NDCG@10", "metric_name": name, "value": _f6(np.mean(vals)), "std_dev": _f6(np.std(vals))})
        for label in TAIL_LABELS:
            ratios = [m["tail_metrics"][label]["tail_to_head_ratio"] for m in records if label in m["tail_metrics"]]
            if ratios:
                w.writerow({"metric_type": "Tail/Head Ratio", "metric_name": label, "value": _f6(np.mean(ratios)),
                            "std_dev": _f6(np.std(ratios))})


def retrieval_metrics(result_lists, relevant, k=10):
    """Mean recall@k, MRR and binary NDCG@k of result lists against sets of relevant item indices."""
    rec, mrr, ndcg = [], [], []
    disc = 1.0 / np.log2(np.arange(2, k + 2))
    for res, rel in zip(result_lists, relevant):
        hits = [1.0 if i in rel else 0.0 for i, _ in res[:k]]
        hits += [0.0] * (k - len(hits))
        rec.append(sum(hits) / max(1, len(rel)))
        first = next((r for r, h in enumerate(hits) if h), None)
        mrr.append(0.0 if first is None else 1.0 / (first + 1))
        ideal = float(disc[: min(k, len(rel))].sum())
        ndcg.append(float((np.array(hits) * disc).sum()) / ideal if ideal > 0 else 0.0)
    return float(np.mean(rec)), float(np.mean(mrr)), float(np.mean(ndcg))


def write_beir_json(path, dataset, dataset_size, tau, cosine_lists, lambda_lists, relevant, k=10):
    """The donor's beir_evaluation_results.json (test_3_beir.py:410-437) from two sets of result lists."""
    c, l = retrieval_metrics(cosine_lists, relevant, k), retrieval_metrics(lambda_lists, relevant, k)
    pct = [100.0 * (b - a) / a if a > 0 else 0.0 for a, b in zip(c, l)]
    doc = {"dataset": dataset, "dataset_size": int(dataset_size), "num_queries": len(cosine_lists), "tau": float(tau),
           "metrics": {"cosine": {"recall@10": c[0], "mrr": c[1], "ndcg@10": c[2]},
                       "lambda_aware": {"recall@10": l[0], "mrr": l[1], "ndcg@10": l[2]},
                       "improvements": {"recall@10_pct": pct[0], "mrr_pct": pct[1], "ndcg@10_pct": pct[2]}}}
    with open(path, "w") as fh:
        json.dump(doc, fh, indent=2)
    return doc
