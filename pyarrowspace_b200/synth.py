"""Seeded synthetic embeddings of the shapes BASELINE.json names.

Generated on the host with numpy so that the CPU oracle and the GPU path see identical bytes:
a mixture of Gaussian clusters (centres N(0,1), within-cluster sigma 0.3), rows L2-normalised,
multiplied by the donor test's scale factor (x100 Quora / MS MARCO, x12 CVE) and shifted by
+0.25*scale so every entry is > 0 (keeps the per-vector median tau away from its 1e-9 floor; the
reference's own toy data, tests/test_0.py:4-10, is all-positive too).
"""
import numpy as np

CONFIGS = {
    # name: (n_items, n_features, seed, scale, graph_params, tau, n_queries)
    "C2": (100_000, 384, 42, 100.0, {"eps": 0.5, "k": 4, "topk": 10, "p": 2.0, "sigma": 0.25}, 0.62, 10_000),
    "C3": (300_000, 768, 43, 12.0, {"eps": 1.31, "k": 25, "topk": 10, "p": 2.0, "sigma": 0.535}, 0.62, 10_000),
    "C4": (1_000_000, 384, 44, 100.0, {"eps": 10.0, "k": 25, "topk": 10, "p": 2.0, "sigma": None}, 0.62, 1_000_000),
}


def make_items(n, f, seed, scale=100.0, n_clusters=256, rows=None, dtype=np.float64, shift=0.25):
    """Rows [rows[0], rows[1]) of the n x f item matrix (whole matrix when rows is None).

    Row i depends only on (seed, i // block), so shards generated on different ranks agree.  shift = 0 gives the
    MEAN-ZERO variant (unit rows x scale, entries of both signs: what sentence-embedding models produce)."""
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((n_clusters, f))
    r0, r1 = (0, n) if rows is None else rows
    out = np.empty((r1 - r0, f), dtype=dtype)
    block = 65536
    for b0 in range((r0 // block) * block, r1, block):
        brng = np.random.default_rng([seed, 1 + b0 // block])
        m = min(block, n - b0)
        lab = brng.integers(0, n_clusters, size=m)
        x = centres[lab] + 0.3 * brng.standard_normal((m, f))
        x /= np.sqrt((x * x).sum(axis=1, keepdims=True))
        lo, hi = max(b0, r0), min(b0 + m, r1)
        if lo < hi:
            out[lo - r0:hi - r0] = x[lo - b0:hi - b0]
    # unit rows in hundreds of dimensions have entries well inside (-0.25, 0.25): a FIXED shift (not the
    # matrix minimum) keeps every entry positive and the matrix independent of how it is sharded
    out *= scale
    if shift:
        out += shift * scale
    return out


def make_queries(items, nq, seed, scale=100.0, noise=0.01):
    """Perturbed, unscaled copies of random items (the donor tests leave queries unscaled)."""
    rng = np.random.default_rng([seed, 7])
    sel = rng.integers(0, items.shape[0], size=nq)
    q = items[sel] / scale + noise * rng.standard_normal((nq, items.shape[1]))
    return np.ascontiguousarray(q), sel


def make_fresh_queries(f, nq, seed, n_clusters=256):
    """Queries that are NOT copies of items: fresh unit-norm draws from the same cluster mixture (cosine ~0.92 to the
    items of their cluster, no exact neighbour), unscaled.  Returns (queries, cluster labels)."""
    centres = np.random.default_rng(seed).standard_normal((n_clusters, f))
    rng = np.random.default_rng([seed, 11])
    lab = rng.integers(0, n_clusters, size=nq)
    q = centres[lab] + 0.3 * rng.standard_normal((nq, f))
    q /= np.sqrt((q * q).sum(axis=1, keepdims=True))
    return np.ascontiguousarray(q), lab


def config(name):
    n, f, seed, scale, gp, tau, nq = CONFIGS[name]
    return dict(n=n, f=f, seed=seed, scale=scale, graph_params=dict(gp), tau=tau, nq=nq)
