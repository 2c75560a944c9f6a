"""search_hybrid without a GPU: the oracle's restatement against an independent numpy one, and the
re-ranking kernels of csrc/hybrid.cuh walked on the CPU (their per-thread bodies compile with g++) against the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _data(n, f, seed, dup=True):
    rng = np.random.default_rng(seed)
    centres = rng.normal(size=(5, f))
    x = centres[rng.integers(0, 5, n)] + 0.3 * rng.normal(size=(n, f)) + 0.4
    if dup and n > 40:
        x[n // 2:n // 2 + 12] = x[:12]                         # exact duplicates: ties in cosine and in the score
    q = x[rng.integers(0, n, 23)] * 1.03 + 0.02 * rng.normal(size=(23, f))
    q[:4] = x[:4] * 2.0                                        # queries that tie duplicates exactly
    return x, q


def _numpy_hybrid(x, lam, nrm, q, lq, tau, topk, pool):
    n = x.shape[0]
    m = min(n, max(topk, pool if pool > 0 else min(2 * topk, 31)))
    out_i = np.full((q.shape[0], topk), -1, dtype=np.int64)
    out_s = np.full((q.shape[0], topk), np.nan)
    for qi in range(q.shape[0]):
        nq = np.sqrt(sum(v * v for v in q[qi]))
        cos = np.empty(n)
        for i in range(n):
            d = 0.0
            for a, b in zip(q[qi], x[i]):
                d += a * b
            den = nq * nrm[i]
            cos[i] = 0.0 if den == 0.0 else d / den
        short = np.lexsort((np.arange(n), -cos))[:m]
        sc = tau * cos[short] + (1.0 - tau) * (1.0 / (1.0 + np.abs(lq[qi] - lam[short])))
        order = np.lexsort((short, -sc))[:min(topk, n)]
        out_i[qi, :len(order)] = short[order]
        out_s[qi, :len(order)] = sc[order]
    return out_i, out_s


@pytest.mark.parametrize("n,f,topk,pool", [(120, 7, 3, 0), (300, 24, 5, 0), (300, 24, 5, 7), (60, 9, 10, 0), (9, 5, 4, 0), (200, 16, 6, 1000)])
def test_oracle_hybrid_equals_numpy(oracle_mod, n, f, topk, pool):
    x, q = _data(n, f, n + f)
    gp = {"eps": 0.7, "k": 4, "topk": topk, "p": 2.0, "sigma": 0.3}
    s, g = oracle_mod.build(gp, x)
    idx, sc, lq = s.search_hybrid_batch(q, g, 0.62, pool)
    ni, ns = _numpy_hybrid(x, s.lambdas(), s.norms(), q, lq, 0.62, topk, pool)
    assert np.array_equal(idx, ni)
    m = ni >= 0
    assert np.array_equal(sc[m], ns[m]) and np.isnan(sc[~m]).all()            # same expression, same order: bit for bit
    if pool >= n:                                                              # the shortlist is everything: plain search
        i2, s2, _ = s.search_batch(q, g, 0.62)
        assert np.array_equal(idx, i2) and np.array_equal(sc[m], s2[m])


def test_oracle_hybrid_has_no_lambda_zero_assertion(oracle_mod):
    """search asserts lambda_q != 0 (src/lib.rs:156-159); search_hybrid does not (src/lib.rs:182-219)."""
    x = np.eye(4) + 0.0
    gp = {"eps": 1e-6, "k": 2, "topk": 2, "p": 2.0, "sigma": 1.0}          # no edges: every lambda is 0
    s, g = oracle_mod.build(gp, x)
    q = np.array([[1.0, 0.5, 0.0, 0.0]])
    with pytest.raises(oracle_mod.OracleError):
        s.search_batch(q, g, 0.9)
    idx, sc, lq = s.search_hybrid_batch(q, g, 0.9)
    assert lq[0] == 0.0 and list(idx[0]) == [0, 1]


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hyb") / "libhybrid_host.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "host_emul", "hybrid_host.cpp")])
    lib = C.CDLL(out)
    vp, i64 = C.c_void_p, C.c_int64
    lib.hyb_emulate.argtypes = [i64, i64, i64, vp, C.c_int, vp, C.c_int, C.c_int, i64, vp, vp, vp, vp, C.c_double, vp, vp, vp, vp]
    lib.hyb_emulate.restype = None
    return lib


@pytest.mark.parametrize("n,f,topk,pool,tau", [(300, 24, 5, 0, 0.62), (300, 23, 5, 7, 0.9), (60, 9, 10, 0, 0.3), (9, 5, 4, 0, 0.62),
                                               (500, 130, 3, 0, 1.0), (200, 16, 6, 1000, 0.0)])
def test_rerank_kernel_bodies_on_the_cpu_equal_the_oracle(oracle_mod, emul, n, f, topk, pool, tau):
    """What asp_search_hybrid_batch does after the shortlist: hybrid_rescore_kernel + hybrid_select_kernel, every thread
    index walked on the CPU with the device's buffer layout (pitched rows, -1 padded shortlists), against orc_search_hybrid.
    The shortlist is the search at tau = 1 with topk = pool, exactly what the device path asks of its validated search."""
    x, q = _data(n, f, 3 * n + f)
    gp = {"eps": 0.7, "k": 4, "topk": topk, "p": 2.0, "sigma": 0.3}
    s, g = oracle_mod.build(gp, x)
    oidx, osc, lq = s.search_hybrid_batch(q, g, tau, pool)
    m = min(n, max(topk, pool if pool > 0 else min(2 * topk, 31)))
    m_dev = max(topk, pool if pool > 0 else min(2 * topk, 31))
    m_dev = min(m_dev, n)
    assert m_dev == m
    s2, g2 = oracle_mod.build(dict(gp, topk=m), x)                          # same graph, shortlist-sized result lists
    pidx, _, lq2 = s2.search_batch(q, g2, 1.0) if (lq != 0).all() else (None, None, None)
    if pidx is None:
        pytest.skip("a query has lambda 0: the oracle's plain search asserts")
    assert np.array_equal(lq, lq2)
    fp = (f + 3) // 4 * 4                                                     # the device pitch: rows padded to 4 doubles
    xp = np.zeros((n, fp)); xp[:, :f] = x
    qp = np.zeros((q.shape[0], fp)); qp[:, :f] = q
    nq = q.shape[0]
    norm_q = np.array([np.sqrt(sum(v * v for v in row)) for row in q])        # left to right, as taumode_kernel's norm chain
    pool_idx = np.ascontiguousarray(pidx, dtype=np.int64).copy()
    pool_score = np.full((nq, m), -7.0)
    out_idx = np.full((nq, topk), 99, dtype=np.int64)
    out_score = np.full((nq, topk), 99.0)
    lam, nrm = s.lambdas(), s.norms()
    emul.hyb_emulate(nq, m, topk, qp.ctypes.data, fp, xp.ctypes.data, fp, f, 0, nrm.ctypes.data, lam.ctypes.data,
                     norm_q.ctypes.data, lq.ctypes.data, tau, pool_idx.ctypes.data, pool_score.ctypes.data,
                     out_idx.ctypes.data, out_score.ctypes.data)
    assert np.array_equal(out_idx, oidx)
    ok = oidx >= 0
    assert np.array_equal(out_score[ok], osc[ok]) and np.isnan(out_score[~ok]).all()


def test_rerank_kernel_bodies_with_a_row_offset_and_padding(emul):
    """Global indices (row0 > 0) and shortlists padded with -1 (fewer items than slots)."""
    rng = np.random.default_rng(5)
    n, f, fp, nq, m, topk, row0 = 6, 3, 4, 2, 8, 7, 1000
    x = np.zeros((n, fp)); x[:, :f] = rng.normal(size=(n, f))
    q = np.zeros((nq, fp)); q[:, :f] = rng.normal(size=(nq, f))
    nrm = np.sqrt((x * x).sum(1)); nq_ = np.sqrt((q * q).sum(1))
    lam = rng.uniform(0.1, 0.9, n); lq = rng.uniform(0.1, 0.9, nq)
    pool_idx = np.full((nq, m), -1, dtype=np.int64)
    pool_idx[:, :n] = row0 + np.arange(n)[::-1]
    pool_score = np.zeros((nq, m))
    out_idx = np.zeros((nq, topk), dtype=np.int64); out_score = np.zeros((nq, topk))
    emul.hyb_emulate(nq, m, topk, q.ctypes.data, fp, x.ctypes.data, fp, f, row0, nrm.ctypes.data, lam.ctypes.data,
                     nq_.ctypes.data, lq.ctypes.data, 0.5, pool_idx.ctypes.data, pool_score.ctypes.data,
                     out_idx.ctypes.data, out_score.ctypes.data)
    for qi in range(nq):
        sc = np.array([0.5 * (q[qi] @ x[i]) / (nq_[qi] * nrm[i]) + 0.5 / (1 + abs(lq[qi] - lam[i])) for i in range(n)])
        order = np.lexsort((np.arange(n), -sc))
        assert list(out_idx[qi, :n]) == list(row0 + order) and out_idx[qi, n] == -1
        np.testing.assert_allclose(out_score[qi, :n], sc[order], rtol=1e-13)
        assert np.isnan(out_score[qi, n])


# ----------------------------------------------------------------------------- properties (hypothesis, CPU)

from hypothesis import HealthCheck, given, settings              # noqa: E402
from hypothesis import strategies as st                           # noqa: E402

FAST = settings(max_examples=30, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])


@FAST
@given(st.integers(1, 40), st.integers(1, 9), st.integers(1, 12), st.integers(0, 60), st.sampled_from([0.0, 0.3, 0.62, 1.0]),
       st.integers(0, 2 ** 31 - 1))
def test_select_and_rescore_bodies_on_random_shapes(emul, n, f, topk, pool, tau, seed):
    """Ragged shapes, heavy ties (values rounded to one decimal), shortlists shorter and longer than the item count: the
    kernel bodies against a direct numpy evaluation of H3 on the same shortlist."""
    rng = np.random.default_rng(seed)
    nq = int(rng.integers(1, 6))
    fp = (f + 3) // 4 * 4
    x = np.zeros((n, fp)); x[:, :f] = np.round(rng.normal(size=(n, f)), 1)
    x[x[:, :f].any(axis=1) == 0, 0] = 1.0                                       # no zero rows
    q = np.zeros((nq, fp)); q[:, :f] = np.round(rng.normal(size=(nq, f)), 1) + 0.05
    nrm = np.array([np.sqrt(sum(v * v for v in row[:f])) for row in x])
    nrq = np.array([np.sqrt(sum(v * v for v in row[:f])) for row in q])
    lam = np.round(rng.uniform(0, 1, n), 1); lq = np.round(rng.uniform(0, 1, nq), 1)
    m = max(topk, pool if pool > 0 else min(2 * topk, 31))                               # slots (may exceed n: padded with -1)
    pool_idx = np.full((nq, m), -1, dtype=np.int64)
    for qi in range(nq):
        take = rng.permutation(n)[:min(n, m)]
        pool_idx[qi, :len(take)] = take
    want_idx = np.full((nq, topk), -1, dtype=np.int64); want_sc = np.full((nq, topk), np.nan)
    for qi in range(nq):
        ids = pool_idx[qi][pool_idx[qi] >= 0]
        sc = []
        for i in ids:
            d = 0.0
            for a, b in zip(q[qi, :f], x[i, :f]):
                d += a * b
            den = nrq[qi] * nrm[i]
            c = 0.0 if den == 0.0 else d / den
            sc.append(tau * c + (1.0 - tau) * (1.0 / (1.0 + abs(lq[qi] - lam[i]))))
        sc = np.array(sc)
        order = np.lexsort((ids, -sc))[:topk]
        want_idx[qi, :len(order)] = ids[order]; want_sc[qi, :len(order)] = sc[order]
    pool_score = np.zeros((nq, m)); out_idx = np.zeros((nq, topk), dtype=np.int64); out_sc = np.zeros((nq, topk))
    work = pool_idx.copy()
    emul.hyb_emulate(nq, m, topk, q.ctypes.data, fp, x.ctypes.data, fp, f, 0, nrm.ctypes.data, lam.ctypes.data, nrq.ctypes.data,
                     lq.ctypes.data, tau, work.ctypes.data, pool_score.ctypes.data, out_idx.ctypes.data, out_sc.ctypes.data)
    assert np.array_equal(out_idx, want_idx)
    ok = want_idx >= 0
    assert np.array_equal(out_sc[ok], want_sc[ok]) and np.isnan(out_sc[~ok]).all()
