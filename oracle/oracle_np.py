"""numpy mirror of the CPU oracle -- TEST INFRASTRUCTURE ONLY.

A second, independent restatement (numpy / pure-Python loops) of the algorithm in
``oracle/oracle.c``; it exists so that the C oracle can be cross-checked on small
cases.  Nothing in the product path (``pyarrowspace_b200/``, ``arrowspace/``) may
import this module.  Only ``tests/`` use it.

Spec: SURVEY.md Appendix A (A1-A9).  Reference call sites it restates:
  build   -> /root/reference/src/lib.rs:270-300  (ArrowSpaceBuilder.build)
  search  -> /root/reference/src/lib.rs:132-174  (ArrowSpace.search)
  params  -> /root/reference/src/helpers.rs:48-77 (sigma default eps*0.5)
  graph   -> /root/reference/GRAPH_VARIABLES.md:3,7-10 (distance, eps, k, kernel)
  lambda  -> /root/reference/TAUMODE.md:8,18-19,24-27  (Rayleigh, bounded transform)
  score   -> /root/reference/TAUMODE.md:33             (alpha cos + beta lambda-prox)

PARITY UNPINNED for everything the two reference KATs cannot see:
the arithmetic lives in the un-vendored crate arrowspace 0.18.0
(/root/reference/Cargo.lock:94-97).
"""
import math
import numpy as np

TAU_FLOOR = 1e-9


def _seq_dot(a, b):
    s = 0.0
    for x, y in zip(a, b):
        s += float(x) * float(y)
    return s


def resolve_sigma(eps, sigma):
    return eps * 0.5 if sigma is None else float(sigma)      # helpers.rs:68-72


def build_graph(nodes, eps, k, p, sigma, kernel="inv_power", laplacian="combinatorial", symmetrise="max",
                k_counts_self=False, topk_prunes=False, topk=None, distance="cosine"):
    """nodes: (M, D) array, one node vector per row. Returns dense W, dense L, edge set.

    The keyword switches are the UNPINNED choices of SURVEY.md 8(c) (same names / meaning as oracle.h)."""
    nodes = np.asarray(nodes, dtype=np.float64)
    M = nodes.shape[0]
    nsq = [_seq_dot(nodes[a], nodes[a]) for a in range(M)]
    norm = [math.sqrt(v) for v in nsq]
    d = np.ones((M, M))
    for a in range(M):
        for b in range(M):
            if a == b:
                continue
            dot = _seq_dot(nodes[a], nodes[b])
            if distance == "cosine":
                den = norm[a] * norm[b]
                c = 0.0 if (norm[a] == 0.0 or norm[b] == 0.0) else dot / den
                d[a, b] = 1.0 - max(0.0, c)                     # GRAPH_VARIABLES.md:7
            else:
                d2 = max(0.0, (nsq[a] + nsq[b]) - 2.0 * dot)
                d[a, b] = math.sqrt(d2) if distance == "l2" else d2
    cap = k
    if topk_prunes and topk is not None:
        cap = min(cap, topk)
    if k_counts_self:
        cap -= 1
    cap = max(0, min(cap, M - 1))
    sel = np.zeros((M, M))                                      # directed selections
    for a in range(M):
        cand = sorted((d[a, b], b) for b in range(M) if b != a and d[a, b] <= eps)
        for dist, b in cand[:cap]:                              # GRAPH_VARIABLES.md:8
            t = (dist / sigma) ** p
            sel[a, b] = 1.0 / (1.0 + t) if kernel == "inv_power" else math.exp(-t)   # :3,9
    if symmetrise == "max":
        W = np.maximum(sel, sel.T)
    elif symmetrise == "avg":
        W = 0.5 * sel + 0.5 * sel.T
    elif symmetrise == "min":
        W = np.minimum(sel, sel.T)
    else:
        W = sel
    deg = np.array([sum(W[a, b] for b in range(M) if W[a, b] != 0.0) for a in range(M)])
    if laplacian == "combinatorial":
        L = -W.copy()
        for a in range(M):
            L[a, a] = deg[a]
    else:                                                        # normalised variants
        L = np.zeros((M, M))
        for a in range(M):
            for b in range(M):
                if a == b:
                    L[a, a] = 1.0 if deg[a] > 0 else 0.0
                elif W[a, b] != 0.0:
                    if laplacian == "sym":
                        L[a, b] = -(W[a, b] / math.sqrt(deg[a] * deg[b])) if (deg[a] > 0 and deg[b] > 0) else 0.0
                    else:
                        L[a, b] = -(W[a, b] / deg[a]) if deg[a] > 0 else 0.0
    edges = sorted((a, b) for a in range(M) for b in range(a + 1, M) if W[a, b] > 0.0 or W[b, a] > 0.0)
    return W, L, edges


def median(v):
    s = sorted(float(x) for x in v)
    n = len(s)
    return s[n // 2] if n % 2 else 0.5 * (s[n // 2 - 1] + s[n // 2])


def tau_of(x, tau_mode="median", tau_fixed=0.0):
    if tau_mode == "median":
        t = median(x)
    elif tau_mode == "median_abs":
        t = median(np.abs(x))
    elif tau_mode == "mean":
        t = sum(float(v) for v in x) / len(x)
    else:
        t = tau_fixed
    return t if t > TAU_FLOOR else TAU_FLOOR


def taumode_lambda(x, L, tau_mode="median", tau_fixed=0.0, lambda_form="bounded"):
    F = len(x)
    num = 0.0
    for a in range(F):
        y = 0.0
        for b in range(F):
            if L[a, b] != 0.0:
                y += L[a, b] * float(x[b])
        num += float(x[a]) * y
    den = _seq_dot(x, x)
    if den == 0.0:
        raise ValueError("all-zero vector")                     # TAUMODE.md:13
    E = num / den
    tau = tau_of(x, tau_mode, tau_fixed)
    Eb = E / (E + tau)                                           # TAUMODE.md:19,25
    if lambda_form == "bounded":
        return Eb
    tot = 0.0
    sq = 0.0
    for a in range(F):
        for b in range(a + 1, F):
            if L[a, b] != 0.0 or L[b, a] != 0.0:
                cab = (-0.5 * L[a, b]) + (-0.5 * L[b, a])       # coefficient of the symmetrised form
                e = cab * (float(x[a]) - float(x[b])) ** 2
                tot += e
                sq += e * e
    G = 0.0 if tot == 0.0 else min(1.0, max(0.0, sq / (tot * tot)))   # TAUMODE.md:26-27
    return tau * Eb + (1.0 - tau) * G                            # TAUMODE.md:8


def search(items, lambdas, q, lam_q, topk, tau):
    nq = math.sqrt(_seq_dot(q, q))
    out = []
    for i, x in enumerate(items):
        nx = math.sqrt(_seq_dot(x, x))
        den = nq * nx
        c = 0.0 if den == 0.0 else _seq_dot(q, x) / den          # README KAT: bit exact
        s = tau * c + (1.0 - tau) * (1.0 / (1.0 + abs(lam_q - lambdas[i])))
        out.append((i, s))
    out.sort(key=lambda t: (-t[1], t[0]))
    return out[: min(topk, len(out))]


def build(items, eps, k, topk, p, sigma=None, nodes="feature_columns", **sw):
    items = np.asarray(items, dtype=np.float64)
    sigma = resolve_sigma(eps, sigma)
    node_mat = items.T if nodes == "feature_columns" else items
    gkw = {k_: sw[k_] for k_ in ("kernel", "laplacian", "symmetrise", "k_counts_self", "topk_prunes", "distance") if k_ in sw}
    gkw["topk"] = topk
    lkw = {k_: sw[k_] for k_ in ("tau_mode", "tau_fixed", "lambda_form") if k_ in sw}
    W, L, edges = build_graph(node_mat, eps, k, p, sigma, **gkw)
    lam = None
    if nodes == "feature_columns":
        lam = np.array([taumode_lambda(x, L, **lkw) for x in items])
    return dict(items=items, W=W, L=L, edges=edges, lambdas=lam, topk=topk, lkw=lkw)
