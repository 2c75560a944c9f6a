"""ctypes front-end of the CPU oracle (``oracle/liboracle.so``) -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``pyarrowspace_b200`` / ``arrowspace``) never does; it fails loudly when its CUDA
library is missing instead of falling back to anything here.

PARITY: pinned by the reference's two known-answer tests (README.md:69 bit exact;
tests/test_0.py:29-61, 11/12 indices), UNPINNED otherwise -- see oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

NODES = {"feature_columns": 0, "items": 1}
KERNEL = {"inv_power": 0, "gaussian": 1}
TAU_MODE = {"median": 0, "median_abs": 1, "mean": 2, "fixed": 3}
LAMBDA_FORM = {"bounded": 0, "synthetic": 1}
SYMMETRISE = {"max": 0, "avg": 1, "min": 2, "none": 3}
LAPLACIAN = {"combinatorial": 0, "sym": 1, "rw": 2}
DISTANCE = {"cosine": 0, "l2": 1, "l2sq": 2}

# Named switch sets.  "kat12" is the set tools/fit_switches.py found to reproduce all 12 indices of the reference's
# tests/test_0.py:29-61 (the default spec of SURVEY.md Appendix A reproduces 11); README.md:69 is tau = 1 and therefore
# bit exact under every profile.
PROFILES = {
    "default": {},
    "kat12": {"symmetrise": "none", "laplacian": "sym", "k_counts_self": True, "topk_prunes": True},
}

ERRORS = {
    1: "items must be non-empty 2D array",
    2: "all-zero vector in taumode lambda",
    3: "The lambdas are zero, check the magnitude of items and eps.",
    4: "bad argument",
    5: "out of memory",
}


class OracleError(RuntimeError):
    def __init__(self, code):
        super().__init__(ERRORS.get(code, "oracle error %d" % code))
        self.code = code


class _Params(C.Structure):
    _fields_ = [("eps", C.c_double), ("k", C.c_int64), ("topk", C.c_int64),
                ("p", C.c_double), ("sigma", C.c_double)]


class _Switches(C.Structure):
    _fields_ = [("nodes", C.c_int32), ("kernel", C.c_int32), ("tau_mode", C.c_int32),
                ("lambda_form", C.c_int32), ("tau_fixed", C.c_double), ("symmetrise", C.c_int32),
                ("laplacian", C.c_int32), ("k_counts_self", C.c_int32), ("topk_prunes", C.c_int32),
                ("distance", C.c_int32)]


class _Reduction(C.Structure):
    _fields_ = [("sample_rate", C.c_double), ("seed", C.c_uint64), ("n_clusters", C.c_int32), ("max_iters", C.c_int32),
                ("probes", C.c_int32), ("reserved", C.c_int32)]


class _ReductionInfo(C.Structure):
    _fields_ = [("n_sampled", C.c_int64), ("n_probes", C.c_int64), ("two_nn_mean_ratio", C.c_double),
                ("intrinsic_dim", C.c_int32), ("n_clusters", C.c_int32), ("iters", C.c_int32), ("converged", C.c_int32)]


def make_reduction(reduction=None):
    """True / None -> the defaults (keep rate 0.6, seed 42, K by rule, 10 iterations, 2048 probes); dict overrides."""
    red = _Reduction()
    lib().orc_default_reduction(C.byref(red))
    if isinstance(reduction, dict):
        for key, val in reduction.items():
            if key not in dict(_Reduction._fields_) or key == "reserved":
                raise ValueError("unknown reduction option %r" % key)
            setattr(red, key, val)
    return red


def build_library(force=False):
    """Compile oracle.c (gcc) -- building the checker is not using it."""
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_library()
        L = C.CDLL(_LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)
        vp = C.c_void_p
        L.orc_default_switches.argtypes = [C.POINTER(_Switches)]
        L.orc_build.argtypes = [vp, C.c_int64, C.c_int32, C.POINTER(_Params), C.POINTER(_Switches),
                                C.POINTER(vp), C.POINTER(vp)]
        L.orc_graph_from_nodes_t.argtypes = [vp, C.c_int64, C.c_int64, C.POINTER(_Params),
                                             C.POINTER(_Switches), C.POINTER(vp)]
        for name, res in (("orc_space_nitems", C.c_int64), ("orc_space_nfeatures", C.c_int32),
                          ("orc_space_items", dp), ("orc_space_lambdas", dp), ("orc_space_norms", dp),
                          ("orc_graph_nnodes", C.c_int64), ("orc_graph_nnz", C.c_int64),
                          ("orc_graph_indptr", ip), ("orc_graph_indices", C.POINTER(C.c_int32)),
                          ("orc_graph_data", dp)):
            fn = getattr(L, name)
            fn.argtypes = [vp]
            fn.restype = res
        L.orc_taumode.argtypes = [vp, C.POINTER(_Switches), vp, C.c_int64, vp, vp, vp]
        L.orc_search.argtypes = [vp, vp, C.POINTER(_Switches), vp, C.c_int64, C.c_double, vp, vp, vp]
        L.orc_search_hybrid.argtypes = [vp, vp, C.POINTER(_Switches), vp, C.c_int64, C.c_double, C.c_int64, vp, vp, vp]
        L.orc_gram_columns.argtypes = [vp, C.c_int64, C.c_int64, vp]
        L.orc_gram_columns.restype = None
        L.orc_scores.argtypes = [vp, vp, C.c_double, C.c_double, vp]
        L.orc_scores.restype = None
        L.orc_default_reduction.argtypes = [C.POINTER(_Reduction)]
        L.orc_default_reduction.restype = None
        L.orc_reduction_sample.argtypes = [C.POINTER(_Reduction), C.c_int64, C.c_int64, vp]
        L.orc_reduction_sample.restype = C.c_int64
        L.orc_reduce.argtypes = [vp, C.c_int64, C.c_int32, C.POINTER(_Reduction), C.c_int64, C.POINTER(_ReductionInfo), vp,
                                 C.c_int64]
        L.orc_build_reduced.argtypes = [vp, C.c_int64, C.c_int32, C.POINTER(_Params), C.POINTER(_Switches),
                                        C.POINTER(_Reduction), C.POINTER(vp), C.POINTER(vp), C.POINTER(_ReductionInfo), vp,
                                        C.c_int64]
        L.orc_free_space.argtypes = [vp]
        L.orc_free_graph.argtypes = [vp]
        L.orc_free_space.restype = None
        L.orc_free_graph.restype = None
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _f64(a, ndim=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if ndim is not None and a.ndim != ndim:
        raise ValueError("expected %d-D array" % ndim)
    return a


def _check(rc):
    if rc != 0:
        raise OracleError(rc)


def make_switches(nodes="feature_columns", kernel="inv_power", tau_mode="median",
                  lambda_form="bounded", tau_fixed=0.0, symmetrise="max", laplacian="combinatorial",
                  k_counts_self=False, topk_prunes=False, distance="cosine", profile=None):
    if profile is not None:
        kw = dict(nodes=nodes, kernel=kernel, tau_mode=tau_mode, lambda_form=lambda_form, tau_fixed=tau_fixed,
                  symmetrise=symmetrise, laplacian=laplacian, k_counts_self=k_counts_self, topk_prunes=topk_prunes,
                  distance=distance)
        kw.update(PROFILES[profile])
        return make_switches(**kw)
    return _Switches(NODES[nodes], KERNEL[kernel], TAU_MODE[tau_mode], LAMBDA_FORM[lambda_form],
                     float(tau_fixed), SYMMETRISE[symmetrise], LAPLACIAN[laplacian], int(bool(k_counts_self)),
                     int(bool(topk_prunes)), DISTANCE[distance])


def resolve_params(gp):
    """helpers.rs:48-77: eps,k,topk,p required; sigma missing/None -> eps*0.5."""
    sigma = gp.get("sigma")
    if sigma is None:
        sigma = gp["eps"] * 0.5
    return _Params(float(gp["eps"]), int(gp["k"]), int(gp["topk"]), float(gp["p"]), float(sigma))


class Graph:
    def __init__(self, handle, params):
        self._h = handle
        self.params = params

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.orc_free_graph(self._h)
            self._h = None

    @property
    def nnodes(self):
        return lib().orc_graph_nnodes(self._h)

    @property
    def nnz(self):
        return lib().orc_graph_nnz(self._h)

    def csr(self):
        m, nnz = self.nnodes, self.nnz
        L = lib()
        indptr = np.ctypeslib.as_array(L.orc_graph_indptr(self._h), shape=(m + 1,)).copy()
        indices = np.ctypeslib.as_array(L.orc_graph_indices(self._h), shape=(nnz,)).copy()
        data = np.ctypeslib.as_array(L.orc_graph_data(self._h), shape=(nnz,)).copy()
        return indptr, indices, data

    def edges(self):
        """Sorted list of (a, b), a < b, with W_ab > 0 (SURVEY.md A7 'edge set')."""
        indptr, indices, data = self.csr()
        rows = np.repeat(np.arange(self.nnodes, dtype=np.int64), np.diff(indptr))
        keep = (indices > rows) & (data < 0.0)
        return np.stack([rows[keep], indices[keep].astype(np.int64)], axis=1)

    def taumode(self, x, switches=None):
        x = _f64(x)
        single = x.ndim == 1
        x2 = x.reshape(1, -1) if single else x
        if x2.shape[1] != self.nnodes:
            raise ValueError("vector length must equal nnodes")
        nq = x2.shape[0]
        e, t, lam = (np.empty(nq) for _ in range(3))
        sw = C.byref(switches) if switches is not None else None
        _check(lib().orc_taumode(self._h, sw, x2.ctypes.data, nq, e.ctypes.data, t.ctypes.data,
                                 lam.ctypes.data))
        return (e[0], t[0], lam[0]) if single else (e, t, lam)


class Space:
    def __init__(self, handle, switches):
        self._h = handle
        self.switches = switches

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.orc_free_space(self._h)
            self._h = None

    @property
    def nitems(self):
        return lib().orc_space_nitems(self._h)

    @property
    def nfeatures(self):
        return lib().orc_space_nfeatures(self._h)

    def lambdas(self):
        return np.ctypeslib.as_array(lib().orc_space_lambdas(self._h), shape=(self.nitems,)).copy()

    def norms(self):
        return np.ctypeslib.as_array(lib().orc_space_norms(self._h), shape=(self.nitems,)).copy()

    def search_batch(self, queries, graph, tau):
        q = _f64(queries, 2)
        if q.shape[1] != self.nfeatures:
            raise ValueError("query length %d must match nfeatures %d" % (q.shape[1], self.nfeatures))
        nq, topk = q.shape[0], graph.params.topk
        idx = np.empty((nq, topk), dtype=np.int64)
        sc = np.empty((nq, topk), dtype=np.float64)
        lq = np.empty(nq, dtype=np.float64)
        _check(lib().orc_search(self._h, graph._h, C.byref(self.switches), q.ctypes.data, nq,
                                float(tau), idx.ctypes.data, sc.ctypes.data, lq.ctypes.data))
        return idx, sc, lq

    def search(self, query, graph, tau):
        idx, sc, _ = self.search_batch(np.asarray(query, dtype=np.float64).reshape(1, -1), graph, tau)
        return [(int(i), float(s)) for i, s in zip(idx[0], sc[0]) if i >= 0]

    def search_hybrid_batch(self, queries, graph, tau, pool=0):
        """orc_search_hybrid (src/lib.rs:182-219 restated, PARITY UNPINNED): cosine shortlist of `pool` items (0: min(2 * topk, 31)),
        re-ranked by the lambda-aware score; no lambda_q != 0 assertion."""
        q = _f64(queries, 2)
        if q.shape[1] != self.nfeatures:
            raise ValueError("query length %d must match nfeatures %d" % (q.shape[1], self.nfeatures))
        nq, topk = q.shape[0], graph.params.topk
        idx = np.empty((nq, topk), dtype=np.int64)
        sc = np.empty((nq, topk), dtype=np.float64)
        lq = np.empty(nq, dtype=np.float64)
        _check(lib().orc_search_hybrid(self._h, graph._h, C.byref(self.switches), q.ctypes.data, nq, float(tau), int(pool),
                                       idx.ctypes.data, sc.ctypes.data, lq.ctypes.data))
        return idx, sc, lq

    def search_hybrid(self, query, graph, tau, pool=0):
        idx, sc, _ = self.search_hybrid_batch(np.asarray(query, dtype=np.float64).reshape(1, -1), graph, tau, pool)
        return [(int(i), float(s)) for i, s in zip(idx[0], sc[0]) if i >= 0]

    def scores(self, query, lambda_q, tau):
        q = _f64(query, 1)
        out = np.empty(self.nitems)
        lib().orc_scores(self._h, q.ctypes.data, float(lambda_q), float(tau), out.ctypes.data)
        return out


def build(graph_params, items, **switch_kw):
    """ArrowSpaceBuilder.build restated on the CPU (src/lib.rs:270-300)."""
    x = _f64(items)
    if x.ndim != 2 or x.shape[0] == 0 or x.shape[1] == 0:
        raise OracleError(1)
    gp = resolve_params(graph_params)
    sw = make_switches(**switch_kw)
    hs, hg = C.c_void_p(), C.c_void_p()
    _check(lib().orc_build(x.ctypes.data, x.shape[0], x.shape[1], C.byref(gp), C.byref(sw),
                           C.byref(hs), C.byref(hg)))
    return Space(hs, sw), Graph(hg, gp)


def _info_dict(info):
    return {name: getattr(info, name) for name, _ in _ReductionInfo._fields_}


def reduction_sample(n, reduction=None, row0=0):
    """R1: the kept rows of [row0, row0 + n) as local indices."""
    red = make_reduction(reduction)
    rows = np.empty(n, dtype=np.int32)
    cnt = lib().orc_reduction_sample(C.byref(red), int(row0), int(n), rows.ctypes.data)
    return rows[:cnt].copy()


def reduce(items, reduction=None, n_total_for_k=0):
    """R1-R4: (centroids K x f, info dict)."""
    x = _f64(items, 2)
    red = make_reduction(reduction)
    cap = min(x.shape[0], 65535)
    cent = np.empty((cap, x.shape[1]))
    info = _ReductionInfo()
    _check(lib().orc_reduce(x.ctypes.data, x.shape[0], x.shape[1], C.byref(red), int(n_total_for_k), C.byref(info),
                            cent.ctypes.data, cap))
    return cent[:info.n_clusters].copy(), _info_dict(info)


def build_reduced(graph_params, items, reduction=None, **switch_kw):
    """ArrowSpaceBuilder.build with the pre-graph reduction: (Space, Graph, centroids, info)."""
    x = _f64(items)
    if x.ndim != 2 or x.shape[0] == 0 or x.shape[1] == 0:
        raise OracleError(1)
    gp = resolve_params(graph_params)
    sw = make_switches(**switch_kw)
    red = make_reduction(reduction)
    cap = min(x.shape[0], 65535)
    cent = np.empty((cap, x.shape[1]))
    info = _ReductionInfo()
    hs, hg = C.c_void_p(), C.c_void_p()
    _check(lib().orc_build_reduced(x.ctypes.data, x.shape[0], x.shape[1], C.byref(gp), C.byref(sw), C.byref(red),
                                   C.byref(hs), C.byref(hg), C.byref(info), cent.ctypes.data, cap))
    return Space(hs, sw), Graph(hg, gp), cent[:info.n_clusters].copy(), _info_dict(info)


def graph_from_nodes(nodes, graph_params, **switch_kw):
    """Graph (A2-A7) whose nodes are the ROWS of `nodes` (m x d)."""
    nt = np.ascontiguousarray(_f64(nodes, 2).T)
    gp = resolve_params(graph_params)
    sw = make_switches(**switch_kw)
    hg = C.c_void_p()
    _check(lib().orc_graph_from_nodes_t(nt.ctypes.data, nt.shape[0], nt.shape[1], C.byref(gp),
                                        C.byref(sw), C.byref(hg)))
    return Graph(hg, gp)


def gram_columns(x):
    x = _f64(x, 2)
    out = np.empty((x.shape[1], x.shape[1]))
    lib().orc_gram_columns(x.ctypes.data, x.shape[0], x.shape[1], out.ctypes.data)
    return out


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))
