"""Persistence of a built (ArrowSpace, GraphLaplacian) pair -- SURVEY.md section 8(f) rank 4.

The reference's pyo3 objects cannot be pickled, so its evaluation scripts rebuild the index on every run
(/root/reference/tests/test_2_CVE_db.py:150-160: 335 s at 81k x 768).  Here a built index is one ``.npz`` file: the stored
items (optional), the per-item lambdas and left-to-right norms, the Laplacian CSR, graph_params and the switches it was
built with.  ``load`` restores device-resident handles through the C ABI (asp_space_create + asp_space_import_lambdas +
asp_graph_from_csr): no kernel of the build runs again, and searches on the restored pair are bit-identical.
"""
import ctypes as C

import numpy as np

from . import _lib

FORMAT_VERSION = 1
_SWITCH_FIELDS = [name for name, _ in _lib.Switches._fields_]


def save(path, aspace, gl, items=None):
    """Write the pair to `path` (.npz).  `items`: the matrix the space was built from (host ndarray); when omitted the
    rows are read back from the device (single-GPU spaces only)."""
    from . import api
    if aspace._grid is not None and (aspace._grid["R"] > 1 or aspace._grid["C"] > 1):
        raise ValueError("save() works on single-GPU spaces; gather the shards first")
    n, f = aspace.nitems, aspace.nfeatures
    if items is None:
        items = np.empty((n, f), dtype=np.float64)
        _lib.check(_lib.load().asp_space_items(aspace._h, items.ctypes.data))
    items = np.ascontiguousarray(items, dtype=np.float64)
    if items.shape != (n, f):
        raise ValueError("items must be the %d x %d matrix the space was built from" % (n, f))
    indptr, indices, data = gl.csr()
    sw = _lib.Switches()
    _lib.check(_lib.load().asp_graph_switches(gl._h, C.byref(sw)))
    gp = gl.graph_params
    extra = {}
    if gl.reduction is not None:                      # built with reduction=: keep the statistics and the centroid matrix
        import json
        extra = {"reduction": np.array(json.dumps(gl.reduction)), "centroids": gl.centroids()}
    np.savez(path, **extra, format_version=FORMAT_VERSION, items=items, lambdas=aspace.lambdas(), norms=aspace.norms(), indptr=indptr,
             indices=indices, data=data, nnodes=gl.nnodes,
             graph_params=np.array([gp["eps"], gp["k"], gp["topk"], gp["p"], gp["sigma"]], dtype=np.float64),
             switches=np.array([float(getattr(sw, name)) for name in _SWITCH_FIELDS], dtype=np.float64),
             switch_fields=np.array(_SWITCH_FIELDS))
    return path


def load(path, device=None):
    """-> (ArrowSpace, GraphLaplacian) restored from `path`; no build kernel runs."""
    from . import api
    z = np.load(path, allow_pickle=False)
    if int(z["format_version"]) != FORMAT_VERSION:
        raise ValueError("unsupported index file version %s" % z["format_version"])
    lib = _lib.load()
    ctx = _lib.context(device)
    items = np.ascontiguousarray(z["items"], dtype=np.float64)
    n, f = items.shape
    eps, k, topk, p, sigma = (float(v) for v in z["graph_params"])
    cgp = _lib.make_params(eps, int(k), int(topk), p, sigma)
    sw = _lib.Switches()
    for name, value in zip([str(s) for s in z["switch_fields"]], z["switches"]):
        setattr(sw, name, float(value) if name == "tau_fixed" else int(value))
    hs, hg = C.c_void_p(), C.c_void_p()
    _lib.check(lib.asp_space_create(ctx, items.ctypes.data, n, f, n, 1, 0, C.byref(hs)))
    aspace = api.ArrowSpace._wrap(hs, ctx)
    lam = np.ascontiguousarray(z["lambdas"], dtype=np.float64)
    nrm = np.ascontiguousarray(z["norms"], dtype=np.float64)
    _lib.check(lib.asp_space_import_lambdas(hs, lam.ctypes.data, nrm.ctypes.data))
    indptr = np.ascontiguousarray(z["indptr"], dtype=np.int64)
    indices = np.ascontiguousarray(z["indices"], dtype=np.int32)
    data = np.ascontiguousarray(z["data"], dtype=np.float64)
    nnodes = int(z["nnodes"])
    _lib.check(lib.asp_graph_from_csr(ctx, nnodes, len(indices), indptr.ctypes.data, indices.ctypes.data, data.ctypes.data,
                                      C.byref(cgp), C.byref(sw), 1 if nnodes == f else 0, C.byref(hg)))
    gl = api.GraphLaplacian._wrap(hg)
    if "reduction" in z.files:
        import json
        gl.reduction = json.loads(str(z["reduction"]))
        cent = np.ascontiguousarray(z["centroids"], dtype=np.float64)
        hc = C.c_void_p()
        _lib.check(lib.asp_space_create(ctx, cent.ctypes.data, cent.shape[0], f, cent.shape[0], 1, 0, C.byref(hc)))
        gl._centroids = api.ArrowSpace._wrap(hc, ctx)
    return aspace, gl
