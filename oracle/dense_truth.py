"""TEST INFRASTRUCTURE ONLY (see oracle/mra_oracle.py's header): Python face of oracle/dense_truth.c.

    dense_truth(locs, obs, R, family, l, sig, nodes, precision="l" | "q")

`nodes` is the node record of `mra_oracle(..., record=True)["nodes"]` (ID, global rows, local knot ids): the
tree is an input, the arithmetic is the dense extended-precision posterior under the MRA-implied covariance
(SURVEY.md 0.8 / App. A).  precision "l" = x87 long double (64-bit mantissa, usable up to ~1e4 locations),
"q" = __float128 (113-bit mantissa, software: keep N below ~3000).
Only tests/, tools/ that write parity records, and oracle/make_truth.py call this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_FAM = {"exp": 0, "matern32": 1, "matern52": 2, "gaussian": 3, "dense": 4}
_libs = {}


def _lib(precision):
    if precision not in ("l", "q"):
        raise ValueError("precision must be 'l' or 'q'")
    if precision not in _libs:
        path = os.path.join(HERE, "_ref", "libmra_truth_%s.so" % precision)
        src = os.path.join(HERE, "dense_truth.c")
        if not os.path.exists(path) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(path)):
            subprocess.check_call(["sh", os.path.join(HERE, "build_truth.sh")])
        lib = C.CDLL(path)
        fn = getattr(lib, "mra_dense_truth_%s" % precision)
        fn.restype = C.c_int
        fn.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_double, C.c_double,
                       C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                       C.POINTER(C.c_double), C.c_void_p, C.c_void_p]
        _libs[precision] = fn
    return _libs[precision]


def tree_arrays(nodes):
    """Flat arrays (parents before children) from the oracle's node record."""
    order = sorted(range(len(nodes)), key=lambda i: (len(nodes[i]["ID"]), nodes[i]["ID"]))
    pos = {nodes[i]["ID"]: k for k, i in enumerate(order)}
    parent = np.array([pos.get(nodes[i]["ID"][:-1], -1) if len(nodes[i]["ID"]) > 1 else -1 for i in order], dtype=np.int32)
    rows = [np.asarray(nodes[i]["rows"], dtype=np.int32) for i in order]
    knots = [rows[k][np.asarray(nodes[i]["kInds"], dtype=np.int64)] for k, i in enumerate(order)]
    rows_off = np.zeros(len(order) + 1, dtype=np.int64)
    knots_off = np.zeros(len(order) + 1, dtype=np.int64)
    rows_off[1:] = np.cumsum([len(a) for a in rows])
    knots_off[1:] = np.cumsum([len(a) for a in knots])
    cat = lambda xs: np.ascontiguousarray(np.concatenate(xs) if xs else np.zeros(0), dtype=np.int32)
    return parent, rows_off, cat(rows), knots_off, cat(knots)


def dense_truth(locs, obs, R, family, l, sig, nodes, precision="l", cov_matrix=None):
    locs = np.ascontiguousarray(np.asarray(locs, dtype=np.float64))
    if locs.ndim == 1:
        locs = locs.reshape(-1, 1)
    N, dim = locs.shape
    obs = np.ascontiguousarray(np.asarray(obs, dtype=np.float64).reshape(N))
    parent, rows_off, rows, knots_off, knots = tree_arrays(nodes)
    lik = C.c_double()
    mean = np.empty(N)
    sd = np.empty(N)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    dense = None
    if family == "dense":
        dense = np.ascontiguousarray(np.asarray(cov_matrix, dtype=np.float64))
        assert dense.shape == (N, N)
    rc = _lib(precision)(N, dim, p(locs), p(obs), float(R), _FAM[family], float(l), float(sig),
                         p(dense) if dense is not None else None, len(parent),
                         p(parent), p(rows_off), p(rows), p(knots_off), p(knots), C.byref(lik), p(mean), p(sd))
    if rc != 0:
        raise np.linalg.LinAlgError("dense truth failed (code %d: >0 = node whose kInv is not SPD, -1 = dense "
                                    "system not SPD, -2 = out of memory)" % rc)
    return dict(lik=float(lik.value), mean=mean, sd=sd, precision=precision)
