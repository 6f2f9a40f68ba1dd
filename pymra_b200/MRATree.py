"""Drop-in `MRATree` for the hot path of pyMRA (pyMRA/MRATree.py:20-94).

Same constructor signature and semantics as the reference class; the work the reference does in
`Node.__init__` (pyMRA/MRANode.py:23-115) is split into
  * the host structure builder (pymra_b200/structure.py; bit-exact tree/knot/partition indexing,
    consumes the global NumPy RNG like the reference), and
  * the device library behind the C ABI (include/pymra_b200.h): prior pass, leaf terms, upward
    pass for the likelihood, downward pass for predictions.
PyTorch only supplies the device arena and the stream.  There is no CPU fallback.
"""
import ctypes as C
import logging

import numpy as np

from . import _ffi
from .covariance import introspect
from .structure import build_structure

logger = logging.getLogger("pymra_b200.MRATree")


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("pymra_b200.MRATree needs a CUDA device (B200); there is no CPU fallback")
    return torch


def resolve_params(N, d, r, M, J, critDepth):
    """J default, M clamp, critDepth/mode resolution exactly as MRATree.py:31-59."""
    if J < 0:
        if d == 2:
            J = 4
        else:
            # MRATree.py:33 compares instead of assigning, so 1-D callers must pass J
            raise AttributeError("'MRATree' object has no attribute 'J'")
    num = np.log(N * J / r + 1)
    denom = np.log(J)
    with np.errstate(divide="ignore"):
        maxM = int(num / denom) - 1       # J == 1 raises OverflowError like the reference
    clamped = None
    if M < 0:
        M = maxM
    elif M > maxM:
        clamped = (M, maxM)
        M = maxM
    if critDepth < 0:
        critDepth = M + 1
    return M, J, critDepth, clamped


class _Root(object):
    """The attributes of the reference's root Node that callers touch."""

    def __init__(self, tree):
        self._tree = tree
        self.children = []
        self.ID = "r"
        self.res = 0
        self.N = tree._N
        self.leaf = tree._structure.node_kind[0] != 0
        self.kInds = tree._structure.node_kinds_local[0]
        self.d = tree._d
        self.u = np.matrix([[tree._u]])

    @property
    def mean(self):
        return self._tree._moments()[0]

    @property
    def var(self):
        return self._tree._moments()[1]


class MRATree(object):

    def __init__(self, locs, r, cov, obs, R, M=-1, J=-1, critDepth=-1, verbose=True, device=None):
        torch = _torch()
        self.locs = locs
        self.d = np.shape(self.locs)[1]
        N = len(locs)
        self._N = N
        self.r = r
        M, J, critDepth, clamped = resolve_params(N, self.d, r, M, J, critDepth)
        if clamped is not None:
            logger.warning("The number of resolutions M=%d you requested is to large for your grid. "
                           "Setting M:=%d" % clamped)
        self.M, self.J = M, J
        if not isinstance(R, float):
            raise TypeError("R (me_scale) must be a Python/NumPy float; the reference's matrix-valued R "
                            "path is broken (MRANode.py:421) and is not accelerated")
        obs_arr = np.asarray(obs, dtype=np.float64)
        if obs_arr.shape != (N, 1):
            raise ValueError("obs must have shape (N, 1) (the reference wraps it in np.matrix, MRATree.py:61)")
        self.obs_inds = np.where(np.logical_not(np.isnan(obs)))[0]
        self._cov = introspect(cov, self.d)
        self._R = float(R)
        logger.debug("r: %d, \tJ: %d,\tM: %d" % (self.r, self.J, self.M))
        logger.debug("mode: %s" % ("serial" if critDepth > self.M else "parallel"))

        locs_c = np.ascontiguousarray(locs, dtype=np.float64).reshape(N, self.d)
        self._structure = build_structure(locs_c, r, M, J, critDepth)

        self._dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self._lib = _ffi.lib()
        self._h = C.c_void_p()
        st = self._lib.mra_create(C.byref(self._h), self._dev.index)
        if st != 0:
            raise _ffi.MraError(st, "mra_create failed (no usable CUDA device?)")
        self._obs_c = np.ascontiguousarray(obs_arr.reshape(N))
        self._locs_c = locs_c
        self._set_structure()
        nbytes = C.c_size_t()
        self._check(self._lib.mra_plan(self._h, self._ptr(self._obs_c), 1, C.byref(nbytes)))
        self._ws = torch.empty(int(nbytes.value) + 256, dtype=torch.uint8, device=self._dev)
        base = self._ws.data_ptr()
        aligned = (base + 255) // 256 * 256
        self._check(self._lib.mra_bind_workspace(self._h, C.c_void_p(aligned), C.c_size_t(int(nbytes.value))))
        self._check(self._lib.mra_upload_data(self._h, self._ptr(self._locs_c), self._ptr(self._obs_c), self._stream()))
        self._mom = None
        self._evaluate()
        self.root = _Root(self)

    # ---- plumbing
    @staticmethod
    def _ptr(a):
        return a.ctypes.data_as(C.POINTER(C.c_double))

    def _stream(self):
        import torch
        return C.c_void_p(torch.cuda.current_stream(self._dev).cuda_stream)

    def _check(self, status):
        _ffi.check(self._h, status)

    def _set_structure(self):
        s = self._structure
        keep = dict(
            node_level=np.ascontiguousarray(s.node_level, dtype=np.int32),
            node_parent=np.ascontiguousarray(s.node_parent, dtype=np.int32),
            node_kind=np.ascontiguousarray(s.node_kind, dtype=np.int32),
            node_row_start=np.ascontiguousarray(s.node_row_start, dtype=np.int64),
            node_row_count=np.ascontiguousarray(s.node_row_count, dtype=np.int64),
            node_child_start=np.ascontiguousarray(s.node_child_start, dtype=np.int32),
            node_child_count=np.ascontiguousarray(s.node_child_count, dtype=np.int32),
            node_knot_off=np.ascontiguousarray(s.node_knot_off, dtype=np.int64),
            knot_rows=np.ascontiguousarray(s.knot_rows, dtype=np.int64),
            level_off=np.ascontiguousarray(s.level_off, dtype=np.int32),
            perm=np.ascontiguousarray(s.perm, dtype=np.int64))
        ms = _ffi.MraStructure()
        ms.n_locs, ms.dim, ms.r, ms.depth, ms.n_nodes = s.N, s.d, s.r, s.depth, s.n_nodes
        ms.n_knot_rows = len(keep["knot_rows"])
        for name, arr in keep.items():
            ctype = C.c_int32 if arr.dtype == np.int32 else C.c_int64
            setattr(ms, name, arr.ctypes.data_as(C.POINTER(ctype)))
        self._check(self._lib.mra_set_structure(self._h, C.byref(ms)))

    def _evaluate(self):
        self._check(self._lib.mra_set_cov(self._h, self._cov.family, self._cov.l, self._cov.sig))
        self._check(self._lib.mra_set_nugget(self._h, self._R))
        out = (C.c_double * 2)()
        self._check(self._lib.mra_run_likelihood(self._h, self._stream(), out))
        self._d, self._u = float(out[0]), float(out[1])
        self._mom = None

    def _moments(self):
        if self._mom is None:
            mean = np.empty(self._N)
            sd = np.empty(self._N)
            self._check(self._lib.mra_run_predict(self._h, self._stream(), self._ptr(mean), self._ptr(sd)))
            self._mom = (np.matrix(mean.reshape(-1, 1)), sd * sd, sd)
        return self._mom

    def _debug_fetch(self, what, node=0, count=None):
        if count is None:
            count = 1 << 24
        buf = np.empty(count)
        n = self._lib.mra_debug_fetch(self._h, what.encode(), int(node), self._ptr(buf), C.c_int64(count))
        if n < 0:
            self._check(int(n))
        return buf[:n].copy()

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.mra_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- reference API
    def getLikelihood(self):
        """root.d + root.u as a 1x1 np.matrix (MRATree.py:82-84)."""
        return np.matrix([[self._d + self._u]])

    def predict(self):
        """(root.mean as (N,1) np.matrix, sqrt(root.var) as (N,) ndarray) (MRATree.py:90-94)."""
        mean, _, sd = self._moments()
        return mean, sd

    # ---- beyond the reference: frozen-structure re-evaluation for MLE loops
    def refit(self, cov=None, R=None):
        """Re-evaluate likelihood (and, lazily, predictions) for new covariance parameters / nugget
        on the SAME tree structure and device-resident data."""
        if cov is not None:
            self._cov = introspect(cov, self.d)
        if R is not None:
            if not isinstance(R, float):
                raise TypeError("R must be a float")
            self._R = float(R)
        self._evaluate()
        self.root = _Root(self)
        return self.getLikelihood()
