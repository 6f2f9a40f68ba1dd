"""Import stub for the un-vendored, unpinned `numpy_indexed` dependency of the
reference (setup.py:11).  TEST INFRASTRUCTURE ONLY: it lets oracle/make_golden.py
import the unmodified reference from /root/reference in the build container.

Only `contains(this, that)` is used by the reference (MRANode.py:45,53,83,187,
203,264,312).  Published semantics: boolean mask over `that`, True where the
element (1-D) or row (2-D) of `that` occurs, bit for bit, in `this`.
"""
import numpy as np


def _as_keys(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if a.ndim <= 1:
        return a.reshape(-1)
    if a.shape[1] == 1:
        return a.reshape(-1)
    return a.view(np.dtype((np.void, a.dtype.itemsize * a.shape[1]))).reshape(-1)


def contains(this, that):
    this = _as_keys(this)
    that = _as_keys(that)
    if this.size == 0:
        return np.zeros(that.shape[0], dtype=bool)
    return np.isin(that, this)
