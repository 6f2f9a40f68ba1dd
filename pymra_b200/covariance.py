"""Covariance closure -> device covariance descriptor.

The reference takes `cov` as an arbitrary Python callable `(locs1, locs2) -> np.matrix`
(pyMRA/MRANode.py:77-80, 384).  The device evaluates two families, mt.ExpCovFun and mt.Matern32
(pyMRA/MRATools.py:265-269, 289-293), possibly scaled by a constant (also Matern52 / GaussianCovFun), and the
reference's other form of `cov`, a dense N x N np.matrix (pyMRA/MRANode.py:73-75, 381-382), which is copied to
the device and looked up.  The closure is probed
numerically, its family and parameters are recovered in closed form, and the result is verified
on random location pairs; anything that does not verify raises (there is no CPU fallback).
"""
import math

import numpy as np

from ._ffi import COV_DENSE, COV_EXP, COV_GAUSSIAN, COV_MATERN32, COV_MATERN52

_S3 = math.sqrt(3.0)
_S5 = math.sqrt(5.0)
_NAMES = {COV_EXP: "exp", COV_MATERN32: "matern32", COV_MATERN52: "matern52", COV_GAUSSIAN: "gaussian",
          COV_DENSE: "dense"}
DENSE_MAX_N = 20000      # 8 N^2 bytes of device memory: 3.2 GB


class CovDescriptor(object):
    __slots__ = ("family", "l", "sig", "matrix")

    def __init__(self, family, l, sig=1.0, matrix=None):
        self.family, self.l, self.sig = int(family), float(l), float(sig)
        self.matrix = matrix          # COV_DENSE: the N x N covariance as a C-contiguous float64 ndarray

    @property
    def name(self):
        return _NAMES[self.family]

    def __call__(self, D):
        D = np.asarray(D, dtype=np.float64)
        if self.family == COV_EXP:
            return self.sig * np.exp(-D / self.l)
        if self.family == COV_GAUSSIAN:
            return self.sig * np.exp(-np.square(D) / (2 * self.l ** 2))
        if self.family == COV_MATERN52:
            t = _S5 * D / self.l
            return self.sig * ((1 + t + (5.0 / 3.0) * np.square(D / self.l)) * np.exp(-t))
        t = _S3 * D / self.l
        return self.sig * ((1 + t) * np.exp(-t))

    def __repr__(self):
        return "CovDescriptor(%s, l=%r, sig=%r)" % (self.name, self.l, self.sig)


def _eval(cov, a, b):
    out = np.asarray(cov(a, b), dtype=np.float64)
    if out.shape != (len(a), len(b)):
        raise ValueError("cov(locs1, locs2) returned shape %s, expected %s" % (out.shape, (len(a), len(b))))
    return out


def _solve_matern_t(g, order=3):
    """t > 0 with p(t) exp(-t) = g, 0 < g < 1; p = 1+t (Matern 3/2) or 1+t+t^2/3 (Matern 5/2)."""
    poly = (lambda t: 1 + t) if order == 3 else (lambda t: 1 + t + t * t / 3.0)
    lo, hi = 0.0, 1.0
    while poly(hi) * math.exp(-hi) > g:
        hi *= 2.0
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if poly(mid) * math.exp(-mid) > g:
            lo = mid
        else:
            hi = mid
    return 0.5 * (lo + hi)


def introspect(cov, d, rtol=1e-12, n_check=192, seed=12345, n_locs=None):
    """Return the CovDescriptor of `cov`, or raise ValueError / NotImplementedError."""
    if isinstance(cov, CovDescriptor):
        return cov
    if isinstance(cov, np.ndarray):      # np.matrix included: the reference slices it (MRANode.py:73-75, 381-382)
        C = np.ascontiguousarray(np.asarray(cov, dtype=np.float64))
        if C.ndim != 2 or C.shape[0] != C.shape[1]:
            raise ValueError("a covariance matrix passed as `cov` must be square (N x N over the rows of locs)")
        if n_locs is not None and C.shape[0] != n_locs:
            raise ValueError("the covariance matrix is %d x %d but there are %d locations" % (C.shape + (n_locs,)))
        if C.shape[0] > DENSE_MAX_N:
            raise NotImplementedError("a dense covariance matrix is supported up to N = %d locations" % DENSE_MAX_N)
        dg = np.diag(C)
        if not (np.all(np.isfinite(dg)) and np.all(dg > 0)):
            raise ValueError("the covariance matrix needs a positive, finite diagonal")
        return CovDescriptor(COV_DENSE, 1.0, float(np.max(dg)), matrix=C)
    if not callable(cov):
        raise TypeError("cov must be a callable (locs1, locs2) -> matrix")
    dists = np.array([0.0] + [10.0 ** e for e in np.arange(-4.0, 2.01, 0.25)])
    pts = np.zeros((len(dists), d))
    pts[:, 0] = dists
    row = _eval(cov, pts[:1], pts)[0]
    c0 = row[0]
    if not (np.isfinite(c0) and c0 > 0):
        raise ValueError("cov(x, x) must be positive and finite")
    g = row / c0
    ok = np.flatnonzero((g > 0.2) & (g < 0.8))
    if not len(ok):
        ok = np.flatnonzero((g > 1e-3) & (g < 1 - 1e-6))
    if not len(ok):
        raise ValueError("could not probe the covariance closure (no distance with 0 < c(d)/c(0) < 1)")
    k = ok[len(ok) // 2]
    cands = [CovDescriptor(COV_EXP, -dists[k] / math.log(g[k]), c0),
             CovDescriptor(COV_MATERN32, _S3 * dists[k] / _solve_matern_t(g[k], 3), c0),
             CovDescriptor(COV_MATERN52, _S5 * dists[k] / _solve_matern_t(g[k], 5), c0),
             CovDescriptor(COV_GAUSSIAN, dists[k] / math.sqrt(-2.0 * math.log(g[k])), c0)]
    rng = np.random.RandomState(seed)
    a = rng.uniform(0, 1, size=(n_check, d))
    b = rng.uniform(0, 1, size=(n_check, d)) * rng.choice([1e-3, 1e-1, 1.0, 5.0], size=(n_check, 1))
    want = np.concatenate((row, np.diag(_eval(cov, a, a + b)),
                           _eval(cov, a[:8], a[:8]).ravel()))
    Dcheck = np.concatenate((dists, np.sqrt((b * b).sum(axis=1)),
                             np.sqrt(((a[:8, None, :] - a[None, :8, :]) ** 2).sum(axis=2)).ravel()))
    errs = []
    for cand in cands:
        got = cand(Dcheck)
        err = np.max(np.abs(got - want) / c0)
        errs.append(err)
        if err <= rtol:
            return cand
    raise ValueError("cov is not an ExpCovFun / Matern32 / Matern52 / GaussianCovFun closure (relative mismatch %s); "
                     "other covariance functions are outside the accelerated path and there is no CPU fallback"
                     % ", ".join("%s %.2e" % (c.name, e) for c, e in zip(cands, errs)))
