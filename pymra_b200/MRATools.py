"""Host-side helpers with the reference's names and semantics (pyMRA/MRATools.py), restricted to
what the hot path and its configurations need: distance, the two covariance families the device
evaluates, and the grid generators used to build the benchmark inputs.

These run on the host (NumPy) and exist so that user code written against
`pyMRA.MRATools` keeps working (`cov = lambda a, b: mt.Matern32(a, b, l=kappa, sig=sig)`): the
closure is never called on the hot path, it is probed once and turned into a device covariance
descriptor (pymra_b200.covariance.introspect).
"""
import numpy as np
from scipy.spatial.distance import cdist, pdist, squareform


def genLocations(NGrid, lb=0, ub=1, random=False):
    """1-D locations, (NGrid, 1) (MRATools.py:176-183)."""
    if random:
        pts = np.random.uniform(lb, ub, NGrid)
    else:
        pts = np.linspace(lb, ub, num=NGrid + 1)[1:]
    return pts.reshape((NGrid, 1))


def genLocations2d(Nx, lbx=0, ubx=1, Ny=0, lby=0, uby=1):
    """2-D grid, x fastest, (Nx*Ny, 2) (MRATools.py:188-201)."""
    if not Ny:
        Ny = Nx
    gx, gy = np.meshgrid(np.linspace(lbx, ubx, num=Nx), np.linspace(lby, uby, num=Ny))
    return np.column_stack((gx.ravel(), gy.ravel()))


def dist(locs, locs2=np.array([]), circular=False):
    """Euclidean distance matrix as np.matrix (MRATools.py:229-245).  circular is not supported."""
    if circular:
        raise NotImplementedError("circular distances are outside the accelerated path")
    locs = locs if np.ndim(locs) == 2 else np.reshape(locs, [len(locs), 1])
    if len(locs2):
        locs2 = locs2 if np.ndim(locs2) == 2 else np.reshape(locs2, [len(locs2), 1])
        return np.matrix(cdist(locs, locs2))
    return np.matrix(squareform(pdist(locs)))


def ExpCovFun(locs, locs2=np.array([]), l=1, circular=False):
    """exp(-D/l) (MRATools.py:265-269)."""
    return np.exp(-dist(locs, locs2, circular) / l)


def Matern32(locs, locs2=np.array([]), l=1, sig=1, circular=False):
    """sig*(1+sqrt(3)D/l)*exp(-sqrt(3)D/l) (MRATools.py:289-293)."""
    D = dist(locs, locs2, circular)
    return np.matrix(sig * np.multiply(1 + np.sqrt(3) * D / l, np.exp(-np.sqrt(3) * D / l)))


def Matern52(locs, locs2=np.array([]), l=1, sig=1, circular=False):
    """sig*(1+sqrt(5)D/l+(5/3)(D/l)^2)*exp(-sqrt(5)D/l) (MRATools.py:281-285)."""
    D = dist(locs, locs2, circular)
    return np.matrix(sig * np.multiply(1 + np.sqrt(5) * D / l + (5 / 3) * np.square(D / l), np.exp(-np.sqrt(5) * D / l)))


def GaussianCovFun(locs, locs2=np.array([]), l=1, sig=1, circular=False):
    """sig*exp(-D^2/(2 l^2)) (MRATools.py:297-301)."""
    D = dist(locs, locs2, circular)
    return np.matrix(sig * np.exp(-np.square(D) / (2 * (l ** 2))))
