#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Runs in the build container, where the reference is
mounted read-only at /root/reference; the fixtures it writes are committed so
that nothing at test/bench time has to read /root/reference.

    PYTHONPATH=oracle/ref_stubs:/root/reference python oracle/make_golden.py [case ...]

The reference is imported as-is with two import stubs (oracle/ref_stubs):
`numpy_indexed.contains` (un-vendored dependency, setup.py:11) and an empty
`matplotlib`.  `gc.collect` inside pyMRA.MRANode is stubbed (MRANode.py:111 calls
it once per node, ~0.1 s each, nothing numeric).  Per-node structure is recorded by
wrapping `Node.calculatePosterior` (MRANode.py:403), which runs exactly once per
node while `self.inds` / `self.kInds` are final.

Every fixture holds the inputs (locs, obs, R, r, M, J, critDepth, covariance
family/parameters, numpy seed used right before construction) and the reference
outputs: likelihood, predictive mean and sd, and the full tree structure
(node IDs in DFS post-order, per-node local knot indices and global row sets).
"""
import os
import sys
import time
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "ref_stubs"))
sys.path.insert(1, "/root/reference")

import pyMRA.MRANode as MN  # noqa: E402

MN.gc.collect = lambda *a, **k: 0
import pyMRA.MRATools as mt  # noqa: E402
from pyMRA.MRATree import MRATree  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
REFDATA = "/root/reference/pyMRA/data"


# --------------------------------------------------------------------------- recording
class Recorder:
    def __init__(self):
        self.nodes = {}
        self._orig = MN.Node.calculatePosterior

    def __enter__(self):
        rec = self

        def wrapped(node, obs, R):
            rec._orig(node, obs, R)
            rec.nodes[node.ID] = dict(
                N=int(node.N), leaf=bool(node.leaf),
                kInds=np.asarray(node.kInds, dtype=np.int64).copy(),
                inds={k: np.asarray(v, dtype=np.int64).copy() for k, v in node.inds.items()},
                d=float(np.asarray(node.d).ravel()[0]), u=float(np.asarray(node.u).ravel()[0]))

        MN.Node.calculatePosterior = wrapped
        return self

    def __exit__(self, *a):
        MN.Node.calculatePosterior = self._orig

    def flatten(self):
        """node list in DFS pre-order with global (original-index) row sets."""
        ids, rows, kinds, leaf, dd, uu = [], [], [], [], [], []

        def visit(ID, grows):
            nd = self.nodes[ID]
            assert nd["N"] == len(grows)
            ids.append(ID); rows.append(grows); kinds.append(nd["kInds"]); leaf.append(nd["leaf"])
            dd.append(nd["d"]); uu.append(nd["u"])
            for chID in sorted(nd["inds"], key=lambda s: int(s[-1])):
                visit(chID, grows[nd["inds"][chID]])

        visit("r", np.arange(self.nodes["r"]["N"], dtype=np.int64))
        return dict(
            node_ids=np.array(ids), node_leaf=np.array(leaf, dtype=bool),
            node_d=np.array(dd), node_u=np.array(uu),
            rows_concat=np.concatenate(rows).astype(np.int32),
            rows_offsets=np.cumsum([0] + [len(x) for x in rows]).astype(np.int64),
            kinds_concat=np.concatenate(kinds).astype(np.int32),
            kinds_offsets=np.cumsum([0] + [len(x) for x in kinds]).astype(np.int64))


def make_cov(family, l, sig):
    if family == "exp":
        return lambda a, b: mt.ExpCovFun(a, b, l=l)
    if family == "matern32":
        return lambda a, b: mt.Matern32(a, b, l=l, sig=sig)
    if family == "matern52":
        return lambda a, b: mt.Matern52(a, b, l=l, sig=sig)
    if family == "gaussian":
        return lambda a, b: mt.GaussianCovFun(a, b, l=l, sig=sig)
    raise ValueError(family)


def dense_recipe(locs, l, sig):
    """A non-stationary covariance that only exists as a matrix: sig * s_i s_j exp(-D_ij / l), s = 1 + 0.3 sin(4x) cos(3y)
    (tests/_util.py restates it; the N x N matrix itself is not stored in the fixture)."""
    s = 1.0 + 0.3 * np.sin(4.0 * locs[:, 0]) * np.cos(3.0 * locs[:, -1])
    return np.matrix(sig * (s[:, None] * s[None, :]) * np.asarray(mt.ExpCovFun(locs, locs, l=l)))


def run_case(name, locs, obs, r, R, family, l, sig=1.0, M=-1, J=-1, critDepth=-1, seed=5,
             note="", positional_M=None, record_basis=False):
    cov = dense_recipe(locs, l, sig) if family == "dense" else make_cov(family, l, sig)
    np.random.seed(seed)
    t0 = time.time()
    with Recorder() as rec:
        if positional_M is not None:      # README.md:35 style call: 6th positional binds to M
            tree = MRATree(locs, r, cov, obs, R, positional_M)
        else:
            tree = MRATree(locs, r, cov, obs, R, M=M, J=J, critDepth=critDepth)
        lik = float(np.asarray(tree.getLikelihood()).ravel()[0])
        xP, sdP = tree.predict()
        basis = {}
        if record_basis:      # MRATree.py:445-511 as the unmodified reference returns it (children are deleted: root only)
            for key, distr, kc in (("bf_prior", "prior", False), ("bf_prior_kc", "prior", True),
                                   ("bf_post", "posterior", False), ("bf_post_kc", "posterior", True)):
                basis[key] = np.asarray(tree.getBasisFunctionsMatrix(distr=distr, timesKC=kc), dtype=np.float64)
    dt = time.time() - t0
    st = rec.flatten() if "r" in rec.nodes and critDepth < 0 or critDepth > tree.M else None
    out = dict(locs=np.asarray(locs, dtype=np.float64), obs=np.asarray(obs, dtype=np.float64),
               r=r, R=R, family=family, l=l, sig=sig, M_req=(M if positional_M is None else positional_M),
               J_req=J, critDepth=critDepth, seed=seed, M_eff=tree.M, J_eff=tree.J,
               lik=lik, mean=np.asarray(xP, dtype=np.float64).reshape(-1),
               sd=np.asarray(sdP, dtype=np.float64).reshape(-1),
               root_kinds=np.asarray(tree.root.kInds, dtype=np.int64), note=note,
               ref_seconds=dt)
    if st is not None:
        out.update(st)
    out.update(basis)
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    nn = len(out["node_ids"]) if st is not None else -1
    print("%-18s N=%6d M=%d J=%d nodes=%5d lik=%.15g  (%.1fs) crc=%08x" % (
        name, len(locs), tree.M, tree.J, nn, lik, dt,
        zlib.crc32(out["mean"].tobytes())), flush=True)


# --------------------------------------------------------------------------- inputs
def readme_1d(cov_family, seed=11):
    """README.md:62-93 data recipe (np.NAN spelled np.nan)."""
    import scipy.linalg as lng
    np.random.seed(seed)
    dim_x = 100; sig = 1.0; me_scale = 1e-2; kappa = 0.3
    locs = mt.genLocations(dim_x)
    if cov_family == "matern32":
        Sig = sig * mt.Matern32(locs, l=kappa, sig=sig)
    else:
        Sig = sig * mt.ExpCovFun(locs, l=kappa)
    SigC = np.matrix(lng.cholesky(Sig))
    x_raw = np.matrix(np.random.normal(size=(locs.shape[0], 1)))
    x = SigC.T * x_raw
    eps = np.sqrt(me_scale) * np.matrix(np.random.normal(size=(locs.shape[0], 1)))
    y = x + eps
    obs_inds = np.sort(np.random.choice(dim_x, int(dim_x * 0.4), replace=False))
    y_obs = np.empty(np.shape(y)); y_obs[:] = np.nan; y_obs[obs_inds] = y[obs_inds]
    return locs, np.asarray(y_obs), me_scale, kappa


def grid_case(nx, ny, frac_obs, data_seed, smooth=True):
    """2-D grid (genLocations2d, MRATools.py:190-203) with a cheap smooth field + noise."""
    locs = mt.genLocations2d(nx, Ny=ny)
    rng = np.random.RandomState(data_seed)
    n = nx * ny
    if smooth:
        f = (np.sin(5.0 * locs[:, 0]) * np.cos(3.0 * locs[:, 1]) + 0.5 * np.sin(11.0 * locs[:, 0] * locs[:, 1]))
        y = f.reshape(-1, 1) + 0.3 * rng.normal(size=(n, 1))
    else:
        y = rng.normal(size=(n, 1))
    k = int(round(frac_obs * n))
    obs = np.full((n, 1), np.nan)
    sel = np.sort(rng.choice(n, k, replace=False))
    obs[sel] = y[sel]
    return locs, obs


CASES = {}


def case(f):
    CASES[f.__name__] = f
    return f


@case
def ka1m():
    locs, y_obs, R, kappa = readme_1d("matern32")
    run_case("ka1m", locs, y_obs, 2, R, "matern32", kappa, 1.0, M=3, J=3, critDepth=4, seed=5,
             note="SURVEY KA-1m: README 1-D example, Matern32")


@case
def ka1e():
    locs, y_obs, R, kappa = readme_1d("matern32")
    run_case("ka1e", locs, y_obs, 2, R, "exp", kappa, 1.0, M=3, J=3, critDepth=4, seed=5,
             note="SURVEY KA-1e: README 1-D example data, ExpCovFun")


@case
def ka2_small():
    locs = np.load(os.path.join(REFDATA, "small", "locs.npy"))
    y_obs = np.load(os.path.join(REFDATA, "small", "y_obs.npy")).reshape(-1, 1)
    run_case("ka2_small", locs, y_obs, 4, 1e-4, "exp", 2.0, seed=5,
             note="SURVEY KA-2: bundled small dataset, KMeans knots/splits")


@case
def ka4_large_serial():
    locs = np.load(os.path.join(REFDATA, "large", "locs.npy"))
    y_obs = np.load(os.path.join(REFDATA, "large", "y_obs.npy")).reshape(-1, 1)
    run_case("ka4_large_serial", locs, y_obs, 4, 1e-4, "exp", 2.0, seed=5,
             note="SURVEY KA-4 serial: bundled large dataset, r0=4, M auto (5)")


@case
def ka4_large_m3():
    locs = np.load(os.path.join(REFDATA, "large", "locs.npy"))
    y_obs = np.load(os.path.join(REFDATA, "large", "y_obs.npy")).reshape(-1, 1)
    run_case("ka4_large_m3", locs, y_obs, 4, 1e-4, "exp", 2.0, M=3, seed=5,
             note="bundled large dataset, r0=4, M=3 (no KMeans nodes: all internal nodes > 100 rows)")


@case
def ka4_large_crit0():
    locs = np.load(os.path.join(REFDATA, "large", "locs.npy"))
    y_obs = np.load(os.path.join(REFDATA, "large", "y_obs.npy")).reshape(-1, 1)
    run_case("ka4_large_crit0", locs, y_obs, 4, 1e-4, "exp", 2.0, critDepth=0, seed=5,
             note="SURVEY KA-4 critDepth=0: forked children share the RNG state")


@case
def g48_m32():
    locs, obs = grid_case(48, 48, 0.4, 1)
    run_case("g48_m32", locs, obs, 8, 1e-2, "matern32", 0.3, 1.0, M=2, seed=5)


@case
def g48_m52():
    locs, obs = grid_case(48, 48, 0.4, 11)
    run_case("g48_m52", locs, obs, 6, 2e-2, "matern52", 0.08, 1.5, M=2, seed=5,
             note="Matern52 (MRATools.py:281-285), SURVEY 8f.3")


@case
def g48_gauss():
    locs, obs = grid_case(48, 48, 0.5, 12)
    run_case("g48_gauss", locs, obs, 4, 5e-2, "gaussian", 0.03, 1.0, M=2, seed=5,
             note="GaussianCovFun (MRATools.py:297-301), SURVEY 8f.3")


@case
def g50_exp():
    locs, obs = grid_case(50, 50, 0.4, 2)
    run_case("g50_exp", locs, obs, 16, 1e-2, "exp", 0.3, 1.0, M=2, seed=6)


@case
def g33x47_m32():
    locs, obs = grid_case(33, 47, 0.6, 3)
    run_case("g33x47_m32", locs, obs, 5, 5e-2, "matern32", 0.5, 2.0, M=2, seed=7,
             note="odd non-square grid, sig=2")


@case
def g64_m3_exp():
    locs, obs = grid_case(64, 64, 0.4, 4)
    run_case("g64_m3_exp", locs, obs, 8, 1e-2, "exp", 0.3, 1.0, M=3, seed=8)


@case
def g64_allobs():
    locs, obs = grid_case(64, 64, 1.0, 5)
    run_case("g64_allobs", locs, obs, 12, 1e-1, "matern32", 0.2, 1.0, M=2, seed=9,
             note="every location observed")


@case
def g64_sparse():
    locs, obs = grid_case(64, 64, 0.01, 6)
    run_case("g64_sparse", locs, obs, 8, 1e-2, "exp", 0.5, 1.0, M=3, seed=10,
             note="1% observed: many leaves have zero observations")


@case
def g30_kmeans():
    locs, obs = grid_case(30, 30, 0.5, 7)
    run_case("g30_kmeans", locs, obs, 4, 1e-2, "exp", 0.4, 1.0, seed=11,
             note="M auto: deep nodes take the KMeans knot/split paths (<=100 rows)")


@case
def g96_m32_r16():
    locs, obs = grid_case(96, 96, 0.4, 8)
    run_case("g96_m32_r16", locs, obs, 16, 1e-2, "matern32", 0.3, 1.0, M=3, seed=12)


@case
def g125_m32_r16():
    locs, obs = grid_case(125, 125, 0.4, 9, smooth=False)
    run_case("g125_m32_r16", locs, obs, 16, 1e-2, "matern32", 0.3, 1.0, M=4, seed=5,
             note="cfg4-like (r0=16) at 125^2, white-noise obs")


@case
def m0_dense():
    locs, obs = grid_case(20, 20, 0.5, 10)
    run_case("m0_dense", locs, obs, 4, 1e-2, "exp", 0.3, 1.0, M=0, seed=5,
             note="M=0: root is a leaf, exact GP")


@case
def g40_dense():
    locs, obs = grid_case(40, 40, 0.4, 13)
    run_case("g40_dense", locs, obs, 6, 2e-2, "dense", 0.25, 1.3, M=2, seed=5, record_basis=True,
             note="cov given as an N x N np.matrix (MRANode.py:73-75, 381-382), SURVEY 8f.3; root basis functions recorded")


@case
def b1d_basis():
    locs, y_obs, R, kappa = readme_1d("matern32")
    run_case("b1d_basis", locs, y_obs, 2, R, "exp", kappa, 1.0, M=3, J=3, critDepth=4, seed=5, record_basis=True,
             note="ka1e inputs with getBasisFunctionsMatrix outputs recorded (MRATree.py:445-511), SURVEY 8f.4")


@case
def g32_basis():
    locs, obs = grid_case(32, 32, 0.4, 14)
    run_case("g32_basis", locs, obs, 5, 1e-2, "matern32", 0.4, 1.0, M=2, seed=5, record_basis=True,
             note="2-D case with getBasisFunctionsMatrix outputs recorded, SURVEY 8f.4")


@case
def readme_literal_small():
    locs = np.load(os.path.join(REFDATA, "small", "locs.npy"))
    y_obs = np.load(os.path.join(REFDATA, "small", "y_obs.npy")).reshape(-1, 1)
    run_case("readme_literal_small", locs, y_obs, 4, 1e-4, "exp", 2.0, positional_M=0, seed=5,
             note="README.md:35 call shape: 6th positional (critDepth=0) binds to M => M=0")


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for nm in names:
        CASES[nm]()
