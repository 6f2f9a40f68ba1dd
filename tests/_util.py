"""Shared helpers for the test-suite (golden loading, oracle/model runners, error metrics)."""
import os
import warnings

import numpy as np

from conftest import GOLDEN, golden_names, load_golden, load_truth  # noqa: F401

warnings.filterwarnings("ignore", category=DeprecationWarning)


def dense_recipe(locs, l, sig):
    """The matrix-only covariance of the `dense` fixtures (oracle/make_golden.py's dense_recipe, restated)."""
    import pymra_b200.MRATools as mt
    locs = np.asarray(locs, dtype=np.float64)
    s = 1.0 + 0.3 * np.sin(4.0 * locs[:, 0]) * np.cos(3.0 * locs[:, -1])
    return np.matrix(sig * (s[:, None] * s[None, :]) * np.asarray(mt.ExpCovFun(locs, locs, l=l)))


def oracle_for(g, **kw):
    from oracle.mra_oracle import mra_oracle
    if str(g["family"]) == "dense":
        kw["cov_matrix"] = dense_recipe(g["locs"], float(g["l"]), float(g["sig"]))
    np.random.seed(int(g["seed"]))
    return mra_oracle(g["locs"], int(g["r"]), str(g["family"]), float(g["l"]), float(g["sig"]), g["obs"],
                      float(g["R"]), M=int(g["M_req"]), J=int(g["J_req"]), critDepth=int(g["critDepth"]), **kw)


def structure_for(g):
    from pymra_b200.MRATree import resolve_params
    from pymra_b200.structure import build_structure
    N, d = g["locs"].shape
    M, J, cd, _ = resolve_params(N, d, int(g["r"]), int(g["M_req"]), int(g["J_req"]), int(g["critDepth"]))
    np.random.seed(int(g["seed"]))
    return build_structure(g["locs"], int(g["r"]), M, J, cd)


def make_cov(g):
    import pymra_b200.MRATools as mt
    l, sig = float(g["l"]), float(g["sig"])
    fam = str(g["family"])
    if fam == "dense":
        return dense_recipe(g["locs"], l, sig)
    if fam == "exp":
        return lambda a, b: mt.ExpCovFun(a, b, l=l)
    if fam == "matern52":
        return lambda a, b: mt.Matern52(a, b, l=l, sig=sig)
    if fam == "gaussian":
        return lambda a, b: mt.GaussianCovFun(a, b, l=l, sig=sig)
    return lambda a, b: mt.Matern32(a, b, l=l, sig=sig)


def tree_for(g):
    from pymra_b200.MRATree import MRATree
    np.random.seed(int(g["seed"]))
    return MRATree(g["locs"], int(g["r"]), make_cov(g), g["obs"], float(g["R"]), M=int(g["M_req"]),
                   J=int(g["J_req"]), critDepth=int(g["critDepth"]))


def errs(lik, mean, sd, ref):
    rl = abs(lik - float(ref["lik"])) / abs(float(ref["lik"]))
    em = float(np.max(np.abs(np.asarray(mean).ravel() - np.asarray(ref["mean"]).ravel())))
    es = float(np.max(np.abs(np.asarray(sd).ravel() - np.asarray(ref["sd"]).ravel())
                      / np.maximum(np.asarray(ref["sd"]).ravel(), 1e-300)))
    return rl, em, es


def golden_structure(g):
    ids = list(g["node_ids"])
    ro, ko = g["rows_offsets"], g["kinds_offsets"]
    return {ids[i]: (g["rows_concat"][ro[i]:ro[i + 1]], g["kinds_concat"][ko[i]:ko[i + 1]], bool(g["node_leaf"][i]))
            for i in range(len(ids))}


# ---------------------------------------------------------------------------------------------------------------
# Parity at the sizes BASELINE.json quotes: CUDA path vs the oracle port on the same seeded inputs
FULLSIZE_CASES = {
    # name: (grid side, r0, M requested, family, l, sig, R, frac_obs)  -- bench.py WORKLOADS use the same tuples
    "cfg4": (500, 16, 7, "matern32", 0.3, 1.0, 1e-2, 0.4),          # BASELINE configs[3] exactly (M 7 -> 6)
    "cfg4_exp": (500, 16, 7, "exp", 0.3, 1.0, 1e-2, 0.4),
    "g700_r16_m7": (700, 16, 7, "matern32", 0.3, 1.0, 1e-2, 0.4),   # M = 7 levels, every internal node > 100 rows
    "g700_r16_m7_exp": (700, 16, 7, "exp", 0.3, 1.0, 1e-2, 0.4),
    "cfg3": (1000, 32, 8, "matern32", 0.3, 1.0, 1e-2, 0.4),         # BASELINE configs[2] exactly (M 8 -> 7)
    "cfg3_exp": (1000, 32, 8, "exp", 0.3, 1.0, 1e-2, 0.4),
    "cfg5": (2000, 64, 10, "matern32", 0.3, 1.0, 1e-2, 0.4),        # BASELINE configs[4] (M 10 -> 7): mean / sd
    "cfg5_exp": (2000, 64, 10, "exp", 0.3, 1.0, 1e-2, 0.4),         # same tree with ExpCovFun: likelihood
}


def fullsize_inputs(n, frac_obs, seed=3):
    """Same synthetic field as bench.py's make_inputs (kept in step by tests/test_bench_contract.py)."""
    import pymra_b200.MRATools as mt
    locs = mt.genLocations2d(n)
    N = len(locs)
    rng = np.random.RandomState(seed)
    f = np.sin(5.0 * locs[:, 0]) * np.cos(3.0 * locs[:, 1]) + 0.5 * np.sin(11.0 * locs[:, 0] * locs[:, 1])
    y = f.reshape(-1, 1) + 0.3 * rng.normal(size=(N, 1))
    obs = np.full((N, 1), np.nan)
    sel = np.sort(rng.choice(N, int(frac_obs * N), replace=False))
    obs[sel] = y[sel]
    return locs, obs


def fullsize_parity(case, seed=5):
    """Runs the CUDA path (through pymra_b200.MRATree) and the oracle port on the same inputs and RNG seed.
    Returns the achieved errors and the wall times."""
    import resource
    import time

    import pymra_b200.MRATools as mt
    from oracle.mra_oracle import mra_oracle
    from pymra_b200.MRATree import MRATree
    n, r, M, family, l, sig, R, frac = FULLSIZE_CASES[case]
    locs, obs = fullsize_inputs(n, frac)
    cov = (lambda a, b: mt.ExpCovFun(a, b, l=l)) if family == "exp" else (lambda a, b: mt.Matern32(a, b, l=l, sig=sig))
    np.random.seed(seed)
    t0 = time.time()
    t = MRATree(locs, r, cov, obs, R, M=M)
    lik = float(np.asarray(t.getLikelihood()).ravel()[0])
    mean, sd = t.predict()
    mean = np.asarray(mean).ravel().copy()
    sd = np.asarray(sd).copy()
    t_gpu = time.time() - t0
    state_after_gpu = np.random.get_state()[1].copy()
    warn = t._session.warnings()
    M_eff, nodes = t.M, int(t._structure.n_nodes)
    del t
    np.random.seed(seed)
    t0 = time.time()
    o = mra_oracle(locs, r, family, l, sig, obs, R, M=M)
    t_cpu = time.time() - t0
    rl, em, es = errs(lik, mean, sd, o)
    scale = float(np.max(np.abs(o["mean"])))
    return dict(case=case, grid=n, n_locs=n * n, r0=r, M_requested=M, M_effective=M_eff, nodes=nodes, family=family,
                l=l, R=R, frac_obs=frac, lik_gpu=lik, lik_port=o["lik"], lik_rel_err=rl, mean_max_abs_err=em,
                mean_scale=scale, sd_max_rel_err=es,
                sd_abs_err_max=float(np.max(np.abs(sd - o["sd"]))), sd_min=float(np.min(sd)),
                min_var_over_c0=float(np.min(sd) ** 2 / sig), warnings=warn,
                rng_state_equal=bool(np.array_equal(state_after_gpu, np.random.get_state()[1])),
                gpu_e2e_s=t_gpu, port_s=t_cpu, port_peak_rss_gb=resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6)


# ---------------------------------------------------------------------------------------------------------------
_SEEDED_CACHE = {}


def seeded_case(dim, n, M, J, r, family, frac, l=0.3, sig=1.0, R=1e-2, seed=7):
    """Seeded synthetic case with its oracle-port result and its extended-precision dense truth (long double)."""
    key = (dim, n, M, J, r, family, frac, l, sig, R, seed)
    if key in _SEEDED_CACHE:
        return _SEEDED_CACHE[key]
    import pymra_b200.MRATools as mt
    from oracle.dense_truth import dense_truth
    from oracle.mra_oracle import mra_oracle
    locs = mt.genLocations2d(n) if dim == 2 else mt.genLocations(n)
    rng = np.random.RandomState(n + r)
    if dim == 2:
        f = np.sin(6 * locs[:, :1]) * np.cos(4 * locs[:, 1:])
    else:
        f = np.sin(9 * locs[:, :1]) + 0.5 * np.cos(31 * locs[:, :1])
    y = f + 0.2 * rng.normal(size=(len(locs), 1))
    obs = np.full_like(y, np.nan)
    sel = np.sort(rng.choice(len(locs), int(frac * len(locs)), replace=False))
    obs[sel] = y[sel]
    cov = (lambda a, b: mt.ExpCovFun(a, b, l=l)) if family == "exp" else (lambda a, b: mt.Matern32(a, b, l=l, sig=sig))
    np.random.seed(seed)
    o = mra_oracle(locs, r, family, l, sig, obs, R, M=M, J=J, record=True)
    T = dense_truth(locs, obs, R, family, l, sig, o["nodes"], "l")
    out = dict(locs=locs, obs=obs, cov=cov, R=R, oracle=o, truth=T)
    _SEEDED_CACHE[key] = out
    return out
