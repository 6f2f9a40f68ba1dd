"""Shared helpers for the test-suite (golden loading, oracle/model runners, error metrics)."""
import os
import warnings

import numpy as np

from conftest import GOLDEN, golden_names, load_golden  # noqa: F401

warnings.filterwarnings("ignore", category=DeprecationWarning)


def oracle_for(g, **kw):
    from oracle.mra_oracle import mra_oracle
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
