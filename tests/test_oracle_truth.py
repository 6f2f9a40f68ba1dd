"""The extended-precision dense ground truth (oracle/dense_truth.c) against the reference's recorded outputs and
the oracle port -- CPU only.

Three independent algorithms meet here: the reference's recursion (golden vectors), the oracle's restatement of
it, and a dense Cholesky posterior under the covariance the tree implies, evaluated in long double / __float128.
Where the covariance is well conditioned (ExpCovFun fixtures) all three agree to FP64 round-off, which pins the
truth to the reference; where it is not (Matern / Gaussian, kappa = 0.3) the truth measures how far the
reference's own FP64 result is from exact arithmetic -- the yardstick tests/test_gpu_parity.py uses."""
import numpy as np
import pytest

from _util import dense_recipe, errs, golden_names, load_golden, load_truth, oracle_for

# likelihood relative, mean absolute, sd relative
TIGHT = (1e-10, 2e-9, 5e-8)      # exponential covariance: everything is round-off
LOOSE = (5e-8, 1e-5, 5e-3)       # ill-conditioned families: the reference's own inv()-noise (measured: <= 1e-8, 2.3e-6, 2.4e-3)


def _bound(g):
    fam = str(g["family"])
    if fam == "exp" and float(g["R"]) >= 1e-3 and int(g["M_eff"]) >= 1:      # M = 0 inverts the raw N x N covariance
        return TIGHT
    return LOOSE


@pytest.mark.parametrize("name", golden_names())
def test_reference_and_oracle_agree_with_truth(name):
    g = load_golden(name)
    T = load_truth(name)
    b = _bound(g)
    ref = errs(float(g["lik"]), g["mean"], g["sd"], T)
    assert all(e <= t for e, t in zip(ref, b)), ("reference vs truth", ref)
    o = oracle_for(g)
    orc = errs(o["lik"], o["mean"], o["sd"], T)
    assert all(e <= t for e, t in zip(orc, b)), ("oracle vs truth", orc)


@pytest.mark.parametrize("name", [n for n in golden_names() if len(load_golden(n)["locs"]) <= 1000])
def test_committed_truth_is_reproducible_in_both_precisions(name):
    from oracle.dense_truth import dense_truth
    g = load_golden(name)
    T = load_truth(name)
    o = oracle_for(g, record=True)
    for prec in ("l", "q"):
        cm = dense_recipe(g["locs"], float(g["l"]), float(g["sig"])) if str(g["family"]) == "dense" else None
        t = dense_truth(g["locs"], g["obs"], float(g["R"]), str(g["family"]), float(g["l"]), float(g["sig"]),
                        o["nodes"], prec, cov_matrix=cm)
        rl, em, es = errs(t["lik"], t["mean"], t["sd"], T)
        assert rl <= 1e-14 and em <= 1e-14 * max(1.0, float(np.max(np.abs(T["mean"])))) and es <= 1e-13, (prec, rl, em, es)


def test_truth_equals_exact_kriging_for_m0():
    """M = 0: one leaf holding every location, so the implied covariance is the covariance itself and the truth
    must equal textbook kriging computed in plain NumPy."""
    from oracle.dense_truth import dense_truth
    from oracle.mra_oracle import exp_cov
    rng = np.random.RandomState(2)
    locs = rng.uniform(size=(150, 2))
    y = rng.normal(size=150)
    obs = y.copy()
    obs[rng.choice(150, 60, replace=False)] = np.nan
    R, l = 0.05, 0.4
    nodes = [dict(ID="r", rows=np.arange(150), kInds=np.arange(150))]
    t = dense_truth(locs, obs, R, "exp", l, 1.0, nodes, "q")
    o = np.isfinite(obs)
    C = exp_cov(locs, locs, l)
    S = C[np.ix_(o, o)] + R * np.eye(o.sum())
    a = np.linalg.solve(S, obs[o])
    mean = C[:, o] @ a
    var = np.diag(C) - np.einsum("ij,ji->i", C[:, o], np.linalg.solve(S, C[o, :]))
    lik = np.linalg.slogdet(S)[1] + obs[o] @ a
    assert abs(t["lik"] - lik) <= 1e-11 * abs(lik)
    assert np.max(np.abs(t["mean"] - mean)) <= 1e-11
    assert np.max(np.abs(t["sd"] - np.sqrt(var))) <= 1e-10
