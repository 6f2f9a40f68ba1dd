#!/usr/bin/env python
"""BASELINE.json configs[3]: Nelder-Mead MLE over the Matern-3/2 range kappa on a 2-D 500 x 500 field
(repeated getLikelihood()), r0 = 16, M = 7 (clamped to 6 by the reference's rule), J = 4.

Two objectives, both through the reference-facing API:
  * "reference semantics": every evaluation constructs a new MRATree (README.md:96-104 of the reference); knots
    are re-drawn from the global NumPy RNG each time, so the objective is noisy -- exactly like the reference;
  * "frozen tree": one MRATree, every evaluation is tree.refit(cov=...) on the same knots and device-resident
    data (pymra_b200 extension).
Prints one JSON line per mode: evaluations, wall time, evaluations/s, final kappa.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def matern32_field(n, kappa, rng):
    """Exact GP sample on the n x n grid of genLocations2d by circulant embedding (torus of 2n x 2n cells)."""
    m = 2 * n
    h = 1.0 / n
    d1 = np.minimum(np.arange(m), m - np.arange(m)) * h
    D = np.sqrt(d1[:, None] ** 2 + d1[None, :] ** 2)
    t = np.sqrt(3.0) * D / kappa
    C = (1.0 + t) * np.exp(-t)
    lam = np.fft.fft2(C).real
    lam[lam < 0] = 0.0
    z = rng.normal(size=(m, m)) + 1j * rng.normal(size=(m, m))
    f = np.fft.fft2(np.sqrt(lam / (m * m)) * z).real
    return f[:n, :n]


def main():
    import logging
    import scipy.optimize as opt
    import torch
    import pymra_b200.MRATools as mt
    from pymra_b200.MRATree import MRATree
    logging.getLogger("pymra_b200.MRATree").setLevel(logging.ERROR)
    n, r0, M, R, kappa_true = 500, 16, 7, 1e-2, 0.3
    rng = np.random.RandomState(1)
    locs = mt.genLocations2d(n)
    N = len(locs)
    field = matern32_field(n, kappa_true, rng)            # field[iy, ix] with x fastest, like genLocations2d
    y = field.reshape(-1, 1) + np.sqrt(R) * rng.normal(size=(N, 1))
    obs = np.full((N, 1), np.nan)
    sel = np.sort(rng.choice(N, int(0.4 * N), replace=False))
    obs[sel] = y[sel]

    def cov_for(kappa):
        return lambda a, b: mt.Matern32(a, b, l=float(kappa), sig=1.0)

    # warm-up (library load, allocator)
    np.random.seed(0)
    MRATree(locs, r0, cov_for(0.3), obs, R, M=M).getLikelihood()

    # ---- reference semantics: a new tree (new knots) per evaluation
    np.random.seed(5)
    count = [0]

    def objective(x):
        count[0] += 1
        if x[0] <= 0:
            return 1e300
        return float(np.asarray(MRATree(locs, r0, cov_for(x[0]), obs, R, M=M).getLikelihood()).ravel()[0])

    torch.cuda.synchronize()
    t0 = time.time()
    res = opt.minimize(objective, [0.5], method="nelder-mead", options={"xatol": 1e-3, "fatol": 1e-2, "maxfev": 120})
    dt = time.time() - t0
    print(json.dumps({"mode": "reference semantics (new MRATree per evaluation)", "evaluations": count[0],
                      "wall_s": dt, "evals_per_s": count[0] / dt, "kappa_hat": float(res.x[0]),
                      "kappa_true": kappa_true, "minus2loglik": float(res.fun)}), flush=True)

    # ---- frozen tree: refit on the same knots; once relaunching every kernel, once replaying the captured CUDA graph
    for graph in ("0", "1"):
        os.environ["PYMRA_B200_GRAPH"] = graph
        frozen_tree(MRATree, locs, r0, cov_for, obs, R, M, kappa_true, opt, torch, graph == "1")


def frozen_tree(MRATree, locs, r0, cov_for, obs, R, M, kappa_true, opt, torch, graph):
    np.random.seed(5)
    tree = MRATree(locs, r0, cov_for(0.5), obs, R, M=M)
    count = [0]

    def objective2(x):
        count[0] += 1
        if x[0] <= 0:
            return 1e300
        return float(np.asarray(tree.refit(cov=cov_for(x[0]))).ravel()[0])

    torch.cuda.synchronize()
    t0 = time.time()
    res = opt.minimize(objective2, [0.5], method="nelder-mead", options={"xatol": 1e-3, "fatol": 1e-2, "maxfev": 120})
    dt = time.time() - t0
    # steady state: 200 more evaluations around the optimum (the graph's one-off capture + instantiation is behind us)
    torch.cuda.synchronize()
    t1 = time.time()
    for i in range(200):
        tree.refit(cov=cov_for(float(res.x[0]) * (1.0 + 1e-3 * (i % 7))))
    torch.cuda.synchronize()
    steady = 200 / (time.time() - t1)
    print(json.dumps({"mode": "frozen tree (tree.refit per evaluation%s)" % (", one CUDA graph launch each" if graph else
                                                                             ", kernels relaunched"),
                      "evaluations": count[0], "wall_s": dt, "steady_state_evals_per_s": steady,
                      "evals_per_s": count[0] / dt, "kappa_hat": float(res.x[0]), "kappa_true": kappa_true,
                      "minus2loglik": float(res.fun)}), flush=True)


if __name__ == "__main__":
    main()
