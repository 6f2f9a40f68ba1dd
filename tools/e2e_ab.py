#!/usr/bin/env python
"""A/B of end-to-end construction variants inside ONE process (same box, same clocks, alternating order):

    python tools/e2e_ab.py --workload cfg5 --reps 6 --env PYMRA_B200_TWO_PART=1 --env PYMRA_B200_TWO_PART=0

Each --env is one variant ("K=V[,K=V..]"); prints the median wall time of MRATree(...) + getLikelihood() + predict() per
variant and the median of every entry of tree.timings."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg5")
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--env", action="append", default=[])
    args = ap.parse_args()
    import bench
    from pymra_b200.MRATree import MRATree
    n, r, Mreq, family, l, sig, R, frac = bench.WORKLOADS[args.workload]
    locs, obs = bench.make_inputs(n, frac)
    cov = bench.make_cov(family, l, sig)
    variants = [dict(kv.split("=", 1) for kv in v.split(",")) for v in (args.env or ["A=0"])]
    walls = [[] for _ in variants]
    tims = [[] for _ in variants]
    for rep in range(args.reps + 1):
        for vi, env in enumerate(variants):
            os.environ.update(env)
            np.random.seed(5)
            t0 = time.time()
            tree = MRATree(locs, r, cov, obs, R, M=Mreq)
            lik = float(np.asarray(tree.getLikelihood()).ravel()[0])
            mean, sd = tree.predict()
            w = time.time() - t0
            if rep > 0:
                walls[vi].append(w)
                tims[vi].append(dict(tree.timings))
            del tree, mean, sd
    for vi, env in enumerate(variants):
        keys = sorted(tims[vi][0])
        med = {k: round(float(np.median([t.get(k, 0.0) for t in tims[vi]])), 4) for k in keys}
        print(json.dumps({"variant": env, "median_s": round(float(np.median(walls[vi])), 4),
                          "min_s": round(min(walls[vi]), 4), "all_s": [round(x, 4) for x in walls[vi]], "lik": lik,
                          "timings_median": med}))


if __name__ == "__main__":
    main()
