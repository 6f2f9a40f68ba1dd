#!/usr/bin/env python
"""One device-resident step (likelihood pass + predict pass) for ncu.

    python tools/profile_step.py [--workload cfg5] [--warm 1]

Warm-up steps run outside the cudaProfilerStart/Stop bracket, ONE step inside it, so
`ncu --profile-from-start off` sees exactly the launches of one step.  Prints the per-kernel
CUDA-event times of that step as well (never quote a time printed under ncu).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg5")
    ap.add_argument("--warm", type=int, default=1)
    args = ap.parse_args()
    import torch
    import bench
    from pymra_b200.covariance import introspect
    from pymra_b200.MRATree import resolve_params
    from pymra_b200.session import DeviceSession
    from pymra_b200.structure import build_structure

    n, r, Mreq, family, l, sig, R, frac = bench.WORKLOADS[args.workload]
    locs, obs = bench.make_inputs(n, frac)
    desc = introspect(bench.make_cov(family, l, sig), 2)
    M, J, critDepth, _ = resolve_params(n * n, 2, r, Mreq, -1, -1)
    np.random.seed(5)
    st = build_structure(locs, r, M, J, critDepth)
    sess = DeviceSession(st, locs, obs, want_predict=True)
    sess.set_params(desc, R)
    for _ in range(args.warm):
        sess.likelihood_async()
        sess.predict_dev()
    torch.cuda.synchronize()
    sess.profile_enable(True)
    torch.cuda.profiler.start()
    sess.likelihood_async()
    sess.predict_dev()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    prof = sess.profile_read()
    d, u = sess.fetch_likelihood()
    print(json.dumps({"workload": args.workload, "likelihood": d + u, "launches": sess.launches(),
                      "kernels_ms": {k: v["ms"] for k, v in prof.items() if v["launches"]}}))


if __name__ == "__main__":
    main()
