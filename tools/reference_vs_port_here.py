#!/usr/bin/env python
"""Times the UNMODIFIED reference (pyMRA, /root/reference, import stubs from oracle/ref_stubs) beside the oracle
port on the inputs bench.py's `cpu_baseline` / `--impl reference` legs use (SAMPLE_GRID^2 locations, r0 = 64).
Runs only in the build container (the reference does not travel to the GPU box); the result is committed as
profiles/r03_reference_vs_port_build_container.json so that the "port" baseline can be related to the real thing.

    python tools/reference_vs_port_here.py [grid]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_stubs"))
sys.path.insert(1, "/root/reference")
import bench  # noqa: E402
from oracle.mra_oracle import mra_oracle  # noqa: E402

import pyMRA.MRANode as MN  # noqa: E402
import pyMRA.MRATools as rmt  # noqa: E402
from pyMRA.MRATree import MRATree as RefTree  # noqa: E402

grid = int(sys.argv[1]) if len(sys.argv) > 1 else bench.SAMPLE_GRID
n, r, Mreq, family, l, sig, R, frac = bench.WORKLOADS["cfg5"]
locs, obs = bench.make_inputs(grid, frac, seed=4)
N = len(locs)
out = {"grid": grid, "n_locs": N, "r0": r, "M_requested": Mreq, "cov": family, "l": l, "cores": os.cpu_count(),
       "where": "build container (no GPU); the GPU box's host is faster, the ratio is what matters"}

np.random.seed(5)
t0 = time.time()
o = mra_oracle(locs, r, family, l, sig, obs, R, M=Mreq)
out["oracle_port_s"] = round(time.time() - t0, 3)

cov = (lambda a, b: rmt.Matern32(a, b, l=l, sig=sig)) if family == "matern32" else (lambda a, b: rmt.ExpCovFun(a, b, l=l))
real_gc = MN.gc.collect
for label, gc_fn in (("reference_gc_stubbed_s", lambda *a, **k: 0), ("reference_as_is_s", real_gc)):
    MN.gc.collect = gc_fn
    np.random.seed(5)
    t0 = time.time()
    t = RefTree(locs, r, cov, obs, R, M=Mreq, verbose=False)
    lik = float(np.asarray(t.getLikelihood()).ravel()[0])
    mean, sd = t.predict()
    out[label] = round(time.time() - t0, 3)
    out["reference_lik"] = lik
    del t
MN.gc.collect = real_gc
out["oracle_lik"] = float(o["lik"])
out["port_speedup_over_reference_gc_stubbed"] = round(out["reference_gc_stubbed_s"] / out["oracle_port_s"], 2)
out["port_speedup_over_reference_as_is"] = round(out["reference_as_is_s"] / out["oracle_port_s"], 2)
print(json.dumps(out))
