"""Where the wall time of predict() goes after a streamed construction (cfg5): per-kernel CUDA-event times against the wall."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from pymra_b200.MRATree import MRATree
n, r, Mreq, family, l, sig, R, frac = bench.WORKLOADS["cfg5"]
locs, obs = bench.make_inputs(n, frac)
cov = bench.make_cov(family, l, sig)
for rep in range(3):
    np.random.seed(5)
    tree = MRATree(locs, r, cov, obs, R, M=Mreq)
    s = tree._session
    s.profile_enable(True)
    import torch
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mean, sd = tree.predict()
    t1 = time.perf_counter()
    prof = s.profile_read()
    print(json.dumps({"predict_wall_ms": (t1 - t0) * 1e3, "alloc_ms": s.timings["predict_alloc"] * 1e3, "call_ms": s.timings["predict_call"] * 1e3, "kernels_ms": {k: round(v["ms"], 3) for k, v in prof.items() if v["launches"]}}))
    del tree, mean, sd
