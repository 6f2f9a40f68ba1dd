"""TEST INFRASTRUCTURE: writes tests/golden/truth/<fixture>.npz -- the extended-precision dense ground truth
(oracle/dense_truth.c) of every reference fixture in tests/golden/, on the tree the oracle records for the
fixture's seed (bit-exact with the reference's tree, tests/test_oracle_golden.py).

    python oracle/make_truth.py [fixture ...]

Precision: __float128 for N <= 2600 locations, x87 long double above (the two agree to 1e-15 wherever both
were run: profiles/r04_parity_table.md).  Takes about ten minutes for all fixtures on 8 cores.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from _util import dense_recipe, golden_names, load_golden, oracle_for  # noqa: E402
from oracle.dense_truth import dense_truth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "truth")


def truth_for(g, precision=None):
    o = oracle_for(g, record=True)
    N = len(g["locs"])
    if precision is None:
        precision = "q" if N <= 2600 else "l"
    cm = dense_recipe(g["locs"], float(g["l"]), float(g["sig"])) if str(g["family"]) == "dense" else None
    t = dense_truth(g["locs"], g["obs"], float(g["R"]), str(g["family"]), float(g["l"]), float(g["sig"]),
                    o["nodes"], precision, cov_matrix=cm)
    return t, o


def main():
    os.makedirs(OUT, exist_ok=True)
    for name in sys.argv[1:] or golden_names():
        g = load_golden(name)
        t0 = time.time()
        t, _ = truth_for(g)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), lik=t["lik"], mean=t["mean"], sd=t["sd"],
                            precision=t["precision"])
        print("%-24s N=%6d precision=%s %.1f s" % (name, len(g["locs"]), t["precision"], time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
