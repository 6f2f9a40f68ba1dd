#!/usr/bin/env python
"""GPU diagnostic: compares device intermediates, phase by phase, with the NumPy blueprint
(tests/_model.py) so a wrong kernel is localised in one run.  Test infrastructure."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import warnings
warnings.filterwarnings("ignore")
from _model import model_run  # noqa: E402
from _util import errs, load_golden, tree_for, structure_for  # noqa: E402


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b)))) if a.size else 0.0


def run(name):
    g = load_golden(name)
    t0 = time.time()
    t = tree_for(g)
    st = t._structure
    r = st.r
    mod = model_run(st, g["locs"], g["obs"], str(g["family"]), float(g["l"]), float(g["sig"]), float(g["R"]), keep=True)
    lik = float(t.getLikelihood())
    msg = ["%-20s N=%d nodes=%d depth=%d" % (name, st.N, st.n_nodes, st.depth)]
    ldv = max(2, (max(st.depth, 1) * r + 1) // 2 * 2)
    V = t._debug_fetch("V", 0, st.N * ldv).reshape(st.N, ldv)
    for m in range(st.depth):
        msg.append("  V level %d rel=%.2e" % (m, relerr(V[:, m * r:(m + 1) * r], mod["V"][:, m * r:(m + 1) * r])))
    internal = [n for n in range(st.n_nodes) if st.node_kind[n] == 0]
    worst = dict(LINV=0.0, LPINV=0.0, GT=0.0)
    for n in internal[:40] + internal[-40:]:
        m = int(st.node_level[n])
        LI = t._debug_fetch("LINV", n, r * r).reshape(r, r)
        worst["LINV"] = max(worst["LINV"], relerr(LI, np.linalg.inv(mod["Lk"][n])))
        LP = t._debug_fetch("LPINV", n, r * r).reshape(r, r)
        worst["LPINV"] = max(worst["LPINV"], relerr(LP, np.linalg.inv(mod["Lp"][n])))
        GT = t._debug_fetch("GT", n, (m * r + 1) * r).reshape(m * r + 1, r)
        Gref = np.vstack((mod["G"][n].T, mod["g"][n][None, :]))
        worst["GT"] = max(worst["GT"], relerr(GT, Gref))
    msg.append("  LINV rel=%.2e LPINV rel=%.2e GT rel=%.2e" % (worst["LINV"], worst["LPINV"], worst["GT"]))
    dn = t._debug_fetch("dnode", 0, st.n_nodes)
    msg.append("  dnode rel=%.2e  d=%.12g (model %.12g) u=%.12g (model %.12g)" % (
        relerr(dn, mod["node_d"]), t._d, mod["d"], t._u, mod["u"]))
    mean, sd = t.predict()
    rl, em, es = errs(lik, mean, sd, g)
    rl2, em2, es2 = errs(lik, mean, sd, mod)
    msg.append("  vs golden: lik_rel=%.2e mean_abs=%.2e sd_rel=%.2e | vs model: %.2e %.2e %.2e  (%.2fs)" % (
        rl, em, es, rl2, em2, es2, time.time() - t0))
    print("\n".join(msg), flush=True)


if __name__ == "__main__":
    names = sys.argv[1:] or ["g48_m32", "ka1e", "g64_m3_exp", "g33x47_m32", "g64_allobs", "g64_sparse",
                             "g30_kmeans", "m0_dense", "ka4_large_serial"]
    for nm in names:
        try:
            run(nm)
        except Exception:
            print("FAILED", nm)
            traceback.print_exc()
