"""NumPy model of the device algorithm (whitened basis + dual leaf elimination).

TEST INFRASTRUCTURE: a CPU blueprint of exactly what the CUDA kernels compute on the
flat TreeStructure, phase by phase, so GPU intermediates can be compared one to one
(`mra_debug_fetch`) and the algebra can be checked against the oracle without a GPU.
It is never imported by the product.

Maths (SURVEY.md App. A rewritten):
  prior     V_m[rows n] = (C(X_n, K_n) - sum_{k<m} V_k[rows n] V_k[K_n]^T) L_n^{-T},
            L_n L_n^T = C(K_n, K_n) - sum_k V_k[K_n] V_k[K_n]^T           (= reference kInv)
            so B_m = V_m L_n^T and B k B^T = V V^T.
  leaf      S = C_res(o,o) + R I = Ls Ls^T, U = Ls^{-1} V_a[o], z = Ls^{-1} y_o
            At = U^T U, wt = U^T z, d = 2 sum log diag Ls, u = z^T z
  upward    A = sum_children At, P = I + A_mm = Lp Lp^T, G = Lp^{-1} A[m,<m], g = Lp^{-1} w_m,
            d = 2 sum log diag Lp + sum d_c, u = -g^T g + sum u_c, At = A[<m,<m] - G^T G, wt = w_<m - G^T g
  predict   leaf: Q = Ls^{-1} C_res(o, X), mean = Q^T z, var = C(0) - |V_a|^2 - colnorm^2(Q), Vt = V_a - Q^T U
            level m (bottom-up): t = Vt_m Lp^{-T}, mean += t g, var += |t|^2, Vt_<m -= t G
"""
import numpy as np
from scipy.linalg import cholesky, solve_triangular
from scipy.spatial.distance import cdist

from pymra_b200.structure import KIND_INTERNAL, KIND_LEAF, KIND_ORPHAN


def cov_fn(family, l, sig):
    if family == "exp":
        return lambda a, b: np.exp(-cdist(a, b) / l), 1.0
    if family == "matern32":
        s3 = np.sqrt(3.0)
        return (lambda a, b: sig * (1 + s3 * cdist(a, b) / l) * np.exp(-s3 * cdist(a, b) / l)), sig
    if family == "matern52":
        s5 = np.sqrt(5.0)
        return (lambda a, b: sig * (1 + s5 * cdist(a, b) / l + (5.0 / 3.0) * np.square(cdist(a, b) / l))
                * np.exp(-s5 * cdist(a, b) / l)), sig
    if family == "gaussian":
        return (lambda a, b: sig * np.exp(-np.square(cdist(a, b)) / (2 * l * l))), sig
    raise ValueError(family)


def model_run(st, locs, obs, family, l, sig, R, want_predict=True, keep=False, shard=None):
    """shard = (s, role, allreduce): emulate one rank of a sharded run (pymra_b200/shard.py): only nodes with
    role != 0 are processed, the summaries of the level-s nodes are exchanged with `allreduce(array)` (in-place
    sum over ranks) and the outputs of predict are zero outside this rank's rows and reduced the same way."""
    cov, c0 = cov_fn(family, l, sig)
    r = st.r
    X = np.asarray(locs, dtype=np.float64).reshape(st.N, st.d)[st.perm]
    y = np.asarray(obs, dtype=np.float64).reshape(-1)[st.perm]
    ncol = max(st.depth, 1) * r
    V = np.zeros((st.N, ncol))
    nn = st.n_nodes
    Lk = [None] * nn     # chol of conditional knot covariance (reference kInv)
    out = {}
    s_lvl, role, allreduce = shard if shard is not None else (0, np.ones(nn, dtype=np.int8), None)

    def rng(n):
        s = int(st.node_row_start[n])
        return slice(s, s + int(st.node_row_count[n]))

    # ---- prior, top-down
    for m in range(st.depth + 1):
        for n in st.nodes_at(m):
            if st.node_kind[n] != KIND_INTERNAL or not role[n]:
                continue
            K = st.knot_rows[st.node_knot_off[n]: st.node_knot_off[n] + r]
            VK = V[K, : m * r]
            kInv = cov(X[K], X[K]) - VK @ VK.T
            L = cholesky(kInv, lower=True)
            Lk[n] = L
            rows = rng(n)
            T = cov(X[rows], X[K]) - V[rows, : m * r] @ VK.T
            V[rows, m * r:(m + 1) * r] = solve_triangular(L, T.T, lower=True).T

    # ---- leaves (dual form)
    At = [None] * nn
    wt = [None] * nn
    dd = np.zeros(nn)
    uu = np.zeros(nn)
    leafstate = {}
    for n in range(nn):
        kind = st.node_kind[n]
        if kind == KIND_INTERNAL or not role[n]:
            continue
        m = int(st.node_level[n])
        rows = rng(n)
        Va = V[rows, : m * r]
        if kind == KIND_ORPHAN:
            At[n] = np.zeros((m * r, m * r)); wt[n] = np.zeros(m * r)
            leafstate[n] = None
            continue
        yl = y[rows]
        o = np.flatnonzero(np.isfinite(yl))
        Xl = X[rows]
        Cres = cov(Xl[o], Xl) - Va[o] @ Va.T
        S = Cres[:, o] + R * np.eye(len(o))
        if len(o):
            Ls = cholesky(S, lower=True)
            U = solve_triangular(Ls, Va[o], lower=True)
            z = solve_triangular(Ls, yl[o], lower=True)
            dd[n] = 2.0 * np.sum(np.log(np.diag(Ls)))
        else:
            Ls = np.zeros((0, 0)); U = np.zeros((0, m * r)); z = np.zeros(0)
        At[n] = U.T @ U
        wt[n] = U.T @ z
        uu[n] = float(z @ z)
        leafstate[n] = (o, Cres, Ls, U, z)

    # ---- upward
    Lp = [None] * nn
    G = [None] * nn
    g = [None] * nn
    for m in range(st.depth, -1, -1):
        if s_lvl and m == s_lvl - 1:
            # exchange: slot per level-s node = [[At, wt], [wt^T, u]] (W x W) followed by d
            lo, hi = int(st.level_off[s_lvl]), int(st.level_off[s_lvl + 1])
            Kv = s_lvl * r
            W = Kv + 1
            buf = np.zeros((hi - lo, W * W + 1))
            for c in range(lo, hi):
                if role[c] != 1:
                    continue
                S = np.zeros((W, W))
                S[:Kv, :Kv] = At[c]; S[:Kv, Kv] = wt[c]; S[Kv, :Kv] = wt[c]; S[Kv, Kv] = uu[c]
                buf[c - lo, :W * W] = S.ravel(); buf[c - lo, W * W] = dd[c]
            allreduce(buf)
            for c in range(lo, hi):
                S = buf[c - lo, :W * W].reshape(W, W)
                At[c] = S[:Kv, :Kv].copy(); wt[c] = S[:Kv, Kv].copy(); uu[c] = float(S[Kv, Kv]); dd[c] = float(buf[c - lo, W * W])
        for n in st.nodes_at(m):
            if st.node_kind[n] != KIND_INTERNAL or not role[n]:
                continue
            cs, cc = int(st.node_child_start[n]), int(st.node_child_count[n])
            A = sum(At[c] for c in range(cs, cs + cc))
            w = sum(wt[c] for c in range(cs, cs + cc))
            P = np.eye(r) + A[m * r:, m * r:]
            Lp[n] = cholesky(P, lower=True)
            G[n] = solve_triangular(Lp[n], A[m * r:, : m * r], lower=True)
            g[n] = solve_triangular(Lp[n], w[m * r:], lower=True)
            dd[n] = 2.0 * np.sum(np.log(np.diag(Lp[n]))) + sum(dd[c] for c in range(cs, cs + cc))
            uu[n] = -float(g[n] @ g[n]) + sum(uu[c] for c in range(cs, cs + cc))
            At[n] = A[: m * r, : m * r] - G[n].T @ G[n]
            wt[n] = w[: m * r] - G[n].T @ g[n]
            for c in range(cs, cs + cc):
                At[c] = None
    out["d"] = float(dd[0]); out["u"] = float(uu[0]); out["lik"] = float(dd[0] + uu[0])
    out["node_d"] = dd; out["node_u"] = uu
    if keep:
        out["V"] = V.copy(); out["Lk"] = Lk; out["Lp"] = Lp; out["G"] = G; out["g"] = g

    if not want_predict:
        return out
    # ---- predict
    mean = np.zeros(st.N)
    var = np.zeros(st.N)
    Vt = V
    emit = np.ones(st.N, dtype=bool)
    if s_lvl:
        from pymra_b200.shard import owned_rows
        emit = owned_rows(st, role)
    for n, ls in leafstate.items():
        m = int(st.node_level[n])
        rows = rng(n)
        if ls is None:
            continue
        o, Cres, Ls, U, z = ls
        Va = Vt[rows, : m * r]
        resid = c0 - np.sum(Va * Va, axis=1)
        if len(o):
            Q = solve_triangular(Ls, Cres, lower=True)
            mean[rows] = Q.T @ z
            var[rows] = resid - np.sum(Q * Q, axis=0)
            Vt[rows, : m * r] = Va - Q.T @ U
        else:
            var[rows] = resid
    for m in range(st.depth, -1, -1):
        for n in st.nodes_at(m):
            if st.node_kind[n] != KIND_INTERNAL or not role[n]:
                continue
            rows = rng(n)
            t = solve_triangular(Lp[n], Vt[rows, m * r:(m + 1) * r].T, lower=True).T
            mean[rows] += t @ g[n]
            var[rows] += np.sum(t * t, axis=1)
            Vt[rows, : m * r] -= t @ G[n]
    inv = np.empty(st.N, dtype=np.int64)
    inv[st.perm] = np.arange(st.N)
    sd = np.sqrt(np.maximum(var, 0.0))
    if s_lvl:
        mean = np.where(emit, mean, 0.0); sd = np.where(emit, sd, 0.0); var = np.where(emit, var, 0.0)
        res = np.stack((mean, sd, var))
        allreduce(res)
        mean, sd, var = res
    out["mean"] = mean[inv]
    out["sd"] = sd[inv]
    out["var"] = var[inv]
    return out
