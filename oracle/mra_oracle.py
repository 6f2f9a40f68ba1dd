"""CPU oracle: NumPy restatement of pyMRA's MRATree hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in `pymra_b200/` may import this module; it is
the checker for the CUDA path (tests/, __graft_entry__.smoke(), bench.py's
cpu_baseline / --impl reference legs), never the thing shipped or measured as
the product.

Parity status: PINNED.  The reference ships no runnable tests for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the unmodified
reference itself, generated in the build container by oracle/make_golden.py and
committed under tests/golden/ (tests/test_oracle_golden.py checks every fixture:
tree structure bit-exact, likelihood / mean / sd to FP64 round-off).

What is restated (reference file:line):
  * parameter resolution, M clamp              pyMRA/MRATree.py:23-59
  * DFS recursion per node                     pyMRA/MRANode.py:23-115
  * knot selection (percentile/random/KMeans)  pyMRA/MRANode.py:179-205
  * partitioning                               pyMRA/MRANode.py:213-242, 289-340
  * prior  B, kInv, k                          pyMRA/MRANode.py:73-80, 378-395
  * posterior A, omega, kTil, d, u, ATil, BTil, mean, var
                                               pyMRA/MRANode.py:403-520
  * covariance functions                       pyMRA/MRATools.py:229-245, 265-269, 289-293
  * fork-at-critDepth RNG semantics            pyMRA/MRANode.py:64-65, 90-104

Deliberate differences, all value-preserving to round-off:
  * the child covariance is not a nested closure (MRANode.py:80 re-evaluates the
    base kernel 3^m times); the residual covariance is evaluated from the stored
    ancestor B matrices.  Same numbers to ~1e-12.
  * knot membership is tracked by row index instead of by bitwise row value
    (numpy_indexed.contains); identical whenever locations are unique, which the
    reference needs anyway.
  * non-leaf log-determinants use slogdet by default; logdet="det" reproduces the
    reference's log(det(.)) (MRANode.py:463) including its under/overflow.
  * no gc.collect per node (MRANode.py:111), no pickling through pipes: forked
    children are emulated by saving/restoring the global NumPy RNG state.
"""
import math

import numpy as np
from scipy.spatial.distance import cdist

__all__ = ["exp_cov", "matern32_cov", "resolve_params", "mra_oracle"]


# ----------------------------------------------------------------------------- covariance
def _as2d(x):
    x = np.asarray(x, dtype=np.float64)
    return x if x.ndim == 2 else x.reshape(len(x), 1)


def exp_cov(a, b, l=1.0):
    """exp(-D/l), D Euclidean (MRATools.py:265-269 with dist at :229-245)."""
    return np.exp(-cdist(_as2d(a), _as2d(b)) / l)


def matern32_cov(a, b, l=1.0, sig=1.0):
    """sig*(1+sqrt(3)D/l)*exp(-sqrt(3)D/l) (MRATools.py:289-293); sig scales the variance."""
    D = cdist(_as2d(a), _as2d(b))
    return sig * np.multiply(1 + np.sqrt(3) * D / l, np.exp(-np.sqrt(3) * D / l))


def matern52_cov(a, b, l=1.0, sig=1.0):
    """sig*(1+sqrt(5)D/l+(5/3)(D/l)^2)*exp(-sqrt(5)D/l) (MRATools.py:281-285)."""
    D = cdist(_as2d(a), _as2d(b))
    return sig * np.multiply(1 + np.sqrt(5) * D / l + (5 / 3) * np.square(D / l), np.exp(-np.sqrt(5) * D / l))


def gaussian_cov(a, b, l=1.0, sig=1.0):
    """sig*exp(-D^2/(2 l^2)) (MRATools.py:297-301)."""
    D = cdist(_as2d(a), _as2d(b))
    return sig * np.exp(-np.square(D) / (2 * (l ** 2)))


def make_cov(family, l, sig=1.0):
    if family == "exp":
        return lambda a, b: exp_cov(a, b, l)
    if family == "matern32":
        return lambda a, b: matern32_cov(a, b, l, sig)
    if family == "matern52":
        return lambda a, b: matern52_cov(a, b, l, sig)
    if family == "gaussian":
        return lambda a, b: gaussian_cov(a, b, l, sig)
    raise ValueError("unknown covariance family %r" % (family,))


# ----------------------------------------------------------------------------- parameters
def resolve_params(N, d, r, M=-1, J=-1, critDepth=-1):
    """MRATree.py:31-59.  Returns (M, J, critDepth, clamped)."""
    if J < 0:
        if d == 2:
            J = 4
        else:
            # MRATree.py:33 is a comparison, not an assignment: J stays undefined
            raise AttributeError("'MRATree' object has no attribute 'J'")
    num = np.log(N * J / r + 1)
    denom = np.log(J)
    with np.errstate(divide="ignore"):
        q = num / denom
    maxM = int(q) - 1            # J == 1 -> inf -> OverflowError, as in the reference
    clamped = False
    if M < 0:
        M = maxM
    elif M > maxM:
        M = maxM
        clamped = True
    if critDepth < 0:
        critDepth = M + 1
    return M, J, critDepth, clamped


# ----------------------------------------------------------------------------- tree node
class _Node(object):
    __slots__ = ("ID", "res", "parent", "rows", "N", "leaf", "kInds", "B", "kInv", "k",
                 "inds", "children", "A", "kTil", "d", "u", "ATil", "omgTil", "BTil",
                 "mean", "var")


class _Ctx(object):
    pass


def _knots(ctx, node, nk_local):
    """Local knot indices of a non-leaf node (MRANode.py:179-205)."""
    X = ctx.locs[node.rows]
    r = ctx.r
    if ctx.d == 1:
        vals = X[nk_local, 0]
        picked = [np.percentile(vals, 100.0 * i / (r + 1), method="nearest") for i in range(r + 2)][1:-1]
        return np.flatnonzero(np.isin(X[:, 0], picked))
    if len(nk_local) > 1e2:
        idx = np.random.choice(np.arange(len(nk_local)), size=r, replace=False)
        return np.sort(nk_local[idx])
    from sklearn.cluster import KMeans
    cand = X[nk_local]
    km = KMeans(n_clusters=r, random_state=0).fit(cand)
    D = cdist(cand, km.cluster_centers_)
    chosen = {int(nk_local[int(np.argmin(D[:, c]))]) for c in range(r)}
    return np.array(sorted(chosen), dtype=np.int64)


def _splits_big(ctx, node):
    """N > 100 (MRANode.py:213-242)."""
    X = ctx.locs[node.rows]
    if ctx.d == 1:
        p = np.percentile(X, (33, 66))
        x = X[:, 0]
        return [np.where(x < p[0])[0], np.where(np.logical_and(x > p[0], x < p[1]))[0], np.where(x > p[1])[0]]
    mu = np.mean(X, axis=0)
    lx = X[:, 0] <= mu[0]
    ly = X[:, 1] <= mu[1]
    return [np.where(lx & ly)[0], np.where(lx & ~ly)[0], np.where(~lx & ly)[0], np.where(~lx & ~ly)[0]]


def _splits_small(ctx, node, J, nk_local):
    """N <= 100 (MRANode.py:289-340).  nk_local: local ids of the remaining not-knots."""
    X = ctx.locs[node.rows]
    N = node.N
    rk = len(node.kInds)
    if J == rk + 1 and ctx.d == 1 and N >= J + rk:
        return np.split(np.arange(N), node.kInds)
    from sklearn.cluster import KMeans
    cand = X[nk_local]
    ncl = min(J, len(nk_local))
    km = KMeans(n_clusters=ncl, random_state=0).fit(cand)
    labels = km.labels_
    used = np.setdiff1d(np.arange(N), nk_local)
    klab = np.argmin(cdist(X[used], km.cluster_centers_), axis=1)
    out = []
    for j in range(J):
        ids = np.sort(np.hstack((used[np.where(klab == j)[0]], nk_local[np.where(labels == j)[0]])))
        if len(ids):
            out.append(ids.astype(np.int64))
    if ctx.d == 1:
        out = sorted(out, key=lambda a: np.min(a))
    return out


def _rows_of_ancestor_B(node, l):
    """Rows of ancestor-l's B at this node's locations (MRANode.py:346-355)."""
    nd = node
    idx = np.arange(node.N)
    while nd.res > l:
        idx = nd.parent.inds[nd.ID][idx]
        nd = nd.parent
    return nd.B[idx, :]


def _prior(ctx, node):
    """B = residual covariance (rows x knots), kInv, k (MRANode.py:73-80, 378-395; App. A1)."""
    if ctx.dense is not None:                 # `cov` given as an np.matrix: slices of it (MRANode.py:73-75, 381-382)
        B = ctx.dense[np.ix_(node.rows, node.rows[node.kInds])]
    else:
        X = ctx.locs[node.rows]
        K = X[node.kInds]
        B = ctx.cov(X, K)
    chain = []
    anc = node.parent
    while anc is not None:
        chain.append(anc)
        anc = anc.parent
    for anc in reversed(chain):                          # root first, like the nested closures
        Ba = _rows_of_ancestor_B(node, anc.res)          # N x r_a
        B = B - Ba @ anc.k @ Ba[node.kInds, :].T
    node.B = B
    node.kInv = B[node.kInds, :]
    node.k = np.linalg.inv(node.kInv)


def _posterior(ctx, node):
    """MRANode.py:403-520."""
    m = node.res
    obs = ctx.obs[node.rows]
    R = ctx.R
    if node.leaf:
        fin = np.isfinite(obs)
        w = fin.astype(np.float64) / R
        z = np.where(fin, obs, 0.0) / R
        Bl = [_rows_of_ancestor_B(node, k) for k in range(m + 1)]
        omg = [Bl[k].T @ z for k in range(m + 1)]
        A = [[Bl[k].T @ (w[:, None] * Bl[l]) for l in range(m + 1)] for k in range(m + 1)]
    else:
        omg = [sum(ch.omgTil[k] for ch in node.children) for k in range(m + 1)]
        A = [[sum(ch.ATil[k][l] for ch in node.children) for l in range(m + 1)] for k in range(m + 1)]
    node.A = A
    kTilInv = node.kInv + A[m][m]
    kTil = np.linalg.inv(kTilInv)
    node.kTil = kTil
    quad = float(omg[m] @ kTil @ omg[m])
    if node.leaf:
        yo = obs[fin]
        node.u = -quad + float(yo @ yo) / R
        node.d = np.linalg.slogdet(kTilInv)[1] - np.linalg.slogdet(node.kInv)[1] + len(yo) * math.log(R)
    else:
        if ctx.logdet == "det":
            with np.errstate(divide="ignore", invalid="ignore"):
                node.d = float(-np.log(np.linalg.det(kTil)) - np.log(np.linalg.det(node.kInv)))
        else:
            node.d = np.linalg.slogdet(kTilInv)[1] - np.linalg.slogdet(node.kInv)[1]
        node.u = -quad
        for ch in node.children:
            node.d += ch.d
            node.u += ch.u
    node.omgTil = [omg[k] - A[k][m] @ kTil @ omg[m] for k in range(m)]
    node.ATil = [[A[k][l] - A[k][m] @ kTil @ A[m][l] for l in range(m + 1)] for k in range(m)]
    if node.leaf:
        BTil = Bl
    else:
        BTil = [np.zeros((node.N, len(node.kInds))) for _ in range(m + 1)]
        for ch in node.children:
            ci = node.inds[ch.ID]
            for k in range(m + 1):
                BTil[k][ci, :] = ch.BTil[k] - ch.BTil[ch.res] @ ch.kTil @ ch.A[ch.res][k]
    node.BTil = BTil
    W, V = np.linalg.eigh(kTil)
    half = V * np.sqrt(np.abs(W))[None, :]
    node.mean = BTil[m] @ (kTil @ omg[m])
    node.var = np.linalg.norm(BTil[m] @ half, axis=1) ** 2
    for ch in node.children:
        ci = node.inds[ch.ID]
        node.mean[ci] += ch.mean
        node.var[ci] += ch.var


def _build(ctx, parent, ID, rows, nk_local, levels_left):
    """One Node.__init__ (MRANode.py:23-115).  rows: global ids (ascending in the parent's
    order); nk_local: local ids of rows that are not knots of any ancestor."""
    node = _Node()
    node.ID = ID
    node.res = len(ID) - 1
    node.parent = parent
    node.rows = rows
    node.N = len(rows)
    node.inds = {}
    node.children = []
    splittable = bool(levels_left) and len(nk_local) > max(ctx.r, ctx.J)
    if splittable:
        node.kInds = _knots(ctx, node, nk_local)
        node.leaf = False
    else:
        node.kInds = nk_local.copy()
        node.leaf = True
    _prior(ctx, node)
    if splittable:
        rest = np.setdiff1d(nk_local, node.kInds)
        minJ = min(ctx.J, len(rest))
        parts = _splits_big(ctx, node) if node.N > 1e2 else _splits_small(ctx, node, minJ, rest)
        fork = node.res == ctx.critDepth
        state = np.random.get_state() if fork else None
        is_rest = np.zeros(node.N, dtype=bool)
        is_rest[rest] = True
        if fork and ctx.processes:
            # MRANode.py:90-104 for real: one forked process per child, every child starts from the parent's RNG
            # state (fork copies it), the finished child Node comes back pickled through a pipe
            import multiprocessing as mp
            mpc = mp.get_context("fork")
            jobs = []
            for j, part in enumerate(parts):
                chID = ID + str(j + 1)
                node.inds[chID] = part
                rx, tx = mpc.Pipe(duplex=False)
                p = mpc.Process(target=_forked_child, args=(ctx, node, chID, rows[part], np.flatnonzero(is_rest[part]),
                                                            levels_left - 1, tx, len(parts)))
                p.start()
                tx.close()
                jobs.append((p, rx))
            for p, rx in jobs:
                ch = rx.recv()
                ch.parent = node
                node.children.append(ch)
                p.join()
        else:
            for j, part in enumerate(parts):
                chID = ID + str(j + 1)
                node.inds[chID] = part
                if fork:
                    np.random.set_state(state)
                ch = _build(ctx, node, chID, rows[part], np.flatnonzero(is_rest[part]), levels_left - 1)
                node.children.append(ch)
            if fork:
                np.random.set_state(state)
    _posterior(ctx, node)
    if ctx.record is not None:
        rec = dict(ID=ID, rows=rows, kInds=node.kInds, leaf=node.leaf, d=node.d, u=node.u)
        if ctx.record_full:       # what MRATree.getBasisFunctionsMatrix reads from a node (MRATree.py:445-511)
            W, U = np.linalg.eigh(node.kTil)
            rec.update(B=node.B.copy(), BTil=np.array(node.BTil[node.res]), kC=np.linalg.cholesky(node.k),
                       kTilC=U * np.sqrt(np.abs(W))[None, :])
        ctx.record.append(rec)
    for ch in node.children:      # release the subtree, as MRANode.py:108-110 does
        ch.B = ch.BTil = ch.A = ch.ATil = None
    node.children = []
    return node


def _forked_child(ctx, parent, chID, rows, nk_local, levels_left, pipe, n_siblings):
    """Body of one forked child process (MRANode.py:92-93, 114-115): builds the subtree and pipes the Node back."""
    try:
        from threadpoolctl import threadpool_limits
        import os
        threadpool_limits(limits=max(1, (os.cpu_count() or 1) // max(1, n_siblings)))
    except Exception:
        pass
    ctx.record = None                      # per-node records do not cross the process boundary
    ch = _build(ctx, parent, chID, rows, nk_local, levels_left)
    ch.parent = None                       # the reference pickles the parent chain too; the slim Node is kinder to the CPU
    pipe.send(ch)
    pipe.close()


def mra_oracle(locs, r, family, l, sig, obs, R, M=-1, J=-1, critDepth=-1, logdet="slogdet",
               record=False, processes=False, cov_matrix=None):
    """Restatement of MRATree(locs, r, cov, obs, R, M, J, critDepth) + getLikelihood() + predict().

    Consumes the global NumPy RNG exactly like the reference.  processes=True runs the children of the
    critDepth node as forked processes like the reference (default: same RNG semantics, one process).
    Returns a dict with
    lik (= root.d + root.u, MRATree.py:82-84), mean, sd (MRATree.py:90-94), M, J and,
    if record, the node list (DFS post-order) with global rows and local knot ids.
    """
    locs = _as2d(locs)
    N, d = locs.shape
    M, J, critDepth, clamped = resolve_params(N, d, r, M, J, critDepth)
    ctx = _Ctx()
    ctx.locs = locs
    ctx.obs = np.asarray(obs, dtype=np.float64).reshape(-1)
    ctx.r, ctx.J, ctx.d, ctx.R = r, J, d, float(R)
    ctx.critDepth = critDepth
    ctx.dense = None
    if family == "dense":                     # the N x N covariance over the rows of locs instead of a closure
        ctx.dense = np.asarray(cov_matrix, dtype=np.float64)
        ctx.cov = None
    else:
        ctx.cov = make_cov(family, l, sig)
    ctx.logdet = logdet
    ctx.record = [] if record else None
    ctx.record_full = record == "full"
    ctx.processes = bool(processes)        # True: really fork one process per child at critDepth (MRANode.py:90-104)
    root = _build(ctx, None, "r", np.arange(N, dtype=np.int64), np.arange(N, dtype=np.int64), M)
    out = dict(lik=float(root.d + root.u), d=float(root.d), u=float(root.u),
               mean=np.asarray(root.mean).reshape(-1), sd=np.sqrt(root.var), M=M, J=J,
               clamped=clamped, root_kinds=root.kInds)
    if record:
        out["nodes"] = ctx.record
    return out
