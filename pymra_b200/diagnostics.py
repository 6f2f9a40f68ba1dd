"""Opt-in export of the per-node basis matrices.

The reference keeps, per node, the prior basis B (pyMRA/MRANode.py:384), its posterior update BTil[res]
(:486-495) and the factors kC = chol(k) (:391) and kTilC (eigen-factor of k~, :504-507), and
`MRATree.getBasisFunctionsMatrix` (pyMRA/MRATree.py:445-511) assembles them level by level into one matrix.
The reference deletes a node's children right after use (MRANode.py:108-110), so on a finished tree that
function sees the root only; here the state stays on the device, and `all_levels=True` returns what the
function is written to produce.

Nothing here is on the hot path: the device's whitened blocks (V, Linv, Lp^-1) are copied to the host through
the C ABI's test hook `mra_debug_fetch` and un-whitened with NumPy,
    B_n   = V_m[rows n] L_n^T,                 L_n  = chol(kInv_n)        (Linv_n is what the device stores)
    B~_n  = t_m[rows n] Lp_n^T L_n^T,          Lp_n = chol(I + A_mm)      (t_m: the predict pass' folded basis)
    k_n   = Linv_n^T Linv_n,   k~_n = Linv_n^T Lp_n^-T Lp_n^-1 Linv_n.
Leaves are kept in dual form on the device (no leaf basis is ever stored), so a leaf's B is evaluated here from
the caller's covariance and the ancestors' whitened rows.
"""
import numpy as np
import scipy.linalg

MAX_FETCH_DOUBLES = 1 << 28      # 2 GiB of basis slab: this is a diagnostic for small problems


class NodeBasis(object):
    """One node's matrices in the reference's conventions (rows in the node's own order = ascending original index)."""
    __slots__ = ("ID", "level", "leaf", "rows", "B", "kC", "BTil", "kTilC", "min_x")


def _cov_rows(tree, ia, ib):
    """cov(locs[ia], locs[ib]) with the caller's covariance (closure or dense matrix)."""
    c = tree._cov_closure
    if isinstance(c, np.ndarray):
        return np.asarray(c)[np.ix_(ia, ib)]
    X = np.asarray(tree.locs, dtype=np.float64).reshape(tree._N, tree.d)
    return np.asarray(c(X[ia], X[ib]), dtype=np.float64)


def export_node_bases(tree, posterior=True):
    """List of NodeBasis, level by level (BFS), for every node of the tree."""
    st, sess = tree._structure, tree._session
    if sess.shard_level:
        raise NotImplementedError("basis export needs an unsharded tree (all blocks on one device)")
    N, r = st.N, st.r
    ldv = max(2, (max(st.depth, 1) * r + 1) // 2 * 2)
    if N * ldv > MAX_FETCH_DOUBLES:
        raise MemoryError("basis export is a diagnostic for small problems (N * depth * r <= %d)" % MAX_FETCH_DOUBLES)
    sess.set_params(tree._cov, tree._R)
    sess.likelihood()                                   # V = whitened prior basis of every level
    V = sess.debug_fetch("V", count=N * ldv).reshape(N, ldv)
    T = None
    if posterior:
        sess.keep_posterior_basis(True)
        try:
            sess.predict()                              # V <- t_j for every level (level 0 included)
            T = sess.debug_fetch("V", count=N * ldv).reshape(N, ldv)
        finally:
            sess.keep_posterior_basis(False)
            sess.likelihood()                           # leave the handle as a plain evaluation left it
    perm = np.asarray(st.perm, dtype=np.int64)
    is_knot = np.zeros(N, dtype=bool)
    is_knot[np.asarray(st.knot_rows, dtype=np.int64)] = True
    obs = np.asarray(tree._obs_ref, dtype=np.float64).reshape(N)
    X = np.asarray(tree.locs, dtype=np.float64).reshape(N, tree.d)
    out = []
    for n in range(st.n_nodes):
        lv = int(st.node_level[n])
        r0, cnt = int(st.node_row_start[n]), int(st.node_row_count[n])
        pos = np.arange(r0, r0 + cnt)                   # tree positions of the node's rows
        orig = perm[pos]
        order = np.argsort(orig, kind="stable")         # the reference's local order: ascending original index
        pos, orig = pos[order], orig[order]
        nb = NodeBasis()
        nb.ID, nb.level, nb.rows = st.node_id[n], lv, orig
        nb.leaf = int(st.node_kind[n]) != 0
        nb.min_x = float(np.min(X[orig, 0]))
        nb.BTil = nb.kTilC = None
        if int(st.node_kind[n]) == 2:                   # orphan rows (1-D tie points): no basis at all
            nb.B = np.zeros((cnt, 0))
            nb.kC = np.zeros((0, 0))
            out.append(nb)
            continue
        if not nb.leaf:
            Linv = sess.debug_fetch("LINV", n, r * r).reshape(r, r)
            L = scipy.linalg.solve_triangular(Linv, np.eye(r), lower=True)
            nb.B = V[pos, lv * r:(lv + 1) * r] @ L.T
            k = Linv.T @ Linv
            nb.kC = np.linalg.cholesky(k)
            if posterior:
                Lpinv = sess.debug_fetch("LPINV", n, r * r).reshape(r, r)
                Lp = scipy.linalg.solve_triangular(Lpinv, np.eye(r), lower=True)
                nb.BTil = T[pos, lv * r:(lv + 1) * r] @ Lp.T @ L.T
                kTil = Linv.T @ (Lpinv.T @ Lpinv) @ Linv
                W, U = np.linalg.eigh(kTil)
                nb.kTilC = U * np.sqrt(np.abs(W))[None, :]
        else:
            kpos = pos[~is_knot[pos]]                   # leaf knots: every location no ancestor took (MRANode.py:41-45)
            korig = perm[kpos]
            Va, Vk = V[pos, :lv * r], V[kpos, :lv * r]
            nb.B = _cov_rows(tree, orig, korig) - Va @ Vk.T
            kInv = nb.B[np.searchsorted(orig, korig), :]
            k = np.linalg.inv(kInv)
            nb.kC = np.linalg.cholesky(k)
            if posterior:
                nb.BTil = nb.B                          # MRANode.py:488-489: a leaf's BTil is its B
                fin = np.isfinite(obs[orig])
                A = nb.B[fin].T @ nb.B[fin] / tree._R
                kTil = np.linalg.inv(kInv + A)
                W, U = np.linalg.eigh(kTil)
                nb.kTilC = U * np.sqrt(np.abs(W))[None, :]
        out.append(nb)
    return out


def basis_functions_matrix(tree, distr="prior", groupByResolution=False, order="root", timesKC=False,
                           all_levels=False):
    """pyMRA/MRATree.py:445-511.  all_levels=False reproduces what the reference returns on a finished tree (its
    children are deleted, so only the root's block is left); all_levels=True assembles every resolution."""
    if distr not in ("prior", "posterior"):
        raise ValueError("distr must be 'prior' or 'posterior'")
    if order not in ("root", "leaves"):
        raise ValueError("order must be 'root' or 'leaves'")
    nodes = export_node_bases(tree, posterior=(distr == "posterior"))

    def block(nb):
        M = nb.B if distr == "prior" else nb.BTil
        if timesKC:
            M = M @ (nb.kC if distr == "prior" else nb.kTilC)
        return np.asarray(M)

    root = nodes[0]
    B = block(root)
    if order == "leaves" and all_levels:
        leaf_order = np.concatenate([nb.rows for nb in _leaves_dfs(tree, nodes)])
        B = B[np.searchsorted(root.rows, leaf_order), :]
    B = np.matrix(B)
    if not all_levels:
        return [B] if groupByResolution else B
    out = [B] if groupByResolution else B
    depth = max(nb.level for nb in nodes)
    for lv in range(1, depth + 1):
        m_nodes = [nb for nb in nodes if nb.level == lv]
        blocks = [block(nb) for nb in m_nodes]
        if order == "root":
            idx = np.argsort(np.array([nb.min_x for nb in m_nodes]))          # MRATree.py:476-477
            blocks = [blocks[i] for i in idx]
        Bm = np.matrix(scipy.linalg.block_diag(*blocks))
        if groupByResolution:
            out.append(Bm)
        else:
            out = np.hstack((out, Bm))
    return out


def _leaves_dfs(tree, nodes):
    st = tree._structure
    by_id = {nb.ID: nb for nb in nodes}
    ids = sorted((i for i in by_id if by_id[i].leaf), key=lambda s: [int(c) for c in s[1:]] if len(s) > 1 else [])
    return [by_id[i] for i in ids]
