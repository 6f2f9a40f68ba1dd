"""Host-side tree/knot/partition builder: reproduces the reference's indexing bit-exactly
and flattens it into the arrays the device library consumes (`mra_set_structure`).

Reference behaviour mirrored (file:line, all under /root/reference/pyMRA):
  * Node recursion order (DFS pre-order, RNG consumed in that order)   MRANode.py:23-98
  * leaf rule: levelsFromLeaves == 0 or len(notKnots) <= max(r, J)     MRANode.py:34-47
  * knots: 1-D percentiles / np.random.choice / KMeans                 MRANode.py:179-205
  * splits: quadrants by <=/> column means (N > 100), 1-D percentiles,
    KMeans or knots-on-boundaries for N <= 100                         MRANode.py:213-242, 289-340
  * fork at res == critDepth: every child starts from the same global
    NumPy RNG state and the parent's state is not advanced             MRANode.py:64-65, 90-104

Output layout ("tree order"): locations are permuted so that every node, at every
level, owns one contiguous row range; a node's range is the concatenation of its
children's ranges in child order, followed by rows that fall in no child (1-D
strict-inequality ties, MRANode.py:224-226), which are wrapped into an "orphan"
pseudo-leaf.  Nodes are numbered level by level (BFS), children of a node are
consecutive.  Inside a leaf rows keep ascending original index, as in the reference.
"""
import numpy as np

KIND_INTERNAL = 0
KIND_LEAF = 1
KIND_ORPHAN = 2      # rows of an internal node that belong to no child: no data, no residual term


class TreeStructure(object):
    """Flat description of one MRA tree (host arrays, see module docstring)."""

    def __init__(self):
        self.N = 0
        self.d = 0
        self.r = 0
        self.J = 0
        self.M = 0                  # effective number of levels below the root (MRATree.M)
        self.depth = 0              # deepest level actually present
        self.perm = None            # (N,) int64   tree position -> original index
        self.node_id = []           # reference ID strings ('r', 'r1', ...), BFS order
        self.node_level = None      # (n_nodes,) int32
        self.node_parent = None     # (n_nodes,) int32, -1 for the root
        self.node_kind = None       # (n_nodes,) int32 KIND_*
        self.node_row_start = None  # (n_nodes,) int64
        self.node_row_count = None  # (n_nodes,) int64
        self.node_child_start = None  # (n_nodes,) int32 (BFS id of first child, -1 if none)
        self.node_child_count = None  # (n_nodes,) int32
        self.node_knot_off = None   # (n_nodes,) int64 offset into knot_rows, -1 for leaves
        self.knot_rows = None       # (n_internal*r,) int64 tree-order row ids, reference knot order
        self.level_off = None       # (depth+2,) int32 node ranges per level
        self.node_kinds_local = []  # per node: local knot ids exactly as the reference's kInds

    @property
    def n_nodes(self):
        return len(self.node_level)

    def nodes_at(self, level):
        return range(int(self.level_off[level]), int(self.level_off[level + 1]))


class StructureError(ValueError):
    pass


def _choose_knots(X, d, r, nk_local):
    if d == 1:
        vals = X[nk_local, 0]
        picked = [np.percentile(vals, 100.0 * i / (r + 1), method="nearest") for i in range(r + 2)][1:-1]
        return np.flatnonzero(np.isin(X[:, 0], picked))
    if len(nk_local) > 1e2:
        idx = np.random.choice(np.arange(len(nk_local)), size=r, replace=False)
        return np.sort(nk_local[idx])
    from scipy.spatial.distance import cdist
    from sklearn.cluster import KMeans
    cand = X[nk_local]
    km = KMeans(n_clusters=r, random_state=0).fit(cand)
    D = cdist(cand, km.cluster_centers_)
    return np.unique(nk_local[np.argmin(D, axis=0)])


def _partition_large(X, d):
    if d == 1:
        p = np.percentile(X, (33, 66))
        x = X[:, 0]
        return [np.flatnonzero(x < p[0]), np.flatnonzero((x > p[0]) & (x < p[1])), np.flatnonzero(x > p[1])]
    mu = np.mean(X, axis=0)          # NumPy reduces axis 0 of a C-contiguous (n,2) array sequentially
    lx = X[:, 0] <= mu[0]
    ly = X[:, 1] <= mu[1]
    return [np.flatnonzero(lx & ly), np.flatnonzero(lx & ~ly), np.flatnonzero(~lx & ly), np.flatnonzero(~lx & ~ly)]


def _partition_small(X, d, J, kinds, rest):
    N = len(X)
    rk = len(kinds)
    if J == rk + 1 and d == 1 and N >= J + rk:
        return np.split(np.arange(N), kinds)
    from scipy.spatial.distance import cdist
    from sklearn.cluster import KMeans
    km = KMeans(n_clusters=min(J, len(rest)), random_state=0).fit(X[rest])
    used = np.setdiff1d(np.arange(N), rest)
    klab = np.argmin(cdist(X[used], km.cluster_centers_), axis=1)
    parts = []
    for j in range(J):
        ids = np.sort(np.concatenate((used[klab == j], rest[km.labels_ == j])))
        if len(ids):
            parts.append(ids)
    if d == 1:
        parts.sort(key=lambda a: a.min())
    return parts


def build_structure(locs, r, M, J, critDepth):
    """Build the tree for already-resolved (M, J, critDepth) (see MRATree.__init__).

    Consumes the global NumPy RNG exactly like the reference constructor would.
    """
    locs = np.ascontiguousarray(locs, dtype=np.float64)
    if locs.ndim != 2:
        raise StructureError("locs must be (N, d)")
    N, d = locs.shape
    if d not in (1, 2):
        raise StructureError("only 1-D and 2-D locations are supported (got d=%d)" % d)

    # DFS pass: record nodes in pre-order
    rec = []           # dict per node
    perm_chunks = []

    def visit(ID, parent_idx, rows, nk_local, levels_left, pos):
        """rows: global ids; returns number of rows placed (== len(rows))."""
        me = len(rec)
        level = len(ID) - 1
        node = dict(ID=ID, parent=parent_idx, level=level, row_start=pos, row_count=len(rows),
                    children=[], kind=KIND_LEAF, kinds=None, knot_global=None)
        rec.append(node)
        internal = bool(levels_left) and len(nk_local) > max(r, J)
        if not internal:
            node["kinds"] = nk_local
            perm_chunks.append(rows)
            return
        X = locs[rows]
        kinds = _choose_knots(X, d, r, nk_local)
        if len(kinds) != r:
            # the reference crashes here as well (MRANode.py:492 assumes len(kInds) is the same
            # on the whole root->leaf path)
            raise StructureError("node %s selected %d distinct knots instead of r=%d" % (ID, len(kinds), r))
        node["kind"] = KIND_INTERNAL
        node["kinds"] = kinds
        node["knot_global"] = rows[kinds]
        rest = np.setdiff1d(nk_local, kinds, assume_unique=True)
        if len(rows) > 1e2:
            parts = _partition_large(X, d)
        else:
            parts = _partition_small(X, d, min(J, len(rest)), kinds, rest)
        if len(parts) > 9:
            raise StructureError("more than 9 children per node (reference IDs are digit strings)")
        is_rest = np.zeros(len(rows), dtype=bool)
        is_rest[rest] = True
        fork = level == critDepth
        state = np.random.get_state() if fork else None
        covered = np.zeros(len(rows), dtype=bool)
        p = pos
        for j, part in enumerate(parts):
            if fork:
                np.random.set_state(state)
            covered[part] = True
            node["children"].append(len(rec))
            visit(ID + str(j + 1), me, rows[part], np.flatnonzero(is_rest[part]), levels_left - 1, p)
            p += len(part)
        if fork:
            np.random.set_state(state)
        if not covered.all():
            orphan_rows = rows[~covered]
            node["children"].append(len(rec))
            rec.append(dict(ID=ID + "o", parent=me, level=level + 1, row_start=p, row_count=len(orphan_rows),
                            children=[], kind=KIND_ORPHAN, kinds=np.zeros(0, dtype=np.int64), knot_global=None))
            perm_chunks.append(orphan_rows)

    import sys
    old = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old, 10000))
    try:
        visit("r", -1, np.arange(N, dtype=np.int64), np.arange(N, dtype=np.int64), M, 0)
    finally:
        sys.setrecursionlimit(old)

    perm = np.concatenate(perm_chunks) if perm_chunks else np.zeros(0, dtype=np.int64)
    if len(perm) != N:
        raise StructureError("internal error: permutation covers %d of %d rows" % (len(perm), N))
    inv = np.empty(N, dtype=np.int64)
    inv[perm] = np.arange(N, dtype=np.int64)

    # renumber BFS (stable by level; DFS pre-order within a level == tree order)
    levels = np.array([n["level"] for n in rec], dtype=np.int32)
    order = np.argsort(levels, kind="stable")
    newid = np.empty(len(rec), dtype=np.int64)
    newid[order] = np.arange(len(rec))
    st = TreeStructure()
    st.N, st.d, st.r, st.J, st.M = N, d, r, J, M
    st.depth = int(levels.max())
    st.perm = perm
    n_nodes = len(rec)
    st.node_level = levels[order]
    st.node_parent = np.full(n_nodes, -1, dtype=np.int32)
    st.node_kind = np.zeros(n_nodes, dtype=np.int32)
    st.node_row_start = np.zeros(n_nodes, dtype=np.int64)
    st.node_row_count = np.zeros(n_nodes, dtype=np.int64)
    st.node_child_start = np.full(n_nodes, -1, dtype=np.int32)
    st.node_child_count = np.zeros(n_nodes, dtype=np.int32)
    st.node_knot_off = np.full(n_nodes, -1, dtype=np.int64)
    knots = []
    koff = 0
    for new, oldi in enumerate(order):
        nd = rec[oldi]
        st.node_id.append(nd["ID"])
        st.node_kinds_local.append(np.asarray(nd["kinds"], dtype=np.int64))
        st.node_parent[new] = -1 if nd["parent"] < 0 else newid[nd["parent"]]
        st.node_kind[new] = nd["kind"]
        st.node_row_start[new] = nd["row_start"]
        st.node_row_count[new] = nd["row_count"]
        if nd["children"]:
            ch = newid[np.array(nd["children"])]
            assert np.all(np.diff(ch) == 1)
            st.node_child_start[new] = ch[0]
            st.node_child_count[new] = len(ch)
        if nd["kind"] == KIND_INTERNAL:
            st.node_knot_off[new] = koff
            knots.append(inv[nd["knot_global"]])
            koff += r
    st.knot_rows = np.concatenate(knots).astype(np.int64) if knots else np.zeros(0, dtype=np.int64)
    st.level_off = np.searchsorted(st.node_level, np.arange(st.depth + 2)).astype(np.int32)
    return st
