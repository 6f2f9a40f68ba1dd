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


def _ids_from_parents(parent, child_start, child_count):
    ids = [""] * len(parent)
    ids[0] = "r"
    for n in range(len(parent)):
        for j in range(int(child_count[n])):
            ids[int(child_start[n]) + j] = ids[n] + str(j + 1)
    return ids


def build_structure_native(locs, r, M, J, critDepth):
    """The C++ builder (csrc/mra_structure.cpp) for the large-node 2-D path; returns None (global
    RNG untouched) when the tree needs the KMeans / 1-D paths."""
    import ctypes as C

    from . import _ffi
    locs = np.ascontiguousarray(locs, dtype=np.float64)
    N, d = locs.shape
    if d != 2 or N >= 2 ** 31 or N <= 100:
        return None
    state = np.random.get_state()
    if state[0] != "MT19937":
        return None
    key = np.ascontiguousarray(state[1], dtype=np.uint32).copy()
    pos = C.c_int32(int(state[2]))
    max_nodes = sum(4 ** m for m in range(M + 1))
    if max_nodes > 50_000_000:
        return None
    i32 = lambda n: np.zeros(max(1, n), dtype=np.int32)
    i64 = lambda n: np.zeros(max(1, n), dtype=np.int64)
    lvl, par, kind, cst, ccnt, dfs = (i32(max_nodes) for _ in range(6))
    rs, rc, koff = (i64(max_nodes) for _ in range(3))
    knots = i64(max_nodes * r)
    kloc = i32(max_nodes * r)
    perm = i64(N)
    nn, depth, nk = C.c_int32(), C.c_int32(), C.c_int64()
    p32 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    p64 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
    rcode = _ffi.lib().mra_build_structure_2d(
        locs.ctypes.data_as(C.POINTER(C.c_double)), N, r, M, J, critDepth,
        key.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(pos), max_nodes, C.byref(nn), C.byref(depth),
        p32(lvl), p32(par), p32(kind), p64(rs), p64(rc), p32(cst), p32(ccnt), p64(koff), p64(knots), p32(kloc),
        C.byref(nk), p64(perm), p32(dfs))
    if rcode == 1:
        return None
    if rcode != 0:
        raise StructureError("native structure builder failed with status %d" % rcode)
    np.random.set_state((state[0], key, int(pos.value), state[3], state[4]))
    n = int(nn.value)
    st = TreeStructure()
    st.N, st.d, st.r, st.J, st.M = N, d, r, J, M
    st.depth = int(depth.value)
    st.perm = perm
    st.node_level, st.node_parent, st.node_kind = lvl[:n].copy(), par[:n].copy(), kind[:n].copy()
    st.node_row_start, st.node_row_count = rs[:n].copy(), rc[:n].copy()
    st.node_child_start, st.node_child_count = cst[:n].copy(), ccnt[:n].copy()
    st.node_knot_off = koff[:n].copy()
    st.knot_rows = knots[: int(nk.value)].copy()
    st.level_off = np.searchsorted(st.node_level, np.arange(st.depth + 2)).astype(np.int32)
    st.node_id = _LazyIds(st)
    st.node_kinds_local = _LazyKinds(st, kloc[: int(nk.value)].copy())
    return st


class StreamBuild(object):
    """Streaming native build (mra_build_stream_* in include/pymra_b200.h): the C++ builder runs on its own
    thread; `wait(0)` returns once node arrays, permutation and the root's knots are final, `wait(1 + c)` once
    the knots of the root's child subtree c are, `finish()` joins the build and advances the global NumPy RNG
    exactly like the reference constructor.  `structure` shares its arrays with the running build."""

    def __init__(self, locs, r, M, J, critDepth):
        import ctypes as C

        from . import _ffi
        self._C, self._lib = C, _ffi.lib()
        self.job = None
        locs = np.ascontiguousarray(locs, dtype=np.float64)
        N, d = locs.shape if locs.ndim == 2 else (0, 0)
        self.state = np.random.get_state()
        if d != 2 or N >= 2 ** 31 or N < 65536 or M < 1 or M > 12 or self.state[0] != "MT19937":
            return
        self._locs = locs
        self._key = np.ascontiguousarray(self.state[1], dtype=np.uint32).copy()
        self._pos = C.c_int32(int(self.state[2]))
        nn = sum(4 ** m for m in range(M + 1))
        nk = (nn - 4 ** M) * r
        st = TreeStructure()
        st.N, st.d, st.r, st.J, st.M, st.depth = N, d, r, J, M, M
        st.node_level, st.node_parent, st.node_kind = (np.zeros(nn, dtype=np.int32) for _ in range(3))
        st.node_child_start, st.node_child_count = (np.zeros(nn, dtype=np.int32) for _ in range(2))
        st.node_row_start, st.node_row_count, st.node_knot_off = (np.zeros(nn, dtype=np.int64) for _ in range(3))
        st.knot_rows = np.zeros(max(1, nk), dtype=np.int64)[:nk]
        self._kloc = np.zeros(max(1, nk), dtype=np.int32)[:nk]
        self._dfs = np.zeros(nn, dtype=np.int32)
        st.perm = np.empty(N, dtype=np.int64)
        st.level_off = np.cumsum([0] + [4 ** m for m in range(M + 1)]).astype(np.int32)
        st.node_id = _LazyIds(st)
        st.node_kinds_local = _LazyKinds(st, self._kloc)
        self.structure = st
        p32 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        p64 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
        job = C.c_void_p()
        rcode = self._lib.mra_build_stream_start(
            locs.ctypes.data_as(C.POINTER(C.c_double)), N, r, M, J, critDepth,
            self._key.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(self._pos), nn,
            p32(st.node_level), p32(st.node_parent), p32(st.node_kind), p64(st.node_row_start),
            p64(st.node_row_count), p32(st.node_child_start), p32(st.node_child_count), p64(st.node_knot_off),
            p64(st.knot_rows), p32(self._kloc), p64(st.perm), p32(self._dfs), C.byref(job))
        if rcode == 0:
            self.job = job
        elif rcode != 1:
            raise StructureError("native streaming builder failed with status %d" % rcode)

    @property
    def started(self):
        return self.job is not None

    def wait(self, event):
        """True once `event` has happened; False if the tree turned out to be outside the streaming path
        (global RNG untouched: the caller falls back to build_structure)."""
        rcode = self._lib.mra_build_stream_wait(self.job, int(event))
        if rcode == 1:
            return False
        if rcode != 0:
            raise StructureError("native streaming builder failed with status %d" % rcode)
        return True

    def finish(self):
        """Joins the build; on success installs the advanced RNG state.  Returns False for an unsupported tree."""
        if self.job is None:
            return False
        job, self.job = self.job, None
        rcode = self._lib.mra_build_stream_finish(job, None, None, None)
        if rcode == 1:
            return False
        if rcode != 0:
            raise StructureError("native streaming builder failed with status %d" % rcode)
        np.random.set_state((self.state[0], self._key, int(self._pos.value), self.state[3], self.state[4]))
        return True

    def __del__(self):
        try:
            if self.job is not None:
                job, self.job = self.job, None
                self._lib.mra_build_stream_finish(job, None, None, None)
        except Exception:
            pass


class _LazyIds(object):
    def __init__(self, st):
        self._st, self._ids = st, None

    def _get(self):
        if self._ids is None:
            self._ids = _ids_from_parents(self._st.node_parent, self._st.node_child_start, self._st.node_child_count)
        return self._ids

    def __getitem__(self, i):
        return self._get()[i]

    def __len__(self):
        return self._st.n_nodes


class _LazyKinds(object):
    """node_kinds_local for natively built trees: internal nodes from the builder's output, leaves
    (all rows that are no ancestor's knot) reconstructed on demand."""

    def __init__(self, st, kloc):
        self._st, self._kloc = st, kloc

    def __getitem__(self, n):
        st = self._st
        if st.node_kind[n] == KIND_INTERNAL:
            o = int(st.node_knot_off[n])
            return self._kloc[o:o + st.r].astype(np.int64)
        s, c = int(st.node_row_start[n]), int(st.node_row_count[n])
        isk = np.zeros(st.N, dtype=bool)
        isk[st.knot_rows] = True
        return np.flatnonzero(~isk[s:s + c]).astype(np.int64)

    def __len__(self):
        return self._st.n_nodes


def build_structure(locs, r, M, J, critDepth, native=True):
    """Build the tree for already-resolved (M, J, critDepth) (see MRATree.__init__).

    Consumes the global NumPy RNG exactly like the reference constructor would.  The native
    builder is used when the whole tree stays on the large-node 2-D path; otherwise (or with
    native=False) the NumPy builder below runs.
    """
    locs = np.ascontiguousarray(locs, dtype=np.float64)
    if native and locs.ndim == 2 and locs.shape[1] == 2:
        st = build_structure_native(locs, r, M, J, critDepth)
        if st is not None:
            return st
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
