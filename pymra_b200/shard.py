"""Subtree sharding plan for multi-GPU runs.

The reference parallelises by forking one process per child at `critDepth` (pyMRA/MRANode.py:90-104) and
pickling the finished child Node back (:114-115).  Here whole subtrees rooted at level `s` go to one rank
(= one GPU) each; the levels above `s` are tiny and replicated on every rank.  The only exchange is a
sum-reduction of the per-subtree summaries (A~_c, d_c) -- what a child hands to its parent at
MRANode.py:432-440 -- and, for predict(), of the zero-padded output vectors.

Pure NumPy / host logic: importable and testable without a GPU.
"""
import numpy as np

from .structure import KIND_INTERNAL

ROLE_OTHER, ROLE_MINE, ROLE_TOP, ROLE_TOP_EMIT = 0, 1, 2, 3


def choose_shard_level(st, world):
    """Smallest level s >= 1 with at least `world` nodes such that every level above it is complete enough
    to be replicated; None when the tree is too shallow to shard (callers then replicate the whole tree)."""
    if world <= 1:
        return None
    for s in range(1, st.depth + 1):
        if int(st.level_off[s + 1]) - int(st.level_off[s]) >= world:
            return s
    return None


def plan_shards(st, world, rank, shard_level=None):
    """Returns (s, role, owner): s the shard level (0 = not sharded), role an int8 array per node
    (ROLE_*), owner the rank of every level-s node (round-robin in tree order)."""
    s = choose_shard_level(st, world) if shard_level is None else shard_level
    n = st.n_nodes
    if not s:
        return 0, np.full(n, ROLE_MINE, dtype=np.int8), np.zeros(0, dtype=np.int32)
    lo, hi = int(st.level_off[s]), int(st.level_off[s + 1])
    owner = (np.arange(hi - lo) % world).astype(np.int32)
    role = np.empty(n, dtype=np.int8)
    role[:lo] = ROLE_TOP_EMIT if rank == 0 else ROLE_TOP
    role[lo:hi] = np.where(owner == rank, ROLE_MINE, ROLE_OTHER)
    parent = np.asarray(st.node_parent)
    for m in range(s + 1, st.depth + 1):          # BFS numbering: parents precede children
        a, b = int(st.level_off[m]), int(st.level_off[m + 1])
        role[a:b] = role[parent[a:b]]
    return s, role, owner


def summary_width(st, s):
    """W of the (W x W) summary block of a level-s node: s*r basis columns + the augmented column."""
    return s * st.r + 1


def owned_rows(st, role):
    """Boolean mask (tree order) of the rows whose predictions this rank emits."""
    mask = np.zeros(st.N, dtype=bool)
    level = np.asarray(st.node_level)
    s_nodes = np.flatnonzero(role == ROLE_MINE)
    if len(s_nodes) == st.n_nodes:       # not sharded
        mask[:] = True
        return mask
    kinds = np.asarray(st.node_kind)
    for n in range(st.n_nodes):
        emit = False
        if role[n] == ROLE_MINE and (st.node_parent[n] < 0 or role[st.node_parent[n]] != ROLE_MINE):
            emit = True                                   # one of my subtree roots
        elif role[n] == ROLE_TOP_EMIT and kinds[n] != KIND_INTERNAL:
            emit = True                                   # replicated top-level leaf, emitted by rank 0
        if emit:
            a = int(st.node_row_start[n])
            mask[a:a + int(st.node_row_count[n])] = True
    return mask
