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


# ---- one build per group: rank 0 draws the knots, everybody receives the tree --------------------------------
_I32_FIELDS = ("node_level", "node_parent", "node_kind", "node_child_start", "node_child_count",
               "node_row_start", "node_row_count", "node_knot_off")


def build_structure_group(locs, r, M, J, critDepth, group=None, async_start=None, device=None):
    """The tree for a process group.  The reference's knot draws are one sequential legacy-RNG stream
    (pyMRA/MRANode.py:191-193 in DFS pre-order), so N ranks building redundantly only fight for the host's
    memory bandwidth.  Instead group rank 0 runs the native builder with all the host threads and broadcasts
    the flat arrays (about 7 bytes per location) together with the advanced MT19937 state, which every rank
    installs -- afterwards all ranks hold the same tree and the same global NumPy RNG state, exactly as if
    each had built it.  Trees outside the native builder's path (small / 1-D / ragged) are built by every
    rank on its own as before.  Works over NCCL (device tensors) and gloo (host tensors).
    async_start: optional callable run on every rank while rank 0 builds (e.g. the H2D copy of the inputs)."""
    import os

    import torch
    import torch.distributed as dist

    from .structure import StreamBuild, TreeStructure, _LazyIds, _LazyKinds, build_structure, build_structure_native
    rank = dist.get_rank(group)
    src = dist.get_global_rank(group, 0) if group is not None else 0
    on_gpu = dist.get_backend(group) == "nccl"
    # communication buffers live on the device the session will use (MRATree's `device`), not on whatever the
    # process' current device happens to be
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device) if on_gpu else torch.device("cpu")
    locs = np.ascontiguousarray(locs, dtype=np.float64)
    N = len(locs)
    header = torch.zeros(8, dtype=torch.int64)
    st = None
    if rank == 0:
        saved = {k: os.environ.get(k) for k in ("MRA_BUILD_THREADS",)}
        ncpu = os.cpu_count() or 8
        os.environ["MRA_BUILD_THREADS"] = str(ncpu if ncpu <= 8 else min(12, ncpu - 1))   # the other ranks wait
        try:
            sb = StreamBuild(locs, r, M, J, critDepth) if locs.ndim == 2 and locs.shape[1] == 2 else None
            if async_start is not None:
                async_start()
            if sb is not None and sb.started and sb.wait(5) and sb.finish():
                st = sb.structure
            else:
                if sb is not None:
                    sb.finish()
                st = build_structure_native(locs, r, M, J, critDepth) if locs.ndim == 2 and locs.shape[1] == 2 else None
        finally:
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        if st is not None:
            state = np.random.get_state()
            header[:6] = torch.tensor([1, st.n_nodes, st.depth, len(st.knot_rows), int(state[2]), N])
    elif async_start is not None:
        async_start()
    header = header.to(dev)
    dist.broadcast(header, src=src, group=group)
    if on_gpu and rank != 0:                # sleep, do not spin, while rank 0 is still building
        ev = torch.cuda.Event(blocking=True)
        ev.record()
        ev.synchronize()
    header = header.cpu()
    if int(header[0]) == 0:                 # not a native tree: every rank builds it (RNG untouched so far)
        return build_structure(locs, r, M, J, critDepth)
    nn, depth, nk, mt_pos = int(header[1]), int(header[2]), int(header[3]), int(header[4])
    total = N + len(_I32_FIELDS) * nn + 2 * nk + 624
    if rank == 0:
        kloc = st.node_kinds_local._kloc
        parts = [st.perm.astype(np.int32)] + [np.asarray(getattr(st, f)).astype(np.int32) for f in _I32_FIELDS]
        parts += [st.knot_rows.astype(np.int32), np.asarray(kloc, dtype=np.int32),
                  np.ascontiguousarray(state[1], dtype=np.uint32).view(np.int32)]
        payload = torch.from_numpy(np.concatenate(parts)).to(dev)
    else:
        payload = torch.empty(total, dtype=torch.int32, device=dev)
    dist.broadcast(payload, src=src, group=group)
    if rank == 0:
        return st
    buf = payload.cpu().numpy()
    o = 0

    def take(n):
        nonlocal o
        a = buf[o:o + n]
        o += n
        return a
    st = TreeStructure()
    st.N, st.d, st.r, st.J, st.M, st.depth = N, locs.shape[1], r, J, M, depth
    st.perm = take(N).astype(np.int64)
    for f in _I32_FIELDS:
        a = take(nn)
        setattr(st, f, a.astype(np.int64) if f in ("node_row_start", "node_row_count", "node_knot_off") else a.copy())
    st.knot_rows = take(nk).astype(np.int64)
    kloc = take(nk).copy()
    key = take(624).view(np.uint32).copy()
    st.level_off = np.searchsorted(st.node_level, np.arange(depth + 2)).astype(np.int32)
    st.node_id = _LazyIds(st)
    st.node_kinds_local = _LazyKinds(st, kloc)
    mine = np.random.get_state()
    np.random.set_state((mine[0], key, mt_pos, mine[3], mine[4]))
    return st


_COMM_STREAMS = {}


def _comm_stream(dev):
    import torch
    key = (dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _COMM_STREAMS:
        _COMM_STREAMS[key] = torch.cuda.Stream(device=dev, priority=-1)
    return _COMM_STREAMS[key]


class GroupStreamBuild(object):
    """StreamBuild (pymra_b200/structure.py) for a process group: rank 0 runs the streaming native builder and
    forwards every progress event with the arrays that became final -- event 0: permutation, node table and the
    root's knots; event 1 + c: the knots of the root's child subtree c; event 5: the advanced MT19937 state.
    Every rank sees the same `wait(event)` / `finish()` results and ends with the same tree and global NumPy RNG
    state.  Broadcasts run on their own high-priority stream so that they never queue behind the evaluation
    kernels of a rank that is still busy with an earlier subtree."""

    def __init__(self, locs, r, M, J, critDepth, group=None, device=None):
        import torch
        import torch.distributed as dist

        from .structure import StreamBuild
        self.torch, self.dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group)
        self.src = dist.get_global_rank(group, 0) if group is not None else 0
        self.on_gpu = dist.get_backend(group) == "nccl"
        self.dev = (torch.device("cuda", torch.cuda.current_device() if device is None else device)
                    if self.on_gpu else torch.device("cpu"))
        # one side stream per device for the whole process: the caching allocator keeps a pool per stream, so a
        # fresh stream per construction would mean fresh cudaMallocs (slow, and synchronising, once NCCL has
        # enabled peer access) for every forwarded payload
        self.comm = _comm_stream(self.dev) if self.on_gpu else None
        self.locs = np.ascontiguousarray(locs, dtype=np.float64)
        self.r, self.M, self.J = r, M, J
        self.sb = None
        self.done = None
        self.structure = None
        ok = 0
        if self.rank == 0:
            import os
            ncpu = os.cpu_count() or 8
            # the other ranks wait: all host threads for this build (the job reads the variable when it starts its
            # partition, so it is restored only after event 0)
            self._saved_threads = ("MRA_BUILD_THREADS", os.environ.get("MRA_BUILD_THREADS"))
            os.environ["MRA_BUILD_THREADS"] = str(ncpu if ncpu <= 8 else min(12, ncpu - 1))
            if self.locs.ndim == 2 and self.locs.shape[1] == 2:
                self.sb = StreamBuild(self.locs, r, M, J, critDepth)
            ok = 1 if (self.sb is not None and self.sb.started) else 0
            if not ok:
                self._restore_threads()
        head = self._bcast(np.array([ok], dtype=np.int32), 1)
        self.started = bool(head[0])
        self.N = len(self.locs)
        self.nn = sum(4 ** m for m in range(M + 1))
        self.nk = (self.nn - 4 ** M) * r
        self.level_off = np.cumsum([0] + [4 ** m for m in range(M + 1)]).astype(np.int64)

    def _restore_threads(self):
        import os
        saved = getattr(self, "_saved_threads", None)
        if saved is not None:
            self._saved_threads = None
            if saved[1] is None:
                os.environ.pop(saved[0], None)
            else:
                os.environ[saved[0]] = saved[1]

    # one int32 payload from rank 0 to everybody; non-root ranks sleep (not spin) while rank 0 is busy
    def _bcast(self, arr, n):
        torch, dist = self.torch, self.dist
        if not self.on_gpu:
            t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int32)) if self.rank == 0 else torch.empty(n, dtype=torch.int32)
            dist.broadcast(t, src=self.src, group=self.group)
            return t.numpy()
        with torch.cuda.stream(self.comm):
            if self.rank == 0:
                t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int32)).to(self.dev, non_blocking=True)
            else:
                t = torch.empty(n, dtype=torch.int32, device=self.dev)
            dist.broadcast(t, src=self.src, group=self.group)
            if self.rank == 0:
                return arr
            ev = torch.cuda.Event(blocking=True)
            ev.record(self.comm)
            ev.synchronize()
            return t.cpu().numpy()

    def _subtree_slots(self, c):
        r = self.r
        out = []
        for L in range(1, self.M):
            lo = int(self.level_off[L]) + c * 4 ** (L - 1)
            out.append((lo * r, (lo + 4 ** (L - 1)) * r))
        return out

    def wait(self, event):
        from .structure import TreeStructure, _LazyIds, _LazyKinds
        rank0 = self.rank == 0
        r, nn, nk, N = self.r, self.nn, self.nk, self.N
        if event == 0:
            size = 1 + N + len(_I32_FIELDS) * nn + 2 * r
            msg = None
            if rank0:
                ok = self.sb.wait(0)
                self._restore_threads()
                st = self.sb.structure
                if ok:
                    msg = np.concatenate([np.array([1], dtype=np.int32), st.perm.astype(np.int32)]
                                         + [np.asarray(getattr(st, f)).astype(np.int32) for f in _I32_FIELDS]
                                         + [st.knot_rows[:r].astype(np.int32), self.sb._kloc[:r]])
                else:
                    msg = np.zeros(size, dtype=np.int32)
            msg = self._bcast(msg, size)
            if not msg[0]:
                return False
            if rank0:
                self.structure = self.sb.structure
                return True
            o = 1
            st = TreeStructure()
            st.N, st.d, st.r, st.J, st.M, st.depth = N, 2, r, self.J, self.M, self.M
            st.perm = msg[o:o + N].astype(np.int64)
            o += N
            for f in _I32_FIELDS:
                a = msg[o:o + nn]
                o += nn
                setattr(st, f, a.astype(np.int64) if f in ("node_row_start", "node_row_count", "node_knot_off") else a.copy())
            st.knot_rows = np.zeros(max(1, nk), dtype=np.int64)[:nk]
            self._kloc = np.zeros(max(1, nk), dtype=np.int32)[:nk]
            st.knot_rows[:r] = msg[o:o + r]
            self._kloc[:r] = msg[o + r:o + 2 * r]
            st.level_off = self.level_off.astype(np.int32)
            st.node_id = _LazyIds(st)
            st.node_kinds_local = _LazyKinds(st, self._kloc)
            self.structure = st
            return True
        if 1 <= event <= 4:
            slots = self._subtree_slots(event - 1)
            cnt = sum(b - a for a, b in slots)
            msg = None
            if rank0:
                ok = self.sb.wait(event)
                st = self.sb.structure
                if ok:
                    msg = np.concatenate([np.array([1], dtype=np.int32)]
                                         + [st.knot_rows[a:b].astype(np.int32) for a, b in slots]
                                         + [self.sb._kloc[a:b] for a, b in slots])
                else:
                    msg = np.zeros(1 + 2 * cnt, dtype=np.int32)
            msg = self._bcast(msg, 1 + 2 * cnt)
            if not msg[0]:
                return False
            if not rank0:
                o = 1
                for a, b in slots:
                    self.structure.knot_rows[a:b] = msg[o:o + b - a]
                    o += b - a
                for a, b in slots:
                    self._kloc[a:b] = msg[o:o + b - a]
                    o += b - a
            return True
        return self.finish()

    def finish(self):
        if self.done is not None:
            return self.done
        if not self.started:
            self.done = False
            return False
        msg = None
        if self.rank == 0:
            ok = self.sb.finish()
            state = np.random.get_state()
            msg = np.concatenate([np.array([1 if ok else 0, int(state[2])], dtype=np.int32),
                                  np.ascontiguousarray(state[1], dtype=np.uint32).view(np.int32)])
        msg = self._bcast(msg, 2 + 624)
        self.done = bool(msg[0])
        if self.done and self.rank != 0:
            mine = np.random.get_state()
            np.random.set_state((mine[0], msg[2:].view(np.uint32).copy(), int(msg[1]), mine[3], mine[4]))
        return self.done
