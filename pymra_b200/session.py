"""DeviceSession: one device handle of the C ABI (include/pymra_b200.h) bound to one tree
structure and one data set.  MRATree uses it for the reference-shaped API; bench.py uses it
directly to time device-resident passes.  PyTorch supplies the arena and the stream only."""
import ctypes as C
import time

import numpy as np

from . import _ffi


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("pymra_b200 needs a CUDA device (B200); there is no CPU fallback")
    return torch


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class DeviceSession(object):
    def __init__(self, structure, locs, obs, want_predict=True, device=None, group=None, emulate=None,
                 gather="all", staged=None, two_part=False):
        """group: a torch.distributed process group (or True for the default group) to shard whole
        subtrees across its ranks, one GPU per rank (pymra_b200/shard.py); None = single GPU.
        emulate=(world, rank): build the shard of `rank` without a process group; the caller drives
        likelihood_local_async / likelihood_top_async and reduces `summary` itself (single-GPU tests).
        staged=(dev_locs, dev_obs): torch tensors already holding locs / obs on the device (caller's order).
        two_part=True (unsharded, needs `staged`): only the tree-dependent half of the set-up runs here (mra_plan_tree /
        mra_bind_tree), so that the prior pass can be launched at once; finish_plan() does the observation half
        (mra_plan_obs / mra_bind_obs) while the device is already busy."""
        torch = _torch()
        self.structure = structure
        self.N = structure.N
        self.dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.lib = _ffi.lib()
        self.h = C.c_void_p()
        self.timings = {}
        self.ws = self.ws_obs = None
        t0 = time.perf_counter()
        st = self.lib.mra_create(C.byref(self.h), self.dev.index)
        if st != 0:
            raise _ffi.MraError(st, "mra_create failed (no usable CUDA device?)")
        if group is not None or emulate is not None:
            self.check(self.lib.mra_expect_shard(self.h))      # the lists are built once, by mra_set_shard
        self._set_structure()
        self.timings["set_structure"] = time.perf_counter() - t0
        self.group, self.world, self.rank, self.shard_level, self.summary = None, 1, 0, 0, None
        self._collective = False    # True when the ranks of a real process group run this session together
        self.gather = gather        # sharded predict(): "all" = every rank gets all N results, "root" = rank 0 only
        if group is not None or emulate is not None:
            from .shard import plan_shards
            if emulate is not None:
                self.world, self.rank = emulate
            else:
                import torch.distributed as dist
                self.group = None if group is True else group
                self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
                self._collective = True
            s, role, _ = plan_shards(structure, self.world, self.rank)
            if s:
                role = np.ascontiguousarray(role, dtype=np.int8)
                self.check(self.lib.mra_set_shard(self.h, s, role.ctypes.data_as(C.POINTER(C.c_int8))))
                n = C.c_int64()
                self.check(self.lib.mra_summary_size(self.h, C.byref(n)))
                self.summary = torch.zeros(int(n.value), dtype=torch.float64, device=self.dev)
                self.shard_level = s
        t0 = time.perf_counter()
        obs_c = np.ascontiguousarray(np.asarray(obs, dtype=np.float64).reshape(self.N))
        locs_c = np.ascontiguousarray(np.asarray(locs, dtype=np.float64).reshape(self.N, structure.d))
        self._pending_obs = None
        if two_part and staged is not None and not self.shard_level:
            nbytes = C.c_size_t()
            self.check(self.lib.mra_plan_tree(self.h, 1 if want_predict else 0, C.byref(nbytes)))
            t1 = time.perf_counter()
            self.ws = torch.empty(int(nbytes.value) + 256, dtype=torch.uint8, device=self.dev)
            aligned = (self.ws.data_ptr() + 255) // 256 * 256
            t2 = time.perf_counter()
            self.check(self.lib.mra_bind_tree(self.h, C.c_void_p(aligned), C.c_size_t(int(nbytes.value)),
                                              C.c_void_p(staged[0].data_ptr()), C.c_void_p(staged[1].data_ptr()),
                                              self.stream()))
            self.workspace_bytes = int(nbytes.value)
            self._pending_obs = obs_c
            self.timings.update(plan_tree=t1 - t0, alloc_bind=t2 - t1, upload_tree=time.perf_counter() - t2)
            return
        nbytes = C.c_size_t()
        self.check(self.lib.mra_plan(self.h, _dptr(obs_c), 1 if want_predict else 0, C.byref(nbytes)))
        t1 = time.perf_counter()
        self.workspace_bytes = int(nbytes.value)
        self.ws = torch.empty(self.workspace_bytes + 256, dtype=torch.uint8, device=self.dev)
        aligned = (self.ws.data_ptr() + 255) // 256 * 256
        self.check(self.lib.mra_bind_workspace(self.h, C.c_void_p(aligned), C.c_size_t(self.workspace_bytes)))
        t2 = time.perf_counter()
        if staged is not None:
            self.check(self.lib.mra_upload_data_dev(self.h, C.c_void_p(staged[0].data_ptr()),
                                                    C.c_void_p(staged[1].data_ptr()), self.stream()))
        else:
            self.upload(locs_c, obs_c)
        self.timings.update(plan=t1 - t0, alloc_bind=t2 - t1, upload=time.perf_counter() - t2)

    def finish_plan(self):
        """Second half of a two_part set-up: NaN scan, the leaves' row lists, the observation-dependent part of the arena."""
        if self._pending_obs is None:
            return
        torch = _torch()
        t0 = time.perf_counter()
        nbytes = C.c_size_t()
        self.check(self.lib.mra_plan_obs(self.h, _dptr(self._pending_obs), C.byref(nbytes)))
        t1 = time.perf_counter()
        self.ws_obs = torch.empty(int(nbytes.value) + 256, dtype=torch.uint8, device=self.dev)
        aligned = (self.ws_obs.data_ptr() + 255) // 256 * 256
        self.check(self.lib.mra_bind_obs(self.h, C.c_void_p(aligned), C.c_size_t(int(nbytes.value)), self.stream()))
        self.workspace_bytes += int(nbytes.value)
        self._pending_obs = None
        self.timings.update(plan_obs=t1 - t0, upload_obs=time.perf_counter() - t1)

    # ---- plumbing
    def check(self, status):
        _ffi.check(self.h, status)

    def stream(self):
        import torch
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def _set_structure(self):
        s = self.structure
        keep = dict(
            node_level=np.ascontiguousarray(s.node_level, dtype=np.int32),
            node_parent=np.ascontiguousarray(s.node_parent, dtype=np.int32),
            node_kind=np.ascontiguousarray(s.node_kind, dtype=np.int32),
            node_row_start=np.ascontiguousarray(s.node_row_start, dtype=np.int64),
            node_row_count=np.ascontiguousarray(s.node_row_count, dtype=np.int64),
            node_child_start=np.ascontiguousarray(s.node_child_start, dtype=np.int32),
            node_child_count=np.ascontiguousarray(s.node_child_count, dtype=np.int32),
            node_knot_off=np.ascontiguousarray(s.node_knot_off, dtype=np.int64),
            knot_rows=np.ascontiguousarray(s.knot_rows, dtype=np.int64),
            level_off=np.ascontiguousarray(s.level_off, dtype=np.int32),
            perm=np.ascontiguousarray(s.perm, dtype=np.int64))
        ms = _ffi.MraStructure()
        ms.n_locs, ms.dim, ms.r, ms.depth, ms.n_nodes = s.N, s.d, s.r, s.depth, s.n_nodes
        ms.n_knot_rows = len(keep["knot_rows"])
        for name, arr in keep.items():
            ctype = C.c_int32 if arr.dtype == np.int32 else C.c_int64
            setattr(ms, name, arr.ctypes.data_as(C.POINTER(ctype)))
        self.check(self.lib.mra_set_structure(self.h, C.byref(ms)))
        self._knot_rows_i64 = keep["knot_rows"]      # no copy when the structure already holds contiguous int64
        self.h2d_structure_bytes = sum(a.nbytes for a in keep.values())

    # ---- data / parameters
    def upload(self, locs_c, obs_c):
        self.check(self.lib.mra_upload_data(self.h, _dptr(locs_c), _dptr(obs_c), self.stream()))

    def set_params(self, cov, R):
        if cov.family == _ffi.COV_DENSE:
            import torch
            if getattr(self, "_dense_src", None) is not cov.matrix:      # the same matrix stays resident across refits
                self._dense = torch.from_numpy(cov.matrix).to(self.dev)
                self._dense_src = cov.matrix
            self.check(self.lib.mra_set_cov_dense(self.h, C.c_void_p(self._dense.data_ptr()), cov.matrix.shape[0],
                                                  float(cov.sig)))
        else:
            self._dense = self._dense_src = None
            self.check(self.lib.mra_set_cov(self.h, cov.family, cov.l, cov.sig))
        self.check(self.lib.mra_set_nugget(self.h, float(R)))

    # ---- passes
    def likelihood(self):
        self.likelihood_async()
        return self.fetch_likelihood()

    def likelihood_async(self):
        if not self.shard_level:
            self.check(self.lib.mra_run_likelihood_async(self.h, self.stream()))
            return
        import torch.distributed as dist
        self.likelihood_local_async()
        dist.all_reduce(self.summary, group=self.group)       # disjoint slots: the sum is exact
        self.likelihood_top_async()

    def likelihood_graph(self, with_predict=False):
        """The whole likelihood pass (optionally + predict pass) as one CUDA graph launch; captured on first use and
        after anything that changes the work lists.  Parameters set with set_params since the capture take effect."""
        self.check(self.lib.mra_run_graph(self.h, self.stream(), 1 if with_predict else 0))

    def likelihood_local_async(self):
        """Sharded step 1: prior + leaves + upward pass of my subtrees; my summaries land in self.summary."""
        self.summary.zero_()
        self.check(self.lib.mra_run_likelihood_local_async(self.h, self.stream(), C.c_void_p(self.summary.data_ptr())))

    def likelihood_top_async(self):
        """Sharded step 3 (after self.summary has been sum-reduced over the ranks): replicated top levels."""
        self.check(self.lib.mra_run_likelihood_top_async(self.h, self.stream(), C.c_void_p(self.summary.data_ptr())))

    # ---- streamed evaluation (overlaps the device passes with a host build that is still drawing knots)
    def n_parts(self):
        n = C.c_int32()
        self.check(self.lib.mra_stream_parts(self.h, C.byref(n)))
        return int(n.value)

    def my_parts(self):
        """Parts this rank evaluates in a streamed pass (all of them when unsharded)."""
        m = C.c_int32()
        self.check(self.lib.mra_stream_my_parts(self.h, C.byref(m)))
        return [p for p in range(31) if (int(m.value) >> p) & 1]

    def _knots_ptr(self):
        # the int64 array mra_set_structure was given; for a streamed build it is the array the builder fills
        return self._knot_rows_i64.ctypes.data_as(C.POINTER(C.c_int64))

    def stream_begin(self):
        if self.shard_level:
            self.summary.zero_()
        self.check(self.lib.mra_stream_begin_async(self.h, self.stream(), self._knots_ptr()))

    def stream_part(self, part):
        self.check(self.lib.mra_stream_part_async(self.h, self.stream(), int(part), self._knots_ptr()))

    def stream_part_prior(self, part):
        """Only the prior levels of the part (allowed before finish_plan)."""
        self.check(self.lib.mra_stream_part_prior_async(self.h, self.stream(), int(part), self._knots_ptr()))

    def stream_end_local(self):
        """Sharded at level 1: my summaries into self.summary (the caller reduces them, then likelihood_top_async)."""
        self.check(self.lib.mra_stream_end_local_async(self.h, self.stream(), C.c_void_p(self.summary.data_ptr())))

    def stream_end(self):
        if not self.shard_level:
            self.check(self.lib.mra_stream_end_async(self.h, self.stream()))
            return
        import torch.distributed as dist
        self.stream_end_local()
        dist.all_reduce(self.summary, group=self.group)       # disjoint slots: the sum is exact
        self.likelihood_top_async()
        self.sync_knots()

    def sync_knots(self):
        """Sharded streamed pass: the knots of the parts other ranks ran reach this device now (the build has
        ended, the table is final) -- a later plain pass (refit) factors every replicated top node."""
        self.check(self.lib.mra_stream_sync_knots(self.h, self.stream(), self._knots_ptr()))

    def fetch_likelihood(self):
        out = (C.c_double * 2)()
        st = self.lib.mra_fetch_likelihood(self.h, self.stream(), out)
        if self.shard_level and self._collective:
            # a failure on one rank only (a non-positive pivot in one shard) must not leave the others running
            # into the next collective on their own: every rank learns about it here and raises together
            import torch
            import torch.distributed as dist
            flag = torch.tensor([1 if st != 0 else 0], dtype=torch.int32, device=self.dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
            if int(flag.item()) and st == 0:
                raise _ffi.MraError(-4, "another rank of the group failed its pass (its own exception has the cause)")
        self.check(st)
        return float(out[0]), float(out[1])

    # ---- sharded predict: every rank packs the rows it owns, the packs are gathered (16 B per location in total)
    def my_rows(self):
        """Tree-order row ranges [(start, count), ..] whose predictions this rank emits."""
        n = C.c_int32()
        self.check(self.lib.mra_predict_rows(self.h, None, 0, C.byref(n)))
        buf = np.zeros(2 * max(1, n.value), dtype=np.int64)
        self.check(self.lib.mra_predict_rows(self.h, buf.ctypes.data_as(C.POINTER(C.c_int64)), n.value, C.byref(n)))
        return [(int(buf[2 * i]), int(buf[2 * i + 1])) for i in range(n.value)]

    def _gather_layout(self):
        """(ranges of every rank, rows per rank, padded pack length) -- exchanged once per session."""
        if getattr(self, "_layout", None) is None:
            import torch
            import torch.distributed as dist
            mine = self.my_rows()
            k = torch.tensor([len(mine)], dtype=torch.int64, device=self.dev)
            ks = [torch.zeros_like(k) for _ in range(self.world)]
            dist.all_gather(ks, k, group=self.group)
            kmax = max(int(x.item()) for x in ks)
            flat = torch.zeros(2 * max(1, kmax), dtype=torch.int64, device=self.dev)
            if mine:
                flat[:2 * len(mine)] = torch.tensor([v for rg in mine for v in rg], dtype=torch.int64)
            allr = [torch.zeros_like(flat) for _ in range(self.world)]
            dist.all_gather(allr, flat, group=self.group)
            ranges = []
            for r in range(self.world):
                a = allr[r].cpu().numpy()
                ranges.append([(int(a[2 * i]), int(a[2 * i + 1])) for i in range(int(ks[r].item()))])
            rows = [sum(c for _, c in rg) for rg in ranges]
            self._layout = (ranges, rows, max(1, max(rows)))
        return self._layout

    def predict_tree_gathered(self, root_only=False):
        """Runs the downward pass for this rank's subtrees and gathers every rank's rows: returns device tensors
        (mean, var) of all N locations in TREE order on every rank (root_only: on group rank 0, None elsewhere)."""
        import torch
        import torch.distributed as dist
        ranges, rows, pad = self._gather_layout()
        pack = torch.zeros(2, pad, dtype=torch.float64, device=self.dev)
        mine = rows[self.rank]
        if mine:
            tmp = torch.empty(2 * mine, dtype=torch.float64, device=self.dev)
            self.check(self.lib.mra_run_predict_pack_dev(self.h, self.stream(), C.c_void_p(tmp.data_ptr()), mine))
            pack[:, :mine] = tmp.view(2, mine)
        if root_only:
            dst = dist.get_global_rank(self.group, 0) if self.group is not None else 0
            got = [torch.empty_like(pack) for _ in range(self.world)] if self.rank == 0 else None
            dist.gather(pack, got, dst=dst, group=self.group)
            if self.rank != 0:
                return None, None
        else:
            got = torch.empty(self.world, 2, pad, dtype=torch.float64, device=self.dev)
            dist.all_gather_into_tensor(got, pack, group=self.group)
        full = torch.empty(2, self.N, dtype=torch.float64, device=self.dev)
        for r in range(self.world):
            off = 0
            for start, cnt in ranges[r]:
                full[:, start:start + cnt] = got[r][:, off:off + cnt]
                off += cnt
        return full[0], full[1]

    def predict(self):
        if self.shard_level:
            import torch
            mt, vt = self.predict_tree_gathered(root_only=(self.gather == "root"))
            self.fetch_likelihood()                           # surfaces a non-SPD status like the plain path (collective)
            if mt is None:
                return None, None
            out = torch.empty(2, self.N, dtype=torch.float64, device=self.dev)
            self.check(self.lib.mra_unpermute_tree_dev(self.h, self.stream(), C.c_void_p(mt.data_ptr()),
                                                       C.c_void_p(vt.data_ptr()), C.c_void_p(out[0].data_ptr()),
                                                       C.c_void_p(out[1].data_ptr())))
            host = torch.empty(2, self.N, dtype=torch.float64, pin_memory=True)
            host.copy_(out)
            res = host.numpy()
            return res[0], res[1]
        # page-locked result buffer from torch's caching host allocator: the device->host copy runs at PCIe
        # rate and the returned arrays are zero-copy views of it (the block is recycled once they are dropped;
        # an extra host copy out of a shared staging buffer was measured slower)
        import torch
        t0 = time.perf_counter()
        out = torch.empty(2, self.N, dtype=torch.float64, pin_memory=True)
        host = out.numpy()
        t1 = time.perf_counter()
        self.check(self.lib.mra_run_predict(self.h, self.stream(), _dptr(host[0]), _dptr(host[1])))
        self.timings.update(predict_alloc=t1 - t0, predict_call=time.perf_counter() - t1)
        return host[0], host[1]

    def predict_dev(self, mean_t=None, sd_t=None, reduce=False):
        """Results into device tensors (caller's order).  Sharded without reduce: each rank fills its own rows and
        zeros elsewhere; reduce=True: the ranks' rows are gathered (16 B per location) so that every rank holds all N
        results."""
        if reduce and self.shard_level and self._collective:
            if mean_t is None or sd_t is None:
                raise ValueError("reduce=True needs caller-owned output tensors")
            mt, vt = self.predict_tree_gathered()
            self.check(self.lib.mra_unpermute_tree_dev(self.h, self.stream(), C.c_void_p(mt.data_ptr()),
                                                       C.c_void_p(vt.data_ptr()), C.c_void_p(mean_t.data_ptr()),
                                                       C.c_void_p(sd_t.data_ptr())))
            return
        pm = C.c_void_p(mean_t.data_ptr()) if mean_t is not None else None
        ps = C.c_void_p(sd_t.data_ptr()) if sd_t is not None else None
        self.check(self.lib.mra_run_predict_dev(self.h, self.stream(), pm, ps))

    def keep_posterior_basis(self, on=True):
        """Diagnostics (pymra_b200/diagnostics.py): the next predict pass keeps every level's folded posterior basis."""
        self.check(self.lib.mra_set_diagnostics(self.h, 1 if on else 0))

    # ---- counters
    def warnings(self):
        """MRA_WARN_* bits the passes have raised since the last likelihood pass started (include/pymra_b200.h)."""
        f = C.c_int32()
        self.check(self.lib.mra_last_warnings(self.h, C.byref(f)))
        return int(f.value)

    def launches(self):
        n = C.c_int64()
        self.check(self.lib.mra_last_launches(self.h, C.byref(n)))
        return int(n.value)

    def flops(self):
        a, b = C.c_double(), C.c_double()
        self.check(self.lib.mra_last_flops(self.h, C.byref(a), C.byref(b)))
        return float(a.value), float(b.value)

    def profile_enable(self, on=True):
        self.check(self.lib.mra_profile_enable(self.h, 1 if on else 0))

    def profile_read(self):
        buf = C.create_string_buffer(1 << 16)
        self.check(self.lib.mra_profile_read(self.h, buf, C.c_size_t(len(buf))))
        out = {}
        for line in buf.value.decode().splitlines():
            name, ms, n, fl, by = line.split()
            out[name] = dict(ms=float(ms), launches=int(n), flops=float(fl), bytes=float(by))
        return out

    def debug_fetch(self, what, node=0, count=None):
        count = (1 << 24) if count is None else int(count)
        buf = np.empty(count)
        n = self.lib.mra_debug_fetch(self.h, what.encode(), int(node), _dptr(buf), C.c_int64(count))
        if n < 0:
            self.check(int(n))
        return buf[:n].copy()

    def close(self):
        if getattr(self, "h", None):
            self.lib.mra_destroy(self.h)
            self.h = None
            self.ws = None
            self.ws_obs = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
