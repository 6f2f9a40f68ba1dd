"""Drop-in `MRATree` for the hot path of pyMRA (pyMRA/MRATree.py:20-94).

Same constructor signature and semantics as the reference class; the work the reference does in
`Node.__init__` (pyMRA/MRANode.py:23-115) is split into
  * the host structure builder (pymra_b200/structure.py; bit-exact tree/knot/partition indexing,
    consumes the global NumPy RNG like the reference), and
  * the device library behind the C ABI (include/pymra_b200.h): prior pass, leaf terms, upward
    pass for the likelihood, downward pass for predictions.
PyTorch only supplies the device arena and the stream.  There is no CPU fallback.

Beyond the reference signature: `device` (CUDA ordinal) and `group` (a torch.distributed process group,
or True for the default one): every rank of the group constructs the same tree from the same inputs and
RNG state, whole subtrees are sharded over the ranks' GPUs (pymra_b200/shard.py), and every rank returns
the full likelihood and predictions (gather="root": predictions only on rank 0, None elsewhere).
"""
import logging
import os
import time
import weakref

import numpy as np

from . import _ffi
from .covariance import introspect
from .session import DeviceSession
from .structure import StreamBuild, build_structure

logger = logging.getLogger("pymra_b200.MRATree")


def resolve_params(N, d, r, M, J, critDepth):
    """J default, M clamp, critDepth/mode resolution exactly as MRATree.py:31-59."""
    if J < 0:
        if d == 2:
            J = 4
        else:
            # MRATree.py:33 compares instead of assigning, so 1-D callers must pass J
            raise AttributeError("'MRATree' object has no attribute 'J'")
    num = np.log(N * J / r + 1)
    denom = np.log(J)
    with np.errstate(divide="ignore"):
        maxM = int(num / denom) - 1       # J == 1 raises OverflowError like the reference
    clamped = None
    if M < 0:
        M = maxM
    elif M > maxM:
        clamped = (M, maxM)
        M = maxM
    if critDepth < 0:
        critDepth = M + 1
    return M, J, critDepth, clamped


class _Root(object):
    """The attributes of the reference's root Node that callers touch."""

    def __init__(self, tree):
        self._tree_ref = weakref.ref(tree)     # no reference cycle: the device arena is released with the tree
        self.children = []
        self.ID = "r"
        self.res = 0
        self.N = tree._N
        self.leaf = tree._structure.node_kind[0] != 0
        self.kInds = tree._structure.node_kinds_local[0]
        self.d = tree._d
        self.u = np.matrix([[tree._u]])

    @property
    def _tree(self):
        t = self._tree_ref()
        if t is None:
            raise ReferenceError("the MRATree this root belonged to has been deleted")
        return t

    @property
    def B(self):
        """The root's prior basis cov(locs, knots) (MRANode.py:384; no ancestors to condition on), evaluated with the
        caller's closure on first use -- what callers of the reference read from tree.root.B."""
        t = self._tree
        if isinstance(t._cov_closure, np.ndarray):          # dense covariance matrix: a column slice (MRANode.py:381-382)
            return np.matrix(np.asarray(t._cov_closure)[:, self.kInds])
        X = np.asarray(t.locs, dtype=np.float64).reshape(t._N, t.d)
        return np.matrix(t._cov_closure(X, X[self.kInds]))

    @property
    def mean(self):
        return self._tree._moments()[0]

    @property
    def var(self):
        sd = self._tree._moments()[1]          # root.var = sd^2, formed only when somebody asks for it
        return None if sd is None else sd * sd


class MRATree(object):

    def __init__(self, locs, r, cov, obs, R, M=-1, J=-1, critDepth=-1, verbose=True, device=None, group=None,
                 gather="all"):
        t0 = time.perf_counter()
        self.timings = {}          # host-side wall-clock breakdown of this construction (seconds)
        self.locs = locs
        self.d = np.shape(self.locs)[1]
        N = len(locs)
        self._N = N
        self.r = r
        M, J, critDepth, clamped = resolve_params(N, self.d, r, M, J, critDepth)
        if clamped is not None:
            logger.warning("The number of resolutions M=%d you requested is to large for your grid. "
                           "Setting M:=%d" % clamped)
        self.M, self.J = M, J
        if isinstance(R, bool) or not isinstance(R, (int, float, np.integer, np.floating)):
            raise TypeError("R (me_scale) must be a real scalar; the reference's matrix-valued R "
                            "path is broken (MRANode.py:421) and is not accelerated")
        obs_arr = np.asarray(obs, dtype=np.float64)
        if obs_arr.shape != (N, 1):
            raise ValueError("obs must have shape (N, 1) (the reference wraps it in np.matrix, MRATree.py:61)")
        self._obs_ref = obs
        self._obs_inds = None
        self._cov_closure = cov
        self._cov = introspect(cov, self.d, n_locs=N)
        self._R = float(R)
        logger.debug("r: %d, \tJ: %d,\tM: %d" % (self.r, self.J, self.M))
        logger.debug("mode: %s" % ("serial" if critDepth > self.M else "parallel"))

        locs_c = np.ascontiguousarray(locs, dtype=np.float64).reshape(N, self.d)
        t1 = time.perf_counter()
        self.timings["args_and_cov"] = t1 - t0
        self._session = None
        if self.d == 2 and os.environ.get("PYMRA_B200_STREAM", "1") != "0":
            self._construct_streamed(locs_c, obs_arr, r, M, J, critDepth, device, group, gather)
        if self._session is None:
            t1 = time.perf_counter()
            staged = None
            if group is not None:
                # one build per group (rank 0 draws, everybody receives); the inputs go to the device meanwhile
                from .shard import build_structure_group
                box = []

                def early_h2d():
                    import torch
                    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
                    box.append((torch.from_numpy(locs_c).to(dev, non_blocking=True),
                                torch.from_numpy(np.ascontiguousarray(obs_arr).reshape(-1)).to(dev, non_blocking=True)))
                self._structure = build_structure_group(locs_c, r, M, J, critDepth, None if group is True else group,
                                                        async_start=early_h2d, device=device)
                staged = box[0] if box else None
            else:
                self._structure = build_structure(locs_c, r, M, J, critDepth)
            t2 = time.perf_counter()
            self._session = DeviceSession(self._structure, locs_c, obs_arr, want_predict=True, device=device,
                                          group=group, gather=gather, staged=staged)
            del staged
            t3 = time.perf_counter()
            self._mom = None
            self._evaluate()
            t4 = time.perf_counter()
            self.timings.update(structure=t2 - t1, session_plan_upload=t3 - t2, likelihood=t4 - t3,
                                **self._session.timings)
        self.root = _Root(self)

    def _construct_streamed(self, locs_c, obs_arr, r, M, J, critDepth, device, group=None, gather="all"):
        """Large regular 2-D trees: the C++ builder runs on its own thread (StreamBuild) and the device passes
        start as soon as the RNG-independent part of the tree (partition, permutation, the root's knots) is
        known; each subtree of the root is evaluated when the sequential knot draw (MRANode.py:191-193, DFS
        pre-order) has left it.  With a process group (up to 16 ranks: sharded at level 1 or 2) rank 0 builds and
        forwards every event (GroupStreamBuild) and each rank evaluates the root subtrees that hold its shards.  Same results and RNG consumption as the plain path; leaves self._session = None
        (global RNG untouched) when the tree is outside this path."""
        import torch
        if not torch.cuda.is_available():
            return                      # the plain path raises the "no CPU fallback" error
        t0 = time.perf_counter()
        if group is None:
            sb = StreamBuild(locs_c, r, M, J, critDepth)
            world, rank = 1, 0
        else:
            import torch.distributed as dist
            from .shard import GroupStreamBuild
            g = None if group is True else group
            world, rank = dist.get_world_size(g), dist.get_rank(g)
            if world > 16 or len(locs_c) < 65536 or M < 1 or M > 12:    # deterministic on every rank
                return
            sb = GroupStreamBuild(locs_c, r, M, J, critDepth, g, device=device)
        if not sb.started:
            return
        session = None
        try:
            # the inputs do not depend on the tree: copy them while the partition is running
            dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
            staged = (torch.from_numpy(locs_c).to(dev), torch.from_numpy(np.ascontiguousarray(obs_arr).reshape(-1)).to(dev))
            t1 = time.perf_counter()
            ok = sb.wait(0)
            t2 = time.perf_counter()
            nparts = 0
            two_part = world == 1 and os.environ.get("PYMRA_B200_TWO_PART", "1") != "0"
            if ok:
                # one GPU: only the tree-dependent half of the set-up now -- the root's prior level (and the first
                # subtree's, once its knots are there) run while the host scans the observations (finish_plan)
                session = DeviceSession(sb.structure, locs_c, obs_arr, want_predict=True, device=device, group=group,
                                        gather=gather, staged=staged, two_part=two_part)
                session.set_params(self._cov, self._R)
                nparts = session.n_parts()
            del staged
            t3 = time.perf_counter()
            streamable = ok and nparts == 4 and session.shard_level <= 2 and (world == 1) == (session.shard_level == 0)
            mine = session.my_parts() if streamable else []
            if streamable:
                session.stream_begin()
            waits = 0.0
            for c in range(4):                       # every rank follows every event (they are collective)
                tw = time.perf_counter()
                ok = sb.wait(1 + c) and ok
                waits += time.perf_counter() - tw
                if c == 0 and session is not None:
                    if ok and streamable and c in mine and two_part:
                        session.stream_part_prior(c)
                    tp = time.perf_counter()
                    session.finish_plan()
                    self.timings["finish_plan"] = time.perf_counter() - tp
                if ok and streamable and c in mine:
                    session.stream_part(c)
            tw = time.perf_counter()
            ok = sb.finish() and ok
            waits += time.perf_counter() - tw
            if not ok:
                return
            if streamable:
                session.stream_end()
            else:                         # built (RNG advanced) but not streamable: plain passes on the finished tree
                session.sync_knots()      # the session was set up at event 0, when only the root's knots were final
                session.likelihood_async()
            self._structure = sb.structure
            self._session, session = session, None
            self._d, self._u = self._session.fetch_likelihood()
            self._mom = None
            t4 = time.perf_counter()
            self.timings.update(structure=t2 - t0, early_h2d=t1 - t0, session_plan_upload=t3 - t2,
                                likelihood=t4 - t3, build_waits_in_likelihood=waits,
                                streamed=1.0 if streamable else 0.0,
                                **self._session.timings)
        finally:
            if session is not None:       # fell out of the streaming path after device work had started
                torch.cuda.synchronize()
                session.close()
            sb.finish()

    @property
    def obs_inds(self):
        """MRATree.py:62, evaluated on first use (it is an O(N) scan no part of the hot path needs)."""
        if self._obs_inds is None:
            self._obs_inds = np.where(np.logical_not(np.isnan(self._obs_ref)))[0]
        return self._obs_inds

    # ---- plumbing
    def _evaluate(self, graph=False):
        self._session.set_params(self._cov, self._R)
        if graph and not self._session.shard_level:
            self._session.likelihood_graph()
            self._d, self._u = self._session.fetch_likelihood()
        else:
            self._d, self._u = self._session.likelihood()
        self._mom = None

    def _moments(self):
        if self._mom is None:
            t0 = time.perf_counter()
            mean, sd = self._session.predict()
            if self._session.warnings() & _ffi.WARN_NEGATIVE_VARIANCE:
                logger.warning("predict(): a predictive variance fell below -1e-12*C(0) by cancellation and was "
                               "clamped to zero (the covariance is close to singular at this resolution)")
            if mean is None:            # sharded run with gather="root" on a non-root rank
                self._mom = (None, None)
            else:
                # asmatrix: np.matrix(...) would COPY the 8 N bytes (8 ms at 4 M locations); the (N, 1) matrix the
                # reference returns (MRATree.py:90-94) is a view of the page-locked result buffer here
                self._mom = (np.asmatrix(mean.reshape(-1, 1)), sd)
            self.timings["predict"] = time.perf_counter() - t0
        return self._mom

    def _debug_fetch(self, what, node=0, count=None):
        return self._session.debug_fetch(what, node, count)

    # ---- reference API
    def getLikelihood(self):
        """root.d + root.u as a 1x1 np.matrix (MRATree.py:82-84)."""
        return np.matrix([[self._d + self._u]])

    def predict(self):
        """(root.mean as (N,1) np.matrix, sqrt(root.var) as (N,) ndarray) (MRATree.py:90-94)."""
        mean, sd = self._moments()
        return mean, sd

    def getBasisFunctionsMatrix(self, distr="prior", groupByResolution=False, order="root", timesKC=False,
                                all_levels=False):
        """MRATree.py:445-511.  On a finished tree the reference only has the root left (children are deleted,
        MRANode.py:108-110), which is what all_levels=False returns; all_levels=True assembles every resolution
        from the state kept on the device (pymra_b200/diagnostics.py).  Diagnostic: not on the hot path."""
        from .diagnostics import basis_functions_matrix
        return basis_functions_matrix(self, distr, groupByResolution, order, timesKC, all_levels)

    # ---- beyond the reference: frozen-structure re-evaluation for MLE loops
    def refit(self, cov=None, R=None):
        """Re-evaluate likelihood (and, lazily, predictions) for new covariance parameters / nugget
        on the SAME tree structure and device-resident data."""
        if cov is not None:
            self._cov_closure = cov
            self._cov = introspect(cov, self.d, n_locs=self._N)
        if R is not None:
            if isinstance(R, bool) or not isinstance(R, (int, float, np.integer, np.floating)):
                raise TypeError("R must be a real scalar")
            self._R = float(R)
        self._evaluate(graph=os.environ.get("PYMRA_B200_GRAPH", "1") != "0")      # one graph launch per evaluation
        self.root = _Root(self)
        return self.getLikelihood()
