"""Host-side logic of the multi-GPU subtree sharding (pymra_b200/shard.py), on CPU:
  * the plan partitions the level-s subtrees and the output rows exactly once over the ranks;
  * a world_size-2 gloo run of the NumPy blueprint of the device algorithm (tests/_model.py) with the
    summary exchange reproduces the unsharded likelihood / mean / sd (the same three-step protocol the
    CUDA path uses: local pass -> all-reduce of summaries -> replicated top)."""
import os
import socket

import numpy as np
import pytest

from _util import load_golden, structure_for

from pymra_b200.shard import (ROLE_MINE, ROLE_OTHER, ROLE_TOP, ROLE_TOP_EMIT, choose_shard_level, owned_rows,
                              plan_shards, summary_width)


@pytest.mark.parametrize("name", ["g96_m32_r16", "g125_m32_r16", "ka4_large_serial", "g30_kmeans"])
@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_plan_partitions_subtrees_and_rows(name, world):
    st = structure_for(load_golden(name))
    s = choose_shard_level(st, world)
    if s is None:
        pytest.skip("tree too shallow to shard %d ways" % world)
    lo, hi = int(st.level_off[s]), int(st.level_off[s + 1])
    assert hi - lo >= world and (s == 1 or int(st.level_off[s]) - int(st.level_off[s - 1]) < world)
    cover = np.zeros(st.N, dtype=int)
    owners = np.zeros(hi - lo, dtype=int)
    for rank in range(world):
        s2, role, owner = plan_shards(st, world, rank)
        assert s2 == s and len(role) == st.n_nodes
        assert np.all(role[:lo] == (ROLE_TOP_EMIT if rank == 0 else ROLE_TOP))
        assert set(np.unique(role[lo:])) <= {ROLE_MINE, ROLE_OTHER}
        owners += (role[lo:hi] == ROLE_MINE)
        # descendants inherit the role of their level-s ancestor
        for n in range(hi, st.n_nodes):
            assert role[n] == role[st.node_parent[n]]
        cover += owned_rows(st, role)
    assert np.all(owners == 1)
    assert np.all(cover == 1)
    assert summary_width(st, s) == s * st.r + 1


def _free_port():
    with socket.socket() as so:
        so.bind(("127.0.0.1", 0))
        return so.getsockname()[1]


def _worker(rank, world, port, name, q):
    import torch
    import torch.distributed as dist
    from _model import model_run
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = load_golden(name)
    st = structure_for(g)
    s, role, _ = plan_shards(st, world, rank)

    def allreduce(a):
        t = torch.from_numpy(a)
        dist.all_reduce(t)

    res = model_run(st, g["locs"], g["obs"], str(g["family"]), float(g["l"]), float(g["sig"]), float(g["R"]),
                    shard=(s, role, allreduce))
    q.put((rank, s, res["lik"], res["mean"], res["sd"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,world", [("g96_m32_r16", 2), ("g125_m32_r16", 2)])
def test_gloo_sharded_model_matches_unsharded(name, world):
    import torch.multiprocessing as mp
    from _model import model_run
    g = load_golden(name)
    st = structure_for(g)
    ref = model_run(st, g["locs"], g["obs"], str(g["family"]), float(g["l"]), float(g["sig"]), float(g["R"]))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(rk, world, port, name, q)) for rk in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, s, lik, mean, sd in got:
        assert s >= 1
        assert abs(lik - ref["lik"]) <= 1e-11 * abs(ref["lik"])
        assert np.max(np.abs(mean - ref["mean"])) <= 1e-10
        assert np.max(np.abs(sd - ref["sd"])) <= 1e-10
    # and the reference itself
    assert abs(got[0][2] - float(g["lik"])) <= 1e-9 * abs(float(g["lik"]))


def _group_build_worker(rank, world, port, q):
    import torch.distributed as dist
    import pymra_b200.MRATools as mt
    from pymra_b200.shard import build_structure_group
    from pymra_b200.structure import build_structure
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = []
    # streamed native build (N >= 65536), plain native build, and a tree outside the native path (KMeans nodes)
    for n, r, M, crit in [(300, 16, 4, -1), (300, 16, 4, 1), (60, 8, 2, -1), (12, 4, 2, -1)]:
        locs = mt.genLocations2d(n)
        crit = M + 1 if crit < 0 else crit
        np.random.seed(21)
        want = build_structure(locs, r, M, 4, crit)
        want_state = np.random.get_state()
        native_tree = n >= 60
        # a native tree is rank 0's (only its RNG state matters); otherwise every rank builds from its own state,
        # which callers keep identical across ranks
        np.random.seed(21 if rank == 0 or not native_tree else 99)
        called = []
        got = build_structure_group(locs, r, M, 4, crit, async_start=lambda: called.append(1))
        got_state = np.random.get_state()
        same = all(np.array_equal(getattr(want, f), getattr(got, f)) for f in
                   ("perm", "node_level", "node_parent", "node_kind", "node_row_start", "node_row_count",
                    "node_child_start", "node_child_count", "node_knot_off", "knot_rows", "level_off"))
        same = same and all(np.array_equal(want.node_kinds_local[k], got.node_kinds_local[k]) for k in (0, 1, want.n_nodes - 1))
        same = same and np.array_equal(want_state[1], got_state[1]) and want_state[2] == got_state[2]
        ok.append(bool(same) and called == [1])
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_group_build_broadcasts_rank0_tree_and_rng_state():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_group_build_worker, args=(rk, 2, port, q)) for rk in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok in got:
        assert ok == [True, True, True, True], (rank, ok)


def _group_stream_worker(rank, world, port, q):
    import torch.distributed as dist
    import pymra_b200.MRATools as mt
    from pymra_b200.shard import GroupStreamBuild
    from pymra_b200.structure import build_structure
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = []
    for n, r, M, crit in [(300, 16, 4, 5), (300, 8, 3, 1)]:
        locs = mt.genLocations2d(n)
        np.random.seed(21)
        want = build_structure(locs, r, M, 4, crit)
        want_state = np.random.get_state()
        np.random.seed(21 if rank == 0 else 99)
        sb = GroupStreamBuild(locs, r, M, 4, crit)
        good = sb.started and sb.wait(0)
        st = sb.structure
        good = good and np.array_equal(st.perm, want.perm) and np.array_equal(st.knot_rows[:r], want.knot_rows[:r])
        for c in range(4):
            good = good and sb.wait(1 + c)
            for L in range(1, M):
                lo = (int(st.level_off[L]) + c * 4 ** (L - 1)) * r
                hi = lo + 4 ** (L - 1) * r
                good = good and np.array_equal(st.knot_rows[lo:hi], want.knot_rows[lo:hi])
        good = good and sb.finish() and sb.finish()
        got_state = np.random.get_state()
        good = good and all(np.array_equal(getattr(want, f), getattr(st, f)) for f in
                            ("perm", "node_level", "node_parent", "node_kind", "node_row_start", "node_row_count",
                             "node_child_start", "node_child_count", "node_knot_off", "knot_rows", "level_off"))
        good = good and all(np.array_equal(want.node_kinds_local[k], st.node_kinds_local[k]) for k in (0, 2, want.n_nodes - 1))
        good = good and np.array_equal(want_state[1], got_state[1]) and want_state[2] == got_state[2]
        ok.append(bool(good))
    # a ragged tree: every rank is told to fall back, nobody's RNG state moves
    rng = np.random.RandomState(1)
    locs = rng.uniform(size=(80000, 2))
    np.random.seed(8)
    before = np.random.get_state()
    sb = GroupStreamBuild(locs, 10, 6, 4, 7)
    res = [sb.wait(e) for e in range(5)] + [sb.finish()]
    after = np.random.get_state()
    ok.append(sb.started and not all(res) and res[-1] is False and np.array_equal(before[1], after[1]) and before[2] == after[2])
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_group_stream_build_forwards_events():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_group_stream_worker, args=(rk, 2, port, q)) for rk in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok in got:
        assert ok == [True, True, True], (rank, ok)
