"""Multi-GPU subtree sharding through the C ABI.

test_emulated_ranks_*: all ranks of a sharded run are emulated as separate handles on ONE GPU (the
summary all-reduce is a plain tensor sum), so the shard work lists, gathered knot tiles, summary export and
replicated top pass are checked wherever a single B200 is available.  test_nccl_*: the real thing, one
process per GPU over NCCL; skipped with fewer than 2 GPUs."""
import os
import socket

import numpy as np
import pytest

from _util import errs, load_golden, make_cov, oracle_for, structure_for, tree_for

pytestmark = pytest.mark.gpu


def _sessions(g, world):
    import torch
    from pymra_b200.covariance import introspect
    from pymra_b200.session import DeviceSession
    st = structure_for(g)
    desc = introspect(make_cov(g), g["locs"].shape[1])
    out = []
    for rank in range(world):
        s = DeviceSession(st, g["locs"], g["obs"], want_predict=True, emulate=(world, rank))
        s.set_params(desc, float(g["R"]))
        out.append(s)
    return st, out


@pytest.mark.parametrize("name,world", [("g96_m32_r16", 2), ("g96_m32_r16", 4), ("g125_m32_r16", 8),
                                        ("ka4_large_serial", 3), ("g64_m3_exp", 4)])
def test_emulated_ranks_match_single_gpu_and_reference(name, world):
    import torch
    g = load_golden(name)
    t = tree_for(g)                                   # single-GPU path
    lik1 = float(t.getLikelihood())
    mean1, sd1 = t.predict()
    st, sess = _sessions(g, world)
    if not sess[0].shard_level:
        pytest.skip("tree too shallow to shard %d ways" % world)
    for s in sess:
        s.likelihood_local_async()
    total = torch.stack([s.summary for s in sess]).sum(0)
    N = st.N
    mean = torch.zeros(N, dtype=torch.float64, device="cuda")
    sd = torch.zeros(N, dtype=torch.float64, device="cuda")
    liks = []
    for s in sess:
        s.summary.copy_(total)
        s.likelihood_top_async()
        d, u = s.fetch_likelihood()
        liks.append(d + u)
        m_k = torch.empty(N, dtype=torch.float64, device="cuda")
        s_k = torch.empty(N, dtype=torch.float64, device="cuda")
        s.predict_dev(m_k, s_k)
        assert int(((m_k != 0) & (mean != 0)).sum()) == 0          # disjoint rows
        mean += m_k
        sd += s_k
    assert all(l == liks[0] for l in liks)                          # replicated top: bitwise equal
    assert abs(liks[0] - lik1) <= 1e-12 * abs(lik1)
    assert float((mean.cpu() - torch.from_numpy(np.asarray(mean1).ravel())).abs().max()) <= 1e-11
    assert float((sd.cpu() - torch.from_numpy(sd1)).abs().max()) <= 1e-11
    o = oracle_for(g)
    fl, fm, fs = errs(o["lik"], o["mean"], o["sd"], g)
    rl, em, es = errs(liks[0], mean.cpu().numpy(), sd.cpu().numpy(), g)
    assert rl <= max(1e-9, 20 * fl) and em <= max(1e-9 * max(1.0, float(np.max(np.abs(g["mean"])))), 20 * fm)
    assert es <= max(1e-9, 20 * fs)


def _free_port():
    with socket.socket() as so:
        so.bind(("127.0.0.1", 0))
        return so.getsockname()[1]


def _nccl_worker(rank, world, port, name, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from pymra_b200.MRATree import MRATree
    g = load_golden(name)
    np.random.seed(int(g["seed"]))
    t = MRATree(g["locs"], int(g["r"]), make_cov(g), g["obs"], float(g["R"]), M=int(g["M_req"]), J=int(g["J_req"]),
                critDepth=int(g["critDepth"]), group=True)
    lik = float(np.asarray(t.getLikelihood()).ravel()[0])
    mean, sd = t.predict()
    q.put((rank, t._session.shard_level, lik, np.asarray(mean).ravel(), sd))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["g96_m32_r16", "g125_m32_r16"])
def test_nccl_two_ranks_match_reference(name):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    g = load_golden(name)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(rk, world, port, name, q)) for rk in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    o = oracle_for(g)
    fl, fm, fs = errs(o["lik"], o["mean"], o["sd"], g)
    for rank, s, lik, mean, sd in got:
        assert s >= 1
        rl, em, es = errs(lik, mean, sd, g)
        assert rl <= max(1e-9, 20 * fl) and em <= max(1e-9, 20 * fm) and es <= max(1e-9, 20 * fs)
    assert got[0][2] == got[1][2]


def _nccl_stream_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import pymra_b200.MRATools as mt
    from pymra_b200.MRATree import MRATree
    locs, y = _stream_case()
    np.random.seed(5 if rank == 0 else 77)                 # a native tree is rank 0's draw
    t = MRATree(locs, 16, lambda a, b: mt.Matern32(a, b, l=0.3, sig=1.0), y, 1e-2, M=4, group=True)
    state = np.random.get_state()
    lik = float(np.asarray(t.getLikelihood()).ravel()[0])
    mean, sd = t.predict()
    q.put((rank, t.timings.get("streamed"), lik, np.asarray(mean).ravel(), sd, state[1][:8].copy(), int(state[2])))
    dist.barrier()
    dist.destroy_process_group()


def _stream_case():
    import pymra_b200.MRATools as mt
    locs = mt.genLocations2d(300)
    N = len(locs)
    rng = np.random.RandomState(3)
    sel = np.sort(rng.choice(N, int(0.4 * N), replace=False))
    y = np.full((N, 1), np.nan)
    y[sel] = np.sin(5.0 * locs[sel, :1]) + 0.3 * rng.normal(size=(len(sel), 1))
    return locs, y


def test_nccl_two_ranks_streamed_group_build(monkeypatch):
    """>= 65536 locations: rank 0 builds and forwards the events, both ranks stream their subtrees; the result
    equals the plain single-GPU construction from rank 0's RNG state, and both ranks end with its RNG state."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import pymra_b200.MRATools as mt
    from pymra_b200.MRATree import MRATree
    locs, y = _stream_case()
    monkeypatch.setenv("PYMRA_B200_STREAM", "0")
    np.random.seed(5)
    t = MRATree(locs, 16, lambda a, b: mt.Matern32(a, b, l=0.3, sig=1.0), y, 1e-2, M=4)
    want_state = np.random.get_state()
    lik1 = float(np.asarray(t.getLikelihood()).ravel()[0])
    mean1, sd1 = t.predict()
    mean1 = np.asarray(mean1).ravel().copy()
    sd1 = sd1.copy()
    del t
    monkeypatch.setenv("PYMRA_B200_STREAM", "1")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_stream_worker, args=(rk, 2, port, q)) for rk in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, streamed, lik, mean, sd, key8, pos in got:
        assert streamed == 1.0
        assert abs(lik - lik1) <= 1e-12 * abs(lik1)
        assert np.max(np.abs(mean - mean1)) <= 1e-11 and np.max(np.abs(sd - sd1)) <= 1e-11
        assert np.array_equal(key8, want_state[1][:8]) and pos == int(want_state[2])
    assert got[0][2] == got[1][2]
