"""Streamed construction (mra_build_stream_* + mra_stream_*): the device passes overlap the host's sequential
knot draw.  It must give bit-identical results and consume the global RNG exactly like the plain path, which
the parity tests pin against the reference (tests/test_gpu_parity.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def make(n_side, frac=0.4, seed=3):
    import pymra_b200.MRATools as mt
    locs = mt.genLocations2d(n_side)
    N = len(locs)
    rng = np.random.RandomState(seed)
    sel = np.sort(rng.choice(N, int(frac * N), replace=False))
    y = np.full((N, 1), np.nan)
    y[sel] = np.sin(5.0 * locs[sel, :1]) + 0.3 * rng.normal(size=(len(sel), 1))
    return locs, y


def construct(locs, y, r, M, stream, monkeypatch, cov=None, crit=-1):
    import pymra_b200.MRATools as mt
    from pymra_b200.MRATree import MRATree
    monkeypatch.setenv("PYMRA_B200_STREAM", "1" if stream else "0")
    cov = cov or (lambda a, b: mt.Matern32(a, b, l=0.3, sig=1.0))
    np.random.seed(5)
    t = MRATree(locs, r, cov, y, 1e-2, M=M, critDepth=crit)
    state = np.random.get_state()
    mean, sd = t.predict()
    return t, float(t.getLikelihood()[0, 0]), np.asarray(mean).ravel().copy(), sd.copy(), state


@pytest.mark.parametrize("n_side,r,M,crit,two_part", [(300, 16, 3, -1, 1), (300, 16, 4, 1, 1), (700, 32, 5, -1, 1),
                                                      (700, 32, 5, -1, 0), (2000, 64, 10, -1, 1)])
def test_streamed_equals_plain(n_side, r, M, crit, two_part, monkeypatch):
    # two_part: the set-up in two halves (mra_plan_tree / mra_bind_tree, then mra_plan_obs / mra_bind_obs while the prior
    # pass is already running, the arena in two allocations) against the one-piece set-up
    monkeypatch.setenv("PYMRA_B200_TWO_PART", str(two_part))
    locs, y = make(n_side)
    tp, lp, mp, sp, statep = construct(locs, y, r, M, False, monkeypatch, crit=crit)
    assert "streamed" not in tp.timings
    kp = tp._structure.knot_rows.copy()
    del tp
    ts, ls, ms, ss, states = construct(locs, y, r, M, True, monkeypatch, crit=crit)
    assert ts.timings.get("streamed") == 1.0
    assert ("plan_obs" in ts.timings) == bool(two_part)
    assert np.array_equal(ts._structure.knot_rows, kp)
    assert np.array_equal(statep[1], states[1]) and statep[2] == states[2]
    assert ls == lp
    assert np.array_equal(ms, mp) and np.array_equal(ss, sp)
    # frozen-structure re-evaluation after a streamed construction takes the plain device path
    import pymra_b200.MRATools as mt
    l2 = float(ts.refit(cov=lambda a, b: mt.Matern32(a, b, l=0.3, sig=1.0))[0, 0])
    assert l2 == lp
    m2, s2 = ts.predict()
    assert np.array_equal(np.asarray(m2).ravel(), mp) and np.array_equal(s2, sp)


def test_stream_parts_in_any_order_and_call_order_errors():
    import pymra_b200.MRATools as mt
    from pymra_b200 import _ffi
    from pymra_b200.covariance import introspect
    from pymra_b200.session import DeviceSession
    from pymra_b200.structure import build_structure
    locs, y = make(200)
    np.random.seed(7)
    st = build_structure(locs, 16, 3, 4, 9)
    cov = introspect(lambda a, b: mt.ExpCovFun(a, b, l=0.3), 2)
    s = DeviceSession(st, locs, y)
    s.set_params(cov, 1e-2)
    want = s.likelihood()
    wm, ws = s.predict()
    wm, ws = wm.copy(), ws.copy()
    assert s.n_parts() == 4
    with pytest.raises(_ffi.MraError):
        s.stream_part(0)                       # begin first
    s.stream_begin()
    for part in (2, 0, 3):
        s.stream_part(part)
    with pytest.raises(_ffi.MraError):
        s.stream_part(3)                       # twice
    with pytest.raises(_ffi.MraError):
        s.stream_end()                         # part 1 missing
    s.stream_part(1)
    s.stream_end()
    assert s.fetch_likelihood() == want
    gm, gs = s.predict()
    assert np.array_equal(gm, wm) and np.array_equal(gs, ws)
    s.close()


@pytest.mark.parametrize("world", [2, 4, 8, 16])
def test_emulated_shards_stream_their_parts(world):
    """Handles sharded at level 1 (2-4 GPUs) or 2 (8-16 GPUs) stream the parts they own: same summaries,
    likelihood and predictions as the plain sharded pass; another rank's part is refused."""
    import torch
    import pymra_b200.MRATools as mt
    from pymra_b200 import _ffi
    from pymra_b200.covariance import introspect
    from pymra_b200.session import DeviceSession
    from pymra_b200.structure import build_structure
    locs, y = make(200)
    np.random.seed(7)
    st = build_structure(locs, 16, 3, 4, 9)
    cov = introspect(lambda a, b: mt.Matern32(a, b, l=0.3, sig=1.0), 2)
    N = st.N
    sess = []
    for rank in range(world):
        s = DeviceSession(st, locs, y, want_predict=True, emulate=(world, rank))
        s.set_params(cov, 1e-2)
        assert s.shard_level == (1 if world <= 4 else 2) and s.n_parts() == 4
        sess.append(s)

    def finish(total):
        liks, mean, sd = [], torch.zeros(N, dtype=torch.float64, device="cuda"), torch.zeros(N, dtype=torch.float64, device="cuda")
        for s in sess:
            s.summary.copy_(total)
            s.likelihood_top_async()
            liks.append(s.fetch_likelihood())
            m_k, s_k = torch.empty_like(mean), torch.empty_like(sd)
            s.predict_dev(m_k, s_k)
            mean += m_k
            sd += s_k
        return liks, mean.cpu().numpy(), sd.cpu().numpy()

    for s in sess:
        s.likelihood_local_async()
    want_total = torch.stack([s.summary for s in sess]).sum(0)
    want = finish(want_total.clone())
    for rank, s in enumerate(sess):
        s.stream_begin()
        mine = s.my_parts()
        assert mine == ([p for p in range(4) if p % world == rank] if world <= 4 else
                        sorted({(k // 4) for k in range(16) if k % world == rank}))
        for other in set(range(4)) - set(mine):
            with pytest.raises(_ffi.MraError):
                s.stream_part(other)                       # no subtree of mine in there
        for part in reversed(mine):
            s.stream_part(part)
        s.stream_end_local()
    got_total = torch.stack([s.summary for s in sess]).sum(0)
    assert torch.equal(got_total, want_total)
    got = finish(got_total)
    assert got[0] == want[0] and all(l == got[0][0] for l in got[0])
    assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
    # a plain pass after the streamed one (what refit() runs) needs the knots of every replicated top node
    for s in sess:
        s.sync_knots()
        s.likelihood_local_async()
    assert torch.equal(torch.stack([s.summary for s in sess]).sum(0), want_total)
    again = finish(want_total.clone())
    assert again[0] == want[0] and np.array_equal(again[1], want[1]) and np.array_equal(again[2], want[2])
    for s in sess:
        s.close()


def test_two_part_setup_call_order_and_equivalence():
    """mra_plan_tree / mra_bind_tree / mra_plan_obs / mra_bind_obs: the prior levels may run before the observation half
    is bound, everything else is refused until then; results equal the one-piece set-up bit for bit, also for a refit
    (plain pass on the split arena) and for parts whose prior levels ran early in any order."""
    import torch
    import pymra_b200.MRATools as mt
    from pymra_b200 import _ffi
    from pymra_b200.covariance import introspect
    from pymra_b200.session import DeviceSession
    from pymra_b200.structure import build_structure
    locs, y = make(200)
    np.random.seed(7)
    st = build_structure(locs, 16, 3, 4, 9)
    cov = introspect(lambda a, b: mt.Matern32(a, b, l=0.3, sig=1.0), 2)
    ref = DeviceSession(st, locs, y)
    ref.set_params(cov, 1e-2)
    want = ref.likelihood()
    wm, ws = ref.predict()
    wm, ws = wm.copy(), ws.copy()
    ref.close()
    dev = torch.device("cuda")
    staged = (torch.from_numpy(np.ascontiguousarray(locs)).to(dev), torch.from_numpy(np.ascontiguousarray(y).reshape(-1)).to(dev))
    s = DeviceSession(st, locs, y, staged=staged, two_part=True)
    assert s._pending_obs is not None and s.ws_obs is None
    s.set_params(cov, 1e-2)
    s.stream_begin()
    with pytest.raises(_ffi.MraError):
        s.stream_part(0)                       # leaf terms need the observation half
    s.stream_part_prior(2)
    s.stream_part_prior(0)
    with pytest.raises(_ffi.MraError):
        s.stream_part_prior(2)                 # twice
    s.finish_plan()
    assert s._pending_obs is None and s.ws_obs is not None
    for part in (3, 0, 2, 1):
        s.stream_part(part)
    s.stream_end()
    assert s.fetch_likelihood() == want
    gm, gs = s.predict()
    assert np.array_equal(gm, wm) and np.array_equal(gs, ws)
    assert s.likelihood() == want              # plain pass on the two allocations
    s.close()
