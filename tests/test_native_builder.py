"""The C++ tree builder (csrc/mra_structure.cpp) is bit-identical to the NumPy builder, which is
itself bit-identical to the reference (tests/test_structure.py): same nodes, rows, knots, knot
order, permutation, and the same final state of the global NumPy RNG."""
import numpy as np
import pytest

import pymra_b200.MRATools as mt
from pymra_b200.structure import build_structure, build_structure_native

CASES = [(40, 40, 8, 2, 5), (50, 50, 16, 2, 6), (33, 47, 5, 2, 7), (64, 64, 8, 3, 8), (96, 96, 16, 3, 12),
         (125, 125, 16, 4, 5), (201, 157, 32, 3, 9), (128, 128, 64, 2, 1), (300, 300, 16, 5, 2),
         (400, 380, 64, 3, 4), (512, 512, 128, 2, 13), (700, 611, 32, 4, 21)]   # the last three take the reverse-traced selection and the threaded two-phase build (cooperative path >= 2^18 rows)


def same(a, b):
    assert a.n_nodes == b.n_nodes and a.depth == b.depth
    for f in ("perm", "node_level", "node_parent", "node_kind", "node_row_start", "node_row_count",
              "node_child_start", "node_child_count", "node_knot_off", "knot_rows", "level_off"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    for n in range(a.n_nodes):
        assert a.node_id[n] == b.node_id[n]
        assert np.array_equal(a.node_kinds_local[n], b.node_kinds_local[n]), n


@pytest.mark.parametrize("nx,ny,r,M,seed", CASES)
@pytest.mark.parametrize("crit", [-1, 0, 1])
def test_native_equals_numpy(nx, ny, r, M, seed, crit):
    locs = mt.genLocations2d(nx, Ny=ny)
    cd = M + 1 if crit < 0 else crit
    np.random.seed(seed)
    a = build_structure(locs, r, M, 4, cd, native=False)
    sa = np.random.get_state()
    np.random.seed(seed)
    b = build_structure_native(locs, r, M, 4, cd)
    sb = np.random.get_state()
    if b is None:
        # the NumPy builder met a small node (KMeans path) -> native must have declined and left the RNG alone
        np.random.seed(seed)
        s0 = np.random.get_state()
        np.random.seed(seed)
        assert build_structure_native(locs, r, M, 4, cd) is None
        assert np.array_equal(np.random.get_state()[1], s0[1])
        assert any(a.node_row_count[n] <= 100 or True for n in range(a.n_nodes))
        return
    same(a, b)
    assert np.array_equal(sa[1], sb[1]) and sa[2] == sb[2]


def test_native_on_scattered_points():
    rng = np.random.RandomState(0)
    locs = rng.uniform(size=(5000, 2))
    np.random.seed(3)
    a = build_structure(locs, 10, 2, 4, 3, native=False)
    np.random.seed(3)
    b = build_structure_native(locs, 10, 2, 4, 3)
    assert b is not None
    same(a, b)


def test_legacy_choice_clone_against_numpy():
    """np.random.choice(np.arange(n), r, replace=False) for the root == sorted knots of a depth-1 tree."""
    for seed, n, r in [(1, 20, 3), (2, 37, 8), (11, 64, 16), (5, 101, 64), (7, 450, 64), (9, 300, 128)]:
        locs = mt.genLocations2d(n)
        np.random.seed(seed)
        want = np.sort(np.random.choice(np.arange(n * n), size=r, replace=False))
        after = np.random.get_state()
        np.random.seed(seed)
        st = build_structure_native(locs, r, 0 + 1, 4, 5)
        if st is None:
            continue
        assert np.array_equal(st.node_kinds_local[0], want)


def test_threaded_path_on_scattered_points_and_fallback():
    """>= 2^16 scattered points: regular tree -> threaded two-phase build; one level deeper the leaves'
    parents fall under 100 rows -> both native paths decline and leave the RNG untouched."""
    rng = np.random.RandomState(1)
    locs = rng.uniform(size=(80000, 2))
    np.random.seed(8)
    a = build_structure(locs, 10, 5, 4, 6, native=False)
    sa = np.random.get_state()
    np.random.seed(8)
    b = build_structure_native(locs, 10, 5, 4, 6)
    sb = np.random.get_state()
    assert b is not None
    same(a, b)
    assert np.array_equal(sa[1], sb[1]) and sa[2] == sb[2]
    np.random.seed(8)
    s0 = np.random.get_state()
    assert build_structure_native(locs, 10, 6, 4, 7) is None
    s1 = np.random.get_state()
    assert np.array_equal(s0[1], s1[1]) and s0[2] == s1[2]


@pytest.mark.parametrize("n,r,M,crit", [(300, 8, 3, 9), (300, 8, 3, 1), (400, 16, 4, 0), (512, 32, 5, 9)])
def test_stream_build_equals_native(n, r, M, crit):
    """mra_build_stream_*: same arrays and RNG consumption as mra_build_structure_2d, and every event's promise
    (root knots + permutation at event 0, the knots of subtree c at event 1 + c) holds when it fires."""
    from pymra_b200.structure import StreamBuild
    locs = mt.genLocations2d(n)
    np.random.seed(5)
    a = build_structure_native(locs, r, M, 4, crit)
    state_a = np.random.get_state()
    np.random.seed(5)
    sb = StreamBuild(locs, r, M, 4, crit)
    assert sb.started
    assert sb.wait(0)
    st = sb.structure
    assert np.array_equal(st.perm, a.perm)
    assert np.array_equal(st.node_row_start, a.node_row_start) and np.array_equal(st.node_kind, a.node_kind)
    assert np.array_equal(st.knot_rows[:r], a.knot_rows[:r])
    for c in range(4):
        assert sb.wait(1 + c)
        for L in range(1, M):
            lo = int(st.level_off[L]) + c * 4 ** (L - 1)
            hi = lo + 4 ** (L - 1)
            assert np.array_equal(st.knot_rows[lo * r:hi * r], a.knot_rows[lo * r:hi * r])
    assert sb.wait(5)
    assert sb.finish()
    state_b = np.random.get_state()
    same(a, sb.structure)
    assert np.array_equal(state_a[1], state_b[1]) and state_a[2] == state_b[2]


def test_stream_build_refuses_ragged_tree_and_small_inputs():
    from pymra_b200.structure import StreamBuild
    np.random.seed(8)
    before = np.random.get_state()
    assert not StreamBuild(mt.genLocations2d(100), 8, 2, 4, 9).started          # N < 65536: no job
    rng = np.random.RandomState(1)
    locs = rng.uniform(size=(80000, 2))
    sb = StreamBuild(locs, 10, 6, 4, 7)            # leaves' parents fall under 100 rows (see the test above)
    assert sb.started
    assert not all(sb.wait(e) for e in range(6))
    assert not sb.finish()
    after = np.random.get_state()
    assert np.array_equal(before[1], after[1]) and before[2] == after[2]
    sb = StreamBuild(locs, 10, 5, 4, 6)            # one level less: regular, and equal to the NumPy builder
    assert sb.started and all(sb.wait(e) for e in range(6)) and sb.finish()
    np.random.seed(8)
    same(build_structure(locs, 10, 5, 4, 6, native=False), sb.structure)
