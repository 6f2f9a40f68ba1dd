"""Host structure builder (pymra_b200/structure.py) vs the reference's recorded tree: bit-exact."""
import numpy as np
import pytest

from _util import golden_names, golden_structure, load_golden, structure_for
from pymra_b200.structure import KIND_INTERNAL, KIND_LEAF, KIND_ORPHAN, build_structure

WITH_STRUCT = [n for n in golden_names() if "node_ids" in load_golden(n)]


@pytest.mark.parametrize("name", WITH_STRUCT)
def test_structure_bit_exact(name):
    g = load_golden(name)
    st = structure_for(g)
    gs = golden_structure(g)
    seen = 0
    for n in range(st.n_nodes):
        if st.node_kind[n] == KIND_ORPHAN:
            continue
        seen += 1
        rows, kinds, leaf = gs[st.node_id[n]]
        mine = st.perm[st.node_row_start[n]: st.node_row_start[n] + st.node_row_count[n]]
        assert np.array_equal(np.sort(mine), rows)
        assert np.array_equal(st.node_kinds_local[n], kinds)
        assert (st.node_kind[n] == KIND_LEAF) == leaf
        if leaf:
            assert np.array_equal(mine, rows)          # leaves keep ascending original order
        else:
            K = st.knot_rows[st.node_knot_off[n]: st.node_knot_off[n] + st.r]
            assert np.array_equal(st.perm[K], rows[kinds])   # knot order == reference kInds order
    assert seen == len(gs)


@pytest.mark.parametrize("name", WITH_STRUCT)
def test_structure_invariants(name):
    st = structure_for(load_golden(name))
    assert sorted(st.perm.tolist()) == list(range(st.N))
    assert st.node_level[0] == 0 and st.node_parent[0] == -1
    for n in range(st.n_nodes):
        if st.node_kind[n] == KIND_INTERNAL:
            cs, cc = st.node_child_start[n], st.node_child_count[n]
            assert cc > 0
            assert st.node_row_start[cs] == st.node_row_start[n]
            assert st.node_row_count[cs:cs + cc].sum() == st.node_row_count[n]
            assert np.all(st.node_parent[cs:cs + cc] == n)
            assert np.all(st.node_level[cs:cs + cc] == st.node_level[n] + 1)
            K = st.knot_rows[st.node_knot_off[n]: st.node_knot_off[n] + st.r]
            assert np.all((K >= st.node_row_start[n]) & (K < st.node_row_start[n] + st.node_row_count[n]))
    assert np.all(np.diff(st.node_level) >= 0)


def test_fork_mode_rng_semantics():
    """critDepth forks (MRANode.py:90-95): children share the parent's RNG state, the parent's
    state is not advanced by them."""
    g = load_golden("ka4_large_crit0")
    np.random.seed(int(g["seed"]))
    before = np.random.get_state()[1].copy()
    st = structure_for(g)
    np.random.seed(int(g["seed"]))
    root_draw = np.random.choice(np.arange(st.N), size=st.r, replace=False)
    after_root = np.random.get_state()
    structure_for(g)
    s2 = np.random.get_state()
    assert np.array_equal(s2[1], after_root[1]) and s2[2] == after_root[2]
    assert np.array_equal(np.sort(root_draw), st.node_kinds_local[0])
    assert not np.array_equal(before, s2[1])


def test_one_dimensional_large_has_orphans_or_not():
    import pymra_b200.MRATools as mt
    locs = mt.genLocations(400)
    np.random.seed(1)
    st = build_structure(locs, 3, 2, 4, 3)
    assert sorted(st.perm.tolist()) == list(range(400))
    # 1-D N>100 splits into three parts by strict inequalities (MRANode.py:220-228)
    assert st.node_child_count[0] in (3, 4)
