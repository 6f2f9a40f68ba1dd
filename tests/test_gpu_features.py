"""SURVEY.md 8f.3 / 8f.4 on the GPU: `cov` given as a dense N x N np.matrix, and the opt-in export of the per-node
basis matrices behind MRATree.getBasisFunctionsMatrix -- against the unmodified reference's recorded outputs
(tests/golden/g40_dense, b1d_basis, g32_basis), the dense truth and the oracle port's per-node matrices."""
import numpy as np
import pytest
import scipy.linalg

from _util import errs, load_golden, load_truth, oracle_for, tree_for

pytestmark = pytest.mark.gpu


def test_dense_matrix_cov_matches_truth_and_reference():
    g = load_golden("g40_dense")
    T = load_truth("g40_dense")
    t = tree_for(g)                                    # cov is an np.matrix here (tests/_util.py dense_recipe)
    assert t._cov.name == "dense"
    lik = float(np.asarray(t.getLikelihood()).ravel()[0])
    mean, sd = t.predict()
    ref = errs(float(g["lik"]), g["mean"], g["sd"], T)
    mine = errs(lik, mean, sd, T)
    for a, b in zip(mine, ref):
        assert a <= max(1e-9, b), (mine, ref)
    assert np.array_equal(t.root.kInds, g["root_kinds"])
    assert np.allclose(np.asarray(t.root.B), g["bf_prior"], rtol=0, atol=1e-15)      # root.B is a column slice of cov
    # refit with the same matrix object: stays resident, same numbers
    l2 = float(np.asarray(t.refit(cov=t._cov_closure)).ravel()[0])
    assert l2 == lik


@pytest.mark.parametrize("name", ["b1d_basis", "g32_basis", "g40_dense"])
def test_root_basis_functions_match_reference(name):
    """What the unmodified reference's getBasisFunctionsMatrix returns on a finished tree: the root's blocks."""
    g = load_golden(name)
    t = tree_for(g)
    scale = float(np.max(np.abs(g["bf_prior"])))
    Bp = np.asarray(t.getBasisFunctionsMatrix(distr="prior"))
    assert Bp.shape == g["bf_prior"].shape and np.max(np.abs(Bp - g["bf_prior"])) <= 1e-12 * scale
    Bk = np.asarray(t.getBasisFunctionsMatrix(distr="prior", timesKC=True))
    assert np.max(np.abs(Bk - g["bf_prior_kc"])) <= 1e-9 * float(np.max(np.abs(g["bf_prior_kc"])))
    Bt = np.asarray(t.getBasisFunctionsMatrix(distr="posterior"))
    assert Bt.shape == g["bf_post"].shape
    assert np.max(np.abs(Bt - g["bf_post"])) <= 1e-8 * float(np.max(np.abs(g["bf_post"])))
    # B~ kTilC is defined up to the eigenvector signs / rotations of k~: compare the Gram (= B~ k~ B~^T)
    Btk = np.asarray(t.getBasisFunctionsMatrix(distr="posterior", timesKC=True))
    G1, G0 = Btk @ Btk.T, g["bf_post_kc"] @ g["bf_post_kc"].T
    assert np.max(np.abs(G1 - G0)) <= 1e-8 * float(np.max(np.abs(G0)))
    # the diagnostics leave the tree usable: same predictions as a fresh tree
    m0, s0 = tree_for(g).predict()
    m1, s1 = t.predict()
    assert np.array_equal(np.asarray(m0), np.asarray(m1)) and np.array_equal(s0, s1)


@pytest.mark.parametrize("name", ["b1d_basis", "g32_basis"])
def test_all_level_basis_matches_oracle_nodes(name):
    """all_levels=True: every resolution, assembled as MRATree.py:445-511 is written, against the per-node matrices
    the oracle port records before it frees them."""
    g = load_golden(name)
    tol = 1e-7 if str(g["family"]) == "exp" else 5e-6      # Matern32: the port's inv()-noise (its lik is 7e-10 off, sd 3e-5)
    o = oracle_for(g, record="full")
    nodes = o["nodes"]
    depth = max(len(n["ID"]) for n in nodes) - 1
    t = tree_for(g)
    for distr, key, fac in (("prior", "B", "kC"), ("posterior", "BTil", "kTilC")):
        got = t.getBasisFunctionsMatrix(distr=distr, groupByResolution=True, all_levels=True)
        gotk = t.getBasisFunctionsMatrix(distr=distr, groupByResolution=True, all_levels=True, timesKC=True)
        assert len(got) == depth + 1
        for lv in range(depth + 1):
            lvl = [n for n in nodes if len(n["ID"]) - 1 == lv]
            lvl = [lvl[i] for i in np.argsort([np.min(g["locs"][n["rows"], 0]) for n in lvl])] if lv else lvl
            want = scipy.linalg.block_diag(*[n[key] for n in lvl])
            wantk = scipy.linalg.block_diag(*[n[key] @ n[fac] for n in lvl])
            G = np.asarray(got[lv])
            assert G.shape == want.shape, (distr, lv, G.shape, want.shape)
            sc = max(1e-300, float(np.max(np.abs(want))))
            assert np.max(np.abs(G - want)) <= tol * sc, (distr, lv, np.max(np.abs(G - want)) / sc)
            Gk = np.asarray(gotk[lv])
            assert np.max(np.abs(Gk @ Gk.T - wantk @ wantk.T)) <= tol * max(1e-300, float(np.max(np.abs(wantk @ wantk.T))))
