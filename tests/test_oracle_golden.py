"""Pins the oracle (oracle/mra_oracle.py) to outputs of the unmodified reference (tests/golden)."""
import numpy as np
import pytest

from _util import errs, golden_names, golden_structure, load_golden, oracle_for

# relative likelihood, absolute mean, relative sd: the reference's own FP64 noise for the
# ill-conditioned fixtures (Exp l=2 with R=1e-4, Matern32 kappa=0.3) is far above 1e-9, see
# SURVEY.md section 0 finding 9; these bounds are ~3x the measured oracle-vs-reference gaps.
LOOSE = {"ka4_large_m3": (3e-8, 1e-5, 1e-4), "readme_literal_small": (2e-8, 5e-7, 1e-2),
         "g33x47_m32": (5e-9, 1e-7, 2e-3), "g48_m32": (5e-9, 1e-7, 2e-3), "g64_allobs": (1e-9, 5e-8, 5e-4),
         "g96_m32_r16": (1e-9, 5e-8, 5e-4), "g125_m32_r16": (1e-9, 5e-8, 1e-4), "m0_dense": (1e-9, 1e-8, 2e-7),
         # smooth kernels: the reference's inv()-based recursion is noisier still
         "g48_m52": (3e-9, 3e-7, 5e-4), "g48_gauss": (1e-8, 5e-7, 1.5e-3), "g32_basis": (3e-9, 1e-8, 1e-4)}


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference(name):
    g = load_golden(name)
    o = oracle_for(g, record=True)
    assert o["M"] == int(g["M_eff"]) and o["J"] == int(g["J_eff"])
    rl, em, es = errs(o["lik"], o["mean"], o["sd"], g)
    tl, tm, ts = LOOSE.get(name, (1e-9, 1e-8, 1e-7))
    assert rl < tl and em < tm and es < ts, (rl, em, es)
    assert np.array_equal(o["root_kinds"], g["root_kinds"])
    if "node_ids" in g:
        gs = golden_structure(g)
        assert len(gs) == len(o["nodes"])
        for nd in o["nodes"]:
            rows, kinds, leaf = gs[nd["ID"]]
            assert np.array_equal(rows, nd["rows"]) and np.array_equal(kinds, nd["kInds"]) and leaf == nd["leaf"]


def test_known_answers_from_survey():
    """SURVEY.md App. C spot values recorded by the surveyor from the live reference."""
    ka = {"ka1m": (-99.3672725177306, [1.714032818067426, 0.050449383352438, 0.221944426002978]),
          "ka1e": (-65.5313750147362, [1.650231512616136, 0.123399278816402, 0.245296065997037])}
    for name, (lik, xp) in ka.items():
        g = load_golden(name)
        assert abs(float(g["lik"]) - lik) < 1e-9
        assert np.allclose(g["mean"][[0, 49, 99]], xp, atol=1e-12)
        o = oracle_for(g)
        assert abs(o["lik"] - lik) < 1e-9 * abs(lik)
    assert abs(float(load_golden("ka2_small")["lik"]) - 966.798500451365) < 1e-8
    assert abs(float(load_golden("ka4_large_serial")["lik"]) - 197006.169579) < 1e-5
    assert abs(float(load_golden("ka4_large_crit0")["lik"]) - 197151.738123) < 1e-5


def test_oracle_logdet_det_mode_matches_slogdet_when_finite():
    g = load_golden("g50_exp")
    a = oracle_for(g)
    b = oracle_for(g, logdet="det")
    assert abs(a["lik"] - b["lik"]) < 1e-9 * abs(a["lik"])
