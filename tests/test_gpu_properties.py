"""Size-independent properties of the MRA posterior checked at BASELINE.json's full size (cfg5: 2000 x 2000
grid, 4M locations, r0 = 64, M 10 -> 7), where neither the reference nor the oracle can run in test time:
  * the predictive mean is linear in the observations, the predictive sd and the log-determinant part d of
    the likelihood do not depend on their values, the quadratic part u scales with the square of y;
  * with no observations at all the posterior is the MRA prior: mean 0, sd^2 = C(0) - sum_levels |V row|^2
    <= C(0), likelihood 0;
  * two constructions from the same RNG state are bitwise identical (deterministic kernels)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_SIDE, R0, M_REQ = 2000, 64, 10


@pytest.fixture(scope="module")
def setup():
    import pymra_b200.MRATools as mt
    locs = mt.genLocations2d(N_SIDE)
    N = len(locs)
    rng = np.random.RandomState(3)
    sel = np.sort(rng.choice(N, int(0.4 * N), replace=False))
    y1 = np.full((N, 1), np.nan)
    y2 = np.full((N, 1), np.nan)
    y1[sel] = np.sin(5.0 * locs[sel, :1]) + 0.3 * rng.normal(size=(len(sel), 1))
    y2[sel] = np.cos(3.0 * locs[sel, 1:]) + 0.3 * rng.normal(size=(len(sel), 1))
    cov = lambda a, b: mt.Matern32(a, b, l=0.3, sig=1.0)
    return locs, y1, y2, cov


def run(locs, cov, obs):
    from pymra_b200.MRATree import MRATree
    np.random.seed(5)                      # same knots for every construction
    t = MRATree(locs, R0, cov, obs, 1e-2, M=M_REQ)
    mean, sd = t.predict()
    out = (t._d, t._u, np.asarray(mean).ravel().copy(), sd.copy())
    del t
    return out


def test_linearity_and_value_independence_at_full_size(setup):
    locs, y1, y2, cov = setup
    a, b = 0.7, -1.9
    d1, u1, m1, s1 = run(locs, cov, y1)
    d2, u2, m2, s2 = run(locs, cov, y2)
    d3, u3, m3, s3 = run(locs, cov, a * y1 + b * y2)
    d4, u4, m4, s4 = run(locs, cov, 3.0 * y1)
    scale = max(np.max(np.abs(m1)), np.max(np.abs(m2)), 1.0)
    assert np.max(np.abs(m3 - (a * m1 + b * m2))) <= 1e-9 * scale
    assert np.array_equal(s1, s2) and np.array_equal(s1, s3)          # sd never sees the values
    assert d1 == d2 == d3 == d4                                       # nor does the log-determinant
    assert abs(u4 - 9.0 * u1) <= 1e-10 * abs(u4)
    assert np.max(np.abs(m4 - 3.0 * m1)) <= 1e-9 * scale
    assert np.all(np.isfinite(m1)) and np.all(s1 > 0) and np.all(s1 <= 1.0 + 1e-12)
    # observed locations are pulled towards their data: posterior sd there is below the nugget-free prior sd
    obs_rows = np.flatnonzero(np.isfinite(y1.ravel()))
    assert np.median(s1[obs_rows]) < np.median(s1)


def test_repeatability_and_empty_data_at_full_size(setup):
    locs, y1, y2, cov = setup
    r1 = run(locs, cov, y1)
    r2 = run(locs, cov, y1)
    assert r1[0] == r2[0] and r1[1] == r2[1]
    assert np.array_equal(r1[2], r2[2]) and np.array_equal(r1[3], r2[3])
    d0, u0, m0, s0 = run(locs, cov, np.full_like(y1, np.nan))
    assert d0 == 0.0 and u0 == 0.0
    assert np.all(m0 == 0.0)
    assert np.all(s0 >= 0.0) and np.all(s0 <= 1.0 + 1e-12)
    assert np.all(s0 >= r1[3] - 1e-12)                                # data can only reduce the variance
