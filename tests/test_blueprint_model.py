"""The NumPy blueprint of the device algorithm (tests/_model.py: whitened basis + dual leaf
elimination on the flat TreeStructure) reproduces the reference outputs."""
import pytest

from _model import model_run
from _util import errs, golden_names, load_golden, structure_for
from test_oracle_golden import LOOSE


@pytest.mark.parametrize("name", golden_names())
def test_blueprint_matches_reference(name):
    g = load_golden(name)
    if str(g["family"]) == "dense":
        pytest.skip("the blueprint evaluates covariance closures; the matrix form is covered by tests/test_gpu_features.py")
    st = structure_for(g)
    o = model_run(st, g["locs"], g["obs"], str(g["family"]), float(g["l"]), float(g["sig"]), float(g["R"]))
    rl, em, es = errs(o["lik"], o["mean"], o["sd"], g)
    tl, tm, ts = LOOSE.get(name, (1e-9, 1e-8, 1e-7))
    assert rl < tl and em < tm and es < ts, (rl, em, es)
