import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    config.addinivalue_line("markers", "slow: minutes of host CPU (oracle port at BASELINE sizes); needs MRA_RUN_SLOW=1")


def pytest_collection_modifyitems(config, items):
    if os.environ.get("MRA_RUN_SLOW") == "1":
        return
    skip = pytest.mark.skip(reason="slow: set MRA_RUN_SLOW=1")
    for item in items:
        if "slow" in item.keywords:
            item.add_marker(skip)


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz"))


def load_golden(name):
    import numpy as np
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: g[k] for k in g.files}


def load_truth(name):
    """Extended-precision dense ground truth of a fixture (oracle/make_truth.py)."""
    import numpy as np
    t = np.load(os.path.join(GOLDEN, "truth", name + ".npz"))
    return {k: t[k] for k in t.files}


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
