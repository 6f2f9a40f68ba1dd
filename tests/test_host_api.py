"""Host logic that needs no GPU: parameter resolution, covariance introspection, C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

import pymra_b200.MRATools as mt
from pymra_b200 import _ffi
from pymra_b200.covariance import introspect
from pymra_b200.MRATree import resolve_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_param_resolution_matches_reference_semantics():
    # MRATree.py:41-48 clamps: cfg3 8->7, cfg4 7->6, cfg5 10->7
    assert resolve_params(1000 * 1000, 2, 32, 8, -1, -1)[:2] == (7, 4)
    assert resolve_params(500 * 500, 2, 16, 7, -1, -1)[:2] == (6, 4)
    M, J, cd, clamped = resolve_params(2000 * 2000, 2, 64, 10, -1, -1)
    assert (M, J, cd, clamped) == (7, 4, 8, (10, 7))
    assert resolve_params(100, 1, 2, 3, 3, 4) == (3, 3, 4, None)
    with pytest.raises(AttributeError):        # MRATree.py:33 '==' bug: 1-D needs explicit J
        resolve_params(100, 1, 2, 3, -1, -1)
    with pytest.raises(OverflowError):         # J == 1 -> log(1) == 0
        resolve_params(100, 1, 2, 0, 1, -1)
    assert resolve_params(10000, 2, 4, 0, -1, -1)[0] == 0      # README literal call binds M=0


def test_generators_match_reference_layout():
    l1 = mt.genLocations(100)
    assert l1.shape == (100, 1) and l1[0, 0] == 0.01 and l1[-1, 0] == 1.0
    l2 = mt.genLocations2d(4, Ny=3)
    assert l2.shape == (12, 2)
    assert np.array_equal(l2[:4, 0], np.linspace(0, 1, 4)) and np.all(l2[:4, 1] == 0)   # x fastest


@pytest.mark.parametrize("d", [1, 2])
def test_cov_introspection(d):
    c = introspect(lambda a, b: mt.Matern32(a, b, l=0.3, sig=1.7), d)
    assert c.name == "matern32" and abs(c.l - 0.3) < 1e-14 and abs(c.sig - 1.7) < 1e-14
    c = introspect(lambda a, b: mt.ExpCovFun(a, b, l=2), d)
    assert c.name == "exp" and abs(c.l - 2) < 1e-13 and c.sig == 1.0
    c = introspect(lambda a, b: 2.5 * mt.ExpCovFun(a, b, l=0.03), d)
    assert c.name == "exp" and abs(c.l - 0.03) < 1e-15 and abs(c.sig - 2.5) < 1e-14
    c = introspect(lambda a, b: mt.Matern52(a, b, l=0.2, sig=0.6), d)
    assert c.name == "matern52" and abs(c.l - 0.2) < 1e-13 and abs(c.sig - 0.6) < 1e-14
    c = introspect(lambda a, b: mt.GaussianCovFun(a, b, l=0.07, sig=1.3), d)
    assert c.name == "gaussian" and abs(c.l - 0.07) < 1e-13 and abs(c.sig - 1.3) < 1e-14
    c = introspect(lambda a, b: np.exp(-np.square(mt.dist(a, b))), d)          # a Gaussian in disguise: 2 l^2 = 1
    assert c.name == "gaussian" and abs(c.l - np.sqrt(0.5)) < 1e-13
    with pytest.raises(ValueError):
        introspect(lambda a, b: 1.0 / (1.0 + np.square(mt.dist(a, b))), d)     # rational quadratic: not supported
    c = introspect(np.matrix(2.0 * np.eye(3)), d, n_locs=3)                    # dense matrix form (MRANode.py:73-75)
    assert c.name == "dense" and c.sig == 2.0 and c.matrix.shape == (3, 3)
    with pytest.raises(ValueError):
        introspect(np.matrix(np.eye(3)), d, n_locs=4)
    with pytest.raises(ValueError):
        introspect(np.matrix(np.ones((3, 2))), d)


def test_capi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "pymra_b200.h")).read()
    declared = set(re.findall(r"\b(mra_[a-z_0-9]+)\s*\(", hdr))
    assert declared and declared == set(_ffi.SIGNATURES), declared ^ set(_ffi.SIGNATURES)
    assert os.path.exists(_ffi.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(_ffi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    lib.mra_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.mra_version()


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _ffi.load_library(str(tmp_path / "nope.so"))


def test_product_never_imports_oracle():
    pk = os.path.join(ROOT, "pymra_b200")
    for dp, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), os.path.join(dp, f)
