"""Deterministic elementary functions (include/lpb_detmath.h) vs numpy: a few ulp, NaN/inf handling.
They replace libm inside functors so that host and device evaluate user functions bit-identically."""
import numpy as np
import pytest

from oracle_lib import detmath

FUNS = {0: np.exp, 1: np.tanh, 2: np.sin, 3: np.cos, 4: np.arccos}


def ulp_err(y, ref):
    return np.abs(y - ref) / np.spacing(np.abs(ref))


@pytest.mark.parametrize("which,lo,hi,tol", [(0, -700, 700, 2), (1, -25, 25, 4), (2, -1e3, 1e3, 2), (3, -1e3, 1e3, 2), (4, -1, 1, 2)])
def test_accuracy(which, lo, hi, tol):
    rng = np.random.Generator(np.random.PCG64(which))
    x = np.concatenate([rng.uniform(lo, hi, 200000), rng.uniform(-1, 1, 50000) * min(1.0, hi), [0.0, lo, hi]])
    y, ref = detmath(which, x), FUNS[which](x)
    ok = np.isfinite(ref) & (np.abs(ref) > 1e-300)
    assert np.max(ulp_err(y[ok], ref[ok])) <= tol


def test_special_values():
    nan, inf = np.nan, np.inf
    assert np.isnan(detmath(2, np.array([nan, inf, -inf, 2e6]))).all()      # sin: NaN outside |x| < 1e6
    assert np.isnan(detmath(3, np.array([nan, inf, -inf, -2e6]))).all()
    assert np.array_equal(detmath(0, np.array([-1e4, 1e4])), [0.0, inf])
    assert np.isnan(detmath(0, np.array([nan]))).all()
    assert np.isnan(detmath(1, np.array([nan]))).all()
    assert np.array_equal(detmath(1, np.array([inf, -inf, 0.0])), [1.0, -1.0, 0.0])
    assert np.isnan(detmath(4, np.array([1.5, -1.5, nan]))).all()
    assert detmath(4, np.array([1.0]))[0] == 0.0
