"""Pins the CPU restatement (oracle/) to the reference's OWN transcription code.

`RefOracle` drives /root/reference/Lpopc/src/Core/{LpNLPWrapper,LpHessian,LpFiniteDifferenceDerive,
RPMGenerator,LpBoundsChecker,LpSizeChecker,LpGuessChecker,LpDerivDependciesChecker}.cpp and
SparseMatrix/*.cpp, compiled unmodified against the Armadillo stand-in of oracle/ref_shim/ into
oracle/_ref/liblpopc_ref.so (oracle/ref_build.mk).  Integers must be equal; values within 1e-12
relative (Armadillo's dense-product summation order is the one thing the stand-in cannot pin).
Skipped when neither /root/reference nor a prebuilt oracle/_ref exists.
"""
import numpy as np
import pytest

import cases
from oracle_lib import Oracle, RefOracle, build_reference

pytestmark = pytest.mark.skipif(build_reference() is None, reason="reference library unavailable")
RTOL = 1e-12


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b))
    return float(np.nanmax(np.abs(a - b) / np.maximum(1.0, np.abs(b)))) if a.size else 0.0


@pytest.mark.parametrize("name", cases.CASES + ["launch/u5x4", "hypersensitive/u40x3", "bryson_denham/u7x6"])
def test_restatement_equals_reference(name):
    op = cases.build(name)
    o, r = Oracle(op), RefOracle(op)
    assert o.nlp_info() == r.nlp_info()
    for a, b in zip(o.jac_structure(), r.jac_structure()):
        assert np.array_equal(a, b)
    for a, b in zip(o.h_structure(), r.h_structure()):
        assert np.array_equal(a, b)
    for a, b in zip(o.bounds(), r.bounds()):
        assert np.array_equal(a, b)
    for ip in range(len(op.phases)):
        to, tr = o.tables(ip), r.tables(ip)
        assert np.array_equal(to["points"], tr["points"]) and np.array_equal(to["weights"], tr["weights"])
        for k in ("D", "Diag", "Doffdiag"):
            for a, b in zip(to[k], tr[k]):
                assert np.array_equal(a, b), k
    guess, x, sigma, lam = cases.inputs(op, o, 7)
    # the two-point guess of the problem mirror = the reference's spline-interpolated guess
    assert rel(guess, r.guess()) <= 1e-14
    for xv in (guess, x):
        assert abs(o.eval_f(xv) - r.eval_f(xv)) <= RTOL * max(1.0, abs(r.eval_f(xv)))
        assert rel(o.eval_grad_f(xv), r.eval_grad_f(xv)) <= RTOL
        assert rel(o.eval_g(xv), r.eval_g(xv)) <= RTOL
        assert rel(o.eval_jac_g(xv), r.eval_jac_g(xv)) <= RTOL
    ho, hr = o.eval_h(x, sigma, lam), r.eval_h(x, sigma, lam)
    assert rel(ho, hr) <= 1e-9  # second differences / h^2: noise-dominated, see test_gpu_parity
    assert float(np.mean(ho == hr)) > 0.9


def test_dependency_probe_equals_reference():
    op = cases.build("launch")
    o, r = Oracle(op), RefOracle(op)
    guess, x, sigma, lam = cases.inputs(op, o, 5)
    assert np.array_equal(o.probe_dependencies(guess), r.probe_dependencies(guess))
    assert o.nlp_info() == r.nlp_info()
    for a, b in zip(o.h_structure(), r.h_structure()):
        assert np.array_equal(a, b)
    assert rel(o.eval_h(x, sigma, lam), r.eval_h(x, sigma, lam)) <= 1e-9
