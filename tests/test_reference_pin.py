"""Pins the CPU restatement (oracle/) to the reference's OWN transcription code.

`RefOracle` drives /root/reference/Lpopc/src/Core/{LpNLPWrapper,LpHessian,LpFiniteDifferenceDerive,
RPMGenerator,LpBoundsChecker,LpSizeChecker,LpGuessChecker,LpDerivDependciesChecker}.cpp and
SparseMatrix/*.cpp, compiled unmodified against the Armadillo stand-in of oracle/ref_shim/ into
oracle/_ref/liblpopc_ref.so (oracle/ref_build.mk).  Integers must be equal; values within 1e-12
relative (Armadillo's dense-product summation order is the one thing the stand-in cannot pin).
Skipped when neither /root/reference nor a prebuilt oracle/_ref exists.
"""
import ctypes as C

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


@pytest.mark.parametrize("name,new", [("hypersensitive/u8x4", "hypersensitive/ragged"), ("bryson_denham/u7x6", "bryson_denham/ragged"),
                                      ("launch/u5x4", "launch/ragged")])
def test_guess_transfer_between_grids_matches_the_reference(name, new):
    """Warm start of the mesh loop: the solution on one grid becomes the user guess (time, states, controls with the
    spline end row; Nlp2OPConverter.cpp:160-193) and the reference's GetGuess interpolates it onto the next grid with its
    natural cubic spline (LpGuessChecker.cpp:130-190, 208-294).  lpopc_b200.adaptive.transfer_guess must land on the
    same starting point."""
    from lpopc_b200 import adaptive
    op = cases.build(name)
    r = RefOracle(op)
    _, x, _, lam = cases.inputs(op, r, 13)
    res, _ = r.nlp2op(x, lam)  # the reference's own converted solution: time grid, states, controls incl. the end row
    old_pts = [r.tables(ip)["points"] for ip in range(len(op.phases))]
    op_new = cases.build(new)
    r_new = RefOracle(op_new)
    for ip, q in enumerate(res):
        t = np.ascontiguousarray(q["time"])
        xs = np.ascontiguousarray(q["state"].T).reshape(-1)
        us = np.ascontiguousarray(q["control"].T).reshape(-1) if q["control"].size else np.zeros(1)
        r_new._check(r_new.L.lpo_set_guess(r_new.h, C.c_int(ip), C.c_int(t.size), t.ctypes.data_as(C.POINTER(C.c_double)),
                                           xs.ctypes.data_as(C.POINTER(C.c_double)), us.ctypes.data_as(C.POINTER(C.c_double))))
    r_new.refresh()
    ref_start = r_new.guess()
    new_pts = [r_new.tables(ip)["points"] for ip in range(len(op.phases))]
    mine = adaptive.transfer_guess(op, x, old_pts, new_pts, [q["control"][-1] if q["control"].size else np.zeros(0) for q in res])
    assert mine.shape == ref_start.shape
    assert np.max(np.abs(mine - ref_start) / np.maximum(1.0, np.abs(ref_start))) <= 1e-13
