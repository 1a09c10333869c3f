"""The reference's OWN example user functions as a second opinion on include/problems/*.h.

Everywhere else both sides of a comparison compile the same functor headers, so a wrong constant, sign or term in
hypersensitive.h / bryson_denham.h / launch.h would be invisible.  oracle/ref_build.mk therefore also compiles the
reference's example programs (Lpopc/example/hypersensitive/HyperSensitive.cpp:74-167, bryson-denham/BrysonDenham.cpp,
launch/Launch.cpp:589-770; main() discarded) into oracle/_ref, oracle/ref_examples.cpp hands their FunctionWrapper
subclasses to the reference's transcription ("ref:<name>"), and oracle/make_golden.py freezes f, grad f, g, the
Jacobian and the Hessian they produce as refex_* in tests/golden.  Here the functor headers -- through the CPU
restatement and through the CUDA path -- are held to those values.

Tolerances: hypersensitive and Bryson-Denham use only + - * / -> bit-identical quotients.  Launch calls exp / acos /
pow through libm in the reference and the deterministic lpb_det_* functions here: function values agree to the last
ulp or two (1e-12 relative on g), and a forward difference amplifies one ulp of f by 1/h = 1e6 (second differences by
1/h^2), so Jacobian entries are compared at 1e-9 and Hessian entries at 1e-3 of the segment scale."""
import ctypes as C
import os

import numpy as np
import pytest

import cases
import golden_lib
import parity
from oracle_lib import Oracle, REF_LIB_PATH, ref_lib

REFEX_CASES = [n for n in golden_lib.GOLDEN_CASES if n.split("/")[0] in ("hypersensitive", "bryson_denham", "launch")]


def scale_rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b))))) if a.size else 0.0


def check_against_refex(name, impl):
    G = golden_lib.load(name)
    x, sigma, lam = G["x"], float(G["sigma"]), G["lam"]
    exact = name.split("/")[0] != "launch"
    assert abs(impl.eval_f(x) - float(G["refex_f"])) <= 1e-12 * max(1.0, abs(float(G["refex_f"])))
    parity.assert_parity(impl.eval_grad_f(x), G["refex_grad"], rtol=1e-12, what=name + " grad")
    g, jac, hess = impl.eval_g(x), impl.eval_jac_g(x), impl.eval_h(x, sigma, lam)
    if exact:
        parity.assert_parity(g, G["refex_g"], rtol=1e-12, min_bit_equal=0.99, what=name + " g")
        parity.assert_parity(jac, G["refex_jac"], rtol=1e-12, min_bit_equal=0.99, what=name + " jac")
        assert scale_rel(hess, G["refex_hess"]) <= 1e-9
    else:
        assert scale_rel(g, G["refex_g"]) <= 1e-12
        assert scale_rel(jac, G["refex_jac"]) <= 1e-9
        assert scale_rel(hess, G["refex_hess"]) <= 1e-3


@pytest.mark.parametrize("name", REFEX_CASES)
def test_functor_headers_match_reference_examples_on_cpu(name):
    check_against_refex(name, Oracle(cases.build(name)))


@pytest.mark.gpu
@pytest.mark.parametrize("name", REFEX_CASES)
def test_cuda_matches_reference_examples(nlp_mod, name):
    g = nlp_mod.TranscribedNLP(cases.build(name))
    check_against_refex(name, golden_lib.CudaAdapter(g))
    g.close()


@pytest.mark.skipif(not os.path.exists(REF_LIB_PATH), reason="oracle/_ref not built (needs /root/reference)")
def test_launch_constants_equal_the_reference_globals():
    """CONSTANTS / scales of Launch.cpp:11-74 (static initialisers, evaluated by the reference's own object file) and
    the six fields main() assigns (:115-127,:148-153) against examples.launch(): the device functor's Consts."""
    from lpopc_b200 import examples
    out = np.zeros(15)
    assert ref_lib().lpo_ref_launch_constants(out.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(15)) == 15
    mine = np.asarray(examples.launch().consts, dtype=np.float64)
    assert mine.size == 14
    assert np.max(np.abs(out[:14] - mine) / np.abs(out[:14])) <= 4e-16
