"""The drop-in boundary under a HOST outer loop (north_star: "IPOPT plus its linear solver stay on the host as the
outer loop"; correctness bar: converged objectives within 1e-8 relative).

IPOPT is not in this image, so SciPy's SLSQP stands in for it as the host NLP solver: it sees exactly what IPOPT's
TNLP would -- bounds, eval_f, eval_grad_f, eval_g, the sparse Jacobian triplets -- and drives the reference's own
single-phase examples (Bryson-Denham: free final time, events; hypersensitive; plus the brachistochrone functor) to
convergence.  The same solver is run on three sets of callbacks:
  * the reference's own code (oracle/_ref, where it is built),
  * the CPU restatement (oracle/),
  * the CUDA path through the C ABI (-m gpu),
and the converged objectives must agree within 1e-8 relative.  Bryson-Denham with l = 1/9 has the literature
optimum J* = 4 / (9 l) = 4 (a sanity check external to the reference, not a parity pin).
"""
import os

import numpy as np
import pytest

import golden_lib
from lpopc_b200 import examples
from oracle_lib import Oracle

REF_LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "liblpopc_ref.so")
CASES = [("bryson_denham", dict(intervals=4, nodes=4), 1e-10), ("brachistochrone", dict(intervals=4, nodes=6), 1e-12),
         ("hypersensitive", dict(intervals=8, nodes=6), 1e-12)]
OTOL = 1e-8


class _AsNLP:
    """Oracle-style callbacks (tests/oracle_lib.Oracle, golden_lib.CudaAdapter) behind the TranscribedNLP method names
    that lpopc_b200.adaptive.slsqp_host_solver drives."""

    def __init__(self, cb):
        self.cb = cb

    def get_nlp_info(self): return self.cb.nlp_info()
    def get_bounds_info(self): return self.cb.bounds()
    def eval_f(self, x): return self.cb.eval_f(x)
    def eval_grad_f(self, x): return self.cb.eval_grad_f(x)
    def eval_g(self, x): return self.cb.eval_g(x)
    def eval_jac_g(self, x=None, values=True): return self.cb.eval_jac_g(x) if values else self.cb.jac_structure()


def host_solve(cb, x0, ftol):
    """The product's host outer solver (SciPy SLSQP on the TNLP surface, lpopc_b200/adaptive.py) on callbacks `cb`:
    (objective, max constraint violation, iterations)."""
    from lpopc_b200 import adaptive
    x, obj, _, nit = adaptive.slsqp_host_solver(ftol=ftol)(_AsNLP(cb), x0)
    xl, xu, gl, gu = cb.bounds()
    g = cb.eval_g(x)
    viol = max(float(np.max(np.maximum(gl - g, 0))), float(np.max(np.maximum(g - gu, 0))))
    return obj, viol, nit


def _start(op, o):
    return op.guess([o.tables(ip)["points"] for ip in range(len(op.phases))])


@pytest.mark.parametrize("name,kw,ftol", CASES)
def test_host_solver_converges_on_restatement_and_reference_callbacks(name, kw, ftol):
    op = getattr(examples, name)(**kw)
    o = Oracle(op)
    obj, viol, nit = host_solve(o, _start(op, o), ftol)
    assert viol <= 1e-6 and nit > 3
    if name == "bryson_denham":
        assert abs(obj - 4.0) <= 1e-7 * 4.0  # literature optimum 4 / (9 l), l = 1/9
    if os.path.exists(REF_LIB):
        from oracle_lib import RefOracle
        r = RefOracle(op)
        obj_ref, viol_ref, _ = host_solve(r, _start(op, r), ftol)
        assert viol_ref <= 1e-6
        assert abs(obj - obj_ref) <= OTOL * max(1.0, abs(obj_ref))


@pytest.mark.gpu
@pytest.mark.parametrize("name,kw,ftol", CASES)
def test_host_solver_on_cuda_callbacks_matches_cpu_objective(name, kw, ftol):
    from lpopc_b200 import nlp
    op = getattr(examples, name)(**kw)
    o = Oracle(op)
    x0 = _start(op, o)
    obj_cpu, _, _ = host_solve(o, x0, ftol)
    g = nlp.TranscribedNLP(op)
    l0 = g.kernel_launches
    obj_gpu, viol, nit = host_solve(golden_lib.CudaAdapter(g), x0, ftol)
    assert g.kernel_launches > l0 + 3 * nit  # every iteration went through the CUDA callbacks
    assert viol <= 1e-6
    assert abs(obj_gpu - obj_cpu) <= OTOL * max(1.0, abs(obj_cpu))
    if name == "bryson_denham":
        assert abs(obj_gpu - 4.0) <= 1e-7 * 4.0
