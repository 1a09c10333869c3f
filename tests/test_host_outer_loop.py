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
import scipy.sparse as sp
from scipy.optimize import minimize

import golden_lib
from lpopc_b200 import examples
from oracle_lib import Oracle

REF_LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "liblpopc_ref.so")
CASES = [("bryson_denham", dict(intervals=4, nodes=4), 1e-10), ("brachistochrone", dict(intervals=4, nodes=6), 1e-12),
         ("hypersensitive", dict(intervals=8, nodes=6), 1e-12)]
OTOL = 1e-8


def host_solve(cb, x0, ftol):
    """SLSQP on the TNLP-shaped callbacks `cb` (nlp_info, bounds, eval_f, eval_grad_f, eval_g, jac triplets)."""
    n, m = cb.nlp_info()[:2]
    jI, jJ = cb.jac_structure()
    xl, xu, gl, gu = cb.bounds()
    eq = gl == gu
    lo_f, hi_f = (~eq) & (gl > -1e19), (~eq) & (gu < 1e19)

    def jac(x):  # duplicate triplets sum, as IPOPT does
        return sp.coo_matrix((cb.eval_jac_g(x), (jI, jJ)), shape=(m, n)).toarray()

    cons = [dict(type="eq", fun=lambda x: cb.eval_g(x)[eq] - gl[eq], jac=lambda x: jac(x)[eq])]
    if lo_f.any():
        cons.append(dict(type="ineq", fun=lambda x: cb.eval_g(x)[lo_f] - gl[lo_f], jac=lambda x: jac(x)[lo_f]))
    if hi_f.any():
        cons.append(dict(type="ineq", fun=lambda x: gu[hi_f] - cb.eval_g(x)[hi_f], jac=lambda x: -jac(x)[hi_f]))
    bnds = [(None if lo < -1e19 else lo, None if hi > 1e19 else hi) for lo, hi in zip(xl, xu)]
    r = minimize(cb.eval_f, x0, jac=cb.eval_grad_f, method="SLSQP", constraints=cons, bounds=bnds, options=dict(ftol=ftol, maxiter=600))
    g = cb.eval_g(r.x)
    viol = max(float(np.max(np.maximum(gl - g, 0))), float(np.max(np.maximum(g - gu, 0))))
    return float(r.fun), viol, int(r.nit)


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
