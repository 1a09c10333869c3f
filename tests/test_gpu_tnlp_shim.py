"""The Ipopt::TNLP adapter of INTEGRATION.md, compiled: include/lpopc_b200_ipopt.hpp against the TNLP interface
(oracle/ref_shim/IpTNLP.hpp, IPOPT's published signatures) and driven through the base-class pointer by
tests/shim_harness.cpp in IPOPT's call order -- structure calls, value calls with new_x flags, finalize_solution ->
lpb_nlp2op.  Everything it returns must equal what the C ABI returns directly (bit for bit)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import cases
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libshim_harness.so")


def harness():
    srcs = [os.path.join(HERE, "shim_harness.cpp"), os.path.join(ROOT, "include", "lpopc_b200_ipopt.hpp"),
            os.path.join(ROOT, "include", "lpopc_b200.h"), os.path.join(ROOT, "oracle", "ref_shim", "IpTNLP.hpp")]
    if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        pkg = os.path.join(ROOT, "lpopc_b200")
        subprocess.run(["g++", "-O2", "-std=c++14", "-fPIC", "-shared", "-Wall", "-Werror=overloaded-virtual", "-I", os.path.join(ROOT, "oracle", "ref_shim"),
                        "-o", LIB, srcs[0], "-L", pkg, "-llpopc_b200", "-Wl,-rpath," + pkg], check=True)
    return C.CDLL(LIB)


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


@pytest.mark.parametrize("name", ["hypersensitive", "bryson_denham", "launch/u5x4", "orbit_raising/u200x10"])
def test_compiled_tnlp_adapter_matches_the_c_abi(nlp_mod, name):
    op = cases.build(name)
    o = Oracle(op)
    g = nlp_mod.TranscribedNLP(op)  # loads liblpopc_b200.so first: the harness binds to the same library instance
    L = harness()
    n, m, nnz, nnzh = g.get_nlp_info()
    guess, x, sigma, lam = cases.inputs(op, o, 9)
    nx = 3
    xs = np.ascontiguousarray(np.stack([guess, x, x + 1e-4 * np.sin(np.arange(n))]))
    info = np.zeros(5, dtype=np.int32)
    xl, xu, gl, gu, x0 = np.empty(n), np.empty(n), np.empty(m), np.empty(m), np.empty(n)
    jI, jJ = np.empty(nnz, dtype=np.int32), np.empty(nnz, dtype=np.int32)
    hI, hJ = np.empty(nnzh, dtype=np.int32), np.empty(nnzh, dtype=np.int32)
    f, grad, gg = np.empty(nx), np.empty((nx, n)), np.empty((nx, m))
    jac, hess = np.empty((nx, nnz)), np.empty((nx, nnzh))
    cap = 64 * (n + m) + 1024
    sol, cost, refused = np.empty(cap), C.c_double(), C.c_int()
    lamc = np.ascontiguousarray(lam)
    rc = L.shim_drive(g.h, _p(guess), nx, _p(xs), _p(lamc), C.c_double(sigma),
                      _p(info, C.c_int), _p(xl), _p(xu), _p(gl), _p(gu), _p(x0), _p(jI, C.c_int), _p(jJ, C.c_int), _p(hI, C.c_int),
                      _p(hJ, C.c_int), _p(f), _p(grad), _p(gg), _p(jac), _p(hess), _p(sol), C.c_longlong(cap), C.byref(cost), C.byref(refused))
    assert rc > 0, rc
    assert tuple(info[:4]) == (n, m, nnz, nnzh) and info[4] == 0  # C_STYLE
    assert refused.value == 1 and np.array_equal(x0, guess)
    for a, b in zip((xl, xu, gl, gu), g.get_bounds_info()):
        assert np.array_equal(a, b)
    for a, b in zip((jI, jJ), g.eval_jac_g(values=False)):
        assert np.array_equal(a, b)
    for a, b in zip((hI, hJ), g.eval_h(values=False)):
        assert np.array_equal(a, b)
    assert np.array_equal(jI, o.jac_structure()[0]) and np.array_equal(hJ, o.h_structure()[1])
    for k in range(nx):
        assert f[k] == g.eval_f(xs[k])
        assert np.array_equal(grad[k], g.eval_grad_f(xs[k])) and np.array_equal(gg[k], g.eval_g(xs[k]))
        assert np.array_equal(jac[k], g.eval_jac_g(xs[k])) and np.array_equal(hess[k], g.eval_h(xs[k], sigma, lam))
    res, tot = g.nlp2op(xs[nx - 1], lam)
    assert tot == cost.value
    shapes = [(int(sum(p.nodesperinterval)) + 1, len(p.statemin), len(p.controlmin), len(p.pathmin)) for p in op.phases]
    via_shim = nlp_mod.unpack_nlp2op(sol[:rc], shapes)
    for q, r in zip(via_shim, res):
        for key in ("time", "state", "control", "costate", "pathmult", "hamiltonian"):
            assert np.array_equal(q[key], r[key], equal_nan=True), key
        assert q["mayer"] == r["mayer"] and q["lagrange"] == r["lagrange"]
    g.close()
