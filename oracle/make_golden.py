#!/usr/bin/env python
"""Generates tests/golden/*.npz from the REFERENCE'S OWN code (oracle/_ref/liblpopc_ref.so =
/root/reference sources compiled against oracle/ref_shim/, see oracle/ref_build.mk).

Runs only where /root/reference is mounted (this container); the fixtures travel with the repo so
that the GPU box can check the CUDA path against reference outputs without the reference.
One file per case of tests/cases.py: sizes, integer triplets, bounds, LGR tables, seeded inputs
(tests/cases.py::inputs, seed 7) and the reference's f / grad / g / Jacobian / Hessian values there.

    python oracle/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from oracle_lib import RefOracle  # noqa: E402

GOLDEN_CASES = cases.CASES + ["launch/u5x4", "hypersensitive/u40x3", "bryson_denham/u7x6"]
SEED = 7
DEP_CASES = ("launch", "quadrotor", "bryson_denham", "orbit_raising", "cartpole")  # NaN dependency probe fixtures


def main():
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    for name in GOLDEN_CASES:
        op = cases.build(name)
        r = RefOracle(op)
        guess, x, sigma, lam = cases.inputs(op, r, SEED)
        jI, jJ = r.jac_structure()
        hI, hJ = r.h_structure()
        xl, xu, gl, gu = r.bounds()
        d = dict(info=np.array(r.nlp_info()), jI=jI, jJ=jJ, hI=hI, hJ=hJ, xl=xl, xu=xu, gl=gl, gu=gu,
                 ref_guess=r.guess(), guess=guess, x=x, sigma=np.array(sigma), lam=lam,
                 f_guess=np.array(r.eval_f(guess)), f=np.array(r.eval_f(x)),
                 grad_guess=r.eval_grad_f(guess), grad=r.eval_grad_f(x),
                 g_guess=r.eval_g(guess), g=r.eval_g(x),
                 jac_guess=r.eval_jac_g(guess), jac=r.eval_jac_g(x),
                 hess=r.eval_h(x, sigma, lam))
        for ip in range(len(op.phases)):
            t = r.tables(ip)
            d["points%d" % ip], d["weights%d" % ip] = t["points"], t["weights"]
            d["doff_vals%d" % ip] = t["Doffdiag"][2]
        # mesh-error estimate and ph refinement decision at x (SolutionErrorChecker / PhMeshRefineAlg)
        for ip, e in enumerate(r.mesh_error(x)):
            d["mesh_rel%d" % ip] = e
        for tol, tag in ((1e-2, "a"), (1e-6, "b")):
            done, meshes = r.refine_ph(x, tol=tol)
            d["refine_%s_done" % tag] = np.array(int(done))
            for ip, (mp, nd) in enumerate(meshes):
                d["refine_%s_mesh%d" % (tag, ip)], d["refine_%s_nodes%d" % (tag, ip)] = mp, nd
        # NLP solution -> optimal-control solution at (x, lam) (Nlp2OpConverter::Nlp2OpControl)
        res, tot = r.nlp2op(x, lam)
        d["n2o_cost"] = np.array(tot)
        for ip, q in enumerate(res):
            for key in ("time", "state", "control", "costate", "pathmult", "hamiltonian"):
                d["n2o_%s%d" % (key, ip)] = q[key]
            d["n2o_costs%d" % ip] = np.array([q["mayer"], q["lagrange"]])
        base = name.split("/")[0]
        if base in ("hypersensitive", "bryson_denham", "launch"):
            # second opinion on include/problems/<base>.h: the same transcription with the reference's OWN example
            # user functions (Armadillo expressions + libm; oracle/ref_examples.cpp)
            op2 = cases.build(name)
            op2.functor = "ref:" + base
            r2 = RefOracle(op2)
            assert r2.nlp_info() == r.nlp_info()
            d["refex_f"], d["refex_grad"] = np.array(r2.eval_f(x)), r2.eval_grad_f(x)
            d["refex_g"], d["refex_jac"] = r2.eval_g(x), r2.eval_jac_g(x)
            d["refex_hess"] = r2.eval_h(x, sigma, lam)
        if name in DEP_CASES:  # dependency probe + the sparse Hessian pattern it implies
            d["dep"] = r.probe_dependencies(guess)
            d["dep_info"] = np.array(r.nlp_info())
            d["dep_hI"], d["dep_hJ"] = r.h_structure()
            d["dep_hess"] = r.eval_h(x, sigma, lam)
        path = os.path.join(out, name.replace("/", "__") + ".npz")
        np.savez_compressed(path, **d)
        print("%-28s n=%-6d m=%-6d nnz_jac=%-7d nnz_h=%-7d %6.1f KB" % ((name,) + r.nlp_info() + (os.path.getsize(path) / 1024,)))


def liu_sequences():
    """hp-Liu refinement sequences of the reference (LiuHpMeshRefineAlg, stateful): LIU_STEPS consecutive decisions on a
    manufactured solution re-sampled on every new mesh -> tests/golden/liu__<case>.npz."""
    out = os.path.join(ROOT, "tests", "golden")
    for name in cases.LIU_CASES:
        op = cases.build(name)
        r = RefOracle(op)
        d = {}
        for step in range(cases.LIU_STEPS):
            pts = [r.tables(ip)["points"] for ip in range(len(op.phases))]
            x = cases.manufactured_x(name, op, pts)
            d["s%d_in_mesh" % step] = np.asarray(op.phases[0].meshpoints, dtype=np.float64)
            d["s%d_in_nodes" % step] = np.asarray(op.phases[0].nodesperinterval, dtype=np.int32)
            d["s%d_rel" % step] = r.mesh_error(x)[0]
            done, meshes = r.refine_hp_liu(x, **cases.LIU_OPTIONS)
            mp, nd = meshes[0]
            d["s%d_done" % step] = np.array(int(done))
            d["s%d_mesh" % step], d["s%d_nodes" % step] = mp, nd
            print("%-24s step %d: %3d intervals, %4d nodes -> %3d intervals, %4d nodes%s" %
                  (name, step, len(op.phases[0].nodesperinterval), int(np.sum(op.phases[0].nodesperinterval)), nd.size, int(nd.sum()), "  (done)" if done else ""))
            op.phases[0].set_mesh(mp, nd)
            r.set_mesh(0, mp, nd)
            r.refresh()
        np.savez_compressed(os.path.join(out, "liu__" + name.replace("/", "__") + ".npz"), **d)


if __name__ == "__main__":
    main()
    liu_sequences()
