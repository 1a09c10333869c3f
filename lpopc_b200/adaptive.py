"""Adaptive mesh loop: solve -> mesh-error estimate -> ph / hp-Liu refinement -> re-transcribe -> warm start.

Mirrors the mesh loop of the reference's LpopcAlgorithm::SolveOptimalControlProblem
(Lpopc/src/Core/LpLpopcAlgorithm.cpp:17-46: `SolveNlp; Nlp2OpControl; while(!RefineMesh()){ UpdateGrid;
GetSizes; GetBounds; GetGuess; SolveNlp; Nlp2OpControl; }`).  Every new grid means new n, m, nnz and
new index maps: `lpb_set_mesh` + `lpb_refresh` rebuild them on the GPU (north_star item 3); the error
estimate between two solves runs on the GPU as well (`lpb_mesh_error`).  The outer
NLP solver is lpopc_b200.solver.BatchedIPM (batch of one, GPU-resident) or a host solver on the TNLP callbacks
(`slsqp_host_solver`), either standing in for IPOPT.

Guess transfer between grids (N3, host side like the reference): the previous solution becomes the
guess (Nlp2OPConverter.cpp:160-193) and is interpolated onto the new LGR nodes with a natural cubic
spline (LpGuessChecker.cpp:130-190, 208-294).
"""
import numpy as np


def natural_spline(xdata, ydata, x):
    """Natural cubic spline through (xdata, ydata) evaluated at x, with the reference's formulas
    (LpGuessChecker::spline_second_derivative / spline_interpolation, LpGuessChecker.cpp:208-262): forward sweep for
    the second derivatives, bisection for the bracketing knots, y = A yl + B yr + C d2l + D d2r.  The reference redoes
    the sweep for every abscissa; here it runs once."""
    xd, yd, x = np.asarray(xdata, dtype=np.float64), np.asarray(ydata, dtype=np.float64), np.asarray(x, dtype=np.float64)
    n = xd.size
    mu, z, c = np.zeros(n), np.zeros(n), np.zeros(n)
    for i in range(1, n - 1):
        him1, hi = xd[i] - xd[i - 1], xd[i + 1] - xd[i]
        alphai = 3.0 / hi * (yd[i + 1] - yd[i]) - 3.0 / him1 * (yd[i] - yd[i - 1])
        li = 2 * (xd[i + 1] - xd[i - 1]) - him1 * mu[i - 1]
        mu[i] = hi / li
        z[i] = (alphai - him1 * z[i - 1]) / li
    for j in range(n - 2, -1, -1):
        c[j] = z[j] - mu[j] * c[j + 1]
    c[1:n - 1] *= 2
    # bisection of the reference (1-based kleft / kright): the last knot <= x on the left, except at the right end
    kr = np.clip(np.searchsorted(xd, x, side="right"), 1, n - 1)
    kl = kr - 1
    h = xd[kr] - xd[kl]
    if np.any(h == 0.0):
        raise ValueError("Bad xdata input to routine spline_interpolation()")
    A, B = (xd[kr] - x) / h, (x - xd[kl]) / h
    C_, D_ = (A ** 3 - A) * (h * h) / 6.0, (B ** 3 - B) * (h * h) / 6.0
    return A * yd[kl] + B * yd[kr] + C_ * c[kl] + D_ * c[kr]


def transfer_guess(op, x, old_points, new_points, control_end=None):
    """Previous solution x on the old mesh -> starting point on the new mesh, as the reference does it: the solution
    becomes the user guess -- time grid, states and controls on the N + 1 points [LGR nodes, +1], the controls' last
    row being the spline end row of Nlp2OpControl (Nlp2OPConverter.cpp:160-193) -- and GetGuess interpolates states onto
    [new nodes, +1] and controls onto the new nodes with natural cubic splines in tau (LpGuessChecker.cpp:130-190).
    control_end[ip][j] = that last control row (lpb_nlp2op's control matrix, row N); None = the last node's value."""
    out, off = [], 0
    for ip, p in enumerate(op.phases):
        ns, nc = len(p.statemin), len(p.controlmin)
        N = len(old_points[ip])
        c0 = off + ns * (N + 1)
        t0, tf = x[c0 + nc * N], x[c0 + nc * N + 1]
        tau_all = np.concatenate([old_points[ip], [1.0]])
        tsol = (tf - t0) * (tau_all + 1) / 2 + t0                 # the solution's time grid (Nlp2OPConverter.cpp:56-58)
        tau_o = 2 * (tsol - tsol[0]) / (tsol[-1] - tsol[0]) - 1    # and back to tau (LpGuessChecker.cpp:137-139)
        tau_n = np.concatenate([new_points[ip], [1.0]])
        for j in range(ns):
            out.append(natural_spline(tau_o, x[off + j * (N + 1): off + (j + 1) * (N + 1)], tau_n))
        for j in range(nc):
            u = x[c0 + j * N: c0 + (j + 1) * N]
            u_end = u[-1] if control_end is None else control_end[ip][j]
            out.append(natural_spline(tau_o, np.concatenate([u, [u_end]]), tau_n[:-1]))
        out.append(np.array([tsol[0], tsol[-1]]))
        off = c0 + nc * N + 2
    return np.concatenate(out)


def slsqp_host_solver(ftol=1e-10, maxiter=600):
    """A HOST outer NLP solver on the TNLP surface (bounds, f, grad f, g, Jacobian triplets): SciPy's SLSQP standing in
    for IPOPT, which is not in this image (north_star keeps IPOPT and its linear solver on the host).  Dense
    quasi-Newton SQP: meant for single problems of a few hundred variables (the reference's shipped examples).
    Returns solve(nlp, x0) -> (x, objective, status, iterations) for solve_adaptive(host_solver=...)."""
    import scipy.sparse as sp
    from scipy.optimize import minimize

    def solve(nlp, x0):
        n, m = nlp.get_nlp_info()[:2]
        jI, jJ = nlp.eval_jac_g(values=False)
        xl, xu, gl, gu = nlp.get_bounds_info()
        eq = gl == gu
        lo_f, hi_f = (~eq) & (gl > -1e19), (~eq) & (gu < 1e19)
        jac = lambda x: sp.coo_matrix((nlp.eval_jac_g(x), (jI, jJ)), shape=(m, n)).toarray()  # duplicates sum, as in IPOPT
        cons = [dict(type="eq", fun=lambda x: nlp.eval_g(x)[eq] - gl[eq], jac=lambda x: jac(x)[eq])]
        if lo_f.any():
            cons.append(dict(type="ineq", fun=lambda x: nlp.eval_g(x)[lo_f] - gl[lo_f], jac=lambda x: jac(x)[lo_f]))
        if hi_f.any():
            cons.append(dict(type="ineq", fun=lambda x: gu[hi_f] - nlp.eval_g(x)[hi_f], jac=lambda x: -jac(x)[hi_f]))
        bnds = [(None if lo < -1e19 else lo, None if hi > 1e19 else hi) for lo, hi in zip(xl, xu)]
        r = minimize(nlp.eval_f, x0, jac=nlp.eval_grad_f, method="SLSQP", constraints=cons, bounds=bnds, options=dict(ftol=ftol, maxiter=maxiter))
        g = nlp.eval_g(r.x)
        viol = max(float(np.max(np.maximum(gl - g, 0))), float(np.max(np.maximum(g - gu, 0))))
        return r.x, float(r.fun), 0 if viol <= 1e-6 else 1, int(r.nit)

    return solve


def solve_adaptive(op, make_nlp, make_evaluator, solver_cls, mesh_tol=1e-6, nmax=16, nmin=4, max_grids=10, ipm_tol=1e-6,
                   max_iter=200, verbose=False, host_solver=None, method="ph", ratio_r=1.2, max_nodes=None):
    """Runs the mesh loop on `op` (its phases' meshes are updated in place).

    make_nlp(op) -> object with lgr_points(), initial_guess(), set_mesh(), refresh(), probe_dependencies(),
    refine_mesh_ph() (lpopc_b200.nlp.TranscribedNLP on the GPU); make_evaluator(nlp) -> evaluator for
    solver_cls (lpopc_b200.solver.CudaEvaluator / BatchedIPM).  host_solver(nlp, x) -> (x, obj, status, iters), if
    given, replaces the GPU-resident solver with a host outer loop on the TNLP callbacks (the reference's own
    arrangement: IPOPT on the host, `slsqp_host_solver()` here).  method = "ph" (PhMeshRefineAlg) or "hp-Liu"
    (LiuHpMeshRefineAlg; option "mesh-refine-methods", LpMeshRefiner.h:44-61).  max_nodes: stop (record "node_cap") instead
    of moving to a mesh with more LGR nodes in total than this -- a tolerance below the finite-difference noise floor
    otherwise doubles the mesh until memory runs out.  Returns (x, history)."""
    nlp = make_nlp(op)
    if method == "hp-Liu":
        nlp.refine_reset()
    x = nlp.initial_guess()
    nlp.probe_dependencies(x)  # once per problem, like LpopcAlgorithm::GetDependecies
    history = []
    for grid in range(1, max_grids + 1):
        if host_solver is not None:
            x, obj, status, iters = host_solver(nlp, x)
            n_, m_, nnzj_, nnzh_ = nlp.get_nlp_info()
        else:
            ev = make_evaluator(nlp)
            res = solver_cls(ev, tol=ipm_tol, max_iter=max_iter).solve(x[None, :])
            x = res["x"][0].cpu().numpy()
            obj, status, iters = float(res["obj"][0]), int(res["status"][0]), int(res["iters"][0])
            n_, m_, nnzj_, nnzh_ = ev.n, ev.m, ev.nnz_jac, ev.nnz_h
        if method == "hp-Liu":
            done, meshes = nlp.refine_mesh_hp_liu(x, tol=mesh_tol, nmax=nmax, ratio_r=ratio_r)
        else:
            done, meshes = nlp.refine_mesh_ph(x, tol=mesh_tol, nmax=nmax, nmin=nmin)
        _, imax = nlp.mesh_error(x)
        rec = {"grid": grid, "n": n_, "m": m_, "nnz_jac": nnzj_, "nnz_h": nnzh_,
               "nodes": [int(np.sum(p.nodesperinterval)) for p in op.phases], "intervals": [len(p.nodesperinterval) for p in op.phases],
               "objective": obj, "status": status, "iters": iters,
               "max_rel_error": float(max(v.max() for v in imax)), "mesh_satisfied": bool(done)}
        history.append(rec)
        if verbose:
            print(rec)
        if done or rec["status"] != 0:
            break
        if all(np.array_equal(mp, p.meshpoints) and np.array_equal(nd, p.nodesperinterval) for p, (mp, nd) in zip(op.phases, meshes)):
            # PhMeshRefineAlg::ModifySegment truncates log(e/tol)/log(N) towards zero (LpPhMeshRefineAlg.cpp:81), so an
            # interval whose error exceeds tol by less than a factor N is "refined" to itself; the reference would
            # repeat the identical solve until max-grid-num -- stop instead and say so
            rec["mesh_stalled"] = True
            break
        if max_nodes is not None and sum(int(np.sum(nd)) for _, nd in meshes) > max_nodes:
            rec["node_cap"] = True
            break
        old_pts = nlp.lgr_points()
        # last control row of the converted solution (spline end rows, k_nlp2op_ends); multipliers do not enter it
        ctrl_end = [q["control"][-1] if q["control"].size else np.zeros(0) for q in nlp.nlp2op(x, np.zeros(m_))[0]]
        for ip, (mp, nd) in enumerate(meshes):
            op.phases[ip].set_mesh(mp, nd)
            nlp.set_mesh(ip, mp, nd)
        nlp.refresh()  # new index maps / tables on the GPU
        x = transfer_guess(op, x, old_pts, nlp.lgr_points(), ctrl_end)
    return x, history
