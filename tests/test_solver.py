"""Batched interior-point solver (lpopc_b200/solver.py): convergence on the CPU
evaluator, and converged objectives of the CUDA path within 1e-8 relative of the same algorithm
run on the CPU restatement of the reference path (north_star tolerance)."""
import numpy as np
import pytest
import torch

from lpopc_b200 import examples, solver
from solver_cpu import OracleEvaluator, mpc_bounds, mpc_instances


def _instances(op, ev_cpu, nb, spread, seed):
    pts = [ev_cpu.o.tables(0)["points"]]
    ph = op.phases[0]
    nominal = np.array([ph.stateguess[j][0] for j in range(len(ph.statemin))])
    rng = np.random.Generator(np.random.PCG64(seed))
    x0s = nominal + spread * rng.uniform(-1, 1, (nb, nominal.size))
    return x0s, mpc_instances(op, pts, x0s)


def test_ipm_converges_on_cartpole_cpu():
    op = examples.cartpole(intervals=4, nodes=5)
    ev = OracleEvaluator(op)
    x0s, X0 = _instances(op, ev, 3, 0.1, 0)
    XL, XU = mpc_bounds(ev, op, x0s)
    r = solver.BatchedIPM(ev, tol=1e-6, max_iter=60).solve(X0, XL, XU)
    assert int(r["status"].abs().sum()) == 0 and float(r["kkt_error"].max()) <= 1e-6
    # feasibility of the returned points and fixed initial states
    g = ev.g(r["x"])
    _, _, gl, gu = ev.bounds()
    eq = gl == gu
    assert float((g[:, eq] - gl[eq]).abs().max()) <= 1e-6
    N = op.phases[0].GetTotalNodes()
    assert np.allclose(r["x"][:, torch.arange(4) * (N + 1)].numpy(), x0s, atol=0, rtol=0)
    # a perturbed start converges to the same objective
    r2 = solver.BatchedIPM(ev, tol=1e-7, max_iter=80).solve(X0 + 0.01, XL, XU)
    r1 = solver.BatchedIPM(ev, tol=1e-7, max_iter=80).solve(X0, XL, XU)
    assert np.allclose(r1["obj"].numpy(), r2["obj"].numpy(), rtol=1e-7)


def test_block_tridiagonal_kkt_matches_dense_cpu():
    op = examples.cartpole(intervals=4, nodes=5)
    ev = OracleEvaluator(op)
    x0s, X0 = _instances(op, ev, 3, 0.2, 1)
    XL, XU = mpc_bounds(ev, op, x0s)
    rd = solver.BatchedIPM(ev, tol=1e-7, max_iter=60).solve(X0, XL, XU)
    ipm = solver.BatchedIPM(ev, tol=1e-7, max_iter=60, var_blocks=solver.interval_blocks(op, ev.n))
    rb = ipm.solve(X0, XL, XU)
    assert ipm.kkt_kind.startswith("block-tridiagonal")
    assert int(rb["status"].abs().sum()) == 0 and np.allclose(rb["obj"].numpy(), rd["obj"].numpy(), rtol=1e-9)
    # a problem whose coupling does not fit (free final time couples every node) falls back to the dense step
    op2 = examples.bryson_denham(intervals=4, nodes=4)
    ev2 = OracleEvaluator(op2)
    ipm2 = solver.BatchedIPM(ev2, max_iter=1, var_blocks=solver.interval_blocks(op2, ev2.n))
    ipm2.solve(op2.guess([ev2.o.tables(0)["points"]])[None, :])
    assert ipm2.kkt_kind == "dense"


def test_ipm_inequality_rows_become_slacks():
    """Hypersensitive (reference example, Lpopc/example/hypersensitive): the duration row t_f - t_0 >= 0 is an
    inequality row; it is carried as an equality with a bounded slack and the solve converges to the
    KKT tolerance.  Free-final-time problems (Bryson-Denham, brachistochrone) are outside what this
    first solver version converges on reliably -- see DESIGN.md."""
    op = examples.hypersensitive(intervals=8, nodes=6)
    ev = OracleEvaluator(op)
    ipm = solver.BatchedIPM(ev, tol=1e-7, max_iter=60)
    assert isinstance(ipm.ev, solver.SlackEvaluator) and ipm.ev.ni == 1
    X0 = op.guess([ev.o.tables(0)["points"]])[None, :]
    r = ipm.solve(X0)
    assert int(r["status"][0]) == 0 and float(r["kkt_error"][0]) <= 1e-7
    assert r["x"].shape[1] == ev.n
    g = ev.g(r["x"])
    _, _, gl, gu = ev.bounds()
    assert bool(((g >= gl - 1e-6) & (g <= gu + 1e-6)).all())


@pytest.mark.gpu
@pytest.mark.parametrize("problem,kw,nb,spread", [("cartpole", dict(intervals=8, nodes=8), 64, 0.1), ("quadrotor", dict(intervals=8, nodes=8), 32, 0.2)])
def test_gpu_solves_match_cpu_reference_objectives(problem, kw, nb, spread):
    from lpopc_b200 import nlp
    op = getattr(examples, problem)(**kw)
    ev_cpu = OracleEvaluator(op, threads=8)
    x0s, X0 = _instances(op, ev_cpu, nb, spread, 5)
    g = nlp.TranscribedNLP(op)
    ev_gpu = solver.CudaEvaluator(g)
    XL, XU = mpc_bounds(ev_gpu, op, x0s)
    l0 = g.kernel_launches
    rg = solver.BatchedIPM(ev_gpu, tol=1e-7, max_iter=80).solve(X0, XL, XU)
    assert g.kernel_launches > l0
    assert int(rg["status"].abs().sum().item()) == 0
    # block-tridiagonal KKT step (mesh structure) against the dense condensed step
    ipm_b = solver.BatchedIPM(ev_gpu, tol=1e-7, max_iter=80, var_blocks=solver.interval_blocks(op, ev_gpu.n))
    rb = ipm_b.solve(X0, XL, XU)
    assert ipm_b.kkt_kind.startswith("block-tridiagonal") and int(rb["status"].abs().sum().item()) == 0
    assert float(((rb["obj"] - rg["obj"]).abs() / rg["obj"].abs().clamp(min=1.0)).max().item()) <= 1e-8
    sample = [0, 1, nb - 1]
    XLc, XUc = mpc_bounds(ev_cpu, op, x0s[sample])
    rc = solver.BatchedIPM(ev_cpu, tol=1e-7, max_iter=80).solve(X0[sample], XLc, XUc)
    assert int(rc["status"].abs().sum()) == 0
    og, oc = rg["obj"].cpu().numpy()[sample], rc["obj"].numpy()
    assert np.max(np.abs(og - oc) / np.maximum(1.0, np.abs(oc))) <= 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("B,K,nb,nbd", [(3, 1, 7, 1), (5, 4, 33, 5), (64, 8, 140, 22), (2, 3, 96, 96), (3, 2, 200, 9)])
def test_fused_block_tridiagonal_solve_matches_library(B, K, nb, nbd):
    """lpb_blocktri_solve (one launch per batched block-tridiagonal solve) against the same substitution through
    torch.linalg triangular solves, on random SPD block-tridiagonal systems; and against the assembled dense system."""
    import ctypes as C
    from lpopc_b200 import nlp
    lib = nlp.load_library()
    dev = torch.device("cuda")
    gen = torch.Generator(device="cpu").manual_seed(B * 1000 + nb)
    bnd = torch.randperm(nb, generator=gen)[:nbd].sort().values
    A = torch.randn(B, K, nb, nb, generator=gen, dtype=torch.float64)
    D = A @ A.transpose(2, 3) + nb * torch.eye(nb, dtype=torch.float64)
    E = 0.3 * torch.randn(B, max(K - 1, 1), nbd, nb, generator=gen, dtype=torch.float64)
    rhs = torch.randn(B, K, nb, generator=gen, dtype=torch.float64)
    # block Cholesky with boundary-row couplings (BlockTridiagKKT._factor)
    Ls, Cs, prev = [], [], None
    for i in range(K):
        Ai = D[:, i].clone()
        if i > 0:
            Cm = torch.linalg.solve_triangular(prev, E[:, i - 1].transpose(1, 2), upper=False).transpose(1, 2).contiguous()
            Cs.append(Cm)
            Ai[:, bnd.unsqueeze(1), bnd.unsqueeze(0)] -= Cm @ Cm.transpose(1, 2)
        prev = torch.linalg.cholesky(Ai)
        Ls.append(prev)
    # dense reference of the same system
    n = K * nb
    full = torch.zeros(B, n, n, dtype=torch.float64)
    for i in range(K):
        full[:, i * nb:(i + 1) * nb, i * nb:(i + 1) * nb] = D[:, i]
        if i > 0:
            rows = i * nb + bnd
            full[:, rows, (i - 1) * nb:i * nb] = E[:, i - 1]
            full[:, (i - 1) * nb:i * nb, rows] = E[:, i - 1].transpose(1, 2)
    xref = torch.linalg.solve(full, rhs.reshape(B, n, 1)).reshape(B, K, nb)
    Ld, Cd = [L.to(dev).contiguous() for L in Ls], [c.to(dev).contiguous() for c in Cs]
    bnd32 = bnd.to(torch.int32).to(dev)
    rd = rhs.to(dev).contiguous()
    out = torch.empty_like(rd)
    Lp = (C.c_void_p * K)(*[L.data_ptr() for L in Ld])
    Cp = (C.c_void_p * max(K - 1, 1))(*([c.data_ptr() for c in Cd] or [0]))
    rc = lib.lpb_blocktri_solve(B, K, nb, nbd, Lp, Cp, nb * nb, nbd * nb, C.c_void_p(bnd32.data_ptr()), C.c_void_p(rd.data_ptr()),
                                C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    err = float(((out.cpu() - xref).abs().max() / xref.abs().max()).item())
    assert err <= 1e-11, err
    # the fused factorisation: same factors as the library recursion, and factor + solve reproduce the dense solution
    Dd, Ed = D.to(dev).contiguous(), E.to(dev).contiguous()
    Lall, Call = torch.empty_like(Dd), torch.empty_like(Ed)
    info = torch.full((B,), -7, dtype=torch.int32, device=dev)
    rc = lib.lpb_blocktri_factor(B, K, nb, nbd, C.c_void_p(Dd.data_ptr()), C.c_void_p(Ed.data_ptr()), C.c_void_p(bnd32.data_ptr()),
                                 C.c_void_p(Lall.data_ptr()), C.c_void_p(Call.data_ptr()), C.c_void_p(info.data_ptr()),
                                 C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    assert int(info.abs().sum().item()) == 0
    for i in range(K):
        assert float((torch.tril(Lall[:, i]).cpu() - Ls[i]).abs().max().item()) <= 1e-10 * float(Ls[i].abs().max().item())
        if i + 1 < K:
            assert float((Call[:, i].cpu() - Cs[i]).abs().max().item()) <= 1e-10 * max(1.0, float(Cs[i].abs().max().item()))
    Lp2 = (C.c_void_p * K)(*[Lall[:, i].data_ptr() for i in range(K)])
    Cp2 = (C.c_void_p * max(K - 1, 1))(*([Call[:, i].data_ptr() for i in range(K - 1)] or [0]))
    out2 = torch.empty_like(rd)
    rc = lib.lpb_blocktri_solve(B, K, nb, nbd, Lp2, Cp2, K * nb * nb, max(K - 1, 1) * nbd * nb, C.c_void_p(bnd32.data_ptr()), C.c_void_p(rd.data_ptr()),
                                C.c_void_p(out2.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    assert float(((out2.cpu() - xref).abs().max() / xref.abs().max()).item()) <= 1e-11
    # lpb_blocktri_solve_masked: instances whose mask byte is 0 are skipped (zero solution), the others are bit-identical
    mask = torch.ones(B, dtype=torch.uint8, device=dev)
    mask[B // 2] = 0
    out3 = torch.full_like(rd, float("nan"))
    rc = lib.lpb_blocktri_solve_masked(B, K, nb, nbd, Lp2, Cp2, K * nb * nb, max(K - 1, 1) * nbd * nb, C.c_void_p(bnd32.data_ptr()),
                                       C.c_void_p(mask.data_ptr()), C.c_void_p(rd.data_ptr()), C.c_void_p(out3.data_ptr()),
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    keep = mask.bool()
    assert torch.equal(out3[keep], out2[keep]) and float(out3[~keep].abs().max()) == 0.0
    # an indefinite instance is reported (inertia test), the others are unaffected, everything stays finite
    if B >= 2:
        Dbad = Dd.clone()
        Dbad[1, K - 1] -= 3.0 * nb * torch.eye(nb, dtype=torch.float64, device=dev)
        rc = lib.lpb_blocktri_factor(B, K, nb, nbd, C.c_void_p(Dbad.data_ptr()), C.c_void_p(Ed.data_ptr()), C.c_void_p(bnd32.data_ptr()),
                                     C.c_void_p(Lall.data_ptr()), C.c_void_p(Call.data_ptr()), C.c_void_p(info.data_ptr()),
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        inf = info.cpu()
        assert rc == 0 and int(inf[1]) > (K - 1) * nb and int(inf[0]) == 0 and bool(torch.isfinite(torch.tril(Lall[1])).all())


@pytest.mark.parametrize("problem,optimum,rtol", [("bryson_denham", 4.0, 1e-6), ("brachistochrone", 0.824338669, 1e-8)])
def test_ipm_filter_line_search_solves_free_final_time_problems_cpu(problem, optimum, rtol):
    """The reference solves its free-final-time examples with IPOPT's filter line search (LpNLPSolver.cpp:27-33 leaves the
    default; Lpopc/example/bryson-denham/BrysonDenham.cpp:77-78).  With the l1 merit function the iteration stalls on
    them (steps of 2^-10); the filter line search converges -- to the literature optimum 4 / (9 l), l = 1/9, of
    Bryson-Denham and to the fine-mesh brachistochrone value -- alone and in a batch, from perturbed starts too."""
    op = getattr(examples, problem)(intervals=2, nodes=6)
    ev = OracleEvaluator(op)
    x0 = op.guess([ev.o.tables(0)["points"]])
    X0 = np.stack([x0, x0, x0 * 1.01])
    r = solver.BatchedIPM(ev, tol=1e-6, max_iter=150).solve(X0)
    assert r["status"].tolist() == [0, 0, 0] and float(r["kkt_error"].max()) <= 1e-6
    assert np.allclose(r["obj"].numpy(), optimum, rtol=rtol, atol=0)
    assert int(r["iters"].max()) <= 60
    stalled = solver.BatchedIPM(ev, tol=1e-6, max_iter=60, linesearch="merit").solve(x0[None, :])
    assert problem != "brachistochrone" or int(stalled["status"][0]) == 1  # what the filter is there for


@pytest.mark.parametrize("problem,kw", [("hypersensitive", dict(intervals=4, nodes=6)), ("brachistochrone", dict(intervals=2, nodes=6)),
                                        ("cartpole", dict(intervals=3, nodes=5))])
def test_ipm_limited_memory_hessian_mode_cpu(problem, kw):
    """hessian = "limited-memory" (the reference's default "hessian-approximation", LpNLPSolver.cpp:30-33): damped L-BFGS
    on the Lagrangian gradient, no eval_h call, same converged objective as the exact-Hessian mode."""
    op = getattr(examples, problem)(**kw)
    ev = OracleEvaluator(op)
    calls = {"h": 0}
    hess = ev.hess
    ev.hess = lambda *a: (calls.__setitem__("h", calls["h"] + 1), hess(*a))[1]
    x0 = op.guess([ev.o.tables(0)["points"]])
    X0 = np.stack([x0, x0 * 1.01])
    lm = solver.BatchedIPM(ev, tol=1e-6, max_iter=300, hessian="limited-memory").solve(X0)
    assert calls["h"] == 0 and lm["status"].tolist() == [0, 0]
    ex = solver.BatchedIPM(ev, tol=1e-6, max_iter=150).solve(X0)
    assert calls["h"] > 0 and ex["status"].tolist() == [0, 0]
    assert np.allclose(lm["obj"].numpy(), ex["obj"].numpy(), rtol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("B,K,nb,nbd,mr", [(5, 4, 44, 8, 30), (3, 8, 140, 22, 96), (4, 1, 17, 3, 9), (6, 3, 36, 6, 40), (3, 2, 170, 10, 40)])
def test_kkt_factor_kernel_matches_library_assembly(B, K, nb, nbd, mr):
    """lpb_kkt_factor (gamma J^T J + boundary coupling + block Cholesky + inertia-correction retry in one launch, one
    8 x 8 tile of the augmented matrix per thread; 6 x 6 for the last shape, whose 8 x 8 tiling needs more than 8 warps) against the library path it replaces (einsum, elementwise assembly,
    batched Cholesky, Python retry loop): same factors, same couplings, same regularisation per instance -- on systems
    that are positive definite as they come, on one that needs the retry, and on a converged (inactive) instance."""
    from lpopc_b200 import nlp
    nlp.load_library()
    dev = torch.device("cuda")
    gen = torch.Generator(device="cpu").manual_seed(B * 977 + nb)
    kkt = object.__new__(solver.BlockTridiagKKT)
    kkt.K, kkt.nb, kkt.nbd, kkt.mr, kkt.gamma = K, nb, nbd, mr, 1e3
    kkt.bnd = torch.randperm(nb, generator=gen)[:nbd].sort().values.to(dev)
    kkt.fused_factor, kkt.fused_solve, kkt.n_factor = False, None, 0
    A = torch.randn(B, K, nb, nb, generator=gen, dtype=torch.float64)
    D = (A @ A.transpose(2, 3) * 0.05 + 0.5 * torch.eye(nb, dtype=torch.float64)).to(dev)
    D[1] = D[1] - 50.0 * torch.eye(nb, dtype=torch.float64, device=dev)  # indefinite without regularisation: the retry must find dw
    E = (0.1 * torch.randn(B, max(K - 1, 1), nbd, nb, generator=gen, dtype=torch.float64)).to(dev)
    Jb = torch.randn(B, K, mr, nb + nbd, generator=gen, dtype=torch.float64)
    Jb[:, -1, :, nb:] = 0.0  # the last interval has no next boundary
    Jb = (Jb * (torch.rand(B, K, mr, nb + nbd, generator=gen) < 0.2)).to(dev)  # sparse like a collocation Jacobian
    kkt.Jb = Jb
    sigma = torch.rand(B, K, nb, generator=gen, dtype=torch.float64).to(dev)
    di = torch.arange(nb, device=dev)
    base = D[:, :, di, di] + sigma
    dw0 = torch.zeros(B, dtype=torch.float64, device=dev)
    dw0[2] = 3e-3
    done = torch.zeros(B, dtype=torch.bool, device=dev)
    done[0] = True
    Lr, Cr, info_r, dw_r = kkt._assemble_and_factor(D.clone(), E.clone(), base, dw0.clone(), done)
    fused = kkt._kkt_factor_fused(D.clone(), E.clone(), sigma, dw0.clone(), done)
    assert fused is not None
    Lf, Cf, info_f, dw_f = fused
    torch.cuda.synchronize()
    assert info_f.tolist() == [0] * B and info_r.tolist() == [0] * B
    assert torch.equal(dw_f, dw_r) and float(dw_f[1]) > 0 and float(dw_f[2]) == 3e-3
    live = ~done
    eye = torch.eye(nb, dtype=torch.float64, device=dev)
    for i in range(K):
        lr, lf = torch.tril(Lr[i]), torch.tril(Lf[i])
        assert float((lf[live] - lr[live]).abs().max() / lr[live].abs().max()) <= 1e-10, i
        assert torch.equal(lf[0], eye)  # the converged instance is not factorised: identity, its step is discarded
    for i in range(K - 1):
        assert float((Cf[i][live] - Cr[i][live]).abs().max() / Cr[i][live].abs().max().clamp(min=1e-300)) <= 1e-9, i
        assert float(Cf[i][0].abs().max()) == 0.0


@pytest.mark.gpu
def test_batched_spmv_matches_dense_products():
    """lpb_batched_spmv (shared CSR structure, values read in place from per-instance triplet arrays through perm, with
    duplicates and an optional diagonal term) against dense matrix-vector products assembled from the same triplets."""
    import ctypes as C
    from lpopc_b200 import nlp
    lib = nlp.load_library()
    dev = torch.device("cuda")
    gen = torch.Generator(device="cpu").manual_seed(5)
    B, nrows, ncols, nnz = 7, 53, 41, 400
    rows = torch.randint(0, nrows, (nnz,), generator=gen)
    cols = torch.randint(0, ncols, (nnz,), generator=gen)
    rows[:ncols] = torch.arange(ncols) % nrows   # some structure, duplicates included by chance
    vals = torch.randn(B, nnz + 9, generator=gen, dtype=torch.float64)  # stride > nnz: values sit inside a longer array
    x = torch.randn(B, ncols, generator=gen, dtype=torch.float64)
    dense = torch.zeros(B, nrows, ncols, dtype=torch.float64)
    dense.index_put_((torch.arange(B).unsqueeze(1).expand(B, nnz), rows.expand(B, nnz), cols.expand(B, nnz)), vals[:, :nnz], accumulate=True)
    rowptr, col, perm = solver.BlockTridiagKKT._csr(rows.to(dev), cols.to(dev), torch.arange(nnz, device=dev), nrows)
    vd, xd = vals.to(dev), x.to(dev)
    y = torch.empty(B, nrows, dtype=torch.float64, device=dev)
    rc = lib.lpb_batched_spmv(B, nrows, ncols, C.c_void_p(rowptr.data_ptr()), C.c_void_p(col.data_ptr()), C.c_void_p(perm.data_ptr()),
                              C.c_void_p(vd.data_ptr()), vd.stride(0), C.c_void_p(xd.data_ptr()), None, C.c_void_p(y.data_ptr()),
                              C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    ref = torch.einsum("bij,bj->bi", dense, x)
    assert float((y.cpu() - ref).abs().max()) <= 1e-12 * float(ref.abs().max())
    # square matrix with the diagonal term
    n = 41
    rows2 = torch.randint(0, n, (nnz,), generator=gen)
    dg = torch.randn(B, n, generator=gen, dtype=torch.float64)
    dense2 = torch.zeros(B, n, n, dtype=torch.float64)
    dense2.index_put_((torch.arange(B).unsqueeze(1).expand(B, nnz), rows2.expand(B, nnz), cols.expand(B, nnz)), vals[:, :nnz], accumulate=True)
    rowptr2, col2, perm2 = solver.BlockTridiagKKT._csr(rows2.to(dev), cols.to(dev), torch.arange(nnz, device=dev), n)
    dgd = dg.to(dev)
    y2 = torch.empty(B, n, dtype=torch.float64, device=dev)
    rc = lib.lpb_batched_spmv(B, n, n, C.c_void_p(rowptr2.data_ptr()), C.c_void_p(col2.data_ptr()), C.c_void_p(perm2.data_ptr()),
                              C.c_void_p(vd.data_ptr()), vd.stride(0), C.c_void_p(xd.data_ptr()), C.c_void_p(dgd.data_ptr()), C.c_void_p(y2.data_ptr()),
                              C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    ref2 = torch.einsum("bij,bj->bi", dense2, x) + dg * x
    assert float((y2.cpu() - ref2).abs().max()) <= 1e-12 * float(ref2.abs().max())
