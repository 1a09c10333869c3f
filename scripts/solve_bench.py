#!/usr/bin/env python
"""Batched OCP solves/s: B quadrotor (or cart-pole) MPC instances with varied initial states solved
to KKT tolerance by lpopc_b200.solver.BatchedIPM on one GPU, every NLP callback a device-resident
call into the transcription kernels.  Prints one JSON line.

  python scripts/solve_bench.py quadrotor 1024 --chunk 1024
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("problem")
    ap.add_argument("nbatch", type=int)
    ap.add_argument("--chunk", type=int, default=1024)
    ap.add_argument("--tol", type=float, default=1e-6)
    ap.add_argument("--probe", type=int, default=1, help="run the NaN dependency probe first (sparser Hessian pattern), like the reference")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--fused", type=int, default=1, help="1: lpb_blocktri_solve (one launch per block-tridiagonal solve), 0: library triangular solves")
    ap.add_argument("--compact-at", type=float, default=0.5, help="compact the batch once fewer than this fraction of it is unconverged")
    ap.add_argument("--refine", type=int, default=3, help="iterative-refinement steps of the block-tridiagonal KKT step")
    ap.add_argument("--gamma", type=float, default=1e6, help="dual regularisation 1/delta of the block-tridiagonal KKT step")
    ap.add_argument("--dense", action="store_true", help="dense condensed KKT step instead of the block-tridiagonal one")
    args = ap.parse_args()
    import torch
    from lpopc_b200 import examples, nlp, solver
    from lpopc_b200 import batch
    op = getattr(examples, args.problem)(intervals=8, nodes=8)
    g = nlp.TranscribedNLP(op)
    pts = g.lgr_points()
    ph = op.phases[0]
    nominal = np.array([ph.stateguess[j][0] for j in range(len(ph.statemin))])
    rng = np.random.Generator(np.random.PCG64(5))
    x0s = nominal + 0.2 * rng.uniform(-1, 1, (args.nbatch, nominal.size))
    X0 = batch.mpc_starting_points(op, pts, x0s)
    if args.probe:
        g.probe_dependencies(X0[0])
    ev = solver.CudaEvaluator(g)
    xl, xu, _, _ = ev.bounds()
    XL, XU = batch.mpc_bounds(xl, xu, op, x0s)
    ipm = solver.BatchedIPM(ev, tol=args.tol, max_iter=150, verbose=args.verbose,
                            var_blocks=None if args.dense else solver.interval_blocks(op, ev.n),
                            kkt_gamma=args.gamma, kkt_refine=args.refine, kkt_fused=bool(args.fused), compact_at=args.compact_at)
    ipm.solve(X0[:8], XL[:8], XU[:8])  # warm-up (cuSOLVER handles, kernels)
    torch.cuda.synchronize()
    l0, t0 = g.kernel_launches, time.perf_counter()
    r = ipm.solve(X0, XL, XU, chunk=args.chunk)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ok = int((r["status"] == 0).sum().item())
    print(json.dumps({"metric": "batched OCP solves/s", "value": ok / dt, "unit": "solves/s", "problem": args.problem, "instances": args.nbatch,
                      "converged": ok, "seconds": dt, "iters_mean": float(r["iters"].double().mean().item()), "iters_max": int(r["iters"].max().item()),
                      "kkt_error_max": float(r["kkt_error"].max().item()), "objective_mean": float(r["obj"].mean().item()),
                      "n": ev.n, "m": ev.m, "nnz_jac": ev.nnz_jac, "nnz_h": ev.nnz_h, "transcription_kernel_launches": g.kernel_launches - l0,
                      "tol": args.tol, "chunk": args.chunk, "kkt": ipm.kkt_kind, "kkt_factorisations": getattr(ipm.kkt, "n_factor", None),
                      "kkt_solves": getattr(ipm.kkt, "n_solve", None), "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}))


if __name__ == "__main__":
    main()
