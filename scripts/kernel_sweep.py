#!/usr/bin/env python
"""Device-time sweep of the fused eval_g + eval_jac_g pass (and optionally eval_h) over kernel
options, inputs resident in HBM.  Development tool: prints one line per configuration with
the step time, the k_cons_jac time (CUDA events inside the library) and the HBM fractions.

  python scripts/kernel_sweep.py quadrotor 4096 --unroll 0 1 --split 0 1 2
  python scripts/kernel_sweep.py synthetic20 1 --intervals 10000 --nodes 10
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("problem")
    ap.add_argument("nbatch", type=int)
    ap.add_argument("--intervals", type=int, default=None)
    ap.add_argument("--nodes", type=int, default=None)
    ap.add_argument("--unroll", type=int, nargs="*", default=[-1])
    ap.add_argument("--split", type=int, nargs="*", default=[0])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--stage", type=int, nargs="*", default=[0], help="shared-memory staged write-out of the Jacobian values (where it applies)")
    ap.add_argument("--rotate", type=int, nargs="*", default=[1], help="thread -> node rotation of the Jacobian kernel (aligned stores)")
    ap.add_argument("--sweep-mode", type=int, nargs="*", default=[0], help="functor sets with sweep hooks: 0 row-parallel (default row warps), 1 per-thread, 2 / 3 other row-warp counts")
    ap.add_argument("--hessian", action="store_true")
    ap.add_argument("--fg", action="store_true", help="also time eval_f and eval_grad_f")
    ap.add_argument("--pair-split", type=int, nargs="*", default=[0])
    args = ap.parse_args()
    import torch
    from lpopc_b200 import examples, nlp
    kw = {}
    if args.intervals:
        kw["intervals"] = args.intervals
    if args.nodes:
        kw["nodes"] = args.nodes
    op = getattr(examples, args.problem)(**kw)
    g = nlp.TranscribedNLP(op)
    g.set_stream(torch.cuda.current_stream().cuda_stream)
    n, m, nnz, nnz_h = g.get_nlp_info()
    nb = args.nbatch
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    rng = np.random.Generator(np.random.PCG64(7))
    xl, xu, _, _ = g.get_bounds_info()
    lo, hi = np.maximum(xl, -1.0), np.minimum(xu, 1.0)
    hi = np.where(hi > lo, hi, lo + 1.0)
    x0 = rng.uniform(lo, hi, (nb, n))
    x0[:, -2], x0[:, -1] = 0.0, 2.0
    xs = [torch.from_numpy(x0 + 1e-3 * k).cuda() for k in range(4)]
    d_g = torch.empty((nb, m), dtype=torch.float64, device="cuda")
    d_v = torch.empty((nb, nnz), dtype=torch.float64, device="cuda")
    step_bytes = 8 * nb * (n + m + nnz)
    print("problem %s nb=%d n=%d m=%d nnz_jac=%d nnz_h=%d step_bytes=%.1f MB" % (args.problem, nb, n, m, nnz, nnz_h, step_bytes / 1e6))
    for un, rot, stg, swm in [(u, r, s_, w) for u in args.unroll for r in args.rotate for s_ in args.stage for w in args.sweep_mode]:
        for sp in args.split:
            g.set_option("sweep_mode", swm)
            g.set_option("unroll_colours", un)
            g.set_option("rotate_nodes", rot)
            g.set_option("stage_values", stg)
            g.set_option("colour_split", sp)
            for k in range(3):
                g.eval_g_jac_dev(nb, xs[k % 4].data_ptr(), d_g.data_ptr(), d_v.data_ptr())
            torch.cuda.synchronize()
            g.set_option("time_kernels", 1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(args.steps):
                g.eval_g_jac_dev(nb, xs[k % 4].data_ptr(), d_g.data_ptr(), d_v.data_ptr())
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            kms, kc = g.kernel_time("cons_jac")
            g.set_option("time_kernels", 0)
            print("unroll=%2d rotate=%d stage=%d sweep_mode=%d split=%2d  step %.4f ms (%.1f%% hbm, %.3e nnz/s)  k_cons_jac %.4f ms" %
                  (un, rot, stg, swm, sp, ms, 100 * step_bytes / (ms * 1e-3) / 1e9 / peak, nnz * nb / (ms * 1e-3), kms / max(kc, 1)))
    if args.fg:
        d_f = torch.empty(nb, dtype=torch.float64, device="cuda")
        d_gr = torch.empty((nb, n), dtype=torch.float64, device="cuda")
        for name, fn, nbytes in (("eval_f", lambda k: g.eval_f_dev(nb, xs[k % 4].data_ptr(), d_f.data_ptr()), 8 * nb * (n + 1)),
                                 ("eval_grad_f", lambda k: g.eval_grad_f_dev(nb, xs[k % 4].data_ptr(), d_gr.data_ptr()), 16 * nb * n)):
            for k in range(3):
                fn(k)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(args.steps):
                fn(k)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            print("%-12s %.4f ms (%.1f%% hbm on %d MB)" % (name, ms, 100 * nbytes / (ms * 1e-3) / 1e9 / peak, nbytes // 1000000))
    if args.hessian:
        lam = torch.from_numpy(rng.uniform(-1, 1, (nb, m))).cuda()
        sg = torch.ones(nb, dtype=torch.float64, device="cuda")
        d_h = torch.empty((nb, nnz_h), dtype=torch.float64, device="cuda")
        hb = 8 * nb * (n + m + nnz_h)
        for ps in args.pair_split:
            g.set_option("pair_split", ps)
            for k in range(2):
                g.eval_h_dev(nb, xs[0].data_ptr(), sg.data_ptr(), lam.data_ptr(), d_h.data_ptr())
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            hs = max(2, args.steps // 4)
            for k in range(hs):
                g.eval_h_dev(nb, xs[k % 4].data_ptr(), sg.data_ptr(), lam.data_ptr(), d_h.data_ptr())
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / hs
            print("hessian pair_split=%d  %.4f ms (%.1f%% hbm, %.3e nnz_h/s)" % (ps, ms, 100 * hb / (ms * 1e-3) / 1e9 / peak, nnz_h * nb / (ms * 1e-3)))


if __name__ == "__main__":
    main()
