#!/usr/bin/env python
"""Device time of eval_h per tiled-kernel variant (tuning build: make -C lpopc_b200/csrc HESS_VARIANTS=1) against the
generic pair-loop kernel, dense and probed pattern, with a bit-for-bit comparison of the values.

  python scripts/dev/hess_sweep.py quadrotor 4096 --variants 0 1 2 3 4 5 6 7
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("problem")
    ap.add_argument("nbatch", type=int)
    ap.add_argument("--intervals", type=int, default=8)
    ap.add_argument("--nodes", type=int, default=8)
    ap.add_argument("--variants", type=int, nargs="*", default=[0])
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    import torch
    from lpopc_b200 import examples, nlp
    op = getattr(examples, args.problem)(intervals=args.intervals, nodes=args.nodes)
    nb = args.nbatch
    rng = np.random.Generator(np.random.PCG64(7))
    for probe in (False, True):
        g = nlp.TranscribedNLP(op)
        g.set_stream(torch.cuda.current_stream().cuda_stream)
        if probe:
            g.probe_dependencies(g.initial_guess())
        n, m, nnz, nnz_h = g.get_nlp_info()
        xl, xu, _, _ = g.get_bounds_info()
        lo, hi = np.maximum(xl, -1.0), np.minimum(xu, 1.0)
        hi = np.where(hi > lo, hi, lo + 1.0)
        x0 = rng.uniform(lo, hi, (nb, n))
        x0[:, -2], x0[:, -1] = 0.0, 2.0
        xs = [torch.from_numpy(x0 + 1e-3 * k).cuda() for k in range(4)]
        lam = torch.from_numpy(rng.uniform(-1, 1, (nb, m))).cuda()
        sg = torch.ones(nb, dtype=torch.float64, device="cuda")
        d_h = torch.empty((nb, nnz_h), dtype=torch.float64, device="cuda")
        hb = 8 * nb * (n + m + nnz_h)
        ref = None
        for v in [-1] + list(args.variants):
            g.set_option("unroll_colours", 0 if v < 0 else -1)
            g.set_option("hess_variant", max(v, 0))
            d_h.zero_()
            for k in range(2):
                g.eval_h_dev(nb, xs[0].data_ptr(), sg.data_ptr(), lam.data_ptr(), d_h.data_ptr())
            torch.cuda.synchronize()
            bits = d_h.view(torch.int64).clone()
            if ref is None:
                ref = bits
            same = bool(torch.equal(bits, ref))
            g.set_option("time_kernels", 1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(args.steps):
                g.eval_h_dev(nb, xs[k % 4].data_ptr(), sg.data_ptr(), lam.data_ptr(), d_h.data_ptr())
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            kms, kc = g.kernel_time("hess_nodes")
            g.set_option("time_kernels", 0)
            print("%s nb=%d probe=%d nnz_h=%d variant=%2d  eval_h %.4f ms  node kernel %.4f ms  %.1f GB/s algorithmic  bit-identical=%s" %
                  (args.problem, nb, int(probe), nnz_h, v, ms, kms / max(kc, 1), hb / (ms * 1e-3) / 1e9, same), flush=True)
        g.close()


if __name__ == "__main__":
    main()
