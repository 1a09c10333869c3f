#!/usr/bin/env python
"""Per-configuration timing table (BASELINE.json configs 1-5): for every TNLP callback the device time of the
CUDA path with inputs resident in HBM, the latency of the host-pointer C-ABI call (what IPOPT would see), and the
reference's own CPU code (oracle/_ref, single thread -- the reference is single-threaded) on the same inputs.

    python scripts/config_table.py [--skip-ref] > profiles/r01_config_table.txt
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _hbm_peak():
    """GB/s: the driver-measured copy bandwidth (MEASURED_PEAKS.json), else the profiling recipe's fallback."""
    import json
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


HBM_PEAK = _hbm_peak()


def dev_ms(fn, reps):
    import torch
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def wall_ms(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return 1e3 * (time.perf_counter() - t0) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-ref", action="store_true")
    ap.add_argument("--only", default="", help="comma-separated config tags to run (C1..C5); default all")
    args = ap.parse_args()
    import torch
    import cases
    from lpopc_b200 import examples, nlp
    from oracle_lib import Oracle, RefOracle
    configs = [
        ("C1 hypersensitive 1x20", examples.hypersensitive(), 1),
        ("C2 orbit raising 200x10", examples.orbit_raising(intervals=200, nodes=10), 1),
        ("C3 launch 4 phases 25x10", examples.launch(intervals=25, nodes=10), 1),
        ("C4 quadrotor 8x8 x4096", examples.quadrotor(intervals=8, nodes=8), 4096),
        ("C5 synthetic20 10000x10", examples.synthetic20(intervals=10000, nodes=10), 1),
    ]
    print("# device = CUDA events, inputs resident; host = wall clock of the host-pointer C-ABI call (pageable numpy arrays);")
    print("# ref = the reference's own sources (oracle/_ref), one thread, same inputs.  Times in ms per call (C4: per 4096-instance batch).")
    print("# host+pin = the same call with option auto_pin (the caller's reused arrays are page-locked on their second use).")
    print("%-26s %9s %9s %10s | %-8s %10s %10s %10s %10s" % ("config", "n", "m", "nnz_jac", "callback", "device", "host", "host+pin", "ref cpu"))
    for name, op, nb in configs:
        if args.only and name.split()[0] not in args.only.split(","):
            continue
        g = nlp.TranscribedNLP(op)
        g.set_stream(torch.cuda.current_stream().cuda_stream)
        n, m, nnz, nnz_h = g.get_nlp_info()
        o = Oracle(op)
        _, x, sigma, lam = cases.inputs(op, o, 3)
        rng = np.random.Generator(np.random.PCG64(11))
        X = np.ascontiguousarray(x[None, :] + 1e-3 * rng.uniform(-1, 1, (nb, n)) * (np.abs(x) + 0.1))
        LAM = np.array(np.broadcast_to(lam, (nb, m)))
        SG = np.full(nb, sigma)
        dx, dl, ds = torch.from_numpy(X).cuda(), torch.from_numpy(LAM).cuda(), torch.from_numpy(SG).cuda()
        df = torch.empty(nb, dtype=torch.float64, device="cuda")
        dgr = torch.empty((nb, n), dtype=torch.float64, device="cuda")
        dg = torch.empty((nb, m), dtype=torch.float64, device="cuda")
        dv = torch.empty((nb, nnz), dtype=torch.float64, device="cuda")
        big_h = nb * nnz_h * 8 > 2e9 or nnz_h > 2e7  # config 5: 1193 dae() per node, minutes on the CPU arm
        dh = None if big_h else torch.empty((nb, nnz_h), dtype=torch.float64, device="cuda")
        reps = 20 if nb * nnz < 5e7 else 5
        # config 5: the reference's dense temporaries (repmat / join_horiz of N x (ns+np) per column) do not fit a
        # bounded-memory run at 100k nodes; the restatement (oracle/, same algorithm) stands in, marked with '*'
        port = name.startswith("C5")
        ref = None if args.skip_ref else (o if port else RefOracle(op))
        hg, hv = np.empty((nb, m)), np.empty((nb, nnz))  # the caller's arrays, reused across calls as IPOPT's are
        rows = [
            ("eval_f", lambda: g.eval_f_dev(nb, dx.data_ptr(), df.data_ptr()), lambda: g.eval_f_batch(X),
             (lambda: [ref.eval_f(X[i]) for i in range(min(nb, 8))]) if ref else None),
            ("grad_f", lambda: g.eval_grad_f_dev(nb, dx.data_ptr(), dgr.data_ptr()), lambda: g.eval_grad_f_batch(X),
             (lambda: [ref.eval_grad_f(X[i]) for i in range(min(nb, 8))]) if ref else None),
            ("g+jac", lambda: g.eval_g_jac_dev(nb, dx.data_ptr(), dg.data_ptr(), dv.data_ptr()), lambda: g.eval_g_jac_batch(X, hg, hv),
             (lambda: [(ref.eval_g(X[i]), ref.eval_jac_g(X[i])) for i in range(min(nb, 8))]) if ref else None),
        ]
        if dh is not None:
            rows.append(("eval_h", lambda: g.eval_h_dev(nb, dx.data_ptr(), ds.data_ptr(), dl.data_ptr(), dh.data_ptr()),
                         lambda: g.eval_h_batch(X, SG, LAM),
                         (lambda: [ref.eval_h(X[i], sigma, lam) for i in range(min(nb, 8))]) if ref else None))
        first = True
        for cb, fdev, fhost, fref in rows:
            d = dev_ms(fdev, reps)
            g.set_option("auto_pin", 0)
            h = wall_ms(fhost, max(2, reps // 4))
            g.set_option("auto_pin", 1)
            fhost(); fhost(); fhost()  # sighting, registration, sparse-return learning
            hp = wall_ms(fhost, max(2, reps // 4))
            g.set_option("auto_pin", 0)
            r = None
            if fref is not None and not (port and cb in ("eval_h", "grad_f")):
                r = wall_ms(fref, 1 if nnz > 1e6 else 3) * (nb / min(nb, 8))
            print("%-26s %9s %9s %10s | %-8s %10.4f %10.3f %10.3f %10s" % (name if first else "", n if first else "", m if first else "", nnz if first else "",
                                                                    cb, d, h, hp, ("%.2f%s" % (r, "*" if port else "")) if r is not None else "-"), flush=True)
            first = False
        if nb == 1 and n + m + nnz <= 262144:
            # what a host IPOPT iteration costs: eval_f, eval_grad_f, eval_g, eval_jac_g (and eval_h) of ONE new x
            # through the single-problem TNLP entry points -- with the fast path (one captured graph per new x, the other
            # callbacks served from the pinned stage) and without (every callback uploads x, launches, downloads, syncs)
            xs = [X[0] + 1e-6 * k for k in range(64)]
            def four(k):
                xx = xs[k % 64]
                g.eval_f(xx); g.eval_grad_f(xx); g.eval_g(xx); g.eval_jac_g(xx)
            def five(k):
                four(k); g.eval_h(xs[k % 64], sigma, lam)
            res = {}
            for mode in (1, 0):
                g.set_option("fast_path", mode)
                for fn, key in ((four, "4"), (five, "5")):
                    for k in range(4):
                        fn(k)
                    t0 = time.perf_counter()
                    for k in range(4, 44):
                        fn(k)
                    res[(mode, key)] = 1e3 * (time.perf_counter() - t0) / 40
            g.set_option("fast_path", -1)
            rr = ""
            if ref is not None:
                t0 = time.perf_counter()
                for k in range(3):
                    ref.eval_f(xs[k]); ref.eval_grad_f(xs[k]); ref.eval_g(xs[k]); ref.eval_jac_g(xs[k])
                rr = ", ref cpu %.3f ms" % (1e3 * (time.perf_counter() - t0) / 3)
            print("%-26s %9s %9s %10s | one new x, f + grad_f + g + jac_g through the TNLP calls: fast path %.3f ms, without %.3f ms%s; "
                  "with eval_h: %.3f ms vs %.3f ms" % ("", "", "", "", res[(1, "4")], res[(0, "4")], rr, res[(1, "5")], res[(0, "5")]), flush=True)
        if nb == 1:
            # per-mesh and per-solve steps around the callbacks (host wall clock, one problem): index maps + tables
            # rebuilt on the GPU, NLP -> optimal-control conversion, mesh-error estimate
            ms_refresh = wall_ms(g.refresh, 3)
            st_ms = g.stat("structure_ns") / 1e6            # k_jac_structure + k_hess_structure, CUDA events inside lpb_refresh
            st_bytes = 8.0 * (nnz + g.get_nlp_info()[3])        # two 32-bit indices per triplet of both index maps
            ms_n2o = wall_ms(lambda: g.nlp2op(X[0], lam), 3)
            ms_err = wall_ms(lambda: g.mesh_error(X[0]), 3)
            print("%-26s %9s %9s %10s | refresh (tables + index maps on the GPU) %.3f ms wall, of which the two index-map kernels %.3f ms = %.0f GB/s of 8 (nnz_jac + nnz_h) bytes (%.2f of the HBM peak); nlp2op %.3f ms, mesh_error %.3f ms (both with pageable host copies of x / lambda / results inside)" % ("", "", "", "", ms_refresh, st_ms, st_bytes / (st_ms * 1e-3) / 1e9 if st_ms > 0 else 0.0, st_bytes / (st_ms * 1e-3) / 1e9 / HBM_PEAK if st_ms > 0 else 0.0, ms_n2o, ms_err), flush=True)
        del g


if __name__ == "__main__":
    main()
