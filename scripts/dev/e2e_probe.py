"""A/B of the host-pointer batch call (bench.py's e2e leg): dense vs sparse return, host thread counts."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from lpopc_b200 import nlp

op = bench.quadrotor_problem()
g = nlp.TranscribedNLP(op)
n, m, nnz, _ = g.get_nlp_info()
nb = bench.INSTANCES_PER_GPU
X = bench.make_inputs(op, g.lgr_points(), 0, nb)
hx = [torch.from_numpy(X + 1e-3 * k).pin_memory() for k in range(2)]
hg = torch.empty((nb, m), dtype=torch.float64).pin_memory()
hv = torch.empty((nb, nnz), dtype=torch.float64).pin_memory()


def run(label, steps=10):
    for k in range(2):
        g.eval_g_jac_batch_ptr(nb, hx[k % 2].data_ptr(), hg.data_ptr(), hv.data_ptr())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(steps):
        g.eval_g_jac_batch_ptr(nb, hx[k % 2].data_ptr(), hg.data_ptr(), hv.data_ptr())
    dt = (time.perf_counter() - t0) / steps
    print("%-40s %.3f ms/call  %.3e nnz/s  sent/inst=%d of %d fixups=%d" % (
        label, 1e3 * dt, nnz * nb / dt, g.stat("sparse_on_doubles"), g.stat("head_doubles"), g.stat("sparse_fixups")), flush=True)


g.set_option("host_threads", 16)
g.set_option("sparse_return", 1)
run("sparse (learn + steady)")
for skip, what in ((1, "no host fill"), (2, "no head return (host fill + g only)"), (3, "kernels + g only")):
    g.set_option("debug_skip", skip)
    for sp in (1, 0):
        g.set_option("sparse_return", sp)
        run("%s, %s" % ("sparse" if sp else "dense", what))
g.set_option("debug_skip", 0)
for thr in (4, 8, 12, 16):
    g.set_option("host_threads", thr)
    g.set_option("sparse_return", 0)
    run("dense return, %d host threads" % thr)
    g.set_option("sparse_return", 1)
    run("sparse return, %d host threads" % thr)
