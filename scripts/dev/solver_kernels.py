"""GPU time per kernel over one batched solve (torch.profiler / CUPTI): where a BatchedIPM solve spends the device."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lpopc_b200 import batch, examples, nlp, solver
from torch.profiler import profile, ProfilerActivity
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
op = examples.quadrotor(intervals=8, nodes=8)
g = nlp.TranscribedNLP(op)
pts = g.lgr_points()
rng = np.random.Generator(np.random.PCG64(5))
x0s = 0.2 * rng.uniform(-1, 1, (B, 12))
X0 = batch.mpc_starting_points(op, pts, x0s)
g.probe_dependencies(X0[0])
ev = solver.CudaEvaluator(g)
xl, xu, _, _ = ev.bounds()
XL, XU = batch.mpc_bounds(xl, xu, op, x0s)
ipm = solver.BatchedIPM(ev, tol=1e-6, max_iter=100, var_blocks=solver.interval_blocks(op, ev.n))
ipm.solve(X0[:8], XL[:8], XU[:8])
torch.cuda.synchronize()
t0 = time.perf_counter()
r = ipm.solve(X0, XL, XU, chunk=B)
torch.cuda.synchronize()
print("unprofiled seconds", time.perf_counter() - t0, "iters mean", float(r["iters"].double().mean()), "max", int(r["iters"].max()))
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    r = ipm.solve(X0, XL, XU, chunk=B)
    torch.cuda.synchronize()
ka = [k for k in prof.key_averages() if k.device_time_total > 0 and k.device_type.name == "CUDA"]
tot = sum(k.device_time_total for k in ka)
print("device time total %.1f ms over %d kernel names" % (tot / 1e3, len(ka)))
for k in sorted(ka, key=lambda k: -k.device_time_total)[:32]:
    print("%8.2f ms %5.1f%% x%-5d %s" % (k.device_time_total / 1e3, 100 * k.device_time_total / tot, k.count, k.key[:110]))
