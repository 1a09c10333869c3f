import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lpopc_b200 import batch, examples, nlp, solver
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
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
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    r = ipm.solve(X0, XL, XU, chunk=B)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
