"""Per-mesh steps at config 5 (synthetic20, 100k nodes): wall time and the device activities inside each call."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lpopc_b200 import examples, nlp
from torch.profiler import profile, ProfilerActivity
K = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
op = examples.synthetic20(intervals=K, nodes=10)
g = nlp.TranscribedNLP(op)
n, m, nnzj, nnzh = g.get_nlp_info()
rng = np.random.Generator(np.random.PCG64(3))
x = g.initial_guess() + 1e-3 * rng.standard_normal(n)
lam = rng.standard_normal(m)
torch.zeros(1, device="cuda")
def wall(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return 1e3 * (time.perf_counter() - t0) / reps
for name, fn in (("refresh", g.refresh), ("nlp2op", lambda: g.nlp2op(x, lam)), ("mesh_error", lambda: g.mesh_error(x))):
    w = wall(fn)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    ka = [k for k in prof.key_averages() if k.device_time_total > 0]
    tot = sum(k.device_time_total for k in ka)
    print("%-10s wall %.2f ms, device activities %.2f ms  (n=%d m=%d nnz_jac=%d)" % (name, w, tot / 1e3, n, m, nnzj))
    for k in sorted(ka, key=lambda k: -k.device_time_total)[:8]:
        print("     %8.3f ms x%-3d %s" % (k.device_time_total / 1e3, k.count, k.key[:100]))
