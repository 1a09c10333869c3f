"""Config 5 (synthetic20, 100k nodes): device time of the Lagrangian Hessian values."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lpopc_b200 import examples, nlp
K, N = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (10000, 10)
op = examples.synthetic20(intervals=K, nodes=N)
g = nlp.TranscribedNLP(op)
n, m, nnzj, nnzh = g.get_nlp_info()
rng = np.random.Generator(np.random.PCG64(3))
x = torch.from_numpy(g.initial_guess() + 1e-3 * rng.standard_normal(n)).cuda()
lam = torch.from_numpy(rng.standard_normal(m)).cuda()
sg = torch.ones(1, dtype=torch.float64, device="cuda")
dh = torch.empty(nnzh, dtype=torch.float64, device="cuda")
g.eval_h_dev(1, x.data_ptr(), sg.data_ptr(), lam.data_ptr(), dh.data_ptr())
print("n=%d m=%d nnz_h=%d" % (n, m, nnzh))
import time
torch.cuda.synchronize(); t0 = time.perf_counter()
g.eval_h_dev(1, x.data_ptr(), sg.data_ptr(), lam.data_ptr(), dh.data_ptr())
torch.cuda.synchronize(); print("wall one call %.3f ms" % (1e3 * (time.perf_counter() - t0)), "nonzero values", int((dh != 0).sum()), "finite", bool(torch.isfinite(dh).all()), "launches", g.kernel_launches)
for sp in (1, 2, 4, 8):
    g.set_option("pair_split", sp)
    g.eval_h_dev(1, x.data_ptr(), sg.data_ptr(), lam.data_ptr(), dh.data_ptr())
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        g.eval_h_dev(1, x.data_ptr(), sg.data_ptr(), lam.data_ptr(), dh.data_ptr())
    torch.cuda.synchronize(); print("pair_split %d: %.3f ms" % (sp, 1e3 * (time.perf_counter() - t0) / 3))
g.set_option("pair_split", 0)
hv = g.eval_h(x.cpu().numpy(), 1.0, lam.cpu().numpy())
print("host call equal:", bool(np.array_equal(hv, dh.cpu().numpy())))
