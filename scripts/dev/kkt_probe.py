"""lpb_kkt_factor against the library assembly + factorisation it replaces, on the quadrotor's shapes."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lpopc_b200 import nlp, solver
nlp.load_library()
B, K, nb, nbd, mr = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 8, 140, 22, 96
dev = torch.device("cuda")
gen = torch.Generator(device="cpu").manual_seed(1)
kkt = object.__new__(solver.BlockTridiagKKT)
kkt.K, kkt.nb, kkt.nbd, kkt.mr, kkt.gamma = K, nb, nbd, mr, 1e6
kkt.bnd = torch.arange(nbd, device=dev)
kkt.fused_factor, kkt.fused_solve, kkt.n_factor = False, None, 0
A = torch.randn(64, K, nb, nb, generator=gen, dtype=torch.float64)
D = (A @ A.transpose(2, 3) * 0.05 + 0.5 * torch.eye(nb, dtype=torch.float64)).to(dev).repeat(B // 64, 1, 1, 1)
E = (0.1 * torch.randn(B, K - 1, nbd, nb, generator=gen, dtype=torch.float64)).to(dev)
Jb = torch.randn(B, K, mr, nb + nbd, generator=gen, dtype=torch.float64).to(dev)
Jb[:, -1, :, nb:] = 0.0
kkt.Jb = Jb
sigma = torch.rand(B, K, nb, generator=gen, dtype=torch.float64).to(dev)
di = torch.arange(nb, device=dev)
base = D[:, :, di, di] + sigma
dw0 = torch.zeros(B, dtype=torch.float64, device=dev)
done = torch.zeros(B, dtype=torch.bool, device=dev)
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
t_lib = timed(lambda: kkt._assemble_and_factor(D.clone(), E.clone(), base, dw0.clone(), done))
t_fused = timed(lambda: kkt._kkt_factor_fused(D, E, sigma, dw0, done))
flops = B * K * (2 * mr * (nb + nbd) ** 2 / 2 + (nb + nbd) ** 3 / 3)
print("B=%d  library %.2f ms   lpb_kkt_factor %.2f ms  (%.1fx)  %.2f TFLOP/s fp64" % (B, t_lib, t_fused, t_lib / t_fused, flops / (t_fused * 1e-3) / 1e12))
