"""Device time of lpb_blocktri_factor / lpb_blocktri_solve against the library recursion (torch.linalg) on random
SPD block-tridiagonal systems of the solver's shapes."""
import ctypes as C
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lpopc_b200 import nlp
lib = nlp.load_library()
dev = torch.device("cuda")


def ms(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for B, K, nb, nbd in ((4096, 8, 140, 22), (1024, 8, 140, 22), (4096, 8, 44, 7), (64, 8, 140, 22)):
    g = torch.Generator(device="cpu").manual_seed(1)
    bnd = torch.randperm(nb, generator=g)[:nbd].sort().values.to(dev)
    A = torch.randn(B, K, nb, nb, dtype=torch.float64, device=dev)
    D = A @ A.transpose(2, 3) + nb * torch.eye(nb, dtype=torch.float64, device=dev)
    E = 0.3 * torch.randn(B, K - 1, nbd, nb, dtype=torch.float64, device=dev)
    rhs = torch.randn(B, K, nb, dtype=torch.float64, device=dev)
    bnd32 = bnd.to(torch.int32)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def lib_factor():
        Ls, Cs, prev = [], [], None
        for i in range(K):
            Ai = D[:, i]
            if i > 0:
                Cm = torch.linalg.solve_triangular(prev, E[:, i - 1].transpose(1, 2), upper=False).transpose(1, 2).contiguous()
                Cs.append(Cm)
                Ai = Ai.clone()
                Ai[:, bnd.unsqueeze(1), bnd.unsqueeze(0)] -= Cm @ Cm.transpose(1, 2)
            prev, _ = torch.linalg.cholesky_ex(Ai)
            Ls.append(prev)
        return Ls, Cs

    Lall, Call = torch.empty_like(D), torch.empty_like(E)
    info = torch.empty(B, dtype=torch.int32, device=dev)

    def fused_factor():
        rc = lib.lpb_blocktri_factor(B, K, nb, nbd, C.c_void_p(D.data_ptr()), C.c_void_p(E.data_ptr()), C.c_void_p(bnd32.data_ptr()),
                                     C.c_void_p(Lall.data_ptr()), C.c_void_p(Call.data_ptr()), C.c_void_p(info.data_ptr()), st)
        assert rc == 0

    out = torch.empty_like(rhs)
    Lp = (C.c_void_p * K)(*[Lall[:, i].data_ptr() for i in range(K)])
    Cp = (C.c_void_p * (K - 1))(*[Call[:, i].data_ptr() for i in range(K - 1)])

    def fused_solve():
        rc = lib.lpb_blocktri_solve(B, K, nb, nbd, Lp, Cp, K * nb * nb, (K - 1) * nbd * nb, C.c_void_p(bnd32.data_ptr()), C.c_void_p(rhs.data_ptr()),
                                    C.c_void_p(out.data_ptr()), st)
        assert rc == 0

    print("B=%d K=%d nb=%d nbd=%d: factor library %.3f ms, fused %.3f ms; solve fused %.3f ms" % (B, K, nb, nbd, ms(lib_factor), ms(fused_factor), ms(fused_solve)), flush=True)
