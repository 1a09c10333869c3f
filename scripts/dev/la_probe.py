import torch, time
B, nf, me = 256, 1036, 768
dev = "cuda"
J = torch.randn(B, me, nf, dtype=torch.float64, device=dev) * (torch.rand(B, me, nf, device=dev) < 0.03)
W = torch.randn(B, nf, nf, dtype=torch.float64, device=dev); W = W @ W.transpose(1, 2) / nf + torch.eye(nf, dtype=torch.float64, device=dev)
def t(name, f, n=3):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = f()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    print("%-28s %8.2f ms" % (name, dt * 1e3)); return r
JtJ = t("JtJ bmm", lambda: torch.bmm(J.transpose(1, 2), J))
H = W + 1e3 * JtJ
L = t("cholesky_ex H", lambda: torch.linalg.cholesky_ex(H)[0])
Y = t("cholesky_solve 769 rhs", lambda: torch.cholesky_solve(J.transpose(1, 2).contiguous(), L))
Y2 = t("solve_triangular x1", lambda: torch.linalg.solve_triangular(L, J.transpose(1, 2), upper=False))
S = t("S bmm", lambda: torch.bmm(J, Y))
S2 = t("S syrk via tri", lambda: torch.bmm(Y2.transpose(1, 2), Y2))
Ls = t("cholesky_ex S", lambda: torch.linalg.cholesky_ex(S + 1e-9 * torch.eye(me, dtype=torch.float64, device=dev))[0])
A = torch.randn(B, 1036, 1036, dtype=torch.float64, device=dev)
t("bmm 1036^3", lambda: torch.bmm(A, A))
print("bmm TFLOPS", 2 * 1036**3 * B / 1e12)
