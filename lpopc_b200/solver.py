"""Batched outer NLP solver for independent MPC-style OCP instances.

The reference hands one NLP at a time to IPOPT on the host (Lpopc/src/Core/LpNLPSolver.cpp:13-54:
tol = "Ipopt-tol" 1e-6, adaptive barrier).  For a batch of instances that share problem, mesh and
sparsity pattern this module runs ONE primal-dual interior-point iteration for all instances in
lockstep on the GPU: every callback an interior-point method needs (objective, gradient,
constraints, Jacobian values, Lagrangian-Hessian values) is a device-resident batched call into the
transcription kernels (`lpb_*_dev`), so x, lambda and the triplet values never leave HBM.

Scope of this first version (stated, not hidden):
  * linear algebra is library code -- dense batched Cholesky / triangular solves through
    torch.linalg (cuSOLVER / cuBLAS), on the condensed system  S = J H^-1 J^T  with
    H = W + Sigma + delta_w I; the transcription kernels are the product, this solver is the
    first consumer of their device API;
  * equality constraints and variable bounds (log barrier, primal-dual); variables with
    x_l == x_u are eliminated; inequality rows (duration t_f - t_0 >= 0, path constraints) become
    equalities with bounded slack variables (SlackEvaluator);
  * l1 merit function with backtracking line search, monotone barrier reduction (Fiacco-McCormick
    as in IPOPT's default), per-instance Hessian regularisation delta_w driven by the Cholesky
    status of each instance.

The evaluator abstraction lets the same algorithm run against the CPU oracle in tests/ (parity of
converged objectives); the product evaluator below is CUDA-only.
"""
import numpy as np
import torch


class CudaEvaluator:
    """Device-resident batched callbacks of one TranscribedNLP (all tensors float64 on cuda)."""

    def __init__(self, nlp):
        self.nlp = nlp
        self.device = torch.device("cuda", torch.cuda.current_device())
        nlp.set_stream(torch.cuda.current_stream().cuda_stream)
        self.n, self.m, self.nnz_jac, self.nnz_h = nlp.get_nlp_info()
        jI, jJ = nlp.eval_jac_g(values=False)
        hI, hJ = nlp.eval_h(values=False)
        self.jI, self.jJ = torch.from_numpy(jI.astype(np.int64)).to(self.device), torch.from_numpy(jJ.astype(np.int64)).to(self.device)
        self.hI, self.hJ = torch.from_numpy(hI.astype(np.int64)).to(self.device), torch.from_numpy(hJ.astype(np.int64)).to(self.device)
        self.evals = {"f": 0, "grad": 0, "g_jac": 0, "g": 0, "hess": 0}

    def bounds(self):
        return [torch.from_numpy(a).to(self.device) for a in self.nlp.get_bounds_info()]

    def tensor(self, a):
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=torch.float64)
        return torch.as_tensor(np.asarray(a, dtype=np.float64)).to(self.device)

    def f(self, X):
        out = torch.empty(X.shape[0], dtype=torch.float64, device=self.device)
        self.nlp.eval_f_dev(X.shape[0], X.data_ptr(), out.data_ptr())
        self.evals["f"] += 1
        return out

    def grad(self, X):
        out = torch.empty_like(X)
        self.nlp.eval_grad_f_dev(X.shape[0], X.data_ptr(), out.data_ptr())
        self.evals["grad"] += 1
        return out

    def g(self, X):
        out = torch.empty((X.shape[0], self.m), dtype=torch.float64, device=self.device)
        self.nlp.eval_g_jac_dev(X.shape[0], X.data_ptr(), out.data_ptr(), None)
        self.evals["g"] += 1
        return out

    def g_jac(self, X):
        g = torch.empty((X.shape[0], self.m), dtype=torch.float64, device=self.device)
        v = torch.empty((X.shape[0], self.nnz_jac), dtype=torch.float64, device=self.device)
        self.nlp.eval_g_jac_dev(X.shape[0], X.data_ptr(), g.data_ptr(), v.data_ptr())
        self.evals["g_jac"] += 1
        return g, v

    def hess(self, X, sigma, lam):
        v = torch.empty((X.shape[0], self.nnz_h), dtype=torch.float64, device=self.device)
        self.nlp.eval_h_dev(X.shape[0], X.data_ptr(), sigma.data_ptr(), lam.data_ptr(), v.data_ptr())
        self.evals["hess"] += 1
        return v


class SlackEvaluator:
    """Presents a problem with inequality rows g_l <= g_i(x) <= g_u as an equality-only problem over
    z = [x; s]:  g_i(x) - s_i = 0,  g_l <= s_i <= g_u  (the slack formulation interior-point methods
    use; IPOPT does the same internally).  Triplet structures are extended by one -1 entry per slack."""

    def __init__(self, ev):
        self.ev, self.device = ev, ev.device
        _, _, gl, gu = ev.bounds()
        self.rows = torch.nonzero(gl != gu).squeeze(1)
        self.ni = int(self.rows.numel())
        self.nx = ev.n
        self.n, self.m = ev.n + self.ni, ev.m
        self.nnz_jac, self.nnz_h = ev.nnz_jac + self.ni, ev.nnz_h
        self.jI = torch.cat([ev.jI, self.rows])
        self.jJ = torch.cat([ev.jJ, ev.n + torch.arange(self.ni, device=self.device)])
        self.hI, self.hJ = ev.hI, ev.hJ

    def tensor(self, a):
        return self.ev.tensor(a)

    def bounds(self):
        xl, xu, gl, gu = self.ev.bounds()
        gl2, gu2 = gl.clone(), gu.clone()
        gl2[self.rows] = 0.0
        gu2[self.rows] = 0.0
        return torch.cat([xl, gl[self.rows]]), torch.cat([xu, gu[self.rows]]), gl2, gu2

    def extend(self, x, xl, xu):
        """[B, nx] points and bounds -> [B, nx + ni] with slacks started at g_i(x)."""
        _, _, gl, gu = self.ev.bounds()
        B = x.shape[0]
        s0 = self.ev.g(x.contiguous())[:, self.rows]
        return (torch.cat([x, s0], 1), torch.cat([xl, gl[self.rows].expand(B, -1)], 1), torch.cat([xu, gu[self.rows].expand(B, -1)], 1))

    def _x(self, Z):
        return Z[:, :self.nx].contiguous()

    def f(self, Z):
        return self.ev.f(self._x(Z))

    def grad(self, Z):
        return torch.cat([self.ev.grad(self._x(Z)), torch.zeros((Z.shape[0], self.ni), dtype=torch.float64, device=Z.device)], 1)

    def g(self, Z):
        g = self.ev.g(self._x(Z))
        g[:, self.rows] -= Z[:, self.nx:]
        return g

    def g_jac(self, Z):
        g, v = self.ev.g_jac(self._x(Z))
        g[:, self.rows] -= Z[:, self.nx:]
        return g, torch.cat([v, torch.full((Z.shape[0], self.ni), -1.0, dtype=torch.float64, device=Z.device)], 1)

    def hess(self, Z, sigma, lam):
        return self.ev.hess(self._x(Z), sigma, lam)


def blocked_cholesky_ex(A, split=512):
    """Batched lower Cholesky factor with a 2 x 2 block recursion above `split` rows: the batched
    library factorisation is several times slower per flop beyond ~1000 rows than below, while the
    off-diagonal work (triangular solve + symmetric update) runs at GEMM speed.  Returns (L, info)
    like torch.linalg.cholesky_ex (info != 0: not positive definite)."""
    n = A.shape[-1]
    if n <= split:
        return torch.linalg.cholesky_ex(A)
    h = n // 2
    L11, i1 = blocked_cholesky_ex(A[:, :h, :h], split)
    ok = (i1 == 0).view(-1, 1, 1)
    L11s = torch.where(ok, L11, torch.eye(h, dtype=A.dtype, device=A.device).expand_as(L11))  # keep failed instances finite
    L21 = torch.linalg.solve_triangular(L11s, A[:, h:, :h].transpose(1, 2), upper=False).transpose(1, 2)
    L22, i2 = blocked_cholesky_ex(A[:, h:, h:] - torch.bmm(L21, L21.transpose(1, 2)), split)
    L = torch.zeros_like(A)
    L[:, :h, :h], L[:, h:, :h], L[:, h:, h:] = L11s, L21, L22
    return L, torch.where(i1 != 0, i1, i2)



class DenseKKT:
    """Condensed dense KKT step: S = J H^-1 J^T with H = W + Sigma + rho J^T J + dw I (library Cholesky)."""

    def __init__(self, ipm):
        self.ipm = ipm

    def set_jac(self, jv):
        self.J = self.ipm._dense_jac(jv)

    def Jt(self, v):
        return torch.bmm(self.J.transpose(1, 2), v.unsqueeze(2)).squeeze(2)

    def step(self, hv, Sigma, rhs1, c, dw_last, done):
        ipm, J = self.ipm, self.J
        nf, me, rho = ipm.nf, ipm.me, ipm.rho
        dev = J.device
        W = hv if hv.dim() == 3 else ipm._dense_hess(hv)  # [B, nf, nf]: a quasi-Newton matrix handed over as it is
        # factor H_rho = W + Sigma + rho J^T J + dw I.  Adding rho J^T (J dx + c) = 0 to the first
        # block row leaves (dx, dlam) unchanged, and by Debreu's lemma H_rho is positive definite for
        # large rho exactly when the Hessian is positive definite on the null space of J -- the
        # inertia condition an interior-point step needs -- so the Cholesky status of each instance
        # drives its own regularisation dw (raised only for genuine negative curvature).
        JtJ = torch.bmm(J.transpose(1, 2), J)
        rhs1 = rhs1 + rho * self.Jt(c)
        dw = torch.where(dw_last > 0, dw_last / 3.0, 0.0)
        dw = torch.where(dw < 1e-9, 0.0, dw)
        Hd = W + rho * JtJ
        diag = torch.arange(nf, device=dev)
        base_diag = Hd[:, diag, diag] + Sigma
        for _try in range(40):
            Hd[:, diag, diag] = base_diag + dw.unsqueeze(1)
            Lh, info = blocked_cholesky_ex(Hd)
            bad = (info != 0) & (~done)
            if not bool(bad.any()):
                break
            dw = torch.where(bad, torch.where(dw == 0, 1e-4, dw * (100.0 if _try < 2 else 8.0)), dw)
        not_pd = (info != 0) & (~done)  # still indefinite after the last try: the step of such an instance is not trusted (status 2)
        # condensed system of [H J^T; J 0][dx; dlam] = -[rhs1; c] with H = L L^T:
        #   Y = L^-1 J^T, y = L^-1 rhs1, S = Y^T Y, S dlam = c - Y^T y, dx = -L^-T (y + Y dlam)
        Yr = torch.linalg.solve_triangular(Lh, torch.cat([J.transpose(1, 2), rhs1.unsqueeze(2)], dim=2), upper=False)
        Y, y = Yr[:, :, :me], Yr[:, :, me:]
        S = torch.bmm(Y.transpose(1, 2), Y)
        dS = torch.arange(me, device=dev)
        S[:, dS, dS] += 1e-12
        Ls, info_s = blocked_cholesky_ex(S)
        bad_s = (info_s != 0) & (~done)
        failed = torch.zeros_like(done)
        if bool(bad_s.any()):  # rank-deficient Jacobian: regularise the (2,2) block harder
            S[:, dS, dS] += (bad_s.to(torch.float64) * 1e-8).unsqueeze(1)
            Ls, info_s = blocked_cholesky_ex(S)
            failed = (info_s != 0) & (~done)
        failed = failed | not_pd
        rhs2 = c.unsqueeze(2) - torch.bmm(Y.transpose(1, 2), y)
        dlam = torch.cholesky_solve(rhs2, Ls)
        dx = -torch.linalg.solve_triangular(Lh.transpose(1, 2), y + torch.bmm(Y, dlam), upper=True).squeeze(2)
        return dx, dlam.squeeze(2), dw, failed


class BlockTridiagKKT:
    """KKT step that exploits the mesh structure of the transcription.

    With the free variables grouped by mesh interval (`var_blocks`), the Lagrangian Hessian couples only
    variables of the same node and every constraint row touches at most two neighbouring intervals (the
    defect rows of interval i read the first node of interval i+1), so H and J^T J are block tridiagonal
    with blocks of about N_k (ns + nc) rows.  The step solves the dual-regularised system
        [H  J^T; J  -delta I] [dx; dlam] = rhs,   delta = 1/gamma,
    by eliminating dlam = gamma (J dx - rhs2):  (H + gamma J^T J) dx = rhs1 + gamma J^T rhs2  -- a block
    tridiagonal positive definite system (batched block Cholesky, O(K nb^3) instead of O((K nb)^3)) -- and
    removes the O(delta) regularisation error with a few steps of iterative refinement on the exact KKT
    residual.  Positive definiteness of H + gamma J^T J is again the inertia test (Debreu), driving the
    per-instance dw."""

    def __init__(self, ipm, blk_free, gamma=1e6, refine=3, fused=True):
        self.ipm, self.gamma, self.refine = ipm, gamma, refine
        self.refine_rtol = 1e-7  # stop refining once every active instance's KKT residual is below this share of its rhs (0: always all passes)
        # lpb_blocktri_factor (one launch, factor in shared memory) beats the library recursion for small blocks
        # (nb = 44: 1.8 vs 3.5 ms per 4096-instance factorisation) and for small batches of large blocks (nb = 140,
        # 64 instances: 2.0 vs 2.9 ms -- the straggler tail), and loses for large batches of large blocks (nb = 140,
        # 4096 instances: 38 vs 29 ms, scripts/dev/blocktri_probe.py); _factor picks per call
        self.fused_factor = bool(fused)
        self.n_factor = self.n_solve = 0  # factorisations / block-tridiagonal solves so far
        self.fused_solve = None if fused else False  # None: not probed yet, True: lpb_blocktri_solve, False: library triangular solves
        ev, dev = ipm.ev, ipm.ev.device
        nf, me = ipm.nf, ipm.me
        K = int(blk_free.max().item()) + 1
        self.K = K
        # position of every free variable inside its block
        order = torch.argsort(blk_free * nf + torch.arange(nf, device=dev))
        counts = torch.bincount(blk_free, minlength=K)
        starts = torch.cumsum(counts, 0) - counts
        pos = torch.empty(nf, dtype=torch.int64, device=dev)
        pos[order] = torch.arange(nf, device=dev) - starts[blk_free[order]]
        nb = int(counts.max().item())
        self.nb = nb
        self.fpos = blk_free * nb + pos                      # free variable -> slot in the padded [K*nb] layout
        # Jacobian rows: group = lowest block among the row's free columns
        colmap = torch.full((ipm.n,), -1, dtype=torch.int64, device=dev)
        colmap[ipm.free] = torch.arange(nf, device=dev)
        rowmap = torch.full((ipm.m,), -1, dtype=torch.int64, device=dev)
        rowmap[ipm.eq] = torch.arange(me, device=dev)
        r, c = rowmap[ev.jI], colmap[ev.jJ]
        sel = torch.nonzero((r >= 0) & (c >= 0)).squeeze(1)
        rr, cb = r[sel], blk_free[c[sel]]
        big = torch.full((me,), K, dtype=torch.int64, device=dev)
        rmin = big.scatter_reduce(0, rr, cb, reduce="amin")
        rmax = torch.full((me,), -1, dtype=torch.int64, device=dev).scatter_reduce(0, rr, cb, reduce="amax")
        empty = rmax < 0
        rmin = torch.where(empty, 0, rmin)
        if bool(((rmax - rmin) > 1).any()):
            raise ValueError("a constraint row couples non-adjacent blocks")
        rorder = torch.argsort(rmin * me + torch.arange(me, device=dev))
        rcounts = torch.bincount(rmin, minlength=K)
        rstarts = torch.cumsum(rcounts, 0) - rcounts
        rpos = torch.empty(me, dtype=torch.int64, device=dev)
        rpos[rorder] = torch.arange(me, device=dev) - rstarts[rmin[rorder]]
        mr = max(1, int(rcounts.max().item()))
        self.mr = mr
        self.dpos = rmin * mr + rpos                          # equality row -> slot in the padded [K*mr] layout
        self.j_sel = sel
        # Hessian entries (lower triangle of the global matrix): same block -> D (both triangles), adjacent -> E
        hr, hc = colmap[ev.hI], colmap[ev.hJ]
        hs = torch.nonzero((hr >= 0) & (hc >= 0)).squeeze(1)
        br, bc = blk_free[hr[hs]], blk_free[hc[hs]]
        if bool(((br - bc).abs() > 1).any()):
            raise ValueError("a Hessian entry couples non-adjacent blocks")
        same = br == bc
        pr, pc = pos[hr[hs]], pos[hc[hs]]
        d1 = hs[same]
        offd = same & (hr[hs] != hc[hs])
        d2 = hs[offd]
        self.hd_sel = torch.cat([d1, d2])
        self.hd_flat = torch.cat([(br[same] * nb + pr[same]) * nb + pc[same], (br[offd] * nb + pc[offd]) * nb + pr[offd]])
        adj = ~same
        lo_is_c = br[adj] > bc[adj]   # row variable in the higher block
        lo = torch.where(lo_is_c, bc[adj], br[adj])
        p_hi = torch.where(lo_is_c, pr[adj], pc[adj])
        p_lo = torch.where(lo_is_c, pc[adj], pr[adj])
        # Boundary variables: the only slots of block i+1 that block i couples to (rows of E).  The defect rows of
        # interval i read just the first node of interval i+1 (column overlap 1 of the composite D matrix,
        # RPMGenerator.cpp:150-160), so E = [nbd x nb] with nbd ~ ns << nb: the off-diagonal solves, the rank
        # update of the next diagonal block and the J^T J product all shrink from nb to nbd rows.
        nxt = (cb - rmin[rr]) == 1
        bmask = torch.zeros(nb, dtype=torch.bool, device=dev)
        bmask[pos[c[sel]][nxt]] = True
        bmask[p_hi] = True
        if not bool(bmask.any()):
            bmask[0] = True                                   # decoupled blocks: one dummy boundary slot (E stays zero)
        self.bnd = torch.nonzero(bmask).squeeze(1)            # slots (same set for every block: union over blocks)
        nbd = int(self.bnd.numel())
        self.nbd = nbd
        bidx = torch.full((nb,), -1, dtype=torch.int64, device=dev)
        bidx[self.bnd] = torch.arange(nbd, device=dev)
        self.ncol = nb + nbd
        jcol = torch.where(nxt, nb + bidx[pos[c[sel]]], pos[c[sel]])
        self.j_flat = self.dpos[rr] * self.ncol + jcol        # Jb[row slot][own block | boundary of the next block]
        self.he_sel = hs[adj]
        self.he_flat = (lo * nbd + bidx[p_hi]) * nb + p_lo  # E[lo][boundary slot of block lo+1][col of block lo]
        real = torch.zeros(K * nb, dtype=torch.bool, device=dev)
        real[self.fpos] = True
        self.pad_diag = (~real).to(torch.float64).view(K, nb)  # 1 on padded slots (keeps the blocks non-singular)
        # sparse products for the refinement (lpb_batched_spmv): J, J^T and H applied straight from the triplet values
        # the transcription kernels wrote, in block-space indices (row slot dpos, column slot fpos) -- no dense block read
        self._sp = None
        if fused and dev.type == "cuda":
            Rj, Cj = self.dpos[rr], self.fpos[c[sel]]
            hi_, hj_ = self.fpos[hr[hs]], self.fpos[hc[hs]]
            off_ = hi_ != hj_
            self._sp = {"J": self._csr(Rj, Cj, sel, K * mr) + (K * mr, K * nb),
                        "Jt": self._csr(Cj, Rj, sel, K * nb) + (K * nb, K * mr),
                        "H": self._csr(torch.cat([hi_, hj_[off_]]), torch.cat([hj_, hi_[off_]]), torch.cat([hs, hs[off_]]), K * nb) + (K * nb, K * nb)}
            # the dense blocks the factorisation kernel reads, assembled by the same kernel as a gather: "row" = dense
            # position, entries = the triplets that sum into it, x = [1] (no zero fill, no atomics: index_add_ took 2x longer)
            zc = lambda t: torch.zeros_like(t)  # noqa: E731
            self._sp["Jb"] = self._csr(self.j_flat, zc(self.j_flat), self.j_sel, K * mr * self.ncol) + (K * mr * self.ncol, 1)
            self._sp["D"] = self._csr(self.hd_flat, zc(self.hd_flat), self.hd_sel, K * nb * nb) + (K * nb * nb, 1)
            if self.he_sel.numel():
                self._sp["E"] = self._csr(self.he_flat, zc(self.he_flat), self.he_sel, max(K - 1, 1) * nbd * nb) + (max(K - 1, 1) * nbd * nb, 1)

    @staticmethod
    def _csr(rows, cols, perm, nrows):
        """CSR structure (rowptr, col, perm as int32 device tensors) of the entries (rows, cols), in (row, perm) order."""
        order = torch.argsort(rows * (int(perm.max().item()) + 1 if perm.numel() else 1) + perm)
        counts = torch.bincount(rows[order], minlength=nrows)
        rowptr = torch.zeros(nrows + 1, dtype=torch.int64, device=rows.device)
        rowptr[1:] = torch.cumsum(counts, 0)
        return (rowptr.to(torch.int32).contiguous(), cols[order].to(torch.int32).contiguous(), perm[order].to(torch.int32).contiguous())

    def _ones(self, B, dev):
        o = getattr(self, "_ones_buf", None)
        if o is None or o.shape[0] < B or o.device != dev:
            o = self._ones_buf = torch.ones((B, 1), dtype=torch.float64, device=dev)
        return o[:B]

    def _spmv(self, which, vals, x, diag=None):
        """y = A x (+ diag .* x) through lpb_batched_spmv; vals [B, nnz] triplet values, x [B, ncols] -> [B, nrows]."""
        import ctypes as C
        rowptr, col, perm, nrows, ncols = self._sp[which]
        if getattr(self, "_lib", None) is None:
            from . import nlp
            self._lib = nlp.load_library()
        B = x.shape[0]
        x = x.reshape(B, ncols).contiguous()
        vals = vals if vals.is_contiguous() else vals.contiguous()
        y = torch.empty((B, nrows), dtype=torch.float64, device=x.device)
        dg = None if diag is None else diag.reshape(B, nrows).contiguous()
        rc = self._lib.lpb_batched_spmv(B, nrows, ncols, C.c_void_p(rowptr.data_ptr()), C.c_void_p(col.data_ptr()), C.c_void_p(perm.data_ptr()),
                                        C.c_void_p(vals.data_ptr()), vals.stride(0), C.c_void_p(x.data_ptr()),
                                        C.c_void_p(dg.data_ptr() if dg is not None else None), C.c_void_p(y.data_ptr()),
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc != 0:
            raise RuntimeError("lpb_batched_spmv failed: %d" % rc)
        return y

    # layout conversions
    def to_blocks(self, v):    # [B, nf] -> [B, K, nb]
        out = torch.zeros((v.shape[0], self.K * self.nb), dtype=torch.float64, device=v.device)
        out[:, self.fpos] = v
        return out.view(-1, self.K, self.nb)

    def from_blocks(self, vb):  # [B, K, nb] -> [B, nf]
        return vb.reshape(vb.shape[0], -1)[:, self.fpos]

    def dual_to_blocks(self, v):  # [B, me] -> [B, K, mr]
        out = torch.zeros((v.shape[0], self.K * self.mr), dtype=torch.float64, device=v.device)
        out[:, self.dpos] = v
        return out.view(-1, self.K, self.mr)

    def dual_from_blocks(self, vb):
        return vb.reshape(vb.shape[0], -1)[:, self.dpos]

    def set_jac(self, jv):
        B = jv.shape[0]
        if self._sp is not None and jv.is_cuda:
            Jb = self._spmv("Jb", jv, self._ones(B, jv.device))
        else:
            Jb = torch.zeros((B, self.K * self.mr * self.ncol), dtype=torch.float64, device=jv.device)
            Jb.index_add_(1, self.j_flat, jv[:, self.j_sel])
        self.Jb = Jb.view(B, self.K, self.mr, self.ncol)
        self.jv = jv  # triplet values as the kernels wrote them: the sparse products read them in place

    def _next_bnd(self, xb):   # [B, K, nb] -> boundary slots of the following block [B, K, nbd] (zeros after the last)
        nxt = xb[:, 1:, self.bnd]
        return torch.cat([nxt, torch.zeros_like(nxt[:, :1])], 1) if self.K > 1 else torch.zeros_like(xb[:, :, self.bnd])

    def _Jmul(self, xb):       # [B, K, nb] -> [B, K, mr]
        if self._sp is not None and xb.is_cuda:
            return self._spmv("J", self.jv, xb).view(-1, self.K, self.mr)
        return torch.einsum("bkmn,bkn->bkm", self.Jb, torch.cat([xb, self._next_bnd(xb)], 2))

    def _Jtmul(self, vb):      # [B, K, mr] -> [B, K, nb]
        if self._sp is not None and vb.is_cuda:
            return self._spmv("Jt", self.jv, vb).view(-1, self.K, self.nb)
        y = torch.einsum("bkmn,bkm->bkn", self.Jb, vb)
        out = y[:, :, :self.nb].clone()
        if self.K > 1:
            out[:, 1:, self.bnd] += y[:, :-1, self.nb:]
        return out

    def _Jtmul2(self, va, vb):  # J^T va and J^T vb in one sweep over the Jacobian blocks
        if self._sp is not None and va.is_cuda:
            return self._Jtmul(va), self._Jtmul(vb)
        y = torch.einsum("bkmn,bkmr->bknr", self.Jb, torch.stack([va, vb], 3))
        out = y[:, :, :self.nb].clone()
        if self.K > 1:
            out[:, 1:, self.bnd] += y[:, :-1, self.nb:]
        return out[..., 0], out[..., 1]

    def Jt(self, v):
        return self.from_blocks(self._Jtmul(self.dual_to_blocks(v)))

    def _Hmul(self, D, E, xb):
        out = torch.einsum("bkij,bkj->bki", D, xb)
        if self.K > 1 and self.he_sel.numel():
            out[:, 1:, self.bnd] += torch.einsum("bkij,bkj->bki", E, xb[:, :-1])
            out[:, :-1] += torch.einsum("bkij,bki->bkj", E, xb[:, 1:, self.bnd])
        return out

    def _factor_fused(self, Dp, Ep):
        """The whole block-tridiagonal Cholesky in ONE launch (lpb_blocktri_factor).  None if not available."""
        if not Dp.is_cuda or self.fused_solve is False:
            return None
        import ctypes as C
        if self.fused_solve is None:
            try:
                from . import nlp
                self._lib = nlp.load_library()
                self._bnd32 = self.bnd.to(torch.int32).contiguous()
                self.fused_solve = True
            except Exception:
                self.fused_solve = False
                return None
        B, K, nb, _ = Dp.shape
        Dp, Ep = Dp.contiguous(), Ep.contiguous()
        Lall, Call = torch.empty_like(Dp), torch.empty_like(Ep)
        info = torch.empty(B, dtype=torch.int32, device=Dp.device)
        rc = self._lib.lpb_blocktri_factor(B, K, nb, self.nbd, C.c_void_p(Dp.data_ptr()), C.c_void_p(Ep.data_ptr()), C.c_void_p(self._bnd32.data_ptr()),
                                           C.c_void_p(Lall.data_ptr()), C.c_void_p(Call.data_ptr()), C.c_void_p(info.data_ptr()),
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc == -1:
            return None
        if rc != 0:
            raise RuntimeError("lpb_blocktri_factor failed: %d" % rc)
        return [Lall[:, i] for i in range(K)], [Call[:, i] for i in range(K - 1)], info

    def _factor(self, Dp, Ep):
        """Block-tridiagonal Cholesky with boundary-row off-diagonal blocks: L_i L_i^T = A_i - (C_i C_i^T on the
        boundary slots), C_i = E_i L_{i-1}^-T [nbd x nb].  Returns (list L_i, list C_i, info)."""
        self.n_factor += 1
        fused = self._factor_fused(Dp, Ep) if self.fused_factor and (self.nb <= 64 or Dp.shape[0] <= 256) else None
        if fused is not None:
            return fused
        Ls, Cs = [], []
        info = torch.zeros(Dp.shape[0], dtype=torch.int32, device=Dp.device)
        eye = torch.eye(self.nb, dtype=torch.float64, device=Dp.device)
        bi = self.bnd
        prev = None
        for i in range(self.K):
            A = Dp[:, i]
            if i > 0:
                C = torch.linalg.solve_triangular(prev, Ep[:, i - 1].transpose(1, 2), upper=False).transpose(1, 2).contiguous()  # E L^-T
                Cs.append(C)
                A = A.clone()
                A[:, bi.unsqueeze(1), bi.unsqueeze(0)] -= torch.bmm(C, C.transpose(1, 2))
            L, inf_i = torch.linalg.cholesky_ex(A)
            bad = inf_i != 0
            info = torch.where((info == 0) & bad, inf_i, info)
            L = torch.where(bad.view(-1, 1, 1), eye.expand_as(L), L)  # keep failed instances finite
            Ls.append(L)
            prev = L
        return Ls, Cs, info

    def _solve_fused(self, Ls, Cs, rb, active=None):
        """The whole forward/backward block substitution in ONE launch (lpb_blocktri_solve, csrc/lpb_blocktri.cu)
        instead of 2K library triangular solves.  None if the extension or the shape is not available."""
        if not rb.is_cuda or self.fused_solve is False:
            return None
        import ctypes as C
        if self.fused_solve is None:
            try:
                from . import nlp
                self._lib = nlp.load_library()
                self._bnd32 = self.bnd.to(torch.int32).contiguous()
                self.fused_solve = True
            except Exception:
                self.fused_solve = False
                return None
        B, K, nb = rb.shape
        rb = rb.contiguous()
        out = torch.empty_like(rb)
        Lp = (C.c_void_p * K)(*[L.data_ptr() for L in Ls])
        Cp = (C.c_void_p * max(K - 1, 1))(*([c.data_ptr() for c in Cs] or [0]))
        sL = Ls[0].stride(0)
        sC = Cs[0].stride(0) if Cs else 0
        if any(L.stride(0) != sL or L.stride(1) != nb or L.stride(2) != 1 for L in Ls) or \
                any(c.stride(0) != sC or c.stride(1) != nb or c.stride(2) != 1 for c in Cs):
            return None
        rc = self._lib.lpb_blocktri_solve_masked(B, K, nb, self.nbd, Lp, Cp, sL, sC, C.c_void_p(self._bnd32.data_ptr()),
                                                 C.c_void_p(active.data_ptr() if active is not None else None), C.c_void_p(rb.data_ptr()),
                                                 C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc == -1:
            self.fused_solve = False  # shape not supported: library solves from now on
            return None
        if rc != 0:
            raise RuntimeError("lpb_blocktri_solve failed: %d" % rc)
        return out

    def _solve(self, Ls, Cs, rb, active=None):
        """active: optional uint8 [B] on the device; the fused kernel skips instances with 0 (their solution is zero)."""
        self.n_solve += 1
        fused = self._solve_fused(Ls, Cs, rb, active)
        if fused is not None:
            return fused
        bi = self.bnd
        ys = []
        for i in range(self.K):
            t = rb[:, i]
            if i > 0:
                t = t.clone()
                t[:, bi] -= torch.bmm(Cs[i - 1], ys[-1].unsqueeze(2)).squeeze(2)
            ys.append(torch.linalg.solve_triangular(Ls[i], t.unsqueeze(2), upper=False).squeeze(2))
        xs = [None] * self.K
        for i in range(self.K - 1, -1, -1):
            t = ys[i]
            if i + 1 < self.K:
                t = t - torch.bmm(Cs[i].transpose(1, 2), xs[i + 1][:, bi].unsqueeze(2)).squeeze(2)
            xs[i] = torch.linalg.solve_triangular(Ls[i].transpose(1, 2), t.unsqueeze(2), upper=True).squeeze(2)
        return torch.stack(xs, 1)

    def _assemble_and_factor(self, D, E, base, dw, done):
        """Library path of the assembly + factorisation (einsum, elementwise assembly, batched Cholesky, one more round
        per inertia-correction retry): the reference of lpb_kkt_factor and the fallback where it does not apply."""
        B, K, nb, nbd, g = D.shape[0], self.K, self.nb, self.nbd, self.gamma
        bi = self.bnd
        di = torch.arange(nb, device=D.device)
        G = torch.einsum("bkmi,bkmj->bkij", self.Jb, self.Jb)  # J_i^T J_i over [own block | boundary of block i+1]
        Dp = D + g * G[:, :, :nb, :nb]
        if K > 1 and nbd > 0:
            Dp[:, 1:, bi.unsqueeze(1), bi.unsqueeze(0)] += g * G[:, :-1, nb:, nb:]
            Ep = E + g * G[:, :-1, nb:, :nb]
        else:
            Ep = E
        Gd = torch.diagonal(G, dim1=2, dim2=3)  # [B, K, nb + nbd]
        gdiag = g * Gd[:, :, :nb].clone()
        if K > 1 and nbd > 0:
            gdiag[:, 1:, bi] += g * Gd[:, :-1, nb:]
        for _try in range(40):
            Dp[:, :, di, di] = base + gdiag + dw.view(-1, 1, 1)
            Ls, Cs, info = self._factor(Dp, Ep)
            bad = (info != 0) & (~done)
            if not bool(bad.any()):
                break
            dw = torch.where(bad, torch.where(dw == 0, 1e-4, dw * (100.0 if _try < 2 else 8.0)), dw)
        return Ls, Cs, info, dw

    def _kkt_factor_fused(self, D, E, diag_add, dw0, done):
        """lpb_kkt_factor: assembly (gamma J^T J, boundary coupling), block Cholesky and the inertia-correction retry in
        ONE launch.  None where it does not apply (CPU tensors, unsupported shape)."""
        if not D.is_cuda or self.fused_solve is False:
            return None
        import ctypes as C
        if self.fused_solve is None:
            try:
                from . import nlp
                self._lib = nlp.load_library()
                self._bnd32 = self.bnd.to(torch.int32).contiguous()
                self.fused_solve = True
            except Exception:
                self.fused_solve = False
                return None
        if getattr(self, "_kkt_fused_off", False):
            return None
        B, K, nb, nbd = D.shape[0], self.K, self.nb, self.nbd
        D, E, Jb, dg = D.contiguous(), E.contiguous(), self.Jb.contiguous(), diag_add.contiguous()
        Lall = torch.empty_like(D)
        Call = torch.empty_like(E)
        info = torch.empty(B, dtype=torch.int32, device=D.device)
        active = (~done).to(torch.uint8).contiguous()
        dw = dw0.clone().contiguous()
        rc = self._lib.lpb_kkt_factor(B, K, nb, nbd, self.mr, C.c_double(self.gamma), C.c_void_p(D.data_ptr()), C.c_void_p(E.data_ptr()),
                                      C.c_void_p(Jb.data_ptr()), C.c_void_p(dg.data_ptr()), C.c_void_p(self._bnd32.data_ptr()),
                                      C.c_void_p(active.data_ptr()), C.c_void_p(dw.data_ptr()), C.c_void_p(Lall.data_ptr()),
                                      C.c_void_p(Call.data_ptr()), C.c_void_p(info.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc == -1:
            self._kkt_fused_off = True
            return None
        if rc != 0:
            raise RuntimeError("lpb_kkt_factor failed: %d" % rc)
        self.n_factor += 1
        return [Lall[:, i] for i in range(K)], [Call[:, i] for i in range(K - 1)], info, dw

    def step(self, hv, Sigma, rhs1, c, dw_last, done):
        B, K, nb, nbd, g = hv.shape[0], self.K, self.nb, self.nbd, self.gamma
        dev = hv.device
        bi = self.bnd
        if self._sp is not None and hv.is_cuda:
            D = self._spmv("D", hv, self._ones(B, dev))
            E = self._spmv("E", hv, self._ones(B, dev)) if "E" in self._sp else torch.zeros((B, max(K - 1, 1) * nbd * nb), dtype=torch.float64, device=dev)
        else:
            D = torch.zeros((B, K * nb * nb), dtype=torch.float64, device=dev)
            D.index_add_(1, self.hd_flat, hv[:, self.hd_sel])
            E = torch.zeros((B, max(K - 1, 1) * nbd * nb), dtype=torch.float64, device=dev)
            if self.he_sel.numel():
                E.index_add_(1, self.he_flat, hv[:, self.he_sel])
        D = D.view(B, K, nb, nb)
        E = E.view(B, max(K - 1, 1), nbd, nb)
        di = torch.arange(nb, device=dev)
        base = D[:, :, di, di] + self.to_blocks(Sigma) + self.pad_diag
        dw = torch.where(dw_last > 0, dw_last / 3.0, 0.0)
        dw = torch.where(dw < 1e-9, 0.0, dw)
        fused = self._kkt_factor_fused(D, E, base - D[:, :, di, di], dw, done) if self.fused_factor else None
        if fused is not None:
            Ls, Cs, info, dw = fused
        else:
            Ls, Cs, info, dw = self._assemble_and_factor(D, E, base, dw, done)
        D[:, :, di, di] = base + dw.view(-1, 1, 1)  # H = W + Sigma + dw I (exact system of the refinement)
        r1b, cb = self.to_blocks(rhs1), self.dual_to_blocks(c)
        active = (~done).to(torch.uint8).contiguous()
        # H = W + Sigma + dw I (+ 1 on padded slots): the triplets carry W, the rest is a diagonal
        hdiag = (self.to_blocks(Sigma) + self.pad_diag + dw.view(-1, 1, 1)) if (self._sp is not None and hv.is_cuda) else None
        # iterative refinement on  [H J^T; J 0] [dx; dl] = -[r1; c].  H dx, J dx and J^T dl are carried along as running
        # sums of the products with the corrections (linear in them), so a pass costs one H product and two sweeps over
        # the dense Jacobian blocks (J ddx, and J^T [ddl, res2] in one) instead of one H product and four sweeps
        dxb, dlb = torch.zeros_like(r1b), torch.zeros_like(cb)
        Hdx, Jdx, Jtdl = torch.zeros_like(r1b), torch.zeros_like(cb), torch.zeros_like(r1b)
        res2 = -cb
        Jt_res2 = self._Jtmul(res2)
        rscale = torch.maximum(r1b.abs().amax((1, 2)), cb.abs().amax((1, 2))) + 1e-300
        for it in range(1 + self.refine):
            res1 = -r1b - (Hdx + Jtdl)
            if it >= 2 and self.refine_rtol > 0:
                # the residual of the exact KKT system drops by 1e-3 .. 1e-5 per pass; once every active instance is below
                # refine_rtol of its right-hand side the remaining passes would only polish digits the line search never sees
                rr = torch.maximum(res1.abs().amax((1, 2)), res2.abs().amax((1, 2))) / rscale
                if not bool((rr[~done] > self.refine_rtol).any()):
                    break
            ddx = self._solve(Ls, Cs, res1 + g * Jt_res2, active)
            Jddx = self._Jmul(ddx)
            ddl = g * (Jddx - res2)
            dxb = dxb + ddx
            dlb = dlb + ddl
            if it < self.refine:
                Hdx = Hdx + (self._spmv("H", hv, ddx, hdiag).view(B, K, nb) if hdiag is not None else self._Hmul(D, E, ddx))
                Jdx = Jdx + Jddx
                res2 = -cb - Jdx
                Jt_ddl, Jt_res2 = self._Jtmul2(ddl, res2)
                Jtdl = Jtdl + Jt_ddl
        failed = (info != 0) & (~done)  # still indefinite after the last regularisation try: status 2, not a silent step
        return self.from_blocks(dxb), self.dual_from_blocks(dlb), dw, failed


def interval_blocks(op, n_total):
    """Block id (mesh interval, phases concatenated) of every NLP variable of `op`; variables beyond the
    transcription's own (slacks) and phase times get -1 (= decided from the rows they appear in)."""
    blk = np.full(n_total, -1, dtype=np.int64)
    off, b0 = 0, 0
    for p in op.phases:
        ns, nc = len(p.statemin), len(p.controlmin)
        nodes = [int(v) for v in p.nodesperinterval]
        N, K = sum(nodes), len(nodes)
        node_blk = np.repeat(np.arange(K), nodes)
        for j in range(ns):
            blk[off + j * (N + 1): off + j * (N + 1) + N] = b0 + node_blk
            blk[off + j * (N + 1) + N] = b0 + K - 1
        c0 = off + ns * (N + 1)
        for j in range(nc):
            blk[c0 + j * N: c0 + (j + 1) * N] = b0 + node_blk
        off = c0 + nc * N + 2
        b0 += K
    return blk


class BatchedIPM:
    """Lockstep primal-dual interior-point method over a batch of instances of one NLP."""

    def __init__(self, ev, tol=1e-6, max_iter=100, mu0=0.1, rho=1e3, verbose=False, var_blocks=None, kkt_gamma=1e6, kkt_refine=3, kkt_fused=True, compact_at=0.5,
                 linesearch="filter", hessian="exact", lbfgs_memory=6):
        """var_blocks: optional block id (mesh interval) per NLP variable (`interval_blocks(op, n)`): selects the
        block-tridiagonal KKT step when the problem's coupling allows it, the dense condensed step otherwise."""
        self.var_blocks = var_blocks
        self.kkt_gamma, self.kkt_refine, self.kkt_fused = kkt_gamma, kkt_refine, kkt_fused
        self.compact_at = compact_at  # resume on the unconverged instances once fewer than this fraction is still active
        self.kkt_kind = "dense"
        _, _, gl, gu = ev.bounds()
        self.user_ev = ev
        if bool((gl != gu).any()):
            ev = SlackEvaluator(ev)  # inequality rows -> equalities with bounded slacks
        self.ev, self.tol, self.max_iter, self.mu0, self.rho, self.verbose = ev, tol, max_iter, mu0, rho, verbose
        self.n, self.m = ev.n, ev.m
        # globalisation: "filter" = filter line search in the manner of IPOPT (Waechter & Biegler 2006, section 2.3:
        # a trial point is acceptable when it improves the constraint violation OR the barrier objective enough and
        # is not dominated by the filter; an Armijo condition on the objective when the switching condition holds),
        # the reference's setting (IPOPT defaults, LpNLPSolver.cpp:27-33); "merit" = l1 exact penalty with Armijo
        self.linesearch = linesearch
        self.filter_slots = 24
        # hessian = "exact": eval_h (the reference's "hessian-approximation" = "exact"); "limited-memory": damped
        # L-BFGS on the Lagrangian gradient with `lbfgs_memory` pairs, IPOPT's default and the reference's
        # (LpNLPSolver.cpp:30-33, LpNLPWrapper.hpp:69-76) -- no eval_h calls at all.  Dense KKT step only.
        self.hessian = hessian
        self.lbfgs_memory = lbfgs_memory

    # ---- problem structure shared by all instances ------------------------------------------------
    def _setup(self, xl, xu, gl, gu):
        ev, dev = self.ev, self.ev.device
        fixed = xl[0] == xu[0]
        if not bool(((xl == xu) == fixed.unsqueeze(0)).all()):
            raise ValueError("the set of fixed variables must be the same for every instance")
        if getattr(self, "_setup_fixed", None) is not None and torch.equal(fixed, self._setup_fixed):
            return  # same structure as the previous chunk / resumed sub-batch: index maps and KKT layout are reused
        self._setup_fixed = fixed.clone()
        self.free = torch.nonzero(~fixed).squeeze(1)
        self.nf = int(self.free.numel())
        colmap = torch.full((self.n,), -1, dtype=torch.int64, device=dev)
        colmap[self.free] = torch.arange(self.nf, device=dev)
        eq = gl == gu  # all rows: inequality rows were turned into equalities with slacks (SlackEvaluator)
        self.eq = torch.nonzero(eq).squeeze(1)
        self.me = int(self.eq.numel())
        rowmap = torch.full((self.m,), -1, dtype=torch.int64, device=dev)
        rowmap[self.eq] = torch.arange(self.me, device=dev)
        r, c = rowmap[ev.jI], colmap[ev.jJ]
        self.j_sel = torch.nonzero((r >= 0) & (c >= 0)).squeeze(1)
        self.j_flat = r[self.j_sel] * self.nf + c[self.j_sel]
        hr, hc = colmap[ev.hI], colmap[ev.hJ]
        sel = torch.nonzero((hr >= 0) & (hc >= 0)).squeeze(1)
        off = sel[hr[sel] != hc[sel]]
        self.h_sel = torch.cat([sel, off])  # lower triangle + its mirror
        self.h_flat = torch.cat([hr[sel] * self.nf + hc[sel], hc[off] * self.nf + hr[off]])
        self.c_target = gl[self.eq]
        self.kkt = DenseKKT(self)
        self.kkt_kind = "dense"
        if self.var_blocks is not None:
            try:
                self.kkt = BlockTridiagKKT(self, self._free_blocks(), gamma=self.kkt_gamma, refine=self.kkt_refine, fused=self.kkt_fused)
                self.kkt_kind = "block-tridiagonal (K=%d, nb=%d, boundary=%d)" % (self.kkt.K, self.kkt.nb, self.kkt.nbd)
            except ValueError:
                pass  # coupling does not fit (free phase times, x0-xf Mayer terms, ...): dense step

    def _free_blocks(self):
        """Block id of every free variable: given ids where known, otherwise the lowest block among the variables
        sharing a constraint row (slacks, which belong to the node of their row; isolated ones go to block 0)."""
        ev, dev = self.ev, self.ev.device
        vb = torch.as_tensor(np.asarray(self.var_blocks), dtype=torch.int64, device=dev)
        blk = torch.full((self.n,), -1, dtype=torch.int64, device=dev)
        blk[: vb.numel()] = vb
        free_mask = torch.zeros(self.n, dtype=torch.bool, device=dev)
        free_mask[self.free] = True
        unknown = torch.nonzero((blk < 0) & free_mask).squeeze(1)
        if unknown.numel():
            big = int(blk.max().item()) + 1
            known = (blk[ev.jJ] >= 0) & free_mask[ev.jJ]
            rowmin = torch.full((self.m,), big, dtype=torch.int64, device=dev).scatter_reduce(0, ev.jI[known], blk[ev.jJ[known]], reduce="amin")
            for v in unknown.tolist():
                rows = ev.jI[ev.jJ == v]
                b = int(rowmin[rows].min().item()) if rows.numel() else big
                blk[v] = 0 if b >= big else b
        out = blk[self.free]
        if bool((out < 0).any()):
            raise ValueError("free variable without a block")
        return out

    def _dense_jac(self, vals):
        B = vals.shape[0]
        J = torch.zeros((B, self.me * self.nf), dtype=torch.float64, device=vals.device)
        J.index_add_(1, self.j_flat, vals[:, self.j_sel])  # duplicate triplets sum (TNLP contract)
        return J.view(B, self.me, self.nf)

    def _dense_hess(self, vals):
        B = vals.shape[0]
        H = torch.zeros((B, self.nf * self.nf), dtype=torch.float64, device=vals.device)
        H.index_add_(1, self.h_flat, vals[:, self.h_sel])
        return H.view(B, self.nf, self.nf)

    # ---- limited-memory quasi-Newton Hessian of the Lagrangian ----------------------------------------
    def _lbfgs_push(self, S, Y, cnt, sk, yk, active):
        """Appends the pair (s, y) of every active instance (Powell-damped against B0 = delta I so that s'y > 0) and
        drops the oldest when the memory is full."""
        sy = (sk * yk).sum(1)
        ss = (sk * sk).sum(1)
        yy = (yk * yk).sum(1)
        delta = torch.where(sy > 0, yy / sy.clamp(min=1e-300), 1.0)
        # Powell damping: y <- theta y + (1 - theta) delta s when s'y < 0.2 delta s's
        th = torch.where(sy < 0.2 * delta * ss, 0.8 * delta * ss / (delta * ss - sy).clamp(min=1e-300), 1.0)
        yk = th.unsqueeze(1) * yk + (1 - th).unsqueeze(1) * delta.unsqueeze(1) * sk
        use = active & (ss > 1e-24) & torch.isfinite(yk).all(1)
        M = S.shape[1]
        full = cnt >= M
        S = torch.where((use & full).view(-1, 1, 1), torch.roll(S, -1, 1), S)
        Y = torch.where((use & full).view(-1, 1, 1), torch.roll(Y, -1, 1), Y)
        slot = torch.where(full, M - 1, cnt)
        rows = torch.nonzero(use).squeeze(1)
        S[rows, slot[rows]] = sk[rows]
        Y[rows, slot[rows]] = yk[rows]
        cnt = torch.where(use & ~full, cnt + 1, cnt)
        return S, Y, cnt

    def _lbfgs_matrix(self, S, Y, cnt):
        """Dense B = delta I - [delta S' Y'] [[delta S S', L], [L', -D]]^-1 [delta S; Y] per instance (compact L-BFGS,
        Byrd / Nocedal / Schnabel 1994), unused slots masked out."""
        B, M, nf = S.shape
        dev = S.device
        valid = (torch.arange(M, device=dev).unsqueeze(0) < cnt.unsqueeze(1)).to(S.dtype)  # [B, M]
        S, Y = S * valid.unsqueeze(2), Y * valid.unsqueeze(2)
        last = (cnt - 1).clamp(min=0)
        rows = torch.arange(B, device=dev)
        sy_last = (S[rows, last] * Y[rows, last]).sum(1)
        yy_last = (Y[rows, last] * Y[rows, last]).sum(1)
        delta = torch.where((cnt > 0) & (sy_last > 0), yy_last / sy_last.clamp(min=1e-300), 1.0)
        SY = torch.bmm(S, Y.transpose(1, 2))                      # s_i' y_j
        L = torch.tril(SY, diagonal=-1)
        D = torch.diagonal(SY, dim1=1, dim2=2)
        SS = torch.bmm(S, S.transpose(1, 2)) * delta.view(-1, 1, 1)
        inval = torch.diag_embed(1.0 - valid)                    # identity on the unused slots keeps the middle matrix regular
        top = torch.cat([SS + inval, L], dim=2)
        bot = torch.cat([L.transpose(1, 2), -torch.diag_embed(D) - inval], dim=2)
        Mid = torch.cat([top, bot], dim=1)                        # [B, 2M, 2M]
        Wt = torch.cat([S * delta.view(-1, 1, 1), Y], dim=1)      # [B, 2M, nf]
        core = torch.linalg.solve(Mid, Wt)
        Bm = -torch.bmm(Wt.transpose(1, 2), core)
        di = torch.arange(nf, device=dev)
        Bm[:, di, di] += delta.unsqueeze(1)
        return Bm

    # ---- solve ---------------------------------------------------------------------------------------
    def solve(self, x0, xl=None, xu=None, chunk=1024):
        """x0: [B, n] starting points; xl/xu: per-instance bounds [B, n] (default: the problem's).
        Instances are processed `chunk` at a time (the dense condensed systems of one chunk must fit
        in HBM: about 40 MB per quadrotor instance).  Returns dict(x, obj, status, iters, kkt_error, lam)."""
        B = len(x0)
        outs = []
        for b0 in range(0, B, chunk):
            sl_ = slice(b0, min(B, b0 + chunk))
            outs.append(self._solve_compacting(x0[sl_], None if xl is None else xl[sl_], None if xu is None else xu[sl_]))
        return {k: torch.cat([o[k] for o in outs]) for k in outs[0]}

    def _solve_compacting(self, x0, xl, xu):
        """Lockstep iterations over the instances still active: whenever fewer than half of the current
        sub-batch is unconverged, the iteration is suspended and resumed on the unconverged instances
        only (state carried over), so stragglers do not pay for the whole batch's linear algebra."""
        ev = self.ev
        B = len(x0)
        bxl, bxu, _, _ = self.user_ev.bounds()
        XL = bxl.expand(B, -1) if xl is None else ev.tensor(xl)
        XU = bxu.expand(B, -1) if xu is None else ev.tensor(xu)
        x0 = ev.tensor(x0)
        nx = x0.shape[1]
        if isinstance(ev, SlackEvaluator):
            x0, XL, XU = ev.extend(x0, XL, XU)
        res = self._solve_chunk(x0, XL, XU, None, 0)
        out = {k: v.clone() for k, v in res.items() if k != "state"}
        ids = torch.arange(B, device=out["x"].device)
        while res["state"] is not None:
            st = res["state"]
            act = st["active"]
            ids = ids[act]
            sub = {k: (v[act] if isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] == act.shape[0] else v) for k, v in st.items()}
            res = self._solve_chunk(sub["X"], XL[ids], XU[ids], sub, st["it"])
            for k in out:
                out[k][ids] = res[k]
        out["x"] = out["x"][:, :nx].contiguous()
        return out

    def _solve_chunk(self, x0, xl, xu, state, it0):
        ev = self.ev
        X = ev.tensor(x0).clone()
        B = X.shape[0]
        _, _, gl, gu = ev.bounds()
        self._setup(xl, xu, gl, gu)
        F, nf, me = self.free, self.nf, self.me
        fixed_mask = torch.ones(self.n, dtype=torch.bool, device=X.device)
        fixed_mask[F] = False
        X[:, fixed_mask] = xl[:, fixed_mask]
        lo, hi = xl[:, F], xu[:, F]
        hasL, hasU = torch.isfinite(lo), torch.isfinite(hi)
        if state is None:
            # push the starting point into the interior (IPOPT bound_push / bound_frac = 1e-2)
            xf = X[:, F]
            pl = torch.minimum(1e-2 * torch.clamp(lo.abs(), min=1.0), 1e-2 * (hi - lo))
            pu = torch.minimum(1e-2 * torch.clamp(hi.abs(), min=1.0), 1e-2 * (hi - lo))
            xf = torch.where(hasL & hasU, torch.minimum(torch.maximum(xf, lo + pl), hi - pu), xf)
            xf = torch.where(hasL & ~hasU, torch.maximum(xf, lo + 1e-2 * torch.clamp(lo.abs(), min=1.0)), xf)
            xf = torch.where(~hasL & hasU, torch.minimum(xf, hi - 1e-2 * torch.clamp(hi.abs(), min=1.0)), xf)
            X[:, F] = xf
            lam = torch.zeros((B, self.m), dtype=torch.float64, device=X.device)
            zL = hasL.to(torch.float64)
            zU = hasU.to(torch.float64)
            mu = torch.full((B,), self.mu0, dtype=torch.float64, device=X.device)
            nu = torch.full((B,), 1.0, dtype=torch.float64, device=X.device)  # l1 penalty
            dw_last = torch.zeros(B, dtype=torch.float64, device=X.device)
            iters = torch.zeros(B, dtype=torch.int64, device=X.device)
            # per-instance filter: (theta, phi) pairs a trial point must not be dominated by; slot 0 holds the upper
            # bound on the constraint violation, set at the first iteration
            f_th = torch.full((B, self.filter_slots), float("inf"), dtype=torch.float64, device=X.device)
            f_ph = torch.full((B, self.filter_slots), float("inf"), dtype=torch.float64, device=X.device)
            f_n = torch.zeros(B, dtype=torch.int64, device=X.device)
            th_min = torch.full((B,), -1.0, dtype=torch.float64, device=X.device)
        else:
            lam, zL, zU, mu, nu, dw_last, iters = (state[k] for k in ("lam", "zL", "zU", "mu", "nu", "dw_last", "iters"))
            f_th, f_ph, f_n, th_min = (state[k] for k in ("f_th", "f_ph", "f_n", "th_min"))
        done = torch.zeros(B, dtype=torch.bool, device=X.device)
        failed = torch.zeros(B, dtype=torch.bool, device=X.device)
        sigma1 = torch.ones(B, dtype=torch.float64, device=X.device)
        kkt0 = torch.full((B,), float("inf"), dtype=torch.float64, device=X.device)
        inf = float("inf")
        sl = lambda x: torch.where(hasL, x - lo, 1.0)  # noqa: E731
        su = lambda x: torch.where(hasU, hi - x, 1.0)  # noqa: E731

        def barrier_obj(Xt, mu_):
            xt = Xt[:, F]
            phi = ev.f(Xt)
            phi = phi - mu_ * (torch.where(hasL, torch.log(sl(xt)), 0.0).sum(1)
                               + torch.where(hasU, torch.log(su(xt)), 0.0).sum(1))
            c = ev.g(Xt)[:, self.eq] - self.c_target
            return phi, c.abs().sum(1)

        suspended = None
        lm = self.hessian == "limited-memory"
        if lm:
            if not isinstance(self.kkt, DenseKKT) and state is None:
                self._setup_fixed = None  # rebuild with the dense step
                vb, self.var_blocks = self.var_blocks, None
                self._setup(xl, xu, gl, gu)
                self.var_blocks = vb
            M = self.lbfgs_memory
            lb_S = torch.zeros((B, M, nf), dtype=torch.float64, device=X.device)
            lb_Y = torch.zeros((B, M, nf), dtype=torch.float64, device=X.device)
            lb_n = torch.zeros(B, dtype=torch.int64, device=X.device)
            x_prev = glag_prev = None
        for it in range(it0, self.max_iter):
            g, jv = ev.g_jac(X)
            gradf = ev.grad(X)[:, F]
            c = g[:, self.eq] - self.c_target
            kkt = self.kkt
            kkt.set_jac(jv)
            xf = X[:, F]
            sL, sU = sl(xf), su(xf)
            lam_eq = lam[:, self.eq]
            Jtlam = kkt.Jt(lam_eq)
            rd = gradf + Jtlam - zL + zU
            if lm:
                # secant pair of the Lagrangian gradient at the CURRENT multipliers (grad L(x_k, lam_k) - grad L(x_{k-1}, lam_k))
                if x_prev is not None:
                    sk = xf - x_prev
                    yk = (gradf + Jtlam) - (gradf_prev + kkt_Jt_prev(lam_eq))
                    lb_S, lb_Y, lb_n = self._lbfgs_push(lb_S, lb_Y, lb_n, sk, yk, ~done)
                x_prev, gradf_prev = xf.clone(), gradf.clone()
                Jprev = kkt.J.clone()
                kkt_Jt_prev = lambda v, Jp=Jprev: torch.bmm(Jp.transpose(1, 2), v.unsqueeze(2)).squeeze(2)  # noqa: E731
            compL = torch.where(hasL, sL * zL, 0.0)
            compU = torch.where(hasU, sU * zU, 0.0)
            # scaled optimality error (IPOPT eq. (5)-(6)), smax = 100
            nz = (hasL.sum(1) + hasU.sum(1)).clamp(min=1)
            s_d = torch.clamp((lam_eq.abs().sum(1) + zL.sum(1) + zU.sum(1)) / (me + nz), min=100.0) / 100.0
            s_c = torch.clamp((zL.sum(1) + zU.sum(1)) / nz, min=100.0) / 100.0

            def err(mu_):
                return torch.stack([rd.abs().amax(1) / s_d, c.abs().amax(1),
                                    torch.maximum((compL - torch.where(hasL, mu_.unsqueeze(1), 0.0)).abs().amax(1),
                                                  (compU - torch.where(hasU, mu_.unsqueeze(1), 0.0)).abs().amax(1)) / s_c]).amax(0)

            e0 = err(torch.zeros_like(mu))
            kkt0 = torch.where(done, kkt0, e0)
            newly = (~done) & (e0 <= self.tol)
            done |= newly
            if self.verbose:
                print("it %3d  active %4d  max E0 %.3e  mu[min,max] %.1e %.1e  |c|max %.2e  |rd|max %.2e  dw max %.1e  alpha[min] %.2e"
                      % (it, int((~done).sum()), float(e0[~done].max()) if (~done).any() else 0.0, float(mu.min()), float(mu.max()),
                         float(c.abs().max()), float(rd.abs().max()), float(dw_last.max()), float(getattr(self, "_last_alpha", torch.ones(1)).min())))
            if bool(done.all()):
                break
            n_act = int((~done).sum())
            if B >= 16 and n_act < self.compact_at * B:  # resume on the unconverged instances only
                suspended = {"X": X, "lam": lam, "zL": zL, "zU": zU, "mu": mu, "nu": nu, "dw_last": dw_last, "iters": iters,
                             "f_th": f_th, "f_ph": f_ph, "f_n": f_n, "th_min": th_min, "active": ~done, "it": it}
                break
            # barrier update (monotone): while E_mu <= kappa_eps * mu
            for _ in range(1):
                emu = err(mu)
                upd = (~done) & (emu <= 10.0 * mu) & (mu > self.tol / 10.0)
                if not bool(upd.any()):
                    break
                mu = torch.where(upd, torch.clamp(torch.minimum(0.2 * mu, mu ** 1.5), min=self.tol / 10.0), mu)
                # a new barrier problem: its filter starts empty (the objective phi_mu changed)
                f_th = torch.where(upd.unsqueeze(1), float("inf"), f_th)
                f_ph = torch.where(upd.unsqueeze(1), float("inf"), f_ph)
                f_n = torch.where(upd, 0, f_n)
            mu_c = mu.unsqueeze(1)
            # Hessian of the Lagrangian (exact, finite differences of the reference scheme) + Sigma
            if self.hessian == "limited-memory":
                hv = self._lbfgs_matrix(lb_S, lb_Y, lb_n)
            else:
                hv = ev.hess(X, sigma1, lam)
            Sigma = torch.where(hasL, zL / sL, 0.0) + torch.where(hasU, zU / sU, 0.0)
            rhs1 = gradf + Jtlam - torch.where(hasL, mu_c / sL, 0.0) + torch.where(hasU, mu_c / sU, 0.0)
            dx, dlam, dw, kfail = kkt.step(hv, Sigma, rhs1, c, dw_last, done)
            failed |= kfail
            dw_last = torch.where(done, dw_last, dw)
            dzL = torch.where(hasL, mu_c / sL - zL - zL / sL * dx, 0.0)
            dzU = torch.where(hasU, mu_c / sU - zU + zU / sU * dx, 0.0)
            # fraction to the boundary
            tau = torch.clamp(1.0 - mu, min=0.99).unsqueeze(1)
            ratio = torch.full_like(dx, inf)
            ratio = torch.where(hasL & (dx < 0), -tau * sL / dx, ratio)
            ratio = torch.minimum(ratio, torch.where(hasU & (dx > 0), tau * sU / dx, inf))
            a_max = torch.clamp(ratio.amin(1), max=1.0)
            rz = torch.full_like(dx, inf)
            rz = torch.where(hasL & (dzL < 0), -tau * zL / dzL, rz)
            rz = torch.minimum(rz, torch.where(hasU & (dzU < 0), -tau * zU / dzU, inf))
            a_z = torch.clamp(rz.amin(1), max=1.0)
            gphi = gradf - torch.where(hasL, mu_c / sL, 0.0) + torch.where(hasU, mu_c / sU, 0.0)
            c1 = c.abs().sum(1)
            gd = (gphi * dx).sum(1)
            phi0, _ = barrier_obj(X, mu)
            alpha = a_max.clone()
            accepted = done.clone()
            Xn = X.clone()
            if self.linesearch == "filter":
                g_th, g_ph, eta, delta, s_th, s_ph = 1e-5, 1e-8, 1e-8, 1.0, 1.1, 2.3
                first = th_min < 0
                th_min = torch.where(first, 1e-4 * torch.clamp(c1, min=1.0), th_min)
                th_max = 1e4 * torch.clamp(c1, min=1.0)
                init = first | (f_n == 0)  # an empty filter only bounds the constraint violation
                f_th[:, 0] = torch.where(init, th_max, f_th[:, 0])
                f_ph[:, 0] = torch.where(init, -float("inf"), f_ph[:, 0])
                f_n = torch.where(init, 1, f_n)
                ftype = torch.zeros_like(done)
                a_min_ls = 1e-9
                for _ls in range(40):
                    Xt = X.clone()
                    Xt[:, F] = xf + alpha.unsqueeze(1) * dx
                    phit, tht = barrier_obj(Xt, mu)
                    finite = torch.isfinite(phit) & torch.isfinite(tht)
                    # not dominated by any filter entry
                    in_filter = ((tht.unsqueeze(1) >= (1 - g_th) * f_th) & (phit.unsqueeze(1) >= f_ph - g_ph * f_th)).any(1)
                    switching = (gd < 0) & (alpha * (-gd).clamp(min=0) ** s_ph > delta * c1 ** s_th) & (c1 <= th_min)
                    armijo = phit <= phi0 + eta * alpha * gd + 1e-13 * phi0.abs()
                    # h-type step: less constraint violation, or a better barrier objective -- the latter only while the
                    # violation at most doubles (there is no restoration phase to recover from a run-away of theta)
                    progress = (tht <= (1 - g_th) * c1) | ((phit <= phi0 - g_ph * c1) & (tht <= torch.maximum(2.0 * c1, th_min)))
                    ok = finite & ~in_filter & torch.where(switching, armijo, progress)
                    take = ok & ~accepted
                    Xn[take] = Xt[take]
                    ftype = torch.where(take, switching & armijo, ftype)
                    accepted |= take
                    if bool(accepted.all()) or float(alpha[~accepted].max()) < a_min_ls:
                        break
                    alpha = torch.where(accepted, alpha, alpha * 0.5)
                # accepted h-type steps (and f-type steps that did not satisfy Armijo) augment the filter
                aug = accepted & ~done & ~ftype
                if bool(aug.any()):
                    slot = torch.where(f_n < self.filter_slots, f_n, 1)  # full: overwrite from slot 1 on (slot 0 = bound)
                    rows = torch.nonzero(aug).squeeze(1)
                    f_th[rows, slot[rows]] = ((1 - g_th) * c1)[rows]
                    f_ph[rows, slot[rows]] = (phi0 - g_ph * c1)[rows]
                    f_n = torch.where(aug, torch.where(f_n < self.filter_slots, f_n + 1, 2), f_n)
            else:
                # l1 merit: phi_mu(x) + nu |c|_1, Armijo backtracking
                lam_new_inf = (lam_eq + dlam).abs().amax(1)
                nu = torch.where(nu < lam_new_inf + 1.0, lam_new_inf * 1.5 + 1.0, nu)
                dphi = gd - nu * c1
                merit0 = phi0 + nu * c1
                for _ls in range(25):
                    Xt = X.clone()
                    Xt[:, F] = xf + alpha.unsqueeze(1) * dx
                    phit, ct = barrier_obj(Xt, mu)
                    ok = (phit + nu * ct <= merit0 + 1e-8 * alpha * torch.clamp(dphi, max=0.0) + 1e-12 * merit0.abs()) & torch.isfinite(phit)
                    take = ok & ~accepted
                    Xn[take] = Xt[take]
                    accepted |= take
                    if bool(accepted.all()):
                        break
                    alpha = torch.where(accepted, alpha, alpha * 0.5)
            # instances whose line search failed: take the tiny step anyway and raise dw next time
            stuck = ~accepted
            if bool(stuck.any()):
                Xt = X.clone()
                Xt[:, F] = xf + alpha.unsqueeze(1) * dx
                Xn[stuck] = Xt[stuck]
                dw_last = torch.where(stuck, torch.clamp(dw_last * 100.0, min=1e-2), dw_last)
                if self.linesearch == "filter":  # restart the filter of a stuck instance
                    f_th = torch.where(stuck.unsqueeze(1), float("inf"), f_th)
                    f_ph = torch.where(stuck.unsqueeze(1), float("inf"), f_ph)
                    f_n = torch.where(stuck, 0, f_n)
            self._last_alpha = alpha
            act = (~done).unsqueeze(1)
            a_col = alpha.unsqueeze(1)
            X = torch.where(act, Xn, X)
            lam_eq_new = lam_eq + a_col * dlam
            lam_full = lam.clone()
            lam_full[:, self.eq] = lam_eq_new
            lam = torch.where(act, lam_full, lam)
            zL = torch.where(act, zL + a_z.unsqueeze(1) * dzL, zL)
            zU = torch.where(act, zU + a_z.unsqueeze(1) * dzU, zU)
            # keep z within [mu/(k s), k mu/s] (IPOPT eq. (16), kappa_Sigma = 1e10)
            xf2 = X[:, F]
            sL2, sU2 = sl(xf2), su(xf2)
            zL = torch.where(hasL, torch.maximum(torch.minimum(zL, 1e10 * mu_c / sL2), mu_c / (1e10 * sL2)), zL)
            zU = torch.where(hasU, torch.maximum(torch.minimum(zU, 1e10 * mu_c / sU2), mu_c / (1e10 * sU2)), zU)
            iters += (~done).to(torch.int64)
        obj = ev.f(X)
        status = torch.where(done, 0, 1)
        status = torch.where(failed & ~done, 2, status)
        return {"x": X, "obj": obj, "status": status, "iters": iters, "kkt_error": kkt0, "lam": lam, "state": suspended}
