"""CPU evaluator for lpopc_b200.solver.BatchedIPM backed by the oracle (TEST INFRASTRUCTURE):
the same interior-point iteration run on the CPU restatement of the reference path gives the
reference objective values the GPU solves are compared with."""
import numpy as np
import torch

from oracle_lib import Oracle


class OracleEvaluator:
    def __init__(self, op, threads=4):
        self.o = Oracle(op)
        self.device = torch.device("cpu")
        self.threads = threads
        self.n, self.m, self.nnz_jac, self.nnz_h = self.o.nlp_info()
        jI, jJ = self.o.jac_structure()
        hI, hJ = self.o.h_structure()
        self.jI, self.jJ = torch.from_numpy(jI.astype(np.int64)), torch.from_numpy(jJ.astype(np.int64))
        self.hI, self.hJ = torch.from_numpy(hI.astype(np.int64)), torch.from_numpy(hJ.astype(np.int64))

    def bounds(self):
        return [torch.from_numpy(a) for a in self.o.bounds()]

    def tensor(self, a):
        return torch.as_tensor(np.asarray(a.cpu() if isinstance(a, torch.Tensor) else a, dtype=np.float64))

    def f(self, X):
        return torch.tensor([self.o.eval_f(x) for x in X.numpy()], dtype=torch.float64)

    def grad(self, X):
        return torch.from_numpy(np.stack([self.o.eval_grad_f(x) for x in X.numpy()]))

    def g(self, X):
        g, _ = self.o.eval_g_jac_batch(X.numpy(), nthreads=self.threads, want_jac=False)
        return torch.from_numpy(g)

    def g_jac(self, X):
        g, v = self.o.eval_g_jac_batch(X.numpy(), nthreads=self.threads)
        return torch.from_numpy(g), torch.from_numpy(v)

    def hess(self, X, sigma, lam):
        return torch.from_numpy(np.stack([self.o.eval_h(x, float(s), l) for x, s, l in zip(X.numpy(), sigma.numpy(), lam.numpy())]))


def mpc_instances(op, lgr_points, x0s):
    from lpopc_b200 import batch
    return batch.mpc_starting_points(op, lgr_points, x0s)


def mpc_bounds(ev, op, x0s):
    from lpopc_b200 import batch
    xl, xu, _, _ = ev.bounds()
    return batch.mpc_bounds(xl, xu, op, x0s)
