"""ctypes binding of the CPU oracle (oracle/liblpopc_oracle.so) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import this.  Builds the library with oracle/Makefile when it is missing.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "oracle", "liblpopc_oracle.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build_oracle(force=False):
    srcs = [os.path.join(ROOT, "oracle", f) for f in os.listdir(os.path.join(ROOT, "oracle")) if f.endswith((".cpp", ".hpp"))]
    hdrs = []
    for d in (os.path.join(ROOT, "include"), os.path.join(ROOT, "include", "problems")):
        hdrs += [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".h")]
    stale = force or not os.path.exists(LIB_PATH)
    if not stale:
        t = os.path.getmtime(LIB_PATH)
        stale = any(os.path.getmtime(s) > t for s in srcs + hdrs)
    if stale:
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        _lib = C.CDLL(LIB_PATH)
        _lib.lpo_create.restype = C.c_void_p
        _lib.lpo_last_error.restype = C.c_char_p
        for name in ("lpo_destroy", "lpo_last_error", "lpo_set_mesh", "lpo_refresh", "lpo_get_nlp_info", "lpo_get_bounds_info",
                     "lpo_eval_f", "lpo_eval_grad_f", "lpo_eval_g", "lpo_eval_jac_g", "lpo_eval_h", "lpo_probe_dependencies",
                     "lpo_get_tables", "lpo_get_coo", "lpo_eval_g_jac_batch"):
            getattr(_lib, name).argtypes = None
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


class Oracle:
    """CPU restatement of the reference path for one OptimalProblem."""

    def _load(self):
        return lib()

    def __init__(self, op):
        self.op = op
        self.L = self._load()
        desc, self._keep = op.to_desc()
        err = C.create_string_buffer(512)
        self.h = C.c_void_p(self.L.lpo_create(C.byref(desc), err, 512))
        if not self.h:
            raise RuntimeError("oracle create failed: " + err.value.decode())
        for ip, p in enumerate(op.phases):
            if p.meshpoints:
                self.set_mesh(ip, p.meshpoints, p.nodesperinterval)
        self._check(self.L.lpo_refresh(self.h))
        self.n, self.m, self.nnz_jac, self.nnz_h = self.nlp_info()

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("oracle: " + self.L.lpo_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.L.lpo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_mesh(self, phase, meshpoints, nodes):
        mp = np.ascontiguousarray(meshpoints, dtype=np.float64)
        nd = np.ascontiguousarray(nodes, dtype=np.int32)
        self._check(self.L.lpo_set_mesh(self.h, C.c_int(phase), C.c_int(len(nd)), _d(mp), _i(nd)))

    def refresh(self):
        self._check(self.L.lpo_refresh(self.h))
        self.n, self.m, self.nnz_jac, self.nnz_h = self.nlp_info()

    def nlp_info(self):
        v = [C.c_int() for _ in range(4)]
        self._check(self.L.lpo_get_nlp_info(self.h, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def bounds(self):
        xl, xu = np.empty(self.n), np.empty(self.n)
        gl, gu = np.empty(self.m), np.empty(self.m)
        self._check(self.L.lpo_get_bounds_info(self.h, _d(xl), _d(xu), _d(gl), _d(gu)))
        return xl, xu, gl, gu

    def eval_f(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        f = C.c_double()
        self._check(self.L.lpo_eval_f(self.h, _d(x), C.byref(f)))
        return f.value

    def eval_grad_f(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        g = np.empty(self.n)
        self._check(self.L.lpo_eval_grad_f(self.h, _d(x), _d(g)))
        return g

    def eval_g(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        g = np.empty(self.m)
        self._check(self.L.lpo_eval_g(self.h, _d(x), _d(g)))
        return g

    def jac_structure(self):
        i, j = np.empty(self.nnz_jac, dtype=np.int32), np.empty(self.nnz_jac, dtype=np.int32)
        self._check(self.L.lpo_eval_jac_g(self.h, None, _i(i), _i(j), None))
        return i, j

    def eval_jac_g(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        v = np.empty(self.nnz_jac)
        self._check(self.L.lpo_eval_jac_g(self.h, _d(x), None, None, _d(v)))
        return v

    def h_structure(self):
        i, j = np.empty(self.nnz_h, dtype=np.int32), np.empty(self.nnz_h, dtype=np.int32)
        self._check(self.L.lpo_eval_h(self.h, None, C.c_double(0), None, _i(i), _i(j), None))
        return i, j

    def eval_h(self, x, sigma, lam):
        x = np.ascontiguousarray(x, dtype=np.float64)
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        v = np.empty(self.nnz_h)
        self._check(self.L.lpo_eval_h(self.h, _d(x), C.c_double(sigma), _d(lam), None, None, _d(v)))
        return v

    def probe_dependencies(self, xguess):
        x = np.ascontiguousarray(xguess, dtype=np.float64)
        tot = sum((len(p.statemin) + len(p.pathmin)) * (len(p.statemin) + len(p.controlmin)) for p in self.op.phases)
        dep = np.zeros(tot, dtype=np.int32)
        self._check(self.L.lpo_probe_dependencies(self.h, _d(x), _i(dep)))
        self.n, self.m, self.nnz_jac, self.nnz_h = self.nlp_info()
        return dep

    def tables(self, phase):
        N = self.op.phases[phase].GetTotalNodes()
        pts, w = np.empty(N), np.empty(N)
        nD, nDiag, nDoff = C.c_int(), C.c_int(), C.c_int()
        self._check(self.L.lpo_get_tables(self.h, C.c_int(phase), _d(pts), _d(w), C.byref(nD), C.byref(nDiag), C.byref(nDoff)))
        out = {"points": pts, "weights": w}
        for which, (name, cnt) in enumerate((("D", nD.value), ("Diag", nDiag.value), ("Doffdiag", nDoff.value))):
            r, c, v = np.empty(cnt, dtype=np.int32), np.empty(cnt, dtype=np.int32), np.empty(cnt)
            self._check(self.L.lpo_get_coo(self.h, C.c_int(phase), C.c_int(which), _i(r), _i(c), _d(v)))
            out[name] = (r, c, v)
        return out

    def eval_g_jac_batch(self, x, nthreads=1, want_g=True, want_jac=True):
        x = np.ascontiguousarray(x, dtype=np.float64)
        nb = x.size // self.n
        g = np.empty((nb, self.m)) if want_g else None
        v = np.empty((nb, self.nnz_jac)) if want_jac else None
        self._check(self.L.lpo_eval_g_jac_batch(self.h, C.c_int(nb), _d(x), _d(g) if want_g else None, _d(v) if want_jac else None, C.c_int(nthreads)))
        return g, v


REF_LIB_PATH = os.path.join(ROOT, "oracle", "_ref", "liblpopc_ref.so")
REFERENCE_SRC = "/root/reference/Lpopc/src"
_ref = None


def build_reference():
    """oracle/_ref/liblpopc_ref.so: the reference's own sources compiled against oracle/ref_shim/ (needs
    /root/reference; on the GPU box only a prebuilt library can be used).  Returns the path or None."""
    if os.path.isdir(REFERENCE_SRC):
        subprocess.run(["make", "-j8", "-f", os.path.join("oracle", "ref_build.mk")], cwd=ROOT, check=True, capture_output=True)
    return REF_LIB_PATH if os.path.exists(REF_LIB_PATH) else None


def ref_lib():
    global _ref
    if _ref is None:
        path = build_reference()
        if path is None:
            raise RuntimeError("reference library unavailable (no /root/reference and no prebuilt oracle/_ref)")
        _ref = C.CDLL(path)
        _ref.lpo_create.restype = C.c_void_p
        _ref.lpo_last_error.restype = C.c_char_p
    return _ref


class RefOracle(Oracle):
    """The reference's OWN transcription code (compiled from /root/reference against the Armadillo
    stand-in of oracle/ref_shim/) behind the same interface as the restatement."""

    def _load(self):
        return ref_lib()

    def __init__(self, op):
        super().__init__(op)
        for ip, p in enumerate(op.phases):  # the reference interpolates the user's guess itself
            if p.timeguess:
                t = np.ascontiguousarray(p.timeguess, dtype=np.float64)
                x = np.ascontiguousarray(p.stateguess, dtype=np.float64).reshape(-1)
                u = np.ascontiguousarray(p.controlguess, dtype=np.float64).reshape(-1) if p.controlguess else np.zeros(1)
                self._check(self.L.lpo_set_guess(self.h, C.c_int(ip), C.c_int(t.size), _d(t), _d(x), _d(u)))
        self.refresh()

    def guess(self):
        x = np.empty(self.n)
        self._check(self.L.lpo_get_guess(self.h, _d(x)))
        return x

    def mesh_error(self, x):
        """SolutionErrorChecker::CheckSolutionDiffError of the reference: relative error matrix per phase."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        P = len(self.op.phases)
        rows = np.zeros(P, dtype=np.int32)
        ns = [len(p.statemin) for p in self.op.phases]
        tot = sum((sum(int(v) + 1 for v in p.nodesperinterval) + 1) * s for p, s in zip(self.op.phases, ns))
        rel = np.empty(tot)
        self._check(self.L.lpo_mesh_error(self.h, _d(x), _d(rel), _i(rows)))
        out, k = [], 0
        for ip in range(P):
            out.append(rel[k:k + rows[ip] * ns[ip]].reshape(ns[ip], rows[ip]).T.copy())
            k += rows[ip] * ns[ip]
        return out

    def nlp2op(self, x, lam):
        """Nlp2OpConverter::Nlp2OpControl of the reference: ([dict(time, state, control, costate, pathmult,
        hamiltonian, mayer, lagrange) per phase], total cost)."""
        from lpopc_b200.nlp import unpack_nlp2op
        x, lam = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(lam, dtype=np.float64)
        shapes = [(int(sum(p.nodesperinterval)) + 1, len(p.statemin), len(p.controlmin), len(p.pathmin)) for p in self.op.phases]
        out = np.empty(sum(M * (2 + 2 * ns + nc + npth) + 2 for M, ns, nc, npth in shapes))
        tot = C.c_double()
        self._check(self.L.lpo_nlp2op(self.h, _d(x), _d(lam), _d(out), C.byref(tot)))
        return unpack_nlp2op(out, shapes), tot.value

    def refine_ph(self, x, tol=1e-6, nmax=16, nmin=4):
        """PhMeshRefineAlg::RefineMesh of the reference: (no_more_refine, [(meshpoints, nodes) per phase])."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        P = len(self.op.phases)
        worst = sum(sum(max(2, (int(n) + 64) // nmin + 2) for n in p.nodesperinterval) for p in self.op.phases) + P
        K = np.zeros(P, dtype=np.int32)
        mesh, nodes = np.empty(worst + P), np.zeros(worst, dtype=np.int32)
        done = C.c_int()
        self._check(self.L.lpo_refine_ph(self.h, _d(x), C.c_double(tol), C.c_int(nmax), C.c_int(nmin), C.byref(done), _i(K), _d(mesh), _i(nodes)))
        out, km, kn = [], 0, 0
        for ip in range(P):
            out.append((mesh[km:km + K[ip] + 1].copy(), nodes[kn:kn + K[ip]].copy()))
            km += K[ip] + 1
            kn += K[ip]
        return bool(done.value), out


def _ref_refine_hp_liu(self, x, tol=1e-6, nmax=16, ratio_r=1.2):
    """LiuHpMeshRefineAlg::RefineMesh of the reference (stateful: the algorithm object lives in the handle):
    (no_more_refine, [(meshpoints, nodes) per phase])."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    P = len(self.op.phases)
    worst = sum(len(p.nodesperinterval) for p in self.op.phases) * 64 + 4096
    K = np.zeros(P, dtype=np.int32)
    mesh, nodes = np.empty(worst + P), np.zeros(worst, dtype=np.int32)
    done = C.c_int()
    self._check(self.L.lpo_refine_hp_liu(self.h, _d(x), C.c_double(tol), C.c_int(nmax), C.c_double(ratio_r), C.byref(done), _i(K), _d(mesh), _i(nodes)))
    out, km, kn = [], 0, 0
    for ip in range(P):
        out.append((mesh[km:km + K[ip] + 1].copy(), nodes[kn:kn + K[ip]].copy()))
        km += K[ip] + 1
        kn += K[ip]
    return bool(done.value), out


RefOracle.refine_hp_liu = _ref_refine_hp_liu


def detmath(which, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    lib().lpo_detmath(C.c_int(which), C.c_int(x.size), _d(x), _d(y))
    return y
