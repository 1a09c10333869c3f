"""ctypes binding of the C ABI (include/lpopc_b200.h) and the TNLP-shaped host wrapper.

`TranscribedNLP` mirrors the callback names and argument meaning of the reference's
`Lpopc::LpopcIpopt` (an `Ipopt::TNLP`, Lpopc/src/Core/LpopcIpopt.h:18-102,
LpopcIpopt.cpp:11-218): get_nlp_info / get_bounds_info / eval_f / eval_grad_f / eval_g /
eval_jac_g / eval_h, with `values=None` meaning "structure" exactly like TNLP.  Every
evaluation runs the CUDA kernels of liblpopc_b200.so; there is no CPU fallback -- if the
library is missing or no GPU is usable, construction raises.
"""
import ctypes as C
import os

import numpy as np

from .problem import lpb_problem_desc

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblpopc_b200.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p

LPB_OK = 0
ERRORS = {-1: "LPB_ERR_INVALID", -2: "LPB_ERR_UNSUPPORTED", -3: "LPB_ERR_CUDA", -4: "LPB_ERR_STATE", -5: "LPB_ERR_UNKNOWN_FUNCTOR"}

# name -> (restype, argtypes); every symbol include/lpopc_b200.h declares
SIGNATURES = {
    "lpb_create": (C.c_int, [C.POINTER(lpb_problem_desc), C.POINTER(_vp)]),
    "lpb_destroy": (C.c_int, [_vp]),
    "lpb_last_error": (C.c_char_p, [_vp]),
    "lpb_set_mesh": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _ip]),
    "lpb_refresh": (C.c_int, [_vp]),
    "lpb_get_nlp_info": (C.c_int, [_vp, _ip, _ip, _ip, _ip]),
    "lpb_get_bounds_info": (C.c_int, [_vp, _dp, _dp, _dp, _dp]),
    "lpb_eval_f": (C.c_int, [_vp, _dp, _dp]),
    "lpb_eval_grad_f": (C.c_int, [_vp, _dp, _dp]),
    "lpb_eval_g": (C.c_int, [_vp, _dp, _dp]),
    "lpb_eval_jac_g": (C.c_int, [_vp, _dp, _ip, _ip, _dp]),
    "lpb_eval_h": (C.c_int, [_vp, _dp, C.c_double, _dp, _ip, _ip, _dp]),
    "lpb_eval_g_jac": (C.c_int, [_vp, _dp, _dp, _dp]),
    "lpb_get_lgr_tables": (C.c_int, [_vp, C.c_int, _dp, _dp]),
    "lpb_mesh_error": (C.c_int, [_vp, _dp, _ip, _dp, _dp]),
    "lpb_refine_mesh_ph": (C.c_int, [_vp, _dp, C.c_double, C.c_int, C.c_int, _ip, _ip, _dp, C.c_int, _ip, C.c_int]),
    "lpb_refine_mesh_hp_liu": (C.c_int, [_vp, _dp, C.c_double, C.c_int, C.c_double, _ip, _ip, _dp, C.c_int, _ip, C.c_int]),
    "lpb_refine_reset": (C.c_int, [_vp]),
    "lpb_probe_dependencies": (C.c_int, [_vp, _dp, _ip]),
    "lpb_eval_f_batch": (C.c_int, [_vp, C.c_int, _dp, _dp]),
    "lpb_eval_grad_f_batch": (C.c_int, [_vp, C.c_int, _dp, _dp]),
    "lpb_eval_g_jac_batch": (C.c_int, [_vp, C.c_int, _dp, _dp, _dp]),
    "lpb_eval_h_batch": (C.c_int, [_vp, C.c_int, _dp, _dp, _dp, _dp]),
    "lpb_set_stream": (C.c_int, [_vp, _vp]),
    "lpb_eval_f_dev": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "lpb_eval_grad_f_dev": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "lpb_eval_g_jac_dev": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp]),
    "lpb_eval_h_dev": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp]),
    "lpb_structure_dev": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "lpb_set_option_int": (C.c_int, [_vp, C.c_char_p, C.c_int]),
    "lpb_kernel_launch_count": (C.c_longlong, [_vp]),
    "lpb_blocktri_solve": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp), C.POINTER(_vp), C.c_longlong, C.c_longlong,
                                     _vp, _vp, _vp, _vp]),
    "lpb_blocktri_solve_masked": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp), C.POINTER(_vp), C.c_longlong, C.c_longlong,
                                            _vp, _vp, _vp, _vp, _vp]),
    "lpb_blocktri_factor": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lpb_batched_spmv": (C.c_int, [C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, C.c_longlong, _vp, _vp, _vp, _vp]),
    "lpb_kkt_factor": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lpb_nlp2op_length": (C.c_longlong, [_vp, C.POINTER(C.c_longlong)]),
    "lpb_nlp2op": (C.c_int, [_vp, _dp, _dp, _dp, _dp]),
    "lpb_nlp2op_dev": (C.c_int, [_vp, _vp, _vp, _vp]),
    "lpb_mesh_error_dev": (C.c_int, [_vp, _vp, _vp, _vp]),
    "lpb_get_stat": (C.c_int, [_vp, C.c_char_p, C.POINTER(C.c_longlong)]),
    "lpb_kernel_time": (C.c_int, [_vp, C.c_char_p, _dp, _ip]),
    "lpb_selftest_fd_division": (C.c_int, [C.c_longlong, C.c_ulonglong, C.POINTER(C.c_longlong)]),
    "lpb_num_functors": (C.c_int, []),
    "lpb_functor_name": (C.c_char_p, [C.c_int]),
}

_lib = None


def load_library():
    """Loads liblpopc_b200.so (built by __graft_entry__.build() / csrc/Makefile).  Raises if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("lpopc_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the hot path is CUDA-only, there is no CPU fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _i(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def selftest_fd_division(n, seed=1):
    """Mismatches between the kernels' shared-reciprocal FD quotient and IEEE division over n random pairs."""
    bad = C.c_longlong(-1)
    rc = load_library().lpb_selftest_fd_division(int(n), int(seed), C.byref(bad))
    if rc != LPB_OK:
        raise RuntimeError("lpb_selftest_fd_division failed: %d" % rc)
    return bad.value


class LpopcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (ERRORS.get(code, str(code)), msg))
        self.code = code


def unpack_nlp2op(out, shapes):
    """Flat lpb_nlp2op output -> per-phase dict; shapes = [(M, ns, nc, np)] with M = N + 1 rows (column-major)."""
    res, k = [], 0
    for M, ns, nc, npth in shapes:
        d = {}
        for key, cols in (("time", None), ("state", ns), ("control", nc), ("costate", ns), ("pathmult", npth), ("hamiltonian", None)):
            if cols is None:
                d[key] = out[k:k + M].copy()
                k += M
            else:
                d[key] = out[k:k + M * cols].reshape(cols, M).T.copy()
                k += M * cols
        d["mayer"], d["lagrange"] = float(out[k]), float(out[k + 1])
        k += 2
        res.append(d)
    return res


class TranscribedNLP:
    """The NLP that the interior-point solver sees, evaluated on the GPU."""

    def __init__(self, op):
        self.lib = load_library()
        self.op = op
        desc, self._keep = op.to_desc()
        h = _vp()
        rc = self.lib.lpb_create(C.byref(desc), C.byref(h))
        if rc != LPB_OK:
            raise LpopcError(rc, (self.lib.lpb_last_error(None) or b"").decode())
        self.h = h
        for ip, p in enumerate(op.phases):
            if p.meshpoints:
                self.set_mesh(ip, p.meshpoints, p.nodesperinterval)
        self.refresh()

    # -- plumbing --
    def _ck(self, rc):
        if rc != LPB_OK:
            raise LpopcError(rc, (self.lib.lpb_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.lpb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_mesh(self, phase, meshpoints, nodes):
        mp, nd = _f64(meshpoints), np.ascontiguousarray(nodes, dtype=np.int32)
        self._ck(self.lib.lpb_set_mesh(self.h, phase, len(nd), _d(mp), _i(nd)))
        self.op.phases[phase].set_mesh(mp, nd)  # the Python mirror sizes output arrays from it

    def refresh(self):
        self._ck(self.lib.lpb_refresh(self.h))
        self.n, self.m, self.nnz_jac, self.nnz_h = self.get_nlp_info()

    def set_option(self, name, value):
        self._ck(self.lib.lpb_set_option_int(self.h, name.encode(), int(value)))

    def set_stream(self, cuda_stream_ptr):
        self._ck(self.lib.lpb_set_stream(self.h, _vp(cuda_stream_ptr)))

    @property
    def kernel_launches(self):
        return int(self.lib.lpb_kernel_launch_count(self.h))

    def nlp2op(self, x, lam):
        """NLP solution + multipliers -> optimal-control solution on the GPU (Nlp2OpConverter::Nlp2OpControl,
        Nlp2OPConverter.cpp:13-196): ([dict(time, state, control, costate, pathmult, hamiltonian, mayer, lagrange)
        per phase], total cost)."""
        x, lam = _f64(x), _f64(lam)
        n = int(self.lib.lpb_nlp2op_length(self.h, None))
        if n < 0:
            self._ck(n)
        out = np.empty(n)
        tot = C.c_double()
        self._ck(self.lib.lpb_nlp2op(self.h, _d(x), _d(lam), _d(out), C.byref(tot)))
        shapes = [(int(sum(p.nodesperinterval)) + 1, len(p.statemin), len(p.controlmin), len(p.pathmin)) for p in self.op.phases]
        return unpack_nlp2op(out, shapes), tot.value

    def nlp2op_dev(self, d_x, d_lam, d_out):
        """lpb_nlp2op with device pointers (ints): the converted solution stays on the GPU (layout of lpb_nlp2op)."""
        self._ck(self.lib.lpb_nlp2op_dev(self.h, C.c_void_p(d_x), C.c_void_p(d_lam), C.c_void_p(d_out)))

    def nlp2op_length(self):
        off = (C.c_longlong * (len(self.op.phases) + 1))()
        return int(self.lib.lpb_nlp2op_length(self.h, off))

    def mesh_error_dev(self, d_x, d_rel=None, d_imax=None):
        """lpb_mesh_error with device pointers (ints; None = not wanted): relative error matrices and interval maxima."""
        self._ck(self.lib.lpb_mesh_error_dev(self.h, C.c_void_p(d_x), C.c_void_p(d_rel), C.c_void_p(d_imax)))

    def stat(self, name):
        """Counter of the handle (lpb_get_stat): sparse_calls, sparse_fixups, sparse_on_doubles, head_doubles."""
        v = C.c_longlong()
        self._ck(self.lib.lpb_get_stat(self.h, name.encode(), C.byref(v)))
        return int(v.value)

    def kernel_time(self, kernel):
        """(total device ms, launches) of the named node kernel since the last call (option time_kernels)."""
        ms, cnt = C.c_double(), C.c_int()
        self._ck(self.lib.lpb_kernel_time(self.h, kernel.encode(), C.byref(ms), C.byref(cnt)))
        return ms.value, cnt.value

    # -- TNLP-shaped interface (LpopcIpopt.cpp:11-218) --
    def get_nlp_info(self):
        v = [C.c_int() for _ in range(4)]
        self._ck(self.lib.lpb_get_nlp_info(self.h, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def get_bounds_info(self):
        xl, xu, gl, gu = np.empty(self.n), np.empty(self.n), np.empty(self.m), np.empty(self.m)
        self._ck(self.lib.lpb_get_bounds_info(self.h, _d(xl), _d(xu), _d(gl), _d(gu)))
        return xl, xu, gl, gu

    def eval_f(self, x):
        x = _f64(x)
        f = np.empty(1)
        self._ck(self.lib.lpb_eval_f(self.h, _d(x), _d(f)))
        return float(f[0])

    def eval_grad_f(self, x):
        x = _f64(x)
        g = np.empty(self.n)
        self._ck(self.lib.lpb_eval_grad_f(self.h, _d(x), _d(g)))
        return g

    def eval_g(self, x):
        x = _f64(x)
        g = np.empty(self.m)
        self._ck(self.lib.lpb_eval_g(self.h, _d(x), _d(g)))
        return g

    def eval_jac_g(self, x=None, values=True):
        """values=False -> (iRow, jCol) like TNLP's values == NULL call; else the value vector."""
        if not values:
            i, j = np.empty(self.nnz_jac, dtype=np.int32), np.empty(self.nnz_jac, dtype=np.int32)
            self._ck(self.lib.lpb_eval_jac_g(self.h, None, _i(i), _i(j), None))
            return i, j
        x = _f64(x)
        v = np.empty(self.nnz_jac)
        self._ck(self.lib.lpb_eval_jac_g(self.h, _d(x), None, None, _d(v)))
        return v

    def eval_h(self, x=None, obj_factor=1.0, lam=None, values=True):
        if not values:
            i, j = np.empty(self.nnz_h, dtype=np.int32), np.empty(self.nnz_h, dtype=np.int32)
            self._ck(self.lib.lpb_eval_h(self.h, None, 0.0, None, _i(i), _i(j), None))
            return i, j
        x, lam = _f64(x), _f64(lam)
        v = np.empty(self.nnz_h)
        self._ck(self.lib.lpb_eval_h(self.h, _d(x), float(obj_factor), _d(lam), None, None, _d(v)))
        return v

    def eval_g_jac(self, x):
        x = _f64(x)
        g, v = np.empty(self.m), np.empty(self.nnz_jac)
        self._ck(self.lib.lpb_eval_g_jac(self.h, _d(x), _d(g), _d(v)))
        return g, v

    def lgr_points(self):
        """Composite LGR nodes of every phase on [-1, 1) (PS[phase]->Points of the current mesh)."""
        out = []
        for ip, p in enumerate(self.op.phases):
            pts = np.empty(int(sum(p.nodesperinterval)))
            self._ck(self.lib.lpb_get_lgr_tables(self.h, ip, _d(pts), None))
            out.append(pts)
        return out

    def initial_guess(self):
        """NLP starting point: the user's guess interpolated onto the LGR nodes (LpGuessChecker.cpp:130-203)."""
        return self.op.guess(self.lgr_points())

    def mesh_error(self, x):
        """(relative error matrices per phase [rows x ns], per-interval maxima per phase) of the solution x
        (SolutionErrorChecker::CheckSolutionDiffError on the GPU)."""
        x = _f64(x)
        P = len(self.op.phases)
        rows = np.zeros(P, dtype=np.int32)
        self._ck(self.lib.lpb_mesh_error(self.h, _d(x), _i(rows), None, None))
        ns = [len(p.statemin) for p in self.op.phases]
        rel = np.empty(int(sum(r * s for r, s in zip(rows, ns))))
        imax = np.empty(int(sum(len(p.nodesperinterval) for p in self.op.phases)))
        self._ck(self.lib.lpb_mesh_error(self.h, _d(x), _i(rows), _d(rel), _d(imax)))
        out_r, out_i, kr, ki = [], [], 0, 0
        for ip, p in enumerate(self.op.phases):
            out_r.append(rel[kr:kr + rows[ip] * ns[ip]].reshape(ns[ip], rows[ip]).T.copy())
            out_i.append(imax[ki:ki + len(p.nodesperinterval)].copy())
            kr += rows[ip] * ns[ip]
            ki += len(p.nodesperinterval)
        return out_r, out_i

    def refine_mesh_ph(self, x, tol=1e-6, nmax=16, nmin=4):
        """ph refinement decision (PhMeshRefineAlg::RefineMesh): (no_more_refine, [(meshpoints, nodes) per phase]).
        Defaults = the reference's "desired-relative-error", "Nmax", "Nmin" (LpMeshRefiner.h:67-80)."""
        x = _f64(x)
        P = len(self.op.phases)
        K = np.zeros(P, dtype=np.int32)
        cap = sum(len(p.nodesperinterval) for p in self.op.phases) * 4 + P  # usually enough; the call reports what it needs
        done = C.c_int()
        for attempt in range(2):
            mesh, nodes = np.empty(cap + P), np.zeros(cap, dtype=np.int32)
            rc = self.lib.lpb_refine_mesh_ph(self.h, _d(x), float(tol), int(nmax), int(nmin), C.byref(done), _i(K), _d(mesh), int(mesh.size),
                                             _i(nodes), int(nodes.size))
            if rc == 0 or attempt == 1 or int(K.sum()) <= cap:
                self._ck(rc)
                break
            cap = int(K.sum())  # K_out holds the interval counts of the refined mesh: size exactly and call again
        out, km, kn = [], 0, 0
        for ip in range(P):
            out.append((mesh[km:km + K[ip] + 1].copy(), nodes[kn:kn + K[ip]].copy()))
            km += K[ip] + 1
            kn += K[ip]
        return bool(done.value), out

    def refine_mesh_hp_liu(self, x, tol=1e-6, nmax=16, ratio_r=1.2):
        """hp-Liu refinement decision (LiuHpMeshRefineAlg::RefineMesh): (no_more_refine, [(meshpoints, nodes) per phase]).
        Stateful (history in the handle, refine_reset() starts over); defaults = the reference's options
        "desired-relative-error", "Nmax", "R" (LpMeshRefiner.h:67-80)."""
        x = _f64(x)
        P = len(self.op.phases)
        K = np.zeros(P, dtype=np.int32)
        cap = sum(len(p.nodesperinterval) for p in self.op.phases) * 4 + P
        done = C.c_int()
        for attempt in range(2):  # a result that does not fit leaves the history untouched and reports the sizes in K
            mesh, nodes = np.empty(cap + P), np.zeros(cap, dtype=np.int32)
            rc = self.lib.lpb_refine_mesh_hp_liu(self.h, _d(x), float(tol), int(nmax), float(ratio_r), C.byref(done), _i(K), _d(mesh),
                                                 int(mesh.size), _i(nodes), int(nodes.size))
            if rc == 0 or attempt == 1 or int(K.sum()) <= cap:
                self._ck(rc)
                break
            cap = int(K.sum())
        out, km, kn = [], 0, 0
        for ip in range(P):
            out.append((mesh[km:km + K[ip] + 1].copy(), nodes[kn:kn + K[ip]].copy()))
            km += K[ip] + 1
            kn += K[ip]
        return bool(done.value), out

    def refine_reset(self):
        self._ck(self.lib.lpb_refine_reset(self.h))

    def probe_dependencies(self, x_guess):
        x = _f64(x_guess)
        tot = sum((len(p.statemin) + len(p.pathmin)) * (len(p.statemin) + len(p.controlmin)) for p in self.op.phases)
        dep = np.zeros(tot, dtype=np.int32)
        self._ck(self.lib.lpb_probe_dependencies(self.h, _d(x), _i(dep)))
        self.n, self.m, self.nnz_jac, self.nnz_h = self.get_nlp_info()
        return dep

    # -- batched independent instances (host arrays; x is [nbatch, n]) --
    def eval_f_batch(self, x):
        x = _f64(x)
        nb = x.size // self.n
        f = np.empty(nb)
        self._ck(self.lib.lpb_eval_f_batch(self.h, nb, _d(x), _d(f)))
        return f

    def eval_grad_f_batch(self, x):
        x = _f64(x)
        nb = x.size // self.n
        g = np.empty((nb, self.n))
        self._ck(self.lib.lpb_eval_grad_f_batch(self.h, nb, _d(x), _d(g)))
        return g

    def eval_g_jac_batch(self, x, g_out=None, v_out=None):
        x = _f64(x)
        nb = x.size // self.n
        g = g_out if g_out is not None else np.empty((nb, self.m))
        v = v_out if v_out is not None else np.empty((nb, self.nnz_jac))
        self._ck(self.lib.lpb_eval_g_jac_batch(self.h, nb, _d(x), _d(g), _d(v)))
        return g, v

    def eval_h_batch(self, x, obj_factor, lam):
        x, lam, sg = _f64(x), _f64(lam), _f64(obj_factor)
        nb = x.size // self.n
        v = np.empty((nb, self.nnz_h))
        self._ck(self.lib.lpb_eval_h_batch(self.h, nb, _d(x), _d(sg), _d(lam), _d(v)))
        return v

    # -- raw-pointer variants: host pinned or device pointers as integers --
    def eval_g_jac_batch_ptr(self, nbatch, x_ptr, g_ptr, v_ptr):
        self._ck(self.lib.lpb_eval_g_jac_batch(self.h, nbatch, C.cast(x_ptr, _dp), C.cast(g_ptr, _dp), C.cast(v_ptr, _dp)))

    def eval_g_jac_dev(self, nbatch, d_x, d_g, d_vals):
        self._ck(self.lib.lpb_eval_g_jac_dev(self.h, nbatch, _vp(d_x), _vp(d_g) if d_g else None, _vp(d_vals) if d_vals else None))

    def eval_f_dev(self, nbatch, d_x, d_f):
        self._ck(self.lib.lpb_eval_f_dev(self.h, nbatch, _vp(d_x), _vp(d_f)))

    def eval_grad_f_dev(self, nbatch, d_x, d_grad):
        self._ck(self.lib.lpb_eval_grad_f_dev(self.h, nbatch, _vp(d_x), _vp(d_grad)))

    def eval_h_dev(self, nbatch, d_x, d_sigma, d_lambda, d_vals):
        self._ck(self.lib.lpb_eval_h_dev(self.h, nbatch, _vp(d_x), _vp(d_sigma), _vp(d_lambda), _vp(d_vals)))

    def structure_dev(self):
        p = [_vp() for _ in range(4)]
        self._ck(self.lib.lpb_structure_dev(self.h, *[C.byref(q) for q in p]))
        return tuple(q.value for q in p)
