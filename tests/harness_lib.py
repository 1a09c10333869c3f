"""ctypes binding of tests/host_harness.cpp (TEST INFRASTRUCTURE: product host logic on the CPU)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libhost_harness.so")
_dp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
_lib = None


def lib():
    global _lib
    if _lib is None:
        csrc = os.path.join(ROOT, "lpopc_b200", "csrc")
        srcs = [os.path.join(HERE, "host_harness.cpp"), os.path.join(csrc, "lpb_tables.cpp"), os.path.join(csrc, "lpb_refine_liu.cpp"),
                os.path.join(csrc, "lpb_structure.hpp"), os.path.join(csrc, "lpb_tables.hpp"), os.path.join(csrc, "lpb_refine_liu.hpp")]
        if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
            subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", LIB, srcs[0], srcs[1], srcs[2]],
                           check=True, capture_output=True)
        _lib = C.CDLL(LIB)
        _lib.lpbt_create.restype = C.c_void_p
        _lib.lpbt_liu_create.restype = C.c_void_p
    return _lib


class Harness:
    def __init__(self, op, dep=None):
        ph = op.phases
        ns, nc, npth = len(ph[0].statemin), len(ph[0].controlmin), len(ph[0].pathmin)
        K = np.array([len(p.nodesperinterval) for p in ph], dtype=np.int32)
        mesh = np.concatenate([np.asarray(p.meshpoints, dtype=np.float64) for p in ph])
        nodes = np.concatenate([np.asarray(p.nodesperinterval, dtype=np.int32) for p in ph]).astype(np.int32)
        ne = np.array([len(p.eventmin) for p in ph], dtype=np.int32)
        left = np.array([l.leftphase - 1 for l in op.links] or [0], dtype=np.int32)
        right = np.array([l.rightphase - 1 for l in op.links] or [0], dtype=np.int32)
        nl = np.array([len(l.linkmin) for l in op.links] or [0], dtype=np.int32)
        err = C.create_string_buffer(256)
        depp = None
        if dep is not None:
            dep = np.ascontiguousarray(dep, dtype=np.int32)
            depp = dep.ctypes.data_as(_ip)
        self.h = C.c_void_p(lib().lpbt_create(ns, nc, npth, len(ph), K.ctypes.data_as(_ip), mesh.ctypes.data_as(_dp),
                                              nodes.ctypes.data_as(_ip), ne.ctypes.data_as(_ip), len(op.links),
                                              left.ctypes.data_as(_ip), right.ctypes.data_as(_ip), nl.ctypes.data_as(_ip), depp, err, 256))
        if not self.h:
            raise RuntimeError(err.value.decode())
        n, m, a, b = C.c_int(), C.c_int(), C.c_longlong(), C.c_longlong()
        lib().lpbt_info(self.h, C.byref(n), C.byref(m), C.byref(a), C.byref(b))
        self.n, self.m, self.nnz_jac, self.nnz_h = n.value, m.value, a.value, b.value
        self.N = [int(sum(p.nodesperinterval)) for p in ph]

    def __del__(self):
        if getattr(self, "h", None):
            lib().lpbt_destroy(self.h)
            self.h = None

    def structure(self):
        jI, jJ = np.empty(self.nnz_jac, dtype=np.int32), np.empty(self.nnz_jac, dtype=np.int32)
        hI, hJ = np.empty(self.nnz_h, dtype=np.int32), np.empty(self.nnz_h, dtype=np.int32)
        lib().lpbt_structure(self.h, jI.ctypes.data_as(_ip), jJ.ctypes.data_as(_ip), hI.ctypes.data_as(_ip), hJ.ctypes.data_as(_ip))
        return jI, jJ, hI, hJ

    def tables(self, phase):
        N = self.N[phase]
        tau, w, dd = np.empty(N), np.empty(N), np.empty(N)
        nd = lib().lpbt_tables(self.h, phase, tau.ctypes.data_as(_dp), w.ctypes.data_as(_dp), dd.ctypes.data_as(_dp))
        a, b, v = np.empty(nd, dtype=np.int32), np.empty(nd, dtype=np.int32), np.empty(nd)
        lib().lpbt_doff(self.h, phase, a.ctypes.data_as(_ip), b.ctypes.data_as(_ip), v.ctypes.data_as(_dp))
        nf = lib().lpbt_dfull(self.h, phase, None, None, None)
        fa, fb, fv = np.empty(nf, dtype=np.int32), np.empty(nf, dtype=np.int32), np.empty(nf)
        lib().lpbt_dfull(self.h, phase, fa.ctypes.data_as(_ip), fb.ctypes.data_as(_ip), fv.ctypes.data_as(_dp))
        return {"points": tau, "weights": w, "ddiag": dd, "Doffdiag": (a, b, v), "D": (fa, fb, fv)}


class LiuHarness:
    """The product's hp-Liu decision logic (lpopc_b200/csrc/lpb_refine_liu.cpp) on the CPU, fed with given error estimates."""

    def __init__(self):
        self.h = C.c_void_p(lib().lpbt_liu_create())

    def __del__(self):
        if getattr(self, "h", None):
            lib().lpbt_liu_destroy(self.h)
            self.h = None

    def refine(self, ns, mesh, nodes, rel, state, tol, nmax, ratio_r):
        mesh = np.ascontiguousarray(mesh, dtype=np.float64)
        nodes = np.ascontiguousarray(nodes, dtype=np.int32)
        rel = np.asfortranarray(rel, dtype=np.float64)      # rows x ns, column-major
        state = np.asfortranarray(state, dtype=np.float64)  # (N + 1) x ns, column-major
        cap = 64 * nodes.size + 1024
        mo, no = np.empty(cap + 1), np.zeros(cap, dtype=np.int32)
        done, K = C.c_int(), C.c_int()
        rc = lib().lpbt_liu_refine(self.h, C.c_int(ns), C.c_int(nodes.size), mesh.ctypes.data_as(_dp), nodes.ctypes.data_as(_ip),
                                   rel.ctypes.data_as(_dp), C.c_int(rel.shape[0]), state.ctypes.data_as(_dp), C.c_double(tol), C.c_int(nmax),
                                   C.c_double(ratio_r), C.byref(done), C.byref(K), mo.ctypes.data_as(_dp), no.ctypes.data_as(_ip), C.c_int(cap))
        if rc != 0:
            raise RuntimeError("hp-Liu harness failed: %d" % rc)
        return bool(done.value), mo[:K.value + 1].copy(), no[:K.value].copy()
