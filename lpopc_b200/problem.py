"""Host-side mirror of the reference's problem-definition classes.

`Phase`, `Linkage` and `OptimalProblem` keep the method names and argument meaning of
Lpopc::Phase / Lpopc::Linkage / Lpopc::OptimalProblem (Lpopc/src/Core/LpOptimalProblem.hpp:30-326)
so that a problem set-up reads like the reference's example programs
(Lpopc/example/*/*.cpp).  The user functions are not Python callables: they are a
registered device functor set (include/problems/*.h), named by `functor`.

`to_desc()` lowers the description to the plain-C `lpb_problem_desc` of
include/lpopc_b200.h, which is what both the CUDA library and (in tests only) the CPU
oracle consume.
"""
import ctypes as C

import numpy as np

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class lpb_phase_desc(C.Structure):
    _fields_ = [
        ("nstates", C.c_int), ("ncontrols", C.c_int), ("nparameters", C.c_int), ("npaths", C.c_int), ("nevents", C.c_int),
        ("state_min0", c_double_p), ("state_min", c_double_p), ("state_minf", c_double_p),
        ("state_max0", c_double_p), ("state_max", c_double_p), ("state_maxf", c_double_p),
        ("control_min", c_double_p), ("control_max", c_double_p),
        ("path_min", c_double_p), ("path_max", c_double_p),
        ("event_min", c_double_p), ("event_max", c_double_p),
        ("t0_min", C.c_double), ("t0_max", C.c_double), ("tf_min", C.c_double), ("tf_max", C.c_double),
        ("has_duration", C.c_int), ("duration_min", C.c_double), ("duration_max", C.c_double),
    ]


class lpb_link_desc(C.Structure):
    _fields_ = [("left_phase", C.c_int), ("right_phase", C.c_int), ("nlinks", C.c_int),
                ("link_min", c_double_p), ("link_max", c_double_p)]


class lpb_problem_desc(C.Structure):
    _fields_ = [("functor", C.c_char_p), ("nphases", C.c_int), ("phases", C.POINTER(lpb_phase_desc)),
                ("nlinkpairs", C.c_int), ("links", C.POINTER(lpb_link_desc)),
                ("consts", c_double_p), ("nconsts", C.c_int), ("fd_tol", C.c_double), ("first_derive", C.c_int)]


def _arr(v):
    a = np.ascontiguousarray(np.asarray(v, dtype=np.float64))
    return a, a.ctypes.data_as(c_double_p)


class Phase:
    """Mirror of Lpopc::Phase (LpOptimalProblem.hpp:30-240)."""

    def __init__(self, phase_index, statenum, controlnum, parameternum, pathnum, eventnum):
        self.phase_index = phase_index
        self.statenum, self.controlnum, self.parameternum = statenum, controlnum, parameternum
        self.pathnum, self.eventnum = pathnum, eventnum
        self.statemin, self.statemax = [], []
        self.controlmin, self.controlmax = [], []
        self.pathmin, self.pathmax = [], []
        self.eventmin, self.eventmax = [], []
        self.timemin = (0.0, 0.0)
        self.timemax = (0.0, 0.0)
        self.duration = None
        self.timeguess, self.stateguess, self.controlguess = [], [], []
        self.meshpoints, self.nodesperinterval = [], []

    def SetTimeMin(self, t0, tf): self.timemin = (float(t0), float(tf))
    def SetTimeMax(self, t0, tf): self.timemax = (float(t0), float(tf))
    def SetStateMin(self, state0, state, statef): self.statemin.append((float(state0), float(state), float(statef)))
    def SetStateMax(self, state0, state, statef): self.statemax.append((float(state0), float(state), float(statef)))
    def SetcontrolMin(self, v): self.controlmin.append(float(v))
    def SetcontrolMax(self, v): self.controlmax.append(float(v))
    def SetpathMin(self, v): self.pathmin.append(float(v))
    def SetpathMax(self, v): self.pathmax.append(float(v))
    def SeteventMin(self, v): self.eventmin.append(float(v))
    def SeteventMax(self, v): self.eventmax.append(float(v))
    def SetDuration(self, dmin, dmax): self.duration = (float(dmin), float(dmax))
    def SetTimeGuess(self, g): self.timeguess.append(float(g))

    def SetStateGuess(self, stateindex, g):  # 1-based like the reference
        while len(self.stateguess) < stateindex:
            self.stateguess.append([])
        self.stateguess[stateindex - 1].append(float(g))

    def SetControlGuess(self, controlindex, g):
        while len(self.controlguess) < controlindex:
            self.controlguess.append([])
        self.controlguess[controlindex - 1].append(float(g))

    def SetMeshPoints(self, p): self.meshpoints.append(float(p))
    def SetNodesPerInterval(self, n): self.nodesperinterval.append(int(n))

    def set_mesh(self, meshpoints, nodes):
        self.meshpoints = [float(v) for v in meshpoints]
        self.nodesperinterval = [int(v) for v in nodes]

    def GetTotalNodes(self): return int(sum(self.nodesperinterval))


class Linkage:
    """Mirror of Lpopc::Linkage (LpOptimalProblem.hpp:242-281); phases are 1-based."""

    def __init__(self, ipair, left, right):
        self.pairindex, self.leftphase, self.rightphase = ipair, left, right
        self.linkmin, self.linkmax = [], []

    def SetLinkMin(self, v): self.linkmin.append(float(v))
    def SetLinkMax(self, v): self.linkmax.append(float(v))


class OptimalProblem:
    """Mirror of Lpopc::OptimalProblem (LpOptimalProblem.hpp:283-323); `functor` names the
    device functor set that plays the role of the FunctionWrapper subclass."""

    def __init__(self, numphase, numlinkage, functor, consts=(), fd_tol=1e-6, first_derive="finite-difference"):
        self.numphase, self.numlink = numphase, numlinkage
        self.functor = functor
        self.consts = np.asarray(consts, dtype=np.float64).ravel()
        self.fd_tol = fd_tol
        self.first_derive = first_derive
        self.phases, self.links = [], []

    def AddPhase(self, p): self.phases.append(p)
    def AddLinkage(self, l): self.links.append(l)
    def GetPhase(self, i): return self.phases[i]
    def GetPhaseNum(self): return len(self.phases)
    def GetLinkageNum(self): return len(self.links)

    def to_desc(self):
        """Returns (lpb_problem_desc, keepalive list)."""
        keep = []
        ph_arr = (lpb_phase_desc * len(self.phases))()
        for i, p in enumerate(self.phases):
            d = ph_arr[i]
            d.nstates, d.ncontrols, d.nparameters = len(p.statemin), len(p.controlmin), p.parameternum
            d.npaths, d.nevents = len(p.pathmin), len(p.eventmin)
            if len(p.statemin) != len(p.statemax) or len(p.controlmin) != len(p.controlmax) \
                    or len(p.pathmin) != len(p.pathmax) or len(p.eventmin) != len(p.eventmax):
                raise ValueError("upper & lower bound MUST be same size in phase %d" % (i + 1))
            smin = np.array(p.statemin, dtype=np.float64).reshape(-1, 3)
            smax = np.array(p.statemax, dtype=np.float64).reshape(-1, 3)
            for name, col, src in (("state_min0", 0, smin), ("state_min", 1, smin), ("state_minf", 2, smin),
                                   ("state_max0", 0, smax), ("state_max", 1, smax), ("state_maxf", 2, smax)):
                a, ptr = _arr(src[:, col] if src.size else [])
                keep.append(a)
                setattr(d, name, ptr)
            for name, src in (("control_min", p.controlmin), ("control_max", p.controlmax), ("path_min", p.pathmin),
                              ("path_max", p.pathmax), ("event_min", p.eventmin), ("event_max", p.eventmax)):
                a, ptr = _arr(src)
                keep.append(a)
                setattr(d, name, ptr)
            d.t0_min, d.tf_min = p.timemin
            d.t0_max, d.tf_max = p.timemax
            d.has_duration = 1 if p.duration is not None else 0
            if p.duration is not None:
                d.duration_min, d.duration_max = p.duration
        lk_arr = (lpb_link_desc * max(1, len(self.links)))()
        for i, l in enumerate(self.links):
            d = lk_arr[i]
            d.left_phase, d.right_phase, d.nlinks = l.leftphase, l.rightphase, len(l.linkmin)
            a, ptr = _arr(l.linkmin); keep.append(a); d.link_min = ptr
            a, ptr = _arr(l.linkmax); keep.append(a); d.link_max = ptr
        desc = lpb_problem_desc()
        desc.functor = self.functor.encode()
        desc.nphases, desc.phases = len(self.phases), ph_arr
        desc.nlinkpairs, desc.links = len(self.links), lk_arr
        a, ptr = _arr(self.consts); keep.append(a)
        desc.consts, desc.nconsts = ptr, int(a.size)
        desc.fd_tol = float(self.fd_tol)
        desc.first_derive = 1 if self.first_derive == "analytic" else 0
        keep += [ph_arr, lk_arr]
        return desc, keep

    # ---- layout helpers ---------------------------------
    def nvars(self):
        return sum(len(p.statemin) * (p.GetTotalNodes() + 1) + len(p.controlmin) * p.GetTotalNodes() + 2 for p in self.phases)

    def guess(self, lgr_points):
        """NLP guess vector for two-point guesses (the natural cubic spline of
        LpGuessChecker.cpp:130-190 through two points is the straight line).
        lgr_points[p] = LGR points of phase p on [-1,1)."""
        out = []
        for ip, p in enumerate(self.phases):
            tau = np.concatenate([np.asarray(lgr_points[ip]), [1.0]])
            t0g, tfg = p.timeguess[0], p.timeguess[-1]
            taug = np.array([2 * (t - t0g) / (tfg - t0g) - 1 for t in p.timeguess])
            for g in p.stateguess:
                out.append(np.interp(tau, taug, g))
            for g in p.controlguess:
                out.append(np.interp(tau[:-1], taug, g))
            out.append(np.array([t0g, tfg]))
        return np.concatenate(out)
