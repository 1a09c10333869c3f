/* Cart-pole MPC instance -- authored against the reference's problem-definition
 * interface (LpFunctionWrapper.h:50-69); not shipped by the reference.
 * BASELINE config 4 (second functor): ns=4 (x, theta, xdot, thetadot), nc=1. */
#ifndef LPB_PROBLEM_CARTPOLE_H
#define LPB_PROBLEM_CARTPOLE_H
#include "../lpb_functor.h"

struct LpbCartpole {
    static constexpr int NS = 4, NC = 1, NPATH = 0, NE_MAX = 0, NL_MAX = 0;
    static constexpr bool HAS_ANALYTIC = false;
    static constexpr bool UNROLL_COLOURS = true; /* compile-time colour unrolling of the FD Jacobian kernel */
    /* variables read per dae row and by the Lagrange integrand, order [x0..x3, u, t] (lpb_functor.h) */
    static constexpr unsigned long long HESS_DEP[NS + NPATH + 1] = {
        lpb_vars({2}), lpb_vars({3}), lpb_vars({1, 3, 4}), lpb_vars({1, 3, 4}), lpb_vars({0, 1, 2, 3, 4})};
    struct Consts { double mc, mp, l, g, qx, qth, qv, ru; };
    static const char* name() { return "cartpole"; }

    LPB_HD static void dae(const Consts& C, int, double, const double* x, const double* u, double* f, double*)
    {
        double s = lpb_det_sin(x[1]), c = lpb_det_cos(x[1]);
        double w2 = x[3] * x[3];
        double den = C.mc + C.mp * (s * s);
        f[0] = x[2];
        f[1] = x[3];
        f[2] = (u[0] + (C.mp * s) * (C.l * w2 + C.g * c)) / den;
        f[3] = (((-u[0]) * c - ((C.mp * C.l) * w2) * (c * s)) - ((C.mc + C.mp) * C.g) * s) / (C.l * den);
    }
    LPB_HD static double lagrange(const Consts& C, int, double, const double* x, const double* u)
    {
        double acc = C.qx * (x[0] * x[0]);
        acc = acc + C.qth * (x[1] * x[1]);
        acc = acc + C.qv * (x[2] * x[2] + x[3] * x[3]);
        acc = acc + C.ru * (u[0] * u[0]);
        return 0.5 * acc;
    }
    LPB_HD static double mayer(const Consts&, int, double, const double*, double, const double*) { return 0.0; }
    LPB_HD static void event(const Consts&, int, double, const double*, double, const double*, double*) {}
    LPB_HD static void link(const Consts&, const double*, const double*, double*) {}
};
#endif
