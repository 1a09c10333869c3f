/* Orbit raising (maximum-radius low-thrust transfer) -- authored against the
 * reference's problem-definition interface (LpFunctionWrapper.h:50-69); the
 * reference ships no definition of it.  BASELINE config 2:
 * single phase, ns=5 (r, theta, vr, vt, m), nc=2 (u1,u2), np=1 (u1^2+u2^2 = 1),
 * ne=2 (vr(tf) = 0, sqrt(mu/r(tf)) - vt(tf) = 0), Mayer = -r(tf). */
#ifndef LPB_PROBLEM_ORBIT_RAISING_H
#define LPB_PROBLEM_ORBIT_RAISING_H
#include "../lpb_functor.h"

struct LpbOrbitRaising {
    static constexpr int NS = 5, NC = 2, NPATH = 1, NE_MAX = 2, NL_MAX = 0;
    static constexpr bool HAS_ANALYTIC = false;
    static constexpr bool UNROLL_COLOURS = true; /* compile-time colour unrolling of the FD Jacobian kernel */
    /* variables read per dae row, per path row and by the Lagrange integrand, order [r, theta, vr, vt, m, u0, u1, t] */
    static constexpr unsigned long long HESS_DEP[NS + NPATH + 1] = {
        lpb_vars({2}), lpb_vars({0, 3}), lpb_vars({0, 3, 4, 5}), lpb_vars({0, 2, 3, 4, 6}), lpb_vars({}),
        lpb_vars({5, 6}), lpb_vars({})};
    struct Consts { double T, mu, mdot; };
    static const char* name() { return "orbit_raising"; }

    LPB_HD static void dae(const Consts& C, int, double, const double* x, const double* u, double* f, double* path)
    {
        const double r = x[0], vr = x[2], vt = x[3], m = x[4];
        double a = C.T / m;
        f[0] = vr;
        f[1] = vt / r;
        f[2] = ((vt * vt) / r - C.mu / (r * r)) + a * u[0];
        f[3] = (-(vr * vt)) / r + a * u[1];
        f[4] = -C.mdot;
        path[0] = u[0] * u[0] + u[1] * u[1];
    }
    LPB_HD static double lagrange(const Consts&, int, double, const double*, const double*) { return 0.0; }
    LPB_HD static double mayer(const Consts&, int, double, const double*, double, const double* xf) { return -xf[0]; }
    LPB_HD static void event(const Consts& C, int, double, const double*, double, const double* xf, double* e)
    {
        e[0] = xf[2];
        e[1] = sqrt(C.mu / xf[0]) - xf[3];
    }
    LPB_HD static void link(const Consts&, const double*, const double*, double*) {}
};
#endif
