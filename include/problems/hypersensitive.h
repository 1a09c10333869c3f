/* Hypersensitive problem -- functor restatement of the reference example
 * Lpopc/example/hypersensitive/HyperSensitive.cpp:73-151
 * (1 phase, ns=1, nc=1; xdot = -x^3 + u; L = (x^2+u^2)/2; Mayer = 0; analytic
 * derivatives shipped with the example at :87-97, :112-122, :136-151). */
#ifndef LPB_PROBLEM_HYPERSENSITIVE_H
#define LPB_PROBLEM_HYPERSENSITIVE_H
#include "../lpb_functor.h"

struct LpbHypersensitive {
    static constexpr int NS = 1, NC = 1, NPATH = 0, NE_MAX = 0, NL_MAX = 0;
    static constexpr bool HAS_ANALYTIC = true;
    static constexpr bool UNROLL_COLOURS = true; /* compile-time colour unrolling of the FD Jacobian kernel */
    /* variables read per dae row and by the Lagrange integrand, order [x, u, t] (lpb_functor.h) */
    static constexpr unsigned long long HESS_DEP[NS + NPATH + 1] = {lpb_vars({0, 1}), lpb_vars({0, 1})};
    struct Consts { double unused; };
    static const char* name() { return "hypersensitive"; }

    LPB_HD static void dae(const Consts&, int, double, const double* x, const double* u, double* f, double*)
    {
        /* stateout = -x%x%x + u  (HyperSensitive.cpp:132) */
        f[0] = ((-x[0]) * x[0]) * x[0] + u[0];
    }
    LPB_HD static double lagrange(const Consts&, int, double, const double* x, const double* u)
    {
        /* 0.5*( x%x + u%u )  (HyperSensitive.cpp:108) */
        return 0.5 * (x[0] * x[0] + u[0] * u[0]);
    }
    LPB_HD static double mayer(const Consts&, int, double, const double*, double, const double*) { return 0.0; }
    LPB_HD static void event(const Consts&, int, double, const double*, double, const double*, double*) {}
    LPB_HD static void link(const Consts&, const double*, const double*, double*) {}

    /* analytic first derivatives; rows = [f..., path...], cols = [x..., u..., t]
     * (layout of deriv_state in LpFiniteDifferenceDerive.cpp:299-323) */
    LPB_HD static void ddae(const Consts&, int, double, const double* x, const double*, double* d)
    {
        d[0] = -3.0 * (x[0] * x[0]); /* df/dx  (HyperSensitive.cpp:143) */
        d[1] = 1.0;                  /* df/du */
        d[2] = 0.0;                  /* df/dt */
    }
    LPB_HD static void dlagrange(const Consts&, int, double, const double* x, const double* u, double* d)
    {
        d[0] = x[0]; d[1] = u[0]; d[2] = 0.0; /* HyperSensitive.cpp:118-121 */
    }
    LPB_HD static void dmayer(const Consts&, int, double, const double*, double, const double*, double* d)
    {
        d[0] = 0.0; d[1] = 0.0; d[2] = 0.0; d[3] = 0.0; /* [x0, t0, xf, tf] */
    }
};
#endif
