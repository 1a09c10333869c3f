/* Bryson-Denham problem -- functor restatement of the reference example
 * Lpopc/example/bryson-denham/BrysonDenham.cpp:100-154
 * (1 phase, ns=3, nc=1, ne=5; Mayer = x3(tf); events = [x1(0),x2(0),x3(0),x1(tf),x2(tf)]). */
#ifndef LPB_PROBLEM_BRYSON_DENHAM_H
#define LPB_PROBLEM_BRYSON_DENHAM_H
#include "../lpb_functor.h"

struct LpbBrysonDenham {
    static constexpr int NS = 3, NC = 1, NPATH = 0, NE_MAX = 5, NL_MAX = 0;
    static constexpr bool HAS_ANALYTIC = false;
    static constexpr bool UNROLL_COLOURS = true; /* compile-time colour unrolling of the FD Jacobian kernel */
    /* variables read per dae row and by the Lagrange integrand, order [x0..x2, u, t] (lpb_functor.h) */
    static constexpr unsigned long long HESS_DEP[NS + NPATH + 1] = {lpb_vars({1}), lpb_vars({3}), lpb_vars({3}), lpb_vars({})};
    struct Consts { double unused; };
    static const char* name() { return "bryson_denham"; }

    LPB_HD static void dae(const Consts&, int, double, const double* x, const double* u, double* f, double*)
    {
        f[0] = x[1];                 /* BrysonDenham.cpp:121 */
        f[1] = u[0];                 /* :122 */
        f[2] = 0.5 * (u[0] * u[0]);  /* :123 */
    }
    LPB_HD static double lagrange(const Consts&, int, double, const double*, const double*) { return 0.0; }
    LPB_HD static double mayer(const Consts&, int, double, const double*, double, const double* xf) { return xf[2]; }
    LPB_HD static void event(const Consts&, int, double, const double* x0, double, const double* xf, double* e)
    {
        e[0] = x0[0]; e[1] = x0[1]; e[2] = x0[2]; e[3] = xf[0]; e[4] = xf[1]; /* :142-153 */
    }
    LPB_HD static void link(const Consts&, const double*, const double*, double*) {}
};
#endif
