/* Brachistochrone -- authored against the reference's problem-definition
 * interface (LpFunctionWrapper.h:50-69); not shipped by the reference.
 * ns=3 (x, y, v), nc=1 (theta); Mayer = tf; endpoints through state bounds. */
#ifndef LPB_PROBLEM_BRACHISTOCHRONE_H
#define LPB_PROBLEM_BRACHISTOCHRONE_H
#include "../lpb_functor.h"

struct LpbBrachistochrone {
    static constexpr int NS = 3, NC = 1, NPATH = 0, NE_MAX = 0, NL_MAX = 0;
    static constexpr bool HAS_ANALYTIC = false;
    static constexpr bool UNROLL_COLOURS = true; /* compile-time colour unrolling of the FD Jacobian kernel */
    /* variables read per dae row and by the Lagrange integrand, order [x0..x2, u, t] (lpb_functor.h) */
    static constexpr unsigned long long HESS_DEP[NS + NPATH + 1] = {lpb_vars({2, 3}), lpb_vars({2, 3}), lpb_vars({3}), lpb_vars({})};
    struct Consts { double g; };
    static const char* name() { return "brachistochrone"; }

    LPB_HD static void dae(const Consts& C, int, double, const double* x, const double* u, double* f, double*)
    {
        double s = lpb_det_sin(u[0]), c = lpb_det_cos(u[0]);
        f[0] = x[2] * s;
        f[1] = (-x[2]) * c;
        f[2] = C.g * c;
    }
    LPB_HD static double lagrange(const Consts&, int, double, const double*, const double*) { return 0.0; }
    LPB_HD static double mayer(const Consts&, int, double, const double*, double tf, const double*) { return tf; }
    LPB_HD static void event(const Consts&, int, double, const double*, double, const double*, double*) {}
    LPB_HD static void link(const Consts&, const double*, const double*, double*) {}
};
#endif
