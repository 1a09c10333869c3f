/* Two-stage transfer with an impulsive stage change -- authored against the reference's problem-definition interface
 * (LpFunctionWrapper.h:50-69) to exercise what none of the reference's three examples does: USER-SUPPLIED first
 * derivatives of events and linkages (first-derive = analytic forwards DerivEvent / DerivLink / DerivMayer to the
 * user, LpAnalyticDerive.hpp:18-47).  2 phases, ns = 2 (position, velocity), nc = 1, one link pair with 2 links.
 *   dynamics   x' = v,  v' = u - k x - c v |v|                 (phase-dependent drag c)
 *   Lagrange   u^2 / 2,  Mayer (phase 2) w (xf_0 - 1)^2
 *   events     phase 1: [x0_0, x0_1]      phase 2: [xf_0^2 + xf_1^2]
 *   link       [x0R_0 - xfL_0,  x0R_1 - r xfL_1 - q xfL_0^2]   (restitution r, position-dependent loss q) */
#ifndef LPB_PROBLEM_TWO_STAGE_H
#define LPB_PROBLEM_TWO_STAGE_H
#include "../lpb_functor.h"

struct LpbTwoStage {
    static constexpr int NS = 2, NC = 1, NPATH = 0, NE_MAX = 2, NL_MAX = 2;
    static constexpr bool HAS_ANALYTIC = true;
    static constexpr bool UNROLL_COLOURS = true;
    static constexpr unsigned long long HESS_DEP[NS + NPATH + 1] = {lpb_vars({1}), lpb_vars({0, 1, 2}), lpb_vars({2})};
    struct Consts { double k, c1, c2, w, r, q; };
    static const char* name() { return "two_stage"; }

    LPB_HD static void dae(const Consts& C, int phase, double, const double* x, const double* u, double* f, double*)
    {
        const double c = (phase == 1) ? C.c1 : C.c2;
        f[0] = x[1];
        f[1] = (u[0] - C.k * x[0]) - c * (x[1] * fabs(x[1]));
    }
    LPB_HD static double lagrange(const Consts&, int, double, const double*, const double* u) { return 0.5 * (u[0] * u[0]); }
    LPB_HD static double mayer(const Consts& C, int phase, double, const double*, double, const double* xf)
    {
        return (phase == 2) ? C.w * ((xf[0] - 1.0) * (xf[0] - 1.0)) : 0.0;
    }
    LPB_HD static void event(const Consts&, int phase, double, const double* x0, double, const double* xf, double* e)
    {
        if (phase == 1) { e[0] = x0[0]; e[1] = x0[1]; }
        else e[0] = xf[0] * xf[0] + xf[1] * xf[1];
    }
    LPB_HD static void link(const Consts& C, const double* xf_left, const double* x0_right, double* out)
    {
        out[0] = x0_right[0] - xf_left[0];
        out[1] = (x0_right[1] - C.r * xf_left[1]) - C.q * (xf_left[0] * xf_left[0]);
    }

    /* user derivatives; rows = [f..., path...], cols = [x..., u..., t] */
    LPB_HD static void ddae(const Consts& C, int phase, double, const double* x, const double*, double* d)
    {
        const double c = (phase == 1) ? C.c1 : C.c2;
        d[0] = 0.0; d[1] = 1.0; d[2] = 0.0; d[3] = 0.0;
        d[4] = -C.k; d[5] = -(2.0 * c) * fabs(x[1]); d[6] = 1.0; d[7] = 0.0;
    }
    LPB_HD static void dlagrange(const Consts&, int, double, const double*, const double* u, double* d)
    {
        d[0] = 0.0; d[1] = 0.0; d[2] = u[0]; d[3] = 0.0;
    }
    /* [x0 (ns) | t0 | xf (ns) | tf] */
    LPB_HD static void dmayer(const Consts& C, int phase, double, const double*, double, const double* xf, double* d)
    {
        for (int i = 0; i < 2 * NS + 2; ++i) d[i] = 0.0;
        if (phase == 2) d[NS + 1] = (2.0 * C.w) * (xf[0] - 1.0);
    }
    /* row-major ne x (2 ns + 2), columns [x0 (ns) | t0 | xf (ns) | tf]  (layout of deriv_event, LpNLPWrapper.cpp:651-668) */
    LPB_HD static void devent(const Consts&, int phase, double, const double*, double, const double* xf, double* d)
    {
        for (int i = 0; i < NE_MAX * (2 * NS + 2); ++i) d[i] = 0.0;
        if (phase == 1) { d[0] = 1.0; d[(2 * NS + 2) + 1] = 1.0; }
        else { d[NS + 1] = 2.0 * xf[0]; d[NS + 2] = 2.0 * xf[1]; }
    }
    /* row-major nl x 2 ns, columns [xf_left (ns) | x0_right (ns)]  (layout of derive_link, LpNLPWrapper.cpp:441-519) */
    LPB_HD static void dlink(const Consts& C, const double* xf_left, const double*, double* d)
    {
        d[0] = -1.0; d[1] = 0.0; d[2] = 1.0; d[3] = 0.0;
        d[4] = -(2.0 * C.q) * xf_left[0]; d[5] = -C.r; d[6] = 0.0; d[7] = 1.0;
    }
};
#endif
