/* Quadrotor MPC instance -- authored against the reference's problem-definition
 * interface (LpFunctionWrapper.h:50-69); not shipped by the reference.
 * BASELINE config 4: ns=12 (p(3), v(3), euler phi/theta/psi, body rates p/q/r),
 * nc=4 rotor thrusts, quadratic tracking Lagrange cost, no events (the MPC
 * initial state enters through the state0 bounds, which is what differs between
 * batch instances). */
#ifndef LPB_PROBLEM_QUADROTOR_H
#define LPB_PROBLEM_QUADROTOR_H
#include "../lpb_functor.h"

struct LpbQuadrotor {
    static constexpr int NS = 12, NC = 4, NPATH = 0, NE_MAX = 0, NL_MAX = 0;
    static constexpr bool HAS_ANALYTIC = false;
    static constexpr bool UNROLL_COLOURS = true; /* compile-time colour unrolling of the FD Jacobian kernel */
    /* variables read per dae row and by the Lagrange integrand, order [x0..x11, u0..u3, t] (lpb_functor.h) */
    static constexpr unsigned long long HESS_DEP[NS + NPATH + 1] = {
        lpb_vars({3}), lpb_vars({4}), lpb_vars({5}),
        lpb_vars({6, 7, 8, 12, 13, 14, 15}), lpb_vars({6, 7, 8, 12, 13, 14, 15}), lpb_vars({6, 7, 12, 13, 14, 15}),
        lpb_vars({6, 7, 9, 10, 11}), lpb_vars({6, 10, 11}), lpb_vars({6, 7, 10, 11}),
        lpb_vars({10, 11, 13, 15}), lpb_vars({9, 11, 12, 14}), lpb_vars({9, 10, 12, 13, 14, 15}),
        lpb_vars({0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15})};
    struct Consts {
        double mass, g, Ixx, Iyy, Izz, arm, kM;
        double qp, qv, qa, qw, ru;
        double pref[3];
    };
    static const char* name() { return "quadrotor"; }

    LPB_HD static void dae(const Consts& C, int, double, const double* x, const double* u, double* f, double*)
    {
        const double phi = x[6], th = x[7], psi = x[8];
        const double p = x[9], q = x[10], r = x[11];
        double sphi = lpb_det_sin(phi), cphi = lpb_det_cos(phi);
        double sth = lpb_det_sin(th), cth = lpb_det_cos(th);
        double spsi = lpb_det_sin(psi), cpsi = lpb_det_cos(psi);
        double T = ((u[0] + u[1]) + u[2]) + u[3];
        double a = T / C.mass;
        f[0] = x[3]; f[1] = x[4]; f[2] = x[5];
        f[3] = a * ((cphi * sth) * cpsi + sphi * spsi);
        f[4] = a * ((cphi * sth) * spsi - sphi * cpsi);
        f[5] = a * (cphi * cth) - C.g;
        double qr = q * sphi + r * cphi;
        f[6] = p + qr * (sth / cth);
        f[7] = q * cphi - r * sphi;
        f[8] = qr / cth;
        f[9] = (C.arm * (u[1] - u[3]) - (C.Izz - C.Iyy) * (q * r)) / C.Ixx;
        f[10] = (C.arm * (u[2] - u[0]) - (C.Ixx - C.Izz) * (p * r)) / C.Iyy;
        f[11] = (C.kM * (((u[0] - u[1]) + u[2]) - u[3]) - (C.Iyy - C.Ixx) * (p * q)) / C.Izz;
    }
    LPB_HD static double lagrange(const Consts& C, int, double, const double* x, const double* u)
    {
        double e0 = x[0] - C.pref[0], e1 = x[1] - C.pref[1], e2 = x[2] - C.pref[2];
        double hov = (C.mass * C.g) * 0.25;
        double d0 = u[0] - hov, d1 = u[1] - hov, d2 = u[2] - hov, d3 = u[3] - hov;
        double acc = C.qp * ((e0 * e0 + e1 * e1) + e2 * e2);
        acc = acc + C.qv * ((x[3] * x[3] + x[4] * x[4]) + x[5] * x[5]);
        acc = acc + C.qa * ((x[6] * x[6] + x[7] * x[7]) + x[8] * x[8]);
        acc = acc + C.qw * ((x[9] * x[9] + x[10] * x[10]) + x[11] * x[11]);
        acc = acc + C.ru * (((d0 * d0 + d1 * d1) + d2 * d2) + d3 * d3);
        return 0.5 * acc;
    }
    LPB_HD static double mayer(const Consts&, int, double, const double*, double, const double*) { return 0.0; }
    LPB_HD static void event(const Consts&, int, double, const double*, double, const double*, double*) {}
    LPB_HD static void link(const Consts&, const double*, const double*, double*) {}
};
#endif
