/* Multi-phase launch-vehicle ascent -- functor restatement of the reference
 * example Lpopc/example/launch/Launch.cpp (4 phases, ns=7, nc=3, np=1,
 * ne=(0,0,0,5), 3 link pairs x 7 links).  Dynamics :660-738, orbital-element
 * events :589-630 and :744-754, linkage :760-765, Mayer :632-643.
 * pow(rad,3) (:683) is written rad*rad*rad and exp/acos use lpb_det_* so host and
 * device agree bitwise (see lpb_detmath.h). */
#ifndef LPB_PROBLEM_LAUNCH_H
#define LPB_PROBLEM_LAUNCH_H
#include "../lpb_functor.h"

struct LpbLaunch {
    static constexpr int NS = 7, NC = 3, NPATH = 1, NE_MAX = 5, NL_MAX = 7;
    static constexpr bool HAS_ANALYTIC = false;
    static constexpr bool UNROLL_COLOURS = true; /* compile-time colour unrolling of the FD Jacobian kernel */
    /* variables read per dae row, per path row and by the Lagrange integrand, order [r(3), v(3), m, u(3), t] */
    static constexpr unsigned long long HESS_DEP[NS + NPATH + 1] = {
        lpb_vars({3}), lpb_vars({4}), lpb_vars({5}),
        lpb_vars({0, 1, 2, 3, 4, 5, 6, 7}), lpb_vars({0, 1, 2, 3, 4, 5, 6, 8}), lpb_vars({0, 1, 2, 3, 4, 5, 6, 9}),
        lpb_vars({}), lpb_vars({7, 8, 9}), lpb_vars({})};
    /* CONSTANTS of Launch.cpp:50-74,148-153 (omega = earthRotRate*scales.time) */
    struct Consts {
        double omega, mu, cd, sa, rho0, H, Re, g0;
        double thrust_srb, thrust_first, thrust_second;
        double ISP_srb, ISP_first, ISP_second;
    };
    static const char* name() { return "launch"; }

    LPB_HD static void dae(const Consts& C, int phase, double, const double* x, const double* u, double* f, double* path)
    {
        const double r0 = x[0], r1 = x[1], r2 = x[2];
        const double v0 = x[3], v1 = x[4], v2 = x[5];
        const double m = x[6];
        double rad = sqrt((r0 * r0 + r1 * r1) + r2 * r2);
        /* omegacrossr = r*trans(omega_matrix), omega_matrix = [0 -w 0; w 0 0; 0 0 0] */
        double ocr0 = r1 * (-C.omega);
        double ocr1 = r0 * C.omega;
        double vr0 = v0 - ocr0, vr1 = v1 - ocr1, vr2 = v2;
        double speedrel = sqrt((vr0 * vr0 + vr1 * vr1) + vr2 * vr2);
        double altitude = rad - C.Re;
        double rho = lpb_det_exp((-altitude) / C.H) * C.rho0;
        double bc = rho / (m * 2.0) * (C.sa * C.cd);
        double bcspeed = bc * speedrel;
        double dragk = bcspeed * (-1.0);
        double muor3 = C.mu / ((rad * rad) * rad);
        double T_tot, mdot;
        if (phase == 1) {
            double T_srb = 6.0 * C.thrust_srb, T_first = C.thrust_first;
            T_tot = T_srb + T_first;
            mdot = (0.0 - T_srb / (C.g0 * C.ISP_srb)) + (0.0 - T_first / (C.g0 * C.ISP_first));
        } else if (phase == 2) {
            double T_srb = 3.0 * C.thrust_srb, T_first = C.thrust_first;
            T_tot = T_srb + T_first;
            mdot = (0.0 - T_srb / (C.g0 * C.ISP_srb)) + (0.0 - T_first / (C.g0 * C.ISP_first));
        } else if (phase == 3) {
            T_tot = C.thrust_first;
            mdot = 0.0 - C.thrust_first / (C.g0 * C.ISP_first);
        } else {
            T_tot = C.thrust_second;
            mdot = 0.0 - C.thrust_second / (C.g0 * C.ISP_second);
        }
        path[0] = (u[0] * u[0] + u[1] * u[1]) + u[2] * u[2];
        double Toverm = T_tot / m;
        f[0] = v0; f[1] = v1; f[2] = v2;
        f[3] = (Toverm * u[0] + dragk * vr0) + (-muor3) * r0;
        f[4] = (Toverm * u[1] + dragk * vr1) + (-muor3) * r1;
        f[5] = (Toverm * u[2] + dragk * vr2) + (-muor3) * r2;
        f[6] = mdot;
    }
    LPB_HD static double lagrange(const Consts&, int, double, const double*, const double*) { return 0.0; }
    LPB_HD static double mayer(const Consts&, int phase, double, const double*, double, const double* xf)
    {
        return (phase == 4) ? -xf[6] : 0.0; /* Launch.cpp:637-643 */
    }
    /* Launchrv2oe, Launch.cpp:589-630: first five orbital elements of (r,v)(tf) */
    LPB_HD static void event(const Consts& C, int phase, double, const double*, double, const double* xf, double* e)
    {
        if (phase != 4) return;
        const double PI = 3.14159265358979311600e+00;
        const double rv0 = xf[0], rv1 = xf[1], rv2 = xf[2];
        const double vv0 = xf[3], vv1 = xf[4], vv2 = xf[5];
        double hv0 = rv1 * vv2 - rv2 * vv1;
        double hv1 = rv2 * vv0 - rv0 * vv2;
        double hv2 = rv0 * vv1 - rv1 * vv0;
        /* nv = cross(K,hv), K = (0,0,1) */
        double nv0 = -hv1, nv1 = hv0, nv2 = 0.0;
        double n = sqrt((nv0 * nv0 + nv1 * nv1) + nv2 * nv2);
        double h2 = (hv0 * hv0 + hv1 * hv1) + hv2 * hv2;
        double v2 = (vv0 * vv0 + vv1 * vv1) + vv2 * vv2;
        double r = sqrt((rv0 * rv0 + rv1 * rv1) + rv2 * rv2);
        double rdv = (rv0 * vv0 + rv1 * vv1) + rv2 * vv2;
        double k1 = v2 - C.mu / r;
        double inv = 1.0 / C.mu;
        double ev0 = (rv0 * k1 - vv0 * rdv) * inv;
        double ev1 = (rv1 * k1 - vv1 * rdv) * inv;
        double ev2 = (rv2 * k1 - vv2 * rdv) * inv;
        double p = h2 / C.mu;
        double ecc = sqrt((ev0 * ev0 + ev1 * ev1) + ev2 * ev2);
        double a = p / (1.0 - ecc * ecc);
        double inc = lpb_det_acos(hv2 / sqrt(h2));
        double Om1 = lpb_det_acos(nv0 / n);
        if (nv1 < 0.0 - 2.220446049250313e-16) Om1 = 2.0 * PI - Om1;
        double Om2 = lpb_det_acos(((nv0 * ev0 + nv1 * ev1) + nv2 * ev2) / n / ecc);
        if (ev2 < 0.0) Om2 = 2.0 * PI - Om2;
        e[0] = a; e[1] = ecc; e[2] = inc; e[3] = Om1; e[4] = Om2;
    }
    LPB_HD static void link(const Consts&, const double* xf_left, const double* x0_right, double* out)
    {
        for (int i = 0; i < NS; ++i) out[i] = x0_right[i] - xf_left[i]; /* Launch.cpp:760-765 */
    }
};
#endif
