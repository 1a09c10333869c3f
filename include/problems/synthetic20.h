/* Synthetic large-mesh stress dynamics -- BASELINE config 5 / SURVEY.md 8(d):
 * ns=20, nc=6, f_s = sum_j A_sj tanh(x_j) + sum_j B_sj u_j - 0.1 x_s^3 with
 * dense A (20x20) and B (20x6); L = (|x|^2+|u|^2)/2.  Authored against the
 * reference's problem-definition interface (LpFunctionWrapper.h:50-69). */
#ifndef LPB_PROBLEM_SYNTHETIC20_H
#define LPB_PROBLEM_SYNTHETIC20_H
#include "../lpb_functor.h"

struct LpbSynthetic20 {
    static constexpr int NS = 20, NC = 6, NPATH = 0, NE_MAX = 0, NL_MAX = 0;
    static constexpr bool HAS_ANALYTIC = false;
    static constexpr bool UNROLL_COLOURS = false; /* compile-time colour unrolling of the FD Jacobian kernel */
    static constexpr bool UNROLL_HESSIAN = false; /* 378 dense pair bodies: keep the run-time pair loops */
    struct Consts { double A[NS * NS]; double B[NS * NC]; }; /* row-major */
    static const char* name() { return "synthetic20"; }

    LPB_HD static void dae(const Consts& C, int, double, const double* x, const double* u, double* f, double*)
    {
        double th[NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) th[j] = lpb_det_tanh(x[j]);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < NS; ++j) acc = acc + C.A[s * NS + j] * th[j];
#pragma unroll
            for (int j = 0; j < NC; ++j) acc = acc + C.B[s * NC + j] * u[j];
            f[s] = acc - 0.1 * ((x[s] * x[s]) * x[s]);
        }
    }
    LPB_HD static double lagrange(const Consts&, int, double, const double* x, const double* u)
    {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < NS; ++j) acc = acc + x[j] * x[j];
#pragma unroll
        for (int j = 0; j < NC; ++j) acc = acc + u[j] * u[j];
        return 0.5 * acc;
    }
    LPB_HD static double mayer(const Consts&, int, double, const double*, double, const double*) { return 0.0; }
    LPB_HD static void event(const Consts&, int, double, const double*, double, const double*, double*) {}
    LPB_HD static void link(const Consts&, const double*, const double*, double*) {}
};
#endif
