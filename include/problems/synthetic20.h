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
    /* no HESS_DEP table: every row reads every state (378 dense pair bodies): the Hessian keeps the run-time pair loops */
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
    /* Optional hook (lpb_functor.h "dae_sweep"): the base evaluation and every single-variable perturbation
     * the forward-difference Jacobian needs, in ONE pass that shares work between them.  Perturbing x_j changes
     * only tanh(x_j) and the j-th term onwards of each row sum, so the sum over the terms before j (pre[s],
     * accumulated in the same order as dae()) is reused and only 20 - j + 6 terms are re-added; a perturbed
     * control re-adds the 6 control terms; nothing depends on t.  Every value handed to the sink is computed
     * by the same IEEE operations in the same order as dae() at the perturbed point, i.e. bit-identical
     * (tests compare against the oracle, which calls dae() column by column like the reference). */
    static constexpr bool HAS_SWEEP = true;
    static constexpr int SWEEP_ROWS = 2;     /* rows in flight per colour (independent summation chains) */
    static constexpr int SWEEP_MIN_CTAS = 1; /* resident CTAs per SM the sweep kernel is compiled for */
    template <class K>
    LPB_HD static void dae_sweep(const Consts& C, int, double t, const double* x, const double* u, double* f, double*, K& k)
    {
        double th[NS], pre[NS], cube[NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) th[j] = lpb_det_tanh(x[j]);
#pragma unroll
        for (int j = 0; j < NS; ++j) cube[j] = 0.1 * ((x[j] * x[j]) * x[j]);
#pragma unroll 1
        for (int s = 0; s < NS; ++s) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < NS; ++j) acc = acc + C.A[s * NS + j] * th[j];
#pragma unroll
            for (int j = 0; j < NC; ++j) acc = acc + C.B[s * NC + j] * u[j];
            f[s] = acc - cube[s];
            pre[s] = 0.0;
        }
#pragma unroll
        for (int j = 0; j < NS; ++j) { /* state colours */
            const double vp = k.begin(j, x[j]);
            const double thp = lpb_det_tanh(vp);
            const double cubep = 0.1 * ((vp * vp) * vp);
#pragma unroll(SWEEP_ROWS)
            for (int s = 0; s < NS; ++s) {
                double acc = pre[s] + C.A[s * NS + j] * thp;
#pragma unroll
                for (int q = j + 1; q < NS; ++q) acc = acc + C.A[s * NS + q] * th[q];
#pragma unroll
                for (int q = 0; q < NC; ++q) acc = acc + C.B[s * NC + q] * u[q];
                k.state_row(j, s, acc - (s == j ? cubep : cube[s]), f[s]);
                pre[s] = pre[s] + C.A[s * NS + j] * th[j];
            }
        }
#pragma unroll
        for (int j = 0; j < NC; ++j) { /* control colours: pre[s] now holds the whole state sum */
            const double vp = k.begin(NS + j, u[j]);
#pragma unroll 1
            for (int s = 0; s < NS; ++s) {
                double acc = pre[s];
#pragma unroll
                for (int q = 0; q < NC; ++q) acc = acc + C.B[s * NC + q] * (q == j ? vp : u[q]);
                k.state_row(NS + j, s, acc - cube[s], f[s]);
            }
        }
        k.begin(NS + NC, t); /* time colour: dae() does not read t */
#pragma unroll 1
        for (int s = 0; s < NS; ++s) k.state_row(NS + NC, s, f[s], f[s]);
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
