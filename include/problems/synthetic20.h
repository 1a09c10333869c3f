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
    /* Optional hook (lpb_functor.h "dae_with_pre"): the second-difference Hessian evaluates dae() at 378 points per node
     * that differ from the node's values in one or two variables; the expensive univariate part of this dynamics --
     * tanh(x_j) -- takes three values per state there (x_j, x_j + h_j, (x_j + h_j) + h_j), so the kernel computes those
     * once per node through pre_value and hands the right one per state to dae_with_pre: the same operations as dae()
     * on the same operands, minus 20 tanh per point. */
    static constexpr bool HAS_DAE_PRE = true;
    LPB_HD static double pre_value(const Consts&, int, int, double v) { return lpb_det_tanh(v); }
    LPB_HD static void dae_with_pre(const Consts& C, int, double, const double* x, const double* u, const double* pv, double* f, double*)
    {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < NS; ++j) acc = acc + C.A[s * NS + j] * pv[j];
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
    /* Optional hooks (lpb_functor.h "row-parallel sweep"): the same sweep split over the rows.  sweep_pre runs once
     * per (node, variable) and parks what every row needs of that variable -- tanh at the base and at the perturbed
     * value -- and sweep_row runs once per (node, row s): the row's 26 products A_sq tanh(x_q), B_sq u_q stay in
     * registers, a perturbed state j re-adds the terms from j onwards in dae()'s order, nothing is recomputed per
     * row.  Bit-identical to dae() column by column, like dae_sweep. */
    static constexpr bool HAS_ROW_SWEEP = true;
    static constexpr int ROW_SWEEP_PRE = 2;
#ifndef LPB_S20_ROW_LO
#define LPB_S20_ROW_LO 13
#endif
#ifdef LPB_S20_REGS
    static constexpr int ROW_SWEEP_REGS = LPB_S20_REGS;
#endif
    static constexpr int ROW_LO = LPB_S20_ROW_LO;     /* products q < ROW_LO live in the thread's shared-memory scratch */
    static constexpr int ROW_SWEEP_SCRATCH = ROW_LO;  /* (they are re-added by few colours), the rest in registers */
    LPB_HD static void sweep_pre(const Consts&, int, int cc, double v, double vp, double* pre, int stride)
    {
        if (cc < NS) {
            pre[0] = lpb_det_tanh(v);
            pre[stride] = lpb_det_tanh(vp);
        }
    }
    template <class ND, class K>
    LPB_HD static double sweep_row(const Consts& C, int, int s, const ND& nd, K& k)
    {
        double pr[NS + NC - ROW_LO]; /* pr[q - ROW_LO] = q-th product of the row sum, q >= ROW_LO */
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < NS; ++q) {
            const double prod = C.A[s * NS + q] * nd.pre(q, 0);
            if (q < ROW_LO) nd.scratch(q) = prod;
            else pr[q - ROW_LO] = prod;
            acc = acc + prod;
        }
#pragma unroll
        for (int q = 0; q < NC; ++q) {
            const double prod = C.B[s * NC + q] * nd.u(q);
            pr[NS + q - ROW_LO] = prod;
            acc = acc + prod;
        }
        const double xs = nd.x(s);
        const double cube = 0.1 * ((xs * xs) * xs);
        const double fs = acc - cube;
        const double xsp = nd.perturbed(s); /* the only colour whose cubic term differs is j == s */
        const double cubep = 0.1 * ((xsp * xsp) * xsp);
        double pre = 0.0;
#pragma unroll
        for (int j = 0; j < NS; ++j) { /* state colours */
            k.begin(j, 0.0);
            double a2 = pre + C.A[s * NS + j] * nd.pre(j, 1);
#pragma unroll
            for (int q = j + 1; q < NS + NC; ++q) a2 = a2 + (q < ROW_LO ? nd.scratch(q) : pr[q < ROW_LO ? 0 : q - ROW_LO]);
            k.state_row(j, s, a2 - (s == j ? cubep : cube), fs);
            pre = pre + (j < ROW_LO ? nd.scratch(j) : pr[j < ROW_LO ? 0 : j - ROW_LO]);
        }
#pragma unroll
        for (int j = 0; j < NC; ++j) { /* control colours: pre now holds the whole state sum */
            const double vp = k.begin(NS + j, 0.0);
            double a2 = pre;
#pragma unroll
            for (int q = 0; q < NC; ++q) a2 = a2 + (q == j ? C.B[s * NC + q] * vp : pr[NS + q - ROW_LO]);
            k.state_row(NS + j, s, a2 - cube, fs);
        }
        k.begin(NS + NC, 0.0); /* time colour: dae() does not read t */
        k.state_row(NS + NC, s, fs, fs);
        return fs;
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
