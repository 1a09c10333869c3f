/* lpb_functor.h -- problem-definition interface of lpopc-b200 (device functors).
 *
 * This header is the B200-native mirror of the reference's problem-definition
 * interface `Lpopc::FunctionWrapper` (Lpopc/src/Core/LpFunctionWrapper.h:50-69):
 *
 *   reference virtual (whole mesh, Armadillo)          functor member (ONE node / endpoint)
 *   ------------------------------------------------   ------------------------------------
 *   MayerCost(SolCost&, double&)                        mayer(c, phase, t0, x0, tf, xf)
 *   LagrangeCost(SolCost&, vec&)        [N values]      lagrange(c, phase, t, x, u)
 *   DaeFunction(SolDae&, mat&, mat&)    [N x ns|np]     dae(c, phase, t, x, u, f, path)
 *   EventFunction(SolEvent&, vec&)                      event(c, phase, t0, x0, tf, xf, e)
 *   LinkFunction(SolLink&, vec&)                        link(c, xf_left, x0_right, out)
 *   Deriv* (analytic, optional)                         ddae / dlagrange (HAS_ANALYTIC)
 *
 * The reference requires every user function to be pointwise in the node index
 * (row k of the output depends only on row k of the inputs; the finite-difference
 * scheme in LpFiniteDifferenceDerive.cpp:194-324 and the whole sparsity pattern
 * assume it).  Here that requirement is the interface: a functor sees one node.
 * `phase` is 1-based like `phase_num_` (LpNLPWrapper.cpp:107).
 *
 * One `link` for all linkage pairs, with the same number of links per pair (the C ABI rejects unequal counts): the
 * reference cannot tell its user function reliably WHICH pair it is asked about -- the constraint evaluation leaves
 * SolLink::ipair unset and passes 0-based phase numbers (LpNLPWrapper.cpp:195-205) while the Jacobian and Hessian
 * paths pass ipair + 1 and 1-based phase numbers (LpNLPWrapper.cpp:420-426, LpHessian.cpp:1046-1052) -- so only a
 * pair-independent link function is well defined there, which is what its own launch example uses.
 *
 * A functor set is a plain struct with compile-time sizes shared by all phases
 * (NS states, NC controls, NPATH path constraints; NE_MAX/NL_MAX upper bounds for
 * events per phase / links per pair) and a POD `Consts` block of doubles that the
 * host passes by value to every kernel (it lands in the constant bank).
 *
 * Numerics contract (see DESIGN.md "bit-exact user functions"): the same header
 * is compiled by g++ -ffp-contract=off (oracle) and nvcc --fmad=false (device);
 * functors must only use + - * / sqrt fabs and the lpb_det_* functions from
 * lpb_detmath.h, which makes f(x) bit-identical on host and device and so keeps
 * forward-difference quotients (cancellation-amplified by 1/h ~ 1e6) within the
 * 1e-12 parity bound.  No fast-math: NaN must propagate (dependency probe,
 * LpDerivDependciesChecker.cpp:61-94).
 *
 * Optional member `dae_sweep` (flagged by `static constexpr bool HAS_SWEEP = true`), the finite-difference
 * counterpart of the reference's optional user derivatives (LpFunctionWrapper.h Deriv*): dynamics with DENSE
 * dependencies can evaluate the base point and all ns+nc+1 single-variable perturbations of the reference's
 * column-by-column scheme (LpFiniteDifferenceDerive.cpp:245-259) in one pass that shares the unchanged work,
 *     template <class K> static void dae_sweep(c, phase, t, x, u, f, path, K& k);
 * filling f/path like dae() and, per colour cc in [x.., u.., t] order, calling
 *     vp = k.begin(cc, v)                  v = the variable's value; returns the perturbed value v + h
 *     k.state_row(cc, s, fp, f[s])         fp = f_s at the perturbed point
 *     k.path_row(cc, i, cp, path[i])
 * once per row.  The values passed must be exactly what dae() returns at the perturbed point (same operations,
 * same order); the kernels do the quotients and the scatter.  Without the hook the kernels call dae() per colour.
 *
 * Optional members `sweep_pre` / `sweep_row` (flagged by `static constexpr bool HAS_ROW_SWEEP = true`, with
 * `ROW_SWEEP_PRE` = doubles per variable): the same sweep split over the function rows, for dynamics whose rows read
 * most variables.  The kernel (k_cons_jac_rows) runs one warp per row over 32 nodes:
 *     static void sweep_pre(c, phase, cc, v, vp, double* pre, int stride)
 * once per (node, variable cc): v = the variable's value, vp = v + h; stores up to ROW_SWEEP_PRE values of that
 * variable at pre[0], pre[stride], ... (e.g. an expensive elementary function at v and at vp);
 *     template <class ND, class K> static double sweep_row(c, phase, s, const ND& nd, K& k)
 * once per (node, row s in [0, NS + NPATH)): nd.x(j), nd.u(j), nd.t the node's variables, nd.perturbed(cc) = vp of
 * variable cc, nd.pre(cc, i) the values sweep_pre stored; k as in dae_sweep (begin / state_row / path_row for row s
 * only).  Returns the row's base value f_s (path rows: the path value).  Same exactness contract as dae_sweep.
 *
 * Optional members `pre_value` / `dae_with_pre` (flagged by `static constexpr bool HAS_DAE_PRE = true`): for the
 * second-difference Hessian of dynamics that apply an expensive univariate function to each state,
 *     static double pre_value(c, phase, j, v)                       that function of state j at the value v
 *     static void dae_with_pre(c, phase, t, x, u, pv, f, path)      dae() with pv[j] in place of pre_value(.., j, x[j])
 * The kernel (k_hess_nodes) evaluates pre_value at the three values a state takes among the stencil points and passes
 * the matching one per state; dae_with_pre must perform dae()'s operations on them in dae()'s order (bit-identical
 * values: the parity tests compare against the oracle, which calls dae()).
 */
#ifndef LPB_FUNCTOR_H
#define LPB_FUNCTOR_H

#if defined(__CUDACC__)
#define LPB_HD __host__ __device__ __forceinline__
#else
#define LPB_HD inline
#endif

#include "lpb_detmath.h"

/* Bit mask of variable indices in the order [states, controls, time], for the optional member
 *     static constexpr unsigned long long HESS_DEP[NS + NPATH + 1];
 * one mask per dae row, per path row, and one for the Lagrange integrand: the variables the row reads.  It is the
 * compile-time counterpart of the reference's NaN dependency probe (LpDerivDependciesChecker.cpp:10-94) and lets
 * the Hessian kernel generate a stencil only for the (row, variable pair) combinations that can be non-zero
 * (lpb_hessian.cuh).  A superset is always safe; a missing bit is rejected by lpb_probe_dependencies, which
 * compares the table with the probe's result.  Functor sets without the table use the generic pair loops. */
#ifdef __cplusplus
#include <initializer_list>
constexpr unsigned long long lpb_vars(std::initializer_list<int> vars)
{
    unsigned long long m = 0;
    for (int v : vars) m |= 1ull << v;
    return m;
}
#endif

#endif /* LPB_FUNCTOR_H */
