// lpb_hessian.cuh -- Lagrangian Hessian values and the NaN dependency probe.
//
//   k_hess_nodes     per-node second differences of dae/path/Lagrange contracted with
//                    lambda/sigma in registers; one N-long block per ordered variable pair
//                    (replaces LpHessianCalculator::CalculatePhaseHessian :1192-1411,:2008-2101
//                     and the I-part of GetPhaseHessian :115-218,:404-535)
//   k_hess_endpoint  Mayer/event second differences (E-part, :282-347,:1502-1930), linkage
//                    part (:1020-1189,:2163-2367) and the three t0/tf scalars of the I-part
//   k_probe          sparse-NaN dependency probe (LpDerivDependciesChecker.cpp:10-94)
//
// Colour group = ordered pair (a, b), a >= b in the variable order [states, controls, time]:
// exactly the pairs the reference scatters (lower triangle).  Stencil per node
// (LpHessian.cpp:1269-1282): f_a = F(v_a+h_a), f_b = F(v_b+h_b), f_ab = F(v_a+h_a, then
// v_b += h_b), value = (((f_ab - f_a) - f_b) + f)/(h_a*h_b), h = tol*(1+|v|) per thread.
// The pair loop runs a-outer / b-inner so f_a stays in registers and f_b is recomputed per
// pair (the reference recomputes it too); gridDim.y splits the a-range.
#pragma once
#include "lpb_kernels.cuh"
#include "lpb_mesherr.cuh"
#include "lpb_convert.cuh"

namespace lpb {

// ------------------------------------------------------------------------------------------
// per-thread node state shared by the two node kernels
// ------------------------------------------------------------------------------------------
template <class P>
struct HessNode {
    typedef Dim<P> D;
    double xs[D::NSa], us[D::NCa], lm[D::NSa], mu[D::NPa], f[D::NSa], c[D::NPa];
    double t, L, tau, w, t0, tf, sg, tol, talpha, tbeta;
    int p, k, N;
    double* __restrict__ vI; // I-part values of this (instance, phase)
};

template <class P>
__device__ __forceinline__ void hess_load_node(const ProblemDev& pd, const typename P::Consts& C, long long gid,
                                               const double* __restrict__ x, const double* __restrict__ sigma,
                                               const double* __restrict__ lambda, double* __restrict__ vals, HessNode<P>& nd)
{
    typedef Dim<P> D;
    const int b_inst = (int)(gid / pd.total_nodes);
    const int gnode = (int)(gid - (long long)b_inst * pd.total_nodes);
    nd.p = find_phase(pd, gnode);
    const PhaseDev& ph = pd.ph[nd.p];
    const int k = gnode - ph.node0, N = ph.N;
    nd.k = k;
    nd.N = N;
    const double* __restrict__ xb = x + (size_t)b_inst * pd.n + ph.var0;
    const double* __restrict__ lam = lambda + (size_t)b_inst * pd.m + ph.con0;
    nd.vI = vals + (size_t)b_inst * pd.nnz_h + ph.hI0;
    nd.sg = sigma[b_inst];
    nd.tol = pd.tol;
#pragma unroll
    for (int j = 0; j < D::NS; ++j) nd.xs[j] = xb[(size_t)j * (N + 1) + k];
#pragma unroll
    for (int j = 0; j < D::NC; ++j) nd.us[j] = xb[(size_t)D::NS * (N + 1) + (size_t)j * N + k];
    nd.t0 = xb[(size_t)D::NS * (N + 1) + (size_t)D::NC * N];
    nd.tf = xb[(size_t)D::NS * (N + 1) + (size_t)D::NC * N + 1];
    const double tspan = nd.tf - nd.t0;
    nd.tau = ph.tau[k];
    nd.w = ph.w[k];
    nd.t = (nd.tau + 1) * (tspan / 2.0) + nd.t0;
#pragma unroll
    for (int s = 0; s < D::NS; ++s) nd.lm[s] = lam[(size_t)s * N + k];
#pragma unroll
    for (int s = 0; s < D::NP; ++s) nd.mu[s] = lam[(size_t)(D::NS + s) * N + k];
    P::dae(C, nd.p + 1, nd.t, nd.xs, nd.us, nd.f, nd.c);
    nd.L = P::lagrange(C, nd.p + 1, nd.t, nd.xs, nd.us);
    nd.talpha = (1 - nd.tau) / 2.0;
    nd.tbeta = (1 + nd.tau) / 2.0;
}

// value of variable v in the order [states, controls, time] (run-time index: select chain)
template <class P>
__device__ __forceinline__ double hess_var(const HessNode<P>& nd, int v)
{
    typedef Dim<P> D;
    double r = nd.t;
#pragma unroll
    for (int j = 0; j < D::NS; ++j) r = (v == j) ? nd.xs[j] : r;
#pragma unroll
    for (int j = 0; j < D::NC; ++j) r = (v == D::NS + j) ? nd.us[j] : r;
    return r;
}

// first-derivative term of the time rows, user-supplied derivatives (derive_fun_ = LpAnalyticDerive, :161,:175)
template <class P>
__device__ __forceinline__ void hess_time_analytic(const typename P::Consts& C, const HessNode<P>& nd, int b, double& acc, double& dLb)
{
    typedef Dim<P> D;
    constexpr int T = D::NS + D::NC;
    if constexpr (P::HAS_ANALYTIC) {
        double dd[D::NROW * D::NCOL], dl[D::NCOL];
        P::ddae(C, nd.p + 1, nd.t, nd.xs, nd.us, dd);
        P::dlagrange(C, nd.p + 1, nd.t, nd.xs, nd.us, dl);
#pragma unroll
        for (int s = 0; s < D::NS; ++s) {
            double dsb = dd[s * D::NCOL + T];
#pragma unroll
            for (int j = 0; j < T; ++j) dsb = (b == j) ? dd[s * D::NCOL + j] : dsb;
            acc += nd.lm[s] * dsb;
        }
        dLb = dl[T];
#pragma unroll
        for (int j = 0; j < T; ++j) dLb = (b == j) ? dl[j] : dLb;
    }
}

// stores of one pair: block (a, b) of the xx/ux/uu part, or the t0/tf rows when a is the time variable
// (:186-189,:202-205,:216-218; layout :468-518)
template <class P>
__device__ __forceinline__ void hess_store_pair(const PhaseDev& ph, const HessNode<P>& nd, bool a_is_time, bool b_is_time, int b, int blk,
                                                double core, double A, double* __restrict__ scr, long long gid, long long tot)
{
    typedef Dim<P> D;
    constexpr int NV = D::NS + D::NC;
    // 32-bit value offsets: an instance fits IPOPT's 32-bit Index (checked in build_layout)
    const unsigned uN = (unsigned)nd.N;
    double* __restrict__ vk = nd.vI + nd.k;
    if (!a_is_time) {
        st_stream(vk + (unsigned)blk * uN, core);
    } else if (!b_is_time) {
        st_stream(vk + (unsigned)(ph.nblkH + b) * uN, 0.5 * A + nd.talpha * core);
        st_stream(vk + ((unsigned)(ph.nblkH + NV + b) * uN + 1u), -0.5 * A + nd.tbeta * core);
    } else {
        // per-node terms of the three time-time scalars
        scr[gid] = nd.talpha * (A + nd.talpha * core);
        scr[tot + gid] = nd.tbeta * (-A + nd.tbeta * core);
        scr[2 * tot + gid] = (nd.tbeta - nd.talpha) * A;
        scr[3 * tot + gid] = nd.talpha * (nd.tbeta * core);
    }
}

// ------------------------------------------------------------------------------------------
// generic pair body: run-time variable indices, every function row, IEEE quotients.  This is the
// arithmetic definition of the node part (the tiled kernel below must match it bit for bit).
// ------------------------------------------------------------------------------------------
template <class P, class = void>
struct has_dae_pre : std::false_type {};
template <class P>
struct has_dae_pre<P, std::enable_if_t<P::HAS_DAE_PRE>> : std::true_type {};

// dae() at a stencil point whose variables a and b (b = -1: none) are perturbed; with the functor set's pre-value hook
// and the node's table pv[times perturbed][state] the expensive univariate part is looked up instead of recomputed
template <class P>
__device__ __forceinline__ void hess_dae(const typename P::Consts& C, int phase1, double t, const double* xq, const double* uq,
                                         const double (*pv)[Dim<P>::NSa], int a, int b, double* f, double* c)
{
    if constexpr (has_dae_pre<P>::value) {
        if (pv != nullptr) {
            double pq[Dim<P>::NSa];
#pragma unroll
            for (int j = 0; j < Dim<P>::NS; ++j) {
                const int cnt = (a == j ? 1 : 0) + (b == j ? 1 : 0);
                pq[j] = cnt == 2 ? pv[2][j] : (cnt == 1 ? pv[1][j] : pv[0][j]);
            }
            P::dae_with_pre(C, phase1, t, xq, uq, pq, f, c);
            return;
        }
    }
    P::dae(C, phase1, t, xq, uq, f, c);
}

template <class P>
__device__ __forceinline__ void hess_pair_generic(const ProblemDev& pd, const typename P::Consts& C, const HessNode<P>& nd, int a, double ha,
                                                  const double* xa, const double* ua, double ta, const double* fa, const double* ca,
                                                  double La, int b, double* __restrict__ scr, long long gid, long long tot,
                                                  const double* fb_cached = nullptr, const double* cb_cached = nullptr, double Lb_cached = 0.0,
                                                  const double (*pv)[Dim<P>::NSa] = nullptr)
{
    typedef Dim<P> D;
    constexpr int T = D::NS + D::NC; // index of the time variable
    constexpr int NV = D::NS + D::NC;
    const PhaseDev& ph = pd.ph[nd.p];
    const int p = nd.p;
    int blk = 0;
    if (a < T) {
        blk = ph.hblk[a * NV + b];
        if (blk < 0) return; // pair absent from the pattern (dependency mask)
    }
    const double hb = nd.tol * (1 + fabs(hess_var(nd, b)));
    double xq[D::NSa], uq[D::NCa], fb[D::NSa], cb[D::NPa], fab[D::NSa], cab[D::NPa];
    double Lb;
    if (a == T && b == T) { // stateoutj = stateouti (:1403,:1407,:2097)
#pragma unroll
        for (int s = 0; s < D::NS; ++s) fb[s] = fa[s];
#pragma unroll
        for (int s = 0; s < D::NP; ++s) cb[s] = ca[s];
        Lb = La;
    } else if (fb_cached != nullptr) { // F(v_b + h_b) was evaluated when b was the outer variable: same point, same values
#pragma unroll
        for (int s = 0; s < D::NS; ++s) fb[s] = fb_cached[s];
#pragma unroll
        for (int s = 0; s < D::NP; ++s) cb[s] = cb_cached[s];
        Lb = Lb_cached;
    } else {
#pragma unroll
        for (int j = 0; j < D::NS; ++j) xq[j] = (b == j) ? nd.xs[j] + hb : nd.xs[j];
#pragma unroll
        for (int j = 0; j < D::NC; ++j) uq[j] = (b == D::NS + j) ? nd.us[j] + hb : nd.us[j];
        const double tq = (b == T) ? nd.t + hb : nd.t;
        hess_dae<P>(C, p + 1, tq, xq, uq, pv, b, -1, fb, cb);
        Lb = P::lagrange(C, p + 1, tq, xq, uq);
    }
    // point (a then b): start from the a-perturbed point, add h_b to variable b
#pragma unroll
    for (int j = 0; j < D::NS; ++j) xq[j] = (b == j) ? xa[j] + hb : xa[j];
#pragma unroll
    for (int j = 0; j < D::NC; ++j) uq[j] = (b == D::NS + j) ? ua[j] + hb : ua[j];
    const double tq2 = (b == T) ? ta + hb : ta;
    hess_dae<P>(C, p + 1, tq2, xq, uq, pv, a, b, fab, cab);
    const double Lab = P::lagrange(C, p + 1, tq2, xq, uq);
    const double den = ha * hb;
    const FdDiv dvp(den); // one reciprocal per pair, IEEE-exact quotients (lpb_kernels.cuh)
    double sdae = 0.0, spath = 0.0;
#pragma unroll
    for (int s = 0; s < D::NS; ++s) sdae += nd.lm[s] * dvp.quot_num(fab[s] - fa[s] - fb[s] + nd.f[s]);
#pragma unroll
    for (int s = 0; s < D::NP; ++s) spath += nd.mu[s] * dvp.quot_num(cab[s] - ca[s] - cb[s] + nd.c[s]);
    const double hl = dvp.quot_num(Lab - La - Lb + nd.L);
    const double sL = nd.sg * nd.w * hl;
    const double core = (nd.tf - nd.t0) / 2.0 * (sL - sdae) + spath; // :123-127
    double A = 0.0;
    if (a == T) {
        // first-derivative term: sum(lambda % df/dv_b) - sigma*w*dL/dv_b  (:184-185,:200-201,:214-215)
        double acc = 0.0, dLb = 0.0;
        bool analytic = false;
        if constexpr (P::HAS_ANALYTIC) analytic = pd.analytic != 0;
        if (analytic) {
            hess_time_analytic<P>(C, nd, b, acc, dLb);
        } else {
            const FdDiv dvb(hb);
#pragma unroll
            for (int s = 0; s < D::NS; ++s) acc += nd.lm[s] * dvb.quot_num(fb[s] - nd.f[s]);
            dLb = dvb.quot_num(Lb - nd.L);
        }
        A = acc - nd.sg * nd.w * dLb;
    }
    hess_store_pair<P>(ph, nd, a == T, b == T, b, blk, core, A, scr, gid, tot);
}

// all pairs (a, b), a in [a0, a1), b in [b0, min(b1, a + 1)), with the generic body
// CACHE: keep F(v + h_a) of every outer variable of this call in thread-local memory and reuse it as F(v + h_b) of the
// later pairs (b <= a, so it is there whenever b >= a0): one dae() per pair instead of two for dense functor sets.
template <class P, bool CACHE = false>
__device__ __forceinline__ void hess_range_generic(const ProblemDev& pd, const typename P::Consts& C, const HessNode<P>& nd,
                                                   int a0, int a1, int b0, int b1, double* __restrict__ scr, long long gid, long long tot)
{
    typedef Dim<P> D;
    constexpr int T = D::NS + D::NC;
    constexpr int NCACHE = CACHE ? D::NCOL : 1;
    double fsv[NCACHE][D::NSa], csv[NCACHE][D::NPa], Lsv[NCACHE];
    constexpr bool PRE = CACHE && has_dae_pre<P>::value;
    double pvs[PRE ? 3 : 1][D::NSa];
    const double (*pv)[D::NSa] = nullptr;
    if constexpr (PRE) { // the three values a state takes among the stencil points: v, v + h, (v + h) + h
#pragma unroll 1
        for (int j = 0; j < D::NS; ++j) {
            const double v = nd.xs[j], hj = nd.tol * (1 + fabs(v));
            pvs[0][j] = P::pre_value(C, nd.p + 1, j, v);
            pvs[1][j] = P::pre_value(C, nd.p + 1, j, v + hj);
            pvs[2][j] = P::pre_value(C, nd.p + 1, j, (v + hj) + hj);
        }
        pv = pvs;
    }
    for (int a = a0; a < a1; ++a) {
        const double ha = nd.tol * (1 + fabs(hess_var(nd, a)));
        double xa[D::NSa], ua[D::NCa], fa[D::NSa], ca[D::NPa];
#pragma unroll
        for (int j = 0; j < D::NS; ++j) xa[j] = (a == j) ? nd.xs[j] + ha : nd.xs[j];
#pragma unroll
        for (int j = 0; j < D::NC; ++j) ua[j] = (a == D::NS + j) ? nd.us[j] + ha : nd.us[j];
        const double ta = (a == T) ? nd.t + ha : nd.t;
        hess_dae<P>(C, nd.p + 1, ta, xa, ua, pv, a, -1, fa, ca);
        const double La = P::lagrange(C, nd.p + 1, ta, xa, ua);
        const int bend = b1 < a + 1 ? b1 : a + 1;
        if constexpr (CACHE) {
#pragma unroll
            for (int s = 0; s < D::NS; ++s) fsv[a][s] = fa[s];
#pragma unroll
            for (int s = 0; s < D::NP; ++s) csv[a][s] = ca[s];
            Lsv[a] = La;
            for (int b = b0; b < bend; ++b) {
                if (b >= a0) hess_pair_generic<P>(pd, C, nd, a, ha, xa, ua, ta, fa, ca, La, b, scr, gid, tot, fsv[b], csv[b], Lsv[b], pv);
                else hess_pair_generic<P>(pd, C, nd, a, ha, xa, ua, ta, fa, ca, La, b, scr, gid, tot, nullptr, nullptr, 0.0, pv);
            }
        } else {
            for (int b = b0; b < bend; ++b) hess_pair_generic<P>(pd, C, nd, a, ha, xa, ua, ta, fa, ca, La, b, scr, gid, tot);
        }
    }
}

// Generic node kernel: run-time pair loops, a-outer so that F(v_a + h_a) stays in registers; gridDim.y splits
// the a-range.  Used for functor sets without a dependency table (dense dynamics) and as the on-device
// reference of the tiled kernel.
template <class P>
__global__ void __launch_bounds__(128)
k_hess_nodes(const __grid_constant__ ProblemDev pd, const __grid_constant__ typename P::Consts C, int nbatch,
             const double* __restrict__ x, const double* __restrict__ sigma, const double* __restrict__ lambda,
             double* __restrict__ vals, double* __restrict__ scr)
{
    typedef Dim<P> D;
    const long long tot = (long long)nbatch * pd.total_nodes;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= tot) return;
    HessNode<P> nd;
    hess_load_node<P>(pd, C, gid, x, sigma, lambda, vals, nd);
    const int nchunk = gridDim.y;
    const int abeg = (int)(((long long)D::NCOL * blockIdx.y) / nchunk);
    const int aend = (int)(((long long)D::NCOL * (blockIdx.y + 1)) / nchunk);
    hess_range_generic<P, true>(pd, C, nd, abeg, aend, 0, D::NCOL, scr, gid, tot);
}

// ------------------------------------------------------------------------------------------
// tiled node kernel for functor sets with a dependency table (P::HESS_DEP, lpb_functor.h)
// ------------------------------------------------------------------------------------------
// Every pair body is generated at compile time (variable indices are constants), so that
//  * a perturbed evaluation shares every subexpression that does not read the perturbed variables with
//    evaluations already made (exact CSE: no fast-math, --fmad=false), and unused rows are dead code;
//  * only rows that read BOTH variables get the 4-point stencil.  For a row that does not read b the
//    stencil is ((fa - fa) - f) + f = +0 exactly (finite values) and is dropped; for a row that reads b but
//    not a it is ((fb - f) - fb) + f, a function of b alone that is zero unless fb - f rounds (|f| below
//    |fb - f|): computed from the single perturbation and added when non-zero.  Same operations in the same
//    order as hess_pair_generic on every row that can contribute -> bit-identical values for finite inputs.
//  * quotients are branch-free (FdDiv::quot_fast); a pair whose quotient leaves the fast range of the
//    division sequence (denormal results) is redone by the generic body.
// The lower triangle of variable pairs is cut into G x G tiles (G = P::HESS_TILE) on gridDim.y: a CTA
// executes one tile, whose straight-line code fits the instruction cache, and all CTAs resident at a time
// run the same tile (blockIdx.x varies fastest).  Single perturbations of the tile's b-range are evaluated
// once per tile, ahead of the pair bodies.
template <int I0, int I1, class F>
__device__ __forceinline__ void static_for(F&& fn)
{
    if constexpr (I0 < I1) {
        fn(std::integral_constant<int, I0>{});
        static_for<I0 + 1, I1>(fn);
    }
}

template <class P, class = void> struct has_hess_dep { static constexpr bool value = false; };
template <class P> struct has_hess_dep<P, std::void_t<decltype(P::HESS_DEP[0])>> { static constexpr bool value = true; };

template <class P, class = void> struct hess_tile_size { static constexpr int value = (Dim<P>::NCOL <= 8 ? Dim<P>::NCOL : 6); };
template <class P> struct hess_tile_size<P, std::enable_if_t<(P::HESS_TILE > 0)>> { static constexpr int value = P::HESS_TILE; };

template <class P, int G_ = 0>
struct HessTiles {
    typedef Dim<P> D;
    static constexpr int G = G_ > 0 ? G_ : hess_tile_size<P>::value;
    static constexpr int NG = (D::NCOL + G - 1) / G;
    static constexpr int NT = NG * (NG + 1) / 2;
    // row r < NROW: dae / path row r; r == NROW: the Lagrange integrand.  v in [states, controls, time].
    template <int R, int V> static constexpr bool dep = ((P::HESS_DEP[R] >> V) & 1ull) != 0;
};

template <class P>
__device__ __noinline__ void hess_tile_slow(const ProblemDev& pd, const typename P::Consts& C, long long gid, long long tot,
                                            const double* __restrict__ x, const double* __restrict__ sigma,
                                            const double* __restrict__ lambda, double* __restrict__ vals, double* __restrict__ scr,
                                            int a0, int a1, int b0, int b1)
{
    HessNode<P> nd;
    hess_load_node<P>(pd, C, gid, x, sigma, lambda, vals, nd);
    hess_range_generic<P>(pd, C, nd, a0, a1, b0, b1, scr, gid, tot);
}

template <class P, int G, int GI, int GJ>
__device__ __forceinline__ bool hess_tile(const ProblemDev& pd, const typename P::Consts& C, const HessNode<P>& nd,
                                          double* __restrict__ scr, long long gid, long long tot)
{
    typedef Dim<P> D;
    typedef HessTiles<P, G> HT;
    constexpr int T = D::NS + D::NC;
    constexpr int NV = D::NS + D::NC;
    constexpr int A0 = GI * G, A1 = (GI + 1) * G < D::NCOL ? (GI + 1) * G : D::NCOL;
    constexpr int B0 = GJ * G, B1x = (GJ + 1) * G < D::NCOL ? (GJ + 1) * G : D::NCOL;
    const PhaseDev& ph = pd.ph[nd.p];
    const int p = nd.p;
    const double t = nd.t, tol = nd.tol;
    bool ok = true; // every quotient of the tile stayed on the fast path of the division sequence
    bool analytic = false;
    if constexpr (P::HAS_ANALYTIC) analytic = pd.analytic != 0;

    // single perturbations of the tile's b-range; they dominate every pair body, so the compiler shares them
    double fbs[G][D::NSa], cbs[G][D::NPa], Lbs[G], hbs[G];
    static_for<B0, B1x>([&](auto bc) {
        constexpr int b = decltype(bc)::value;
        double vb = t;
        if constexpr (b < D::NS) vb = nd.xs[b < D::NS ? b : 0];
        else if constexpr (b < T) vb = nd.us[(b >= D::NS && b < T) ? b - D::NS : 0];
        const double hb = tol * (1 + fabs(vb));
        hbs[b - B0] = hb;
        double xq[D::NSa], uq[D::NCa];
#pragma unroll
        for (int j = 0; j < D::NS; ++j) xq[j] = (b == j) ? nd.xs[j] + hb : nd.xs[j];
#pragma unroll
        for (int j = 0; j < D::NC; ++j) uq[j] = (b == D::NS + j) ? nd.us[j] + hb : nd.us[j];
        const double tq = (b == T) ? t + hb : t;
        P::dae(C, p + 1, tq, xq, uq, fbs[b - B0], cbs[b - B0]);
        Lbs[b - B0] = P::lagrange(C, p + 1, tq, xq, uq);
    });

    static_for<A0, A1>([&](auto ac) {
        constexpr int a = decltype(ac)::value;
        double va = t;
        if constexpr (a < D::NS) va = nd.xs[a < D::NS ? a : 0];
        else if constexpr (a < T) va = nd.us[(a >= D::NS && a < T) ? a - D::NS : 0];
        const double ha = tol * (1 + fabs(va));
        double xa[D::NSa], ua[D::NCa], fa[D::NSa], ca[D::NPa];
#pragma unroll
        for (int j = 0; j < D::NS; ++j) xa[j] = (a == j) ? nd.xs[j] + ha : nd.xs[j];
#pragma unroll
        for (int j = 0; j < D::NC; ++j) ua[j] = (a == D::NS + j) ? nd.us[j] + ha : nd.us[j];
        const double ta = (a == T) ? t + ha : t;
        double La;
        if constexpr (a >= B0 && a < B1x) { // diagonal tile: F(v_a + h_a) is one of the singles
#pragma unroll
            for (int s = 0; s < D::NS; ++s) fa[s] = fbs[a - B0][s];
#pragma unroll
            for (int s = 0; s < D::NP; ++s) ca[s] = cbs[a - B0][s];
            La = Lbs[a - B0];
        } else {
            P::dae(C, p + 1, ta, xa, ua, fa, ca);
            La = P::lagrange(C, p + 1, ta, xa, ua);
        }
        constexpr int B1 = B1x < a + 1 ? B1x : a + 1;
        HessRow hr{~0ull, 0, 0};
        if constexpr (a < T) hr = ph.hrow[a]; // one load per row: which pairs (a, b) the pattern holds, and where
        static_for<B0, B1>([&](auto bc) {
            constexpr int b = decltype(bc)::value;
            if ((hr.mask >> b) & 1ull) { // else: pair absent from the pattern (dependency mask)
                const int blk = hr.blk0 + __popcll(hr.mask & ((1ull << b) - 1ull));
                const double hb = hbs[b - B0];
                const double* fb = fbs[b - B0];
                const double* cb = cbs[b - B0];
                const double Lb = Lbs[b - B0];
                double xq[D::NSa], uq[D::NCa], fab[D::NSa], cab[D::NPa];
#pragma unroll
                for (int j = 0; j < D::NS; ++j) xq[j] = (b == j) ? xa[j] + hb : xa[j];
#pragma unroll
                for (int j = 0; j < D::NC; ++j) uq[j] = (b == D::NS + j) ? ua[j] + hb : ua[j];
                const double tq2 = (b == T) ? ta + hb : ta;
                P::dae(C, p + 1, tq2, xq, uq, fab, cab);
                const double den = ha * hb;
                const FdDiv dvp(den);
                double sdae = 0.0, spath = 0.0;
                static_for<0, D::NS>([&](auto sc) {
                    constexpr int s = decltype(sc)::value;
                    if constexpr (HT::template dep<s, a> && HT::template dep<s, b>) {
                        sdae += nd.lm[s] * dvp.quot_fast(fab[s] - fa[s] - fb[s] + nd.f[s], ok);
                    } else if constexpr (HT::template dep<s, b>) {
                        const double tb = fb[s] - nd.f[s] - fb[s] + nd.f[s];
                        if (tb != 0.0) sdae += nd.lm[s] * dvp.quot_num(tb);
                    }
                });
                static_for<0, D::NP>([&](auto sc) {
                    constexpr int s = decltype(sc)::value;
                    if constexpr (HT::template dep<D::NS + s, a> && HT::template dep<D::NS + s, b>) {
                        spath += nd.mu[s] * dvp.quot_fast(cab[s] - ca[s] - cb[s] + nd.c[s], ok);
                    } else if constexpr (HT::template dep<D::NS + s, b>) {
                        const double tb = cb[s] - nd.c[s] - cb[s] + nd.c[s];
                        if (tb != 0.0) spath += nd.mu[s] * dvp.quot_num(tb);
                    }
                });
                double hl = 0.0;
                if constexpr (HT::template dep<D::NROW, a> && HT::template dep<D::NROW, b>) {
                    const double Lab = P::lagrange(C, p + 1, tq2, xq, uq);
                    hl = dvp.quot_fast(Lab - La - Lb + nd.L, ok);
                } else if constexpr (HT::template dep<D::NROW, b>) {
                    const double tb = Lb - nd.L - Lb + nd.L;
                    if (tb != 0.0) hl = dvp.quot_num(tb);
                }
                const double sL = nd.sg * nd.w * hl;
                const double core = (nd.tf - nd.t0) / 2.0 * (sL - sdae) + spath; // :123-127
                double A = 0.0;
                if constexpr (a == T) {
                    double acc = 0.0, dLb = 0.0;
                    if (analytic) {
                        hess_time_analytic<P>(C, nd, b, acc, dLb);
                    } else {
                        const FdDiv dvb(hb);
                        static_for<0, D::NS>([&](auto sc) {
                            constexpr int s = decltype(sc)::value;
                            if constexpr (HT::template dep<s, b>) acc += nd.lm[s] * dvb.quot_fast(fb[s] - nd.f[s], ok);
                        });
                        if constexpr (HT::template dep<D::NROW, b>) dLb = dvb.quot_fast(Lb - nd.L, ok);
                    }
                    A = acc - nd.sg * nd.w * dLb;
                }
                hess_store_pair<P>(ph, nd, a == T, b == T, b, blk, core, A, scr, gid, tot);
            }
        });
    });
    return ok;
}

template <class P, int G, int TI, int GI, int GJ>
__device__ __forceinline__ void hess_dispatch(int tile, const ProblemDev& pd, const typename P::Consts& C, const typename P::Consts& Cparam,
                                              const HessNode<P>& nd, long long gid, long long tot,
                                              const double* __restrict__ x, const double* __restrict__ sigma,
                                              const double* __restrict__ lambda, double* __restrict__ vals, double* __restrict__ scr)
{
    typedef Dim<P> D;
    if constexpr (TI < HessTiles<P, G>::NT) {
        if (tile == TI) {
            if (!hess_tile<P, G, GI, GJ>(pd, C, nd, scr, gid, tot)) {
                constexpr int A1 = (GI + 1) * G < D::NCOL ? (GI + 1) * G : D::NCOL;
                constexpr int B1 = (GJ + 1) * G < D::NCOL ? (GJ + 1) * G : D::NCOL;
                hess_tile_slow<P>(pd, Cparam, gid, tot, x, sigma, lambda, vals, scr, GI * G, A1, GJ * G, B1);
            }
        } else
            hess_dispatch<P, G, TI + 1, (GJ == GI ? GI + 1 : GI), (GJ == GI ? 0 : GJ + 1)>(tile, pd, C, Cparam, nd, gid, tot, x, sigma, lambda,
                                                                                          vals, scr);
    }
}

template <class P, int G, int MINB>
__global__ void __launch_bounds__(128, MINB)
k_hess_tiled(const __grid_constant__ ProblemDev pd, const __grid_constant__ typename P::Consts Cg, int nbatch,
             const double* __restrict__ x, const double* __restrict__ sigma, const double* __restrict__ lambda,
             double* __restrict__ vals, double* __restrict__ scr)
{
    const long long tot = (long long)nbatch * pd.total_nodes;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= tot) return;
    // register copy of the functor constants: the value stores between the pair bodies must not make the
    // compiler reload them (and lose the CSE against the base evaluation)
    const typename P::Consts C = Cg;
    HessNode<P> nd;
    hess_load_node<P>(pd, C, gid, x, sigma, lambda, vals, nd);
    hess_dispatch<P, G, 0, 0, 0>((int)blockIdx.y, pd, C, Cg, nd, gid, tot, x, sigma, lambda, vals, scr);
}

// endpoint variable index e: [0,NS) x0, [NS,2NS) xf, 2NS t0, 2NS+1 tf
template <class P>
__device__ __forceinline__ void bump_endpoint(int e, double h, double* x0, double* xf, double& t0, double& tf)
{
    typedef Dim<P> D;
#pragma unroll
    for (int j = 0; j < D::NS; ++j) {
        if (e == j) x0[j] += h;
        if (e == D::NS + j) xf[j] += h;
    }
    if (e == 2 * D::NS) t0 += h;
    if (e == 2 * D::NS + 1) tf += h;
}

// grid.x = instance, grid.y = P + Lp roles, 128 threads; launched after the node kernel
template <class P>
__global__ void __launch_bounds__(128)
k_hess_endpoint(const __grid_constant__ ProblemDev pd, const __grid_constant__ typename P::Consts C, int nbatch,
                const double* __restrict__ x, const double* __restrict__ sigma, const double* __restrict__ lambda,
                double* __restrict__ vals, const double* __restrict__ scr,
                const HessEntry* __restrict__ eent, int n_eent, const HessEntry* __restrict__ lent, int n_lent)
{
    typedef Dim<P> D;
    __shared__ double sm[128];
    const int b_inst = blockIdx.x;
    const int role = blockIdx.y;
    const double* __restrict__ xi = x + (size_t)b_inst * pd.n;
    const double* __restrict__ lami = lambda + (size_t)b_inst * pd.m;
    double* __restrict__ vi = vals + (size_t)b_inst * pd.nnz_h;
    const double sg = sigma[b_inst];
    const double tol = pd.tol;
    const long long tot = (long long)nbatch * pd.total_nodes;
    if (role < pd.P) {
        const PhaseDev& ph = pd.ph[role];
        const int N = ph.N;
        constexpr int NV = D::NS + D::NC;
        // three time-time scalars of the I-part: deterministic block sums of the node terms
        const size_t off = (size_t)b_inst * pd.total_nodes + ph.node0;
        const double d1 = block_sum<128>(scr + off, N, sm);
        const double d2 = block_sum<128>(scr + tot + off, N, sm);
        const double d3 = block_sum<128>(scr + 2 * tot + off, N, sm);
        const double d4 = block_sum<128>(scr + 3 * tot + off, N, sm);
        if (threadIdx.x == 0) {
            double* vI = vi + ph.hI0;
            vI[(size_t)(ph.nblkH + NV) * N] = d1;                 // t0t0 (:489)
            vI[(size_t)(ph.nblkH + 2 * NV) * N + 1] = 0.5 * d3 + d4; // tft0 (:522)
            vI[(size_t)(ph.nblkH + 2 * NV) * N + 2] = d2;         // tftf (:531)
        }
        const double* xb = xi + ph.var0;
        double x0[D::NSa], xf[D::NSa];
#pragma unroll
        for (int j = 0; j < D::NS; ++j) { x0[j] = xb[(size_t)j * (N + 1)]; xf[j] = xb[(size_t)j * (N + 1) + N]; }
        const double t0 = xb[(size_t)D::NS * (N + 1) + (size_t)D::NC * N];
        const double tf = xb[(size_t)D::NS * (N + 1) + (size_t)D::NC * N + 1];
        const double M = P::mayer(C, role + 1, t0, x0, tf, xf);
        double ev[D::NEa];
#pragma unroll
        for (int q = 0; q < D::NEa; ++q) ev[q] = 0.0;
        if (ph.ne > 0) P::event(C, role + 1, t0, x0, tf, xf, ev);
        const double* lev = lami + ph.con0 + (size_t)(D::NS + D::NP) * N;
        for (int idx = threadIdx.x; idx < n_eent; idx += blockDim.x) {
            const HessEntry en = eent[idx];
            auto val_of = [&](int e) {
                double v = (e == 2 * D::NS) ? t0 : tf;
#pragma unroll
                for (int j = 0; j < D::NS; ++j) { v = (e == j) ? x0[j] : v; v = (e == D::NS + j) ? xf[j] : v; }
                return v;
            };
            const double ha = tol * (1 + fabs(val_of(en.a)));
            const double hb = tol * (1 + fabs(val_of(en.b)));
            const double den = (tol * (1 + fabs(val_of(en.da)))) * (tol * (1 + fabs(val_of(en.db)))); // quirk Q10 via (da, db)
            double x0a[D::NSa], xfa[D::NSa], x0b[D::NSa], xfb[D::NSa];
            double t0a = t0, tfa = tf, t0b = t0, tfb = tf;
#pragma unroll
            for (int j = 0; j < D::NS; ++j) { x0a[j] = x0[j]; xfa[j] = xf[j]; x0b[j] = x0[j]; xfb[j] = xf[j]; }
            bump_endpoint<P>(en.a, ha, x0a, xfa, t0a, tfa);
            bump_endpoint<P>(en.b, hb, x0b, xfb, t0b, tfb);
            const double Ma = P::mayer(C, role + 1, t0a, x0a, tfa, xfa);
            const double Mb = P::mayer(C, role + 1, t0b, x0b, tfb, xfb);
            double ea[D::NEa], eb[D::NEa], eab[D::NEa];
            if (ph.ne > 0) {
#pragma unroll
                for (int q = 0; q < D::NEa; ++q) { ea[q] = 0.0; eb[q] = 0.0; eab[q] = 0.0; }
                P::event(C, role + 1, t0a, x0a, tfa, xfa, ea);
                P::event(C, role + 1, t0b, x0b, tfb, xfb, eb);
            }
            bump_endpoint<P>(en.b, hb, x0a, xfa, t0a, tfa); // (a then b)
            const double Mab = P::mayer(C, role + 1, t0a, x0a, tfa, xfa);
            double lsum = 0.0;
            if (ph.ne > 0) {
                P::event(C, role + 1, t0a, x0a, tfa, xfa, eab);
                for (int q = 0; q < ph.ne; ++q) lsum += ((eab[q] - ea[q] - eb[q] + ev[q]) / (den * 1.0)) * lev[q];
            }
            vi[ph.hE0 + idx] = sg * ((Mab - Ma - Mb + M) / den) + lsum; // :293-347
        }
    } else {
        const LinkDev& lk = pd.lk[role - pd.P];
        const PhaseDev& pl = pd.ph[lk.left];
        const PhaseDev& pr = pd.ph[lk.right];
        double xl[D::NSa], xr[D::NSa], lo[D::NLa];
#pragma unroll
        for (int j = 0; j < D::NS; ++j) {
            xl[j] = xi[pl.var0 + (size_t)j * (pl.N + 1) + pl.N];
            xr[j] = xi[pr.var0 + (size_t)j * (pr.N + 1)];
        }
#pragma unroll
        for (int q = 0; q < D::NLa; ++q) lo[q] = 0.0;
        P::link(C, xl, xr, lo);
        const double* ll = lami + lk.lam0;
        for (int idx = threadIdx.x; idx < n_lent; idx += blockDim.x) {
            const HessEntry en = lent[idx];
            auto val_of = [&](int e) {
                double v = 0.0;
#pragma unroll
                for (int j = 0; j < D::NS; ++j) { v = (e == j) ? xl[j] : v; v = (e == D::NS + j) ? xr[j] : v; }
                return v;
            };
            const double ha = tol * (1 + fabs(val_of(en.a)));
            const double hb = tol * (1 + fabs(val_of(en.b)));
            double la[D::NSa], ra[D::NSa], lb[D::NSa], rb[D::NSa];
#pragma unroll
            for (int j = 0; j < D::NS; ++j) { la[j] = xl[j]; ra[j] = xr[j]; lb[j] = xl[j]; rb[j] = xr[j]; }
#pragma unroll
            for (int j = 0; j < D::NS; ++j) {
                if (en.a == j) la[j] += ha;
                if (en.a == D::NS + j) ra[j] += ha;
                if (en.b == j) lb[j] += hb;
                if (en.b == D::NS + j) rb[j] += hb;
            }
            double oa[D::NLa], ob[D::NLa], oab[D::NLa];
#pragma unroll
            for (int q = 0; q < D::NLa; ++q) { oa[q] = 0.0; ob[q] = 0.0; oab[q] = 0.0; }
            P::link(C, la, ra, oa);
            P::link(C, lb, rb, ob);
#pragma unroll
            for (int j = 0; j < D::NS; ++j) {
                if (en.b == j) la[j] += hb;
                if (en.b == D::NS + j) ra[j] += hb;
            }
            P::link(C, la, ra, oab);
            double acc = 0.0;
            for (int q = 0; q < lk.nl; ++q) acc += ((oab[q] - oa[q] - ob[q] + lo[q]) / (ha * hb)) * ll[q];
            vi[lk.h0 + idx] = acc;
        }
    }
}

// LpDerivDependciesChecker.cpp:10-94: one thread per (phase, variable): set the variable to NaN
// at node index 1 of the guess, evaluate dae at that node, flag non-finite outputs.
template <class P>
__global__ void k_probe(const __grid_constant__ ProblemDev pd, const __grid_constant__ typename P::Consts C,
                        const double* __restrict__ x, int* __restrict__ dep)
{
    typedef Dim<P> D;
    constexpr int NV = D::NS + D::NC;
    const int p = blockIdx.x;
    const int v = threadIdx.x;
    if (p >= pd.P || v >= NV) return;
    const PhaseDev& ph = pd.ph[p];
    const int N = ph.N, k = 1;
    const double* xb = x + ph.var0;
    double xs[D::NSa], us[D::NCa], f[D::NSa], c[D::NPa];
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
    for (int j = 0; j < D::NS; ++j) xs[j] = (v == j) ? nan : xb[(size_t)j * (N + 1) + k];
#pragma unroll
    for (int j = 0; j < D::NC; ++j) us[j] = (v == D::NS + j) ? nan : xb[(size_t)D::NS * (N + 1) + (size_t)j * N + k];
    const double t0 = xb[(size_t)D::NS * (N + 1) + (size_t)D::NC * N];
    const double tf = xb[(size_t)D::NS * (N + 1) + (size_t)D::NC * N + 1];
    const double t = (ph.tau[k] + 1) * ((tf - t0) / 2.0) + t0;
    P::dae(C, p + 1, t, xs, us, f, c);
    int* d = dep + (size_t)p * D::NROW * NV + (size_t)v * D::NROW; // column-major (ns+np) x (ns+nc)
#pragma unroll
    for (int r = 0; r < D::NS; ++r) d[r] = isfinite(f[r]) ? 0 : 1;
#pragma unroll
    for (int r = 0; r < D::NP; ++r) d[D::NS + r] = isfinite(c[r]) ? 0 : 1;
}

// variants of the tiled kernel compiled for a functor set: (tile size, resident CTAs per SM).  Entry 0 is the
// default; LaunchOpts::hess_variant picks another one (tuning runs, scripts/kernel_sweep.py).
template <class P, class = void> struct hess_min_ctas { static constexpr int value = 3; };
template <class P> struct hess_min_ctas<P, std::enable_if_t<(P::HESS_MIN_CTAS > 0)>> { static constexpr int value = P::HESS_MIN_CTAS; };

template <class P, int G, int MINB>
void launch_hess_tiled(const ProblemDev& pd, const typename P::Consts& C, cudaStream_t st, unsigned gx, int nbatch, const double* x,
                       const double* sigma, const double* lambda, double* vals, double* scratch)
{
    k_hess_tiled<P, G, MINB><<<dim3(gx, HessTiles<P, G>::NT), 128, 0, st>>>(pd, C, nbatch, x, sigma, lambda, vals, scratch);
}

template <class P>
int launch_hessian(const ProblemDev& pd, const void* consts, cudaStream_t st, const LaunchOpts& o,
                   int nbatch, const double* x, const double* sigma, const double* lambda, double* vals, double* scratch)
{
    typedef Dim<P> D;
    const typename P::Consts& C = *static_cast<const typename P::Consts*>(consts);
    const long long tot = (long long)nbatch * pd.total_nodes;
    const int block = 128;
    const unsigned gx = (unsigned)((tot + block - 1) / block);
    if (o.ev_begin) cudaEventRecord(o.ev_begin, st);
    bool tiled = false;
    if constexpr (has_hess_dep<P>::value) tiled = o.unroll_colours != 0;
    if (tiled) {
        if constexpr (has_hess_dep<P>::value) {
            constexpr int G0 = hess_tile_size<P>::value, M0 = hess_min_ctas<P>::value;
#ifdef LPB_HESS_VARIANTS
            switch (o.hess_variant) {
            case 1: launch_hess_tiled<P, (D::NCOL < 6 ? D::NCOL : 6), 2>(pd, C, st, gx, nbatch, x, sigma, lambda, vals, scratch); break;
            case 2: launch_hess_tiled<P, (D::NCOL < 6 ? D::NCOL : 6), 4>(pd, C, st, gx, nbatch, x, sigma, lambda, vals, scratch); break;
            case 3: launch_hess_tiled<P, (D::NCOL < 5 ? D::NCOL : 5), 3>(pd, C, st, gx, nbatch, x, sigma, lambda, vals, scratch); break;
            case 4: launch_hess_tiled<P, (D::NCOL < 9 ? D::NCOL : 9), 2>(pd, C, st, gx, nbatch, x, sigma, lambda, vals, scratch); break;
            default: launch_hess_tiled<P, G0, M0>(pd, C, st, gx, nbatch, x, sigma, lambda, vals, scratch); break;
            }
#else
            launch_hess_tiled<P, G0, M0>(pd, C, st, gx, nbatch, x, sigma, lambda, vals, scratch);
#endif
        }
    } else {
        int split = o.pair_split;
        if (split <= 0) {
            // every chunk of the outer range re-evaluates the single perturbations below its first variable (the cache
            // of hess_range_generic only covers its own chunk): split only when the node blocks alone leave SMs idle
            // (config 5, 782 blocks: 25.6 ms unsplit against 31.6 ms in four chunks)
            const long long want = 4LL * o.sm_count * 4;
            split = gx >= 2LL * o.sm_count ? 1 : (int)((want + gx - 1) / gx);
        }
        if (split < 1) split = 1;
        if (split > D::NCOL) split = D::NCOL;
        k_hess_nodes<P><<<dim3(gx, split), block, 0, st>>>(pd, C, nbatch, x, sigma, lambda, vals, scratch);
    }
    if (o.ev_end) cudaEventRecord(o.ev_end, st);
    // grid.x = instance, grid.y = P + Lp roles: any batch size
    k_hess_endpoint<P><<<dim3(nbatch, pd.P + pd.Lp), 128, 0, st>>>(pd, C, nbatch, x, sigma, lambda, vals, scratch,
                                                                  pd.eent, pd.n_eent, pd.lent, pd.n_lent);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 2 : cuda_fail(e);
}

// the functor set's declared dependency masks for the host (validated against the NaN probe), or null
template <class P>
const unsigned long long* hess_dep_table()
{
    if constexpr (has_hess_dep<P>::value) {
        static_assert(sizeof(P::HESS_DEP) / sizeof(P::HESS_DEP[0]) == Dim<P>::NROW + 1,
                      "HESS_DEP needs one mask per dae row, per path row, and one for the Lagrange integrand");
        static_assert(Dim<P>::NCOL <= 64 && Dim<P>::NROW < 64, "HESS_DEP masks hold at most 64 variables");
        static unsigned long long copy[Dim<P>::NROW + 1];
        for (int r = 0; r <= Dim<P>::NROW; ++r) copy[r] = P::HESS_DEP[r];
        return copy;
    } else
        return nullptr;
}

template <class P>
int launch_probe(const ProblemDev& pd, const void* consts, cudaStream_t st, const double* x, int* dep_out)
{
    typedef Dim<P> D;
    const typename P::Consts& C = *static_cast<const typename P::Consts*>(consts);
    int thr = D::NS + D::NC;
    thr = (thr + 31) / 32 * 32;
    k_probe<P><<<pd.P, thr, 0, st>>>(pd, C, x, dep_out);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 1 : cuda_fail(e);
}

template <class P>
size_t scratch_doubles(const ProblemDev& pd, int nbatch)
{
    // objective: 1 array; gradient: 3 arrays + P per instance; Hessian: 4 arrays
    return (size_t)4 * nbatch * pd.total_nodes + (size_t)nbatch * pd.P + 16;
}

template <class P>
const FunctorVTable* make_vtable()
{
    static_assert(FunctorConcept<P>::value, "functor set does not satisfy include/lpb_functor.h");
    static const FunctorVTable vt = {
        P::name(), P::NS, P::NC, P::NPATH, P::NE_MAX, P::NL_MAX,
        (int)(sizeof(typename P::Consts) / sizeof(double)), P::HAS_ANALYTIC ? 1 : 0,
        &launch_cons_jac<P>, &launch_objective<P>, &launch_gradient<P>, &launch_hessian<P>, &launch_probe<P>,
        &scratch_doubles<P>, &launch_mesh_error<P>, &launch_nlp2op<P>, hess_dep_table<P>()};
    return &vt;
}

} // namespace lpb

// One translation unit per functor set instantiates every kernel for it.
#define LPB_DEFINE_FUNCTOR(TYPE) \
    extern "C" const lpb::FunctorVTable* lpb_vtable_##TYPE() { return lpb::make_vtable<TYPE>(); }
