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

// UNROLL: the pair loops carry full-unroll pragmas; as in the Jacobian kernel, perturbed
// evaluations F(v_b+h_b) and F(v_a+h_a, v_b+h_b) whose variable indices are compile-time constants
// share every subexpression that does not depend on the perturbed variables with evaluations
// already made (exact CSE).  The compiler unrolls as far as its code-size budget allows; forcing
// all (ns+nc+1)(ns+nc+2)/2 pair bodies through template recursion was measured and is slower for
// 17 variables (instruction-cache bound: 4.4 ms vs 2.2 ms per 4096 quadrotor instances).
template <class P, bool UNROLL>
__global__ void __launch_bounds__(128)
k_hess_nodes(const __grid_constant__ ProblemDev pd, const __grid_constant__ typename P::Consts C, int nbatch,
             const double* __restrict__ x, const double* __restrict__ sigma, const double* __restrict__ lambda,
             double* __restrict__ vals, double* __restrict__ scr)
{
    typedef Dim<P> D;
    constexpr int T = D::NS + D::NC; // index of the time variable
    constexpr int NV = D::NS + D::NC;
    const long long tot = (long long)nbatch * pd.total_nodes;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= tot) return;
    const int b_inst = (int)(gid / pd.total_nodes);
    const int gnode = (int)(gid - (long long)b_inst * pd.total_nodes);
    const int p = find_phase(pd, gnode);
    const PhaseDev& ph = pd.ph[p];
    const int k = gnode - ph.node0, N = ph.N;
    const double* __restrict__ xb = x + (size_t)b_inst * pd.n + ph.var0;
    const double* __restrict__ lam = lambda + (size_t)b_inst * pd.m + ph.con0;
    double* __restrict__ vI = vals + (size_t)b_inst * pd.nnz_h + ph.hI0;
    const double sg = sigma[b_inst];
    const double tol = pd.tol;

    double xs[D::NSa], us[D::NCa];
#pragma unroll
    for (int j = 0; j < D::NS; ++j) xs[j] = xb[(size_t)j * (N + 1) + k];
#pragma unroll
    for (int j = 0; j < D::NC; ++j) us[j] = xb[(size_t)D::NS * (N + 1) + (size_t)j * N + k];
    const double t0 = xb[(size_t)D::NS * (N + 1) + (size_t)D::NC * N];
    const double tf = xb[(size_t)D::NS * (N + 1) + (size_t)D::NC * N + 1];
    const double tspan = tf - t0;
    const double tau = ph.tau[k], w = ph.w[k];
    const double t = (tau + 1) * (tspan / 2.0) + t0;
    double lm[D::NSa], mu[D::NPa];
#pragma unroll
    for (int s = 0; s < D::NS; ++s) lm[s] = lam[(size_t)s * N + k];
#pragma unroll
    for (int s = 0; s < D::NP; ++s) mu[s] = lam[(size_t)(D::NS + s) * N + k];

    double f[D::NSa], c[D::NPa];
    P::dae(C, p + 1, t, xs, us, f, c);
    const double L = P::lagrange(C, p + 1, t, xs, us);
    const double talpha = (1 - tau) / 2.0, tbeta = (1 + tau) / 2.0;

    // analytic first derivatives for the time rows (derive_fun_ = LpAnalyticDerive, :161,:175)
    double dd[P::HAS_ANALYTIC ? D::NROW * D::NCOL : 1], dl[P::HAS_ANALYTIC ? D::NCOL : 1];
    bool analytic = false;
    if constexpr (P::HAS_ANALYTIC) {
        if (pd.analytic) {
            analytic = true;
            P::ddae(C, p + 1, t, xs, us, dd);
            P::dlagrange(C, p + 1, t, xs, us, dl);
        }
    }

    const int nchunk = gridDim.y;
    const int abeg = (int)(((long long)D::NCOL * blockIdx.y) / nchunk);
    const int aend = (int)(((long long)D::NCOL * (blockIdx.y + 1)) / nchunk);
#pragma unroll(UNROLL ? D::NCOL : 1)
    for (int a = 0; a < D::NCOL; ++a) {
        if (a < abeg || a >= aend) continue;
        double va = t;
#pragma unroll
        for (int j = 0; j < D::NS; ++j) va = (a == j) ? xs[j] : va;
#pragma unroll
        for (int j = 0; j < D::NC; ++j) va = (a == D::NS + j) ? us[j] : va;
        const double ha = tol * (1 + fabs(va));
        double xa[D::NSa], ua[D::NCa], fa[D::NSa], ca[D::NPa];
#pragma unroll
        for (int j = 0; j < D::NS; ++j) xa[j] = (a == j) ? xs[j] + ha : xs[j];
#pragma unroll
        for (int j = 0; j < D::NC; ++j) ua[j] = (a == D::NS + j) ? us[j] + ha : us[j];
        const double ta = (a == T) ? t + ha : t;
        P::dae(C, p + 1, ta, xa, ua, fa, ca);
        const double La = P::lagrange(C, p + 1, ta, xa, ua);
#pragma unroll(UNROLL ? D::NCOL : 1)
        for (int b = 0; b < D::NCOL; ++b) {
            if (b > a) continue;
            int blk = 0;
            if (a < T) {
                blk = ph.hblk[a * NV + b];
                if (blk < 0) continue; // pair absent from the pattern (dependency mask)
            }
            double vb = t;
#pragma unroll
            for (int j = 0; j < D::NS; ++j) vb = (b == j) ? xs[j] : vb;
#pragma unroll
            for (int j = 0; j < D::NC; ++j) vb = (b == D::NS + j) ? us[j] : vb;
            const double hb = tol * (1 + fabs(vb));
            double xq[D::NSa], uq[D::NCa], fb[D::NSa], cb[D::NPa], fab[D::NSa], cab[D::NPa];
            double Lb;
            if (a == T && b == T) { // stateoutj = stateouti (:1403,:1407,:2097)
#pragma unroll
                for (int s = 0; s < D::NS; ++s) fb[s] = fa[s];
#pragma unroll
                for (int s = 0; s < D::NP; ++s) cb[s] = ca[s];
                Lb = La;
            } else {
#pragma unroll
                for (int j = 0; j < D::NS; ++j) xq[j] = (b == j) ? xs[j] + hb : xs[j];
#pragma unroll
                for (int j = 0; j < D::NC; ++j) uq[j] = (b == D::NS + j) ? us[j] + hb : us[j];
                const double tq = (b == T) ? t + hb : t;
                P::dae(C, p + 1, tq, xq, uq, fb, cb);
                Lb = P::lagrange(C, p + 1, tq, xq, uq);
            }
            // point (a then b): start from the a-perturbed point, add h_b to variable b
#pragma unroll
            for (int j = 0; j < D::NS; ++j) xq[j] = (b == j) ? xa[j] + hb : xa[j];
#pragma unroll
            for (int j = 0; j < D::NC; ++j) uq[j] = (b == D::NS + j) ? ua[j] + hb : ua[j];
            const double tq2 = (b == T) ? ta + hb : ta;
            P::dae(C, p + 1, tq2, xq, uq, fab, cab);
            const double Lab = P::lagrange(C, p + 1, tq2, xq, uq);
            const double den = ha * hb;
            const FdDiv dvp(den); // one reciprocal per pair, IEEE-exact quotients (lpb_kernels.cuh)
            double sdae = 0.0, spath = 0.0;
#pragma unroll
            for (int s = 0; s < D::NS; ++s) sdae += lm[s] * dvp.quot_num(fab[s] - fa[s] - fb[s] + f[s]);
#pragma unroll
            for (int s = 0; s < D::NP; ++s) spath += mu[s] * dvp.quot_num(cab[s] - ca[s] - cb[s] + c[s]);
            const double hl = dvp.quot_num(Lab - La - Lb + L);
            const double sL = sg * w * hl;
            const double core = (tf - t0) / 2.0 * (sL - sdae) + spath; // :123-127
            if (a < T) {
                st_stream(vI + (size_t)blk * N + k, core);
            } else {
                // first-derivative term: sum(lambda % df/dv_b) - sigma*w*dL/dv_b  (:184-185,:200-201,:214-215)
                double acc = 0.0, dLb;
                if (analytic) {
                    if constexpr (P::HAS_ANALYTIC) {
#pragma unroll
                        for (int s = 0; s < D::NS; ++s) {
                            double dsb = dd[s * D::NCOL + T];
#pragma unroll
                            for (int j = 0; j < T; ++j) dsb = (b == j) ? dd[s * D::NCOL + j] : dsb;
                            acc += lm[s] * dsb;
                        }
                        dLb = dl[T];
#pragma unroll
                        for (int j = 0; j < T; ++j) dLb = (b == j) ? dl[j] : dLb;
                    }
                } else {
                    const FdDiv dvb(hb);
#pragma unroll
                    for (int s = 0; s < D::NS; ++s) acc += lm[s] * dvb.quot_num(fb[s] - f[s]);
                    dLb = dvb.quot_num(Lb - L);
                }
                const double A = acc - sg * w * dLb;
                if (b < T) {
                    // t0 row block b, tf row block b (:186-189,:202-205; layout :468-518)
                    st_stream(vI + (size_t)(ph.nblkH + b) * N + k, 0.5 * A + talpha * core);
                    st_stream(vI + (size_t)(ph.nblkH + NV) * N + 1 + (size_t)b * N + k, -0.5 * A + tbeta * core);
                } else {
                    // per-node terms of the three time-time scalars (:216-218)
                    scr[gid] = talpha * (A + talpha * core);
                    scr[tot + gid] = tbeta * (-A + tbeta * core);
                    scr[2 * tot + gid] = (tbeta - talpha) * A;
                    scr[3 * tot + gid] = talpha * (tbeta * core);
                }
            }
        }
    }
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

// grid.x = P + Lp roles, grid.y = instance, 128 threads; launched after k_hess_nodes
template <class P>
__global__ void __launch_bounds__(128)
k_hess_endpoint(const __grid_constant__ ProblemDev pd, const __grid_constant__ typename P::Consts C, int nbatch,
                const double* __restrict__ x, const double* __restrict__ sigma, const double* __restrict__ lambda,
                double* __restrict__ vals, const double* __restrict__ scr,
                const HessEntry* __restrict__ eent, int n_eent, const HessEntry* __restrict__ lent, int n_lent)
{
    typedef Dim<P> D;
    __shared__ double sm[128];
    const int b_inst = blockIdx.y;
    const int role = blockIdx.x;
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

template <class P>
int launch_hessian(const ProblemDev& pd, const void* consts, cudaStream_t st, const LaunchOpts& o,
                   int nbatch, const double* x, const double* sigma, const double* lambda, double* vals, double* scratch)
{
    typedef Dim<P> D;
    const typename P::Consts& C = *static_cast<const typename P::Consts*>(consts);
    const long long tot = (long long)nbatch * pd.total_nodes;
    const int block = 128;
    const unsigned gx = (unsigned)((tot + block - 1) / block);
    int split = o.pair_split;
    if (split <= 0) {
        const long long want = 4LL * o.sm_count * 4;
        split = (int)((want + gx - 1) / gx);
    }
    if (split < 1) split = 1;
    if (split > D::NCOL) split = D::NCOL;
    if (o.ev_begin) cudaEventRecord(o.ev_begin, st);
    // the unrolled variant is only instantiated for functor sets that ask for it (its code size and
    // compile time grow with the square of the variable count)
    bool unroll = false;
    if constexpr (P::UNROLL_HESSIAN) unroll = o.unroll_colours != 0;
    if (unroll) {
        if constexpr (P::UNROLL_HESSIAN)
            k_hess_nodes<P, true><<<dim3(gx, split), block, 0, st>>>(pd, C, nbatch, x, sigma, lambda, vals, scratch);
    } else
        k_hess_nodes<P, false><<<dim3(gx, split), block, 0, st>>>(pd, C, nbatch, x, sigma, lambda, vals, scratch);
    if (o.ev_end) cudaEventRecord(o.ev_end, st);
    k_hess_endpoint<P><<<dim3(pd.P + pd.Lp, nbatch), 128, 0, st>>>(pd, C, nbatch, x, sigma, lambda, vals, scratch,
                                                                  pd.eent, pd.n_eent, pd.lent, pd.n_lent);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 2 : cuda_fail(e);
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
    static const FunctorVTable vt = {
        P::name(), P::NS, P::NC, P::NPATH, P::NE_MAX, P::NL_MAX,
        (int)(sizeof(typename P::Consts) / sizeof(double)), P::HAS_ANALYTIC ? 1 : 0,
        &launch_cons_jac<P>, &launch_objective<P>, &launch_gradient<P>, &launch_hessian<P>, &launch_probe<P>,
        &scratch_doubles<P>, &launch_mesh_error<P>, &launch_nlp2op<P>};
    return &vt;
}

} // namespace lpb

// One translation unit per functor set instantiates every kernel for it.
#define LPB_DEFINE_FUNCTOR(TYPE) \
    extern "C" const lpb::FunctorVTable* lpb_vtable_##TYPE() { return lpb::make_vtable<TYPE>(); }
