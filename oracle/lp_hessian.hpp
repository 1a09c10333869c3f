// oracle/lp_hessian.hpp -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of Lpopc::LpHessianCalculator (exact Lagrangian Hessian by
// second-order forward differences) and of the sparse-NaN dependency probe.
// Follows Lpopc/src/Core/LpHessian.cpp:
//   :12-599     GetPhaseHessian (sigma/lambda contraction + scatter, I-part | E-part)
//   :601-876    GetPhaseHessianSparsity
//   :878-1018   GetHessian              :1020-1189  GetLinkHessian
//   :1192-2161  CalculatePhaseHessian   :2163-2367  CalculateLinkHessain
//   :2369-2508  GetLinkHessianSparsity  :2510-2599  GetHessianSparsity
// and Lpopc/src/Core/LpDerivDependciesChecker.cpp:10-94.
// Stencil (LpHessian.cpp:1269-1282): f_i = F(v_i+h_i), f_j = F(v_j+h_j),
// f_ij = F((v_i+h_i), then v_j += h_j), value = (((f_ij - f_i) - f_j) + f)/(h_i*h_j),
// h = tol*(1+|v|) per node.  Quirks kept: Q6 (link_indices), Q7 (left node count
// for right-phase columns), Q10 (x0.xf / xf.xf endpoint denominators use index i).
// nq = 0 (Q3 fenced); Q8 (uninitialised last multiplier) is not replicated.
// Armadillo sum(A%B,1) is restated left-to-right over columns; accu() and the
// (1xN)*(Nx1) products as sequential sums.
#pragma once
#include "lp_nlp.hpp"
#include <limits>

namespace lpo {

class LpHessianCalculator {
public:
    LpHessianCalculator(std::shared_ptr<FunctionWrapper> userfun, LpCalculateData* data, OptimalProblem* optpro, std::shared_ptr<OptDerive> derive_fun, double tolerance)
        : fun_(userfun), Data_(data), optpro_(optpro), derive_fun_(derive_fun), tol(tolerance), nlp_(userfun, data, optpro, derive_fun) {}

    // per-node second derivatives: field(v_row, v_col) of N x ns (dae), N x np (path), N (lagrange).
    // variable index v: [0,ns) states, [ns,ns+nc) controls, ns+nc time.
    struct PhaseHessain {
        int nv = 0;
        std::vector<Mat> hDae, hPath;  // nv*nv, filled only where the reference fills
        std::vector<Vec> hLagrange;    // nv*nv
        // endpoint variable index e: [0,ns) x0, [ns,2ns) xf, 2ns t0, 2ns+1 tf
        int ne_var = 0;
        std::vector<Vec> hEvents;      // ne_var*ne_var, each nevents
        std::vector<double> hMayer;    // ne_var*ne_var
    };

    static int nnz(const Mat& m) { return NLPWrapper::nnz(m); }

    // LpHessian.cpp:2532-2536 / :900-903: depH = trans(dep)*dep, diag forced to 1
    static Mat HessDependencies(const Mat& dep)
    {
        int nv = dep.n_cols;
        Mat t(nv, nv, 0.0);
        for (int i = 0; i < nv; ++i)
            for (int j = 0; j < nv; ++j) {
                double acc = 0;
                for (int r = 0; r < dep.n_rows; ++r) acc += dep(r, i) * dep(r, j);
                t(i, j) = acc;
            }
        for (int i = 0; i < nv; ++i) t(i, i) = 1;
        return t;
    }

    // ---- CalculatePhaseHessian :1192-2161 ------------------------------------------------
    void CalculatePhaseHessian(int iphase, const Vec& x_all, PhaseHessain& H)
    {
        NLPWrapper::PhaseVars v = nlp_.Unpack(iphase, x_all);
        int N = v.sumnodes, ns = v.nstates, nc = v.ncontrols, np = v.npaths, nevents = v.nevents;
        int nv = ns + nc + 1;
        H.nv = nv;
        H.hDae.assign((size_t)nv * nv, Mat()); H.hPath.assign((size_t)nv * nv, Mat()); H.hLagrange.assign((size_t)nv * nv, Vec());
        SolDae mySolDae = NLPWrapper::MakeSolDae(v, iphase + 1);
        Mat stateOut, pathOut;
        fun_->DaeFunction(mySolDae, stateOut, pathOut);
        // perturbations :1242-1259
        Vec pertTime(N);
        for (int k = 0; k < N; ++k) pertTime[k] = tol * (1 + std::fabs(mySolDae.time_[k]));
        Mat pertState(N, ns), pertControl(N, nc);
        for (size_t e = 0; e < pertState.a.size(); ++e) pertState.a[e] = tol * (1 + std::fabs(mySolDae.state_.a[e]));
        for (size_t e = 0; e < pertControl.a.size(); ++e) pertControl.a[e] = tol * (1 + std::fabs(mySolDae.contol_.a[e]));
        auto pertcol = [&](int var) -> const double* {
            if (var < ns) return &pertState.a[(size_t)var * N];
            if (var < ns + nc) return &pertControl.a[(size_t)(var - ns) * N];
            return pertTime.data();
        };
        auto bumpDae = [&](SolDae& s, int var) {
            const double* p = pertcol(var);
            if (var < ns) for (int k = 0; k < N; ++k) s.state_(k, var) += p[k];
            else if (var < ns + nc) for (int k = 0; k < N; ++k) s.contol_(k, var - ns) += p[k];
            else for (int k = 0; k < N; ++k) s.time_[k] += p[k];
        };
        Mat stateouti, stateoutj, stateoutij, pathouti, pathoutj, pathoutij;
        auto pairDae = [&](const SolDae& isoldae, int a, int b, bool same_as_i) {
            SolDae ijsoldae = isoldae;
            if (same_as_i) { stateoutj = stateouti; pathoutj = pathouti; } // time.time :1403,:1407
            else { SolDae jsoldae = mySolDae; bumpDae(jsoldae, b); fun_->DaeFunction(jsoldae, stateoutj, pathoutj); }
            bumpDae(ijsoldae, b);
            fun_->DaeFunction(ijsoldae, stateoutij, pathoutij);
            const double* pa = pertcol(a); const double* pb = pertcol(b);
            Mat hd(N, ns), hp(N, np);
            for (int s = 0; s < ns; ++s)
                for (int k = 0; k < N; ++k) hd(k, s) = (stateoutij(k, s) - stateouti(k, s) - stateoutj(k, s) + stateOut(k, s)) / (pa[k] * pb[k]);
            for (int s = 0; s < np; ++s)
                for (int k = 0; k < N; ++k) hp(k, s) = (pathoutij(k, s) - pathouti(k, s) - pathoutj(k, s) + pathOut(k, s)) / (pa[k] * pb[k]);
            H.hDae[(size_t)a * nv + b] = hd;
            H.hPath[(size_t)a * nv + b] = hp;
        };
        for (int i = 0; i < ns; ++i) { // state.state :1267-1291
            SolDae isoldae = mySolDae; bumpDae(isoldae, i);
            fun_->DaeFunction(isoldae, stateouti, pathouti);
            for (int j = 0; j < ns; ++j) pairDae(isoldae, i, j, false);
        }
        for (int i = 0; i < nc; ++i) { // control.state, control.control :1305-1346
            SolDae isoldae = mySolDae; bumpDae(isoldae, ns + i);
            fun_->DaeFunction(isoldae, stateouti, pathouti);
            for (int j = 0; j < ns; ++j) pairDae(isoldae, ns + i, j, false);
            for (int j = 0; j < nc; ++j) pairDae(isoldae, ns + i, ns + j, false);
        }
        { // time.* :1358-1411
            SolDae tsoldae = mySolDae; bumpDae(tsoldae, ns + nc);
            fun_->DaeFunction(tsoldae, stateouti, pathouti);
            for (int j = 0; j < ns; ++j) pairDae(tsoldae, ns + nc, j, false);
            for (int j = 0; j < nc; ++j) pairDae(tsoldae, ns + nc, ns + j, false);
            pairDae(tsoldae, ns + nc, ns + nc, true);
        }
        // ---- Lagrange :2008-2101 (same loop structure on SolCost)
        SolCost mysolcost = NLPWrapper::MakeSolCost(v, iphase + 1);
        Vec lagrange, lagrangei, lagrangej, lagrangeij;
        fun_->LagrangeCost(mysolcost, lagrange);
        auto bumpCost = [&](SolCost& s, int var) {
            const double* p = pertcol(var);
            if (var < ns) for (int k = 0; k < N; ++k) s.state_(k, var) += p[k];
            else if (var < ns + nc) for (int k = 0; k < N; ++k) s.control_(k, var - ns) += p[k];
            else for (int k = 0; k < N; ++k) s.time_[k] += p[k];
        };
        auto pairL = [&](const SolCost& isolcost, int a, int b, bool same_as_i) {
            SolCost ijsolcost = isolcost;
            if (same_as_i) lagrangej = lagrangei;
            else { SolCost jsolcost = mysolcost; bumpCost(jsolcost, b); fun_->LagrangeCost(jsolcost, lagrangej); }
            bumpCost(ijsolcost, b);
            fun_->LagrangeCost(ijsolcost, lagrangeij);
            const double* pa = pertcol(a); const double* pb = pertcol(b);
            Vec h(N);
            for (int k = 0; k < N; ++k) h[k] = (lagrangeij[k] - lagrangei[k] - lagrangej[k] + lagrange[k]) / (pa[k] * pb[k]);
            H.hLagrange[(size_t)a * nv + b] = h;
        };
        for (int i = 0; i < ns; ++i) {
            SolCost isolcost = mysolcost; bumpCost(isolcost, i);
            fun_->LagrangeCost(isolcost, lagrangei);
            for (int j = 0; j < ns; ++j) pairL(isolcost, i, j, false);
        }
        for (int i = 0; i < nc; ++i) {
            SolCost isolcost = mysolcost; bumpCost(isolcost, ns + i);
            fun_->LagrangeCost(isolcost, lagrangei);
            for (int j = 0; j < ns; ++j) pairL(isolcost, ns + i, j, false);
            for (int j = 0; j < nc; ++j) pairL(isolcost, ns + i, ns + j, false);
        }
        {
            SolCost tsolcost = mysolcost; bumpCost(tsolcost, ns + nc);
            fun_->LagrangeCost(tsolcost, lagrangei);
            for (int j = 0; j < ns; ++j) pairL(tsolcost, ns + nc, j, false);
            for (int j = 0; j < nc; ++j) pairL(tsolcost, ns + nc, ns + j, false);
            pairL(tsolcost, ns + nc, ns + nc, true);
        }
        // ---- endpoint functions: events :1503-1692, Mayer :1750-1930
        int nE = 2 * ns + 2;
        H.ne_var = nE;
        H.hEvents.assign((size_t)nE * nE, Vec(nevents, 0.0));
        H.hMayer.assign((size_t)nE * nE, 0.0);
        double pert0 = tol * (1 + std::fabs(v.t0)), pertf = tol * (1 + std::fabs(v.tf));
        Vec pertx0(ns), pertxf(ns);
        for (int i = 0; i < ns; ++i) { pertx0[i] = tol * (1 + std::fabs(v.x0[i])); pertxf[i] = tol * (1 + std::fabs(v.xf[i])); }
        auto epert = [&](int e) { return e < ns ? pertx0[e] : (e < 2 * ns ? pertxf[e - ns] : (e == 2 * ns ? pert0 : pertf)); };
        // list of (a, b, denominator) in the reference's evaluation order
        struct Pair { int a, b; double den; };
        std::vector<Pair> pairs;
        for (int i = 0; i < ns; ++i) {
            for (int j = 0; j < ns; ++j) {
                pairs.push_back({i, j, pertx0[i] * pertx0[j]});        // x0.x0 :1580
                pairs.push_back({i, ns + j, pertx0[i] * pertxf[i]});   // x0.xf :1588 (Q10)
            }
            for (int j = 0; j < ns; ++j) {
                pairs.push_back({ns + i, j, pertxf[i] * pertx0[j]});      // xf.x0 :1604
                pairs.push_back({ns + i, ns + j, pertxf[i] * pertxf[i]}); // xf.xf :1612 (Q10)
            }
        }
        for (int j = 0; j < ns; ++j) { pairs.push_back({2 * ns, j, pert0 * pertx0[j]}); pairs.push_back({2 * ns, ns + j, pert0 * pertxf[j]}); }
        pairs.push_back({2 * ns, 2 * ns, pert0 * pert0});
        for (int j = 0; j < ns; ++j) { pairs.push_back({2 * ns + 1, j, pertf * pertx0[j]}); pairs.push_back({2 * ns + 1, ns + j, pertf * pertxf[j]}); }
        pairs.push_back({2 * ns + 1, 2 * ns, pertf * pert0});
        pairs.push_back({2 * ns + 1, 2 * ns + 1, pertf * pertf});
        if (nevents > 0) {
            SolEvent base;
            base.initial_time_ = v.t0; base.initial_state_ = v.x0; base.terminal_time_ = v.tf; base.terminal_state_ = v.xf; base.phase_num_ = iphase + 1;
            auto bump = [&](SolEvent& s, int e) {
                if (e < ns) s.initial_state_[e] += epert(e);
                else if (e < 2 * ns) s.terminal_state_[e - ns] += epert(e);
                else if (e == 2 * ns) s.initial_time_ += pert0;
                else s.terminal_time_ += pertf;
            };
            Vec events(nevents, 0.0), ei, ej, eij;
            fun_->EventFunction(base, events);
            for (auto& p : pairs) {
                SolEvent si = base; bump(si, p.a); fun_->EventFunction(si, ei);
                SolEvent sj = base; bump(sj, p.b); fun_->EventFunction(sj, ej);
                SolEvent sij = si; bump(sij, p.b); fun_->EventFunction(sij, eij);
                Vec h(nevents);
                for (int e = 0; e < nevents; ++e) h[e] = (eij[e] - ei[e] - ej[e] + events[e]) / (p.den * 1.0);
                H.hEvents[(size_t)p.a * nE + p.b] = h;
            }
        }
        {
            auto bump = [&](SolCost& s, int e) {
                if (e < ns) s.initial_state_[e] += epert(e);
                else if (e < 2 * ns) s.terminal_state_[e - ns] += epert(e);
                else if (e == 2 * ns) s.initial_time_ += pert0;
                else s.terminal_time_ += pertf;
            };
            double mayer = 0, mi = 0, mj = 0, mij = 0;
            fun_->MayerCost(mysolcost, mayer);
            for (auto& p : pairs) {
                SolCost si = mysolcost; bump(si, p.a); fun_->MayerCost(si, mi);
                SolCost sj = mysolcost; bump(sj, p.b); fun_->MayerCost(sj, mj);
                SolCost sij = si; bump(sij, p.b); fun_->MayerCost(sij, mij);
                H.hMayer[(size_t)p.a * nE + p.b] = (mij - mi - mj + mayer) / p.den;
            }
        }
    }

    // ---- GetPhaseHessian :12-599 ------------------------------------------------------------
    void GetPhaseHessian(int iphase, double sigma, const Vec& lambada, const Vec& x_all, const Mat& idependencies, Vec& Hessian_V)
    {
        PhaseHessain H;
        CalculatePhaseHessian(iphase, x_all, H);
        NLPWrapper::PhaseVars v = nlp_.Unpack(iphase, x_all);
        int N = v.sumnodes, ns = v.nstates, nc = v.ncontrols, np = v.npaths, nevents = v.nevents;
        int nv = H.nv, nE = H.ne_var;
        double t0 = v.t0, tf = v.tf;
        SolDae mySolDae = NLPWrapper::MakeSolDae(v, iphase + 1);
        SolCost mysolcost = NLPWrapper::MakeSolCost(v, iphase + 1);
        size_t con_index_start = Data_->constraint_indices[iphase].front() - 1;
        Mat diff_lambda(N, ns), path_lambda(N, np);
        Vec event_lambda(nevents);
        size_t ls = con_index_start; // :87-106
        for (int s = 0; s < ns; ++s) for (int k = 0; k < N; ++k) diff_lambda(k, s) = lambada[ls++];
        for (int s = 0; s < np; ++s) for (int k = 0; k < N; ++k) path_lambda(k, s) = lambada[ls++];
        for (int e = 0; e < nevents; ++e) event_lambda[e] = lambada[ls++];
        const Vec& rpm_w = Data_->PS[iphase].Weights;
        const Vec& rpm_tau = Data_->PS[iphase].Points;
        auto rowsum = [&](const Mat& lam, const Mat& h, int k) { // sum(lam % h, 1)
            double acc = 0.0;
            for (int s = 0; s < h.n_cols; ++s) acc += lam(k, s) * h(k, s);
            return acc;
        };
        // (tf-t0)/2*(sigma*w%hL - sum(lam%hdae)) + sum(mu%hpath)  :123-127
        auto core = [&](int a, int b) {
            Vec r(N);
            const Mat& hd = H.hDae[(size_t)a * nv + b];
            const Mat& hp = H.hPath[(size_t)a * nv + b];
            const Vec& hl = H.hLagrange[(size_t)a * nv + b];
            for (int k = 0; k < N; ++k) {
                double sdae = rowsum(diff_lambda, hd, k);
                double spath = np > 0 ? rowsum(path_lambda, hp, k) : 0.0;
                double sL = sigma * rpm_w[k] * hl[k];
                r[k] = (tf - t0) / 2.0 * (sL - sdae) + spath;
            }
            return r;
        };
        Mat dstate, dpath, dLagrange;
        derive_fun_->DerivDae(mySolDae, dstate, dpath);       // :161
        derive_fun_->DerivLagrange(mysolcost, dLagrange);      // :175
        Vec talpha(N), tbeta(N);
        for (int k = 0; k < N; ++k) { talpha[k] = (1 - rpm_tau[k]) / 2.0; tbeta[k] = (1 + rpm_tau[k]) / 2.0; }
        // first-derivative term: sum(lam % reshape(dstate.col(c),N,ns),1) - sigma*w%dL.col(c)
        auto first = [&](int c) {
            Vec r(N);
            for (int k = 0; k < N; ++k) {
                double acc = 0.0;
                for (int s = 0; s < ns; ++s) acc += diff_lambda(k, s) * dstate(s * N + k, c);
                r[k] = acc - sigma * rpm_w[k] * dLagrange(k, c);
            }
            return r;
        };
        int T = ns + nc; // time variable index
        std::vector<Vec> hLI_t0(ns + nc), hLI_tf(ns + nc);
        for (int c = 0; c < ns + nc; ++c) { // :178-206
            Vec B = core(T, c), A = first(c);
            hLI_t0[c].resize(N); hLI_tf[c].resize(N);
            for (int k = 0; k < N; ++k) {
                hLI_t0[c][k] = 0.5 * A[k] + talpha[k] * B[k];
                hLI_tf[c][k] = -0.5 * A[k] + tbeta[k] * B[k];
            }
        }
        double hLI_t0t0 = 0, hLI_tftf = 0, hLI_tft0 = 0;
        { // :210-218
            Vec B = core(T, T), A = first(T);
            double d1 = 0, d2 = 0, d3 = 0, d4 = 0;
            for (int k = 0; k < N; ++k) d1 += talpha[k] * (A[k] + talpha[k] * B[k]);
            for (int k = 0; k < N; ++k) d2 += tbeta[k] * (-A[k] + tbeta[k] * B[k]);
            for (int k = 0; k < N; ++k) d3 += (tbeta[k] - talpha[k]) * A[k];
            for (int k = 0; k < N; ++k) d4 += talpha[k] * (tbeta[k] * B[k]);
            hLI_t0t0 = d1; hLI_tftf = d2; hLI_tft0 = 0.5 * d3 + d4;
        }
        // endpoint part :282-347
        auto hE = [&](int a, int b) {
            double lambda_sum_event = 0.0;
            if (nevents > 0) for (int e = 0; e < nevents; ++e) lambda_sum_event += H.hEvents[(size_t)a * nE + b][e] * event_lambda[e];
            return sigma * H.hMayer[(size_t)a * nE + b] + lambda_sum_event;
        };
        int nDependancies = nnz(idependencies);
        size_t hessainI_nonzeros = (size_t)N * ((nDependancies - ns - nc) / 2 + ns + nc) + 2 * (ns + nc) * N + 1 + 2;
        size_t hessainE_nonzeros = (size_t)(ns + ns) * (ns + ns - 1) / 2 + ns + ns + 2 * 2 * ns + 1 + 2;
        Vec SI(hessainI_nonzeros, 0.0), SE(hessainE_nonzeros, 0.0);
        size_t I = 0, E = 0;
        auto put = [&](const Vec& r) { for (int k = 0; k < N; ++k) SI[I + k] = r[k]; I += N; };
        for (int i = 0; i < ns; ++i) // :409-433
            for (int j = 0; j <= i; ++j) {
                if (idependencies(i, j)) put(core(i, j));
                SE[E++] = hE(i, j);
                if (i != j) SE[E++] = hE(i, ns + j);
                SE[E++] = hE(ns + i, j);
                SE[E++] = hE(ns + i, ns + j);
            }
        for (int i = 0; i < nc; ++i) { // :436-461
            for (int j = 0; j < ns; ++j) if (idependencies(i + ns, j)) put(core(ns + i, j));
            for (int j = 0; j <= i; ++j) if (idependencies(i + ns, j + ns)) put(core(ns + i, ns + j));
        }
        for (int i = 0; i < ns; ++i) { put(hLI_t0[i]); SE[E++] = hE(2 * ns, i); SE[E++] = hE(2 * ns, ns + i); } // :468-479
        for (int i = 0; i < nc; ++i) put(hLI_t0[ns + i]);                                                     // :481-486
        SI[I++] = hLI_t0t0; SE[E++] = hE(2 * ns, 2 * ns);                                                     // :489-493
        for (int i = 0; i < ns; ++i) { put(hLI_tf[i]); SE[E++] = hE(2 * ns + 1, i); SE[E++] = hE(2 * ns + 1, ns + i); } // :500-511
        for (int i = 0; i < nc; ++i) put(hLI_tf[ns + i]);                                                     // :513-518
        SI[I++] = hLI_tft0; SE[E++] = hE(2 * ns + 1, 2 * ns);                                                 // :522-526
        SI[I++] = hLI_tftf; SE[E++] = hE(2 * ns + 1, 2 * ns + 1);                                             // :531-535
        Hessian_V = SI;
        Hessian_V.insert(Hessian_V.end(), SE.begin(), SE.end());
    }

    // ---- GetPhaseHessianSparsity :601-876 ---------------------------------------------------
    void GetPhaseHessianSparsity(int iphase, const Mat& idependencies, Vec& Hessian_I, Vec& Hessian_J)
    {
        int N = optpro_->Phases_[iphase].GetTotalNodes();
        int ns = Data_->SIZES_[iphase][0], nc = Data_->SIZES_[iphase][1];
        Vec II, IJ, EI, EJ;
        auto run = [&](int rs, int cs, bool fillrow) { for (int k = 0; k < N; ++k) { II.push_back(fillrow ? rs : k + rs); IJ.push_back(k + cs); } };
        auto e = [&](int r, int c) { EI.push_back(r); EJ.push_back(c); };
        int rowstart, colstart;
        for (int i = 0; i < ns; ++i) {
            rowstart = i * (N + 1);
            for (int j = 0; j <= i; ++j) {
                colstart = j * (N + 1);
                if (idependencies(i, j)) run(rowstart, colstart, false);
                e(rowstart, colstart);
                if (i != j) e(rowstart, colstart + N);
                e(rowstart + N, colstart);
                e(rowstart + N, colstart + N);
            }
        }
        int rowshift = ns * (N + 1);
        for (int i = 0; i < nc; ++i) {
            rowstart = rowshift + i * N;
            for (int j = 0; j < ns; ++j) { colstart = j * (N + 1); if (idependencies(i + ns, j)) run(rowstart, colstart, false); }
            int colshift = ns * (N + 1);
            for (int j = 0; j <= i; ++j) { colstart = colshift + j * N; if (idependencies(i + ns, j + ns)) run(rowstart, colstart, false); }
        }
        rowstart = ns * (N + 1) + nc * N; // t0 row
        for (int i = 0; i < ns; ++i) { colstart = i * (N + 1); run(rowstart, colstart, true); e(rowstart, colstart); e(rowstart, colstart + N); }
        for (int i = 0; i < nc; ++i) { colstart = i * N + ns * (N + 1); run(rowstart, colstart, true); }
        colstart = ns * (N + 1) + nc * N;
        II.push_back(rowstart); IJ.push_back(colstart); e(rowstart, colstart);
        rowstart++; // tf row
        for (int i = 0; i < ns; ++i) { colstart = i * (N + 1); run(rowstart, colstart, true); e(rowstart, colstart); e(rowstart, colstart + N); }
        for (int i = 0; i < nc; ++i) { colstart = i * N + ns * (N + 1); run(rowstart, colstart, true); }
        colstart = ns * (N + 1) + nc * N;
        II.push_back(rowstart); IJ.push_back(colstart); e(rowstart, colstart);
        colstart = ns * (N + 1) + nc * N + 1;
        II.push_back(rowstart); IJ.push_back(colstart); e(rowstart, colstart);
        int nDependancies = nnz(idependencies);
        size_t hessainI_nonzeros = (size_t)N * ((nDependancies - ns - nc) / 2 + ns + nc) + 2 * (ns + nc) * N + 1 + 2; // :647-648
        size_t hessainE_nonzeros = (size_t)(ns + ns) * (ns + ns - 1) / 2 + ns + ns + 2 * 2 * ns + 1 + 2;             // :652-653
        if (II.size() != hessainI_nonzeros || EI.size() != hessainE_nonzeros) throw LpoError("Hessian pattern count mismatch");
        Hessian_I = II; Hessian_I.insert(Hessian_I.end(), EI.begin(), EI.end());
        Hessian_J = IJ; Hessian_J.insert(Hessian_J.end(), EJ.begin(), EJ.end());
    }

    // ---- CalculateLinkHessain :2163-2367 + GetLinkHessian :1020-1189 (nq = 0) ----------------
    void GetLinkHessian(int ipair, const std::vector<NLPWrapper::PhaseVars>& pv, const Vec& lambada, Vec& out)
    {
        const Linkage& lk = optpro_->Linkage_[ipair];
        int left_index = lk.LeftPhase(), right_index = lk.RightPhase();
        size_t link_start = Data_->link_indices[ipair].front() - 1; // quirk Q6
        size_t link_end = Data_->link_indices[ipair].back() - 1;
        Vec link_lambda(lambada.begin() + link_start, lambada.begin() + link_end + 1);
        SolLink my;
        my.left_state_ = pv[left_index].xf; my.left_phase_num_ = left_index + 1;
        my.right_state_ = pv[right_index].x0; my.right_phase_num_ = right_index + 1;
        my.ipair = ipair + 1;
        Vec Linkout;
        fun_->LinkFunction(my, Linkout);
        int nL = (int)my.left_state_.size(), nR = (int)my.right_state_.size();
        Vec pL(nL), pR(nR);
        for (int i = 0; i < nL; ++i) pL[i] = tol * (1 + std::fabs(my.left_state_[i]));
        for (int i = 0; i < nR; ++i) pR[i] = tol * (1 + std::fabs(my.right_state_[i]));
        // variable index: [0,nL) xf_left, [nL,nL+nR) x0_right
        auto bump = [&](SolLink& s, int var) { if (var < nL) s.left_state_[var] += pL[var]; else s.right_state_[var - nL] += pR[var - nL]; };
        auto pert = [&](int var) { return var < nL ? pL[var] : pR[var - nL]; };
        auto hl = [&](int a, int b) {
            SolLink si = my; bump(si, a);
            Vec li, lj, lij;
            fun_->LinkFunction(si, li);
            SolLink sj = my; bump(sj, b); fun_->LinkFunction(sj, lj);
            SolLink sij = si; bump(sij, b); fun_->LinkFunction(sij, lij);
            double acc = 0.0;
            for (size_t l = 0; l < Linkout.size(); ++l) acc += ((lij[l] - li[l] - lj[l] + Linkout[l]) / (pert(a) * pert(b))) * link_lambda[l];
            return acc;
        };
        out.clear();
        for (int i = 0; i < nL; ++i) for (int j = 0; j <= i; ++j) out.push_back(hl(i, j));          // :1086-1096
        for (int i = 0; i < nR; ++i) {
            for (int j = 0; j < nL; ++j) out.push_back(hl(j, nL + i));                              // hLink_xfL_x0R(jstate, istate) :1131
            for (int j = 0; j <= i; ++j) out.push_back(hl(nL + i, nL + j));                         // :1145
        }
    }

    // ---- GetLinkHessianSparsity :2369-2508 (nq = 0) -----------------------------------------
    void GetLinkHessianSparsity(int ipair, Vec& LI, Vec& LJ)
    {
        const Linkage& lk = optpro_->Linkage_[ipair];
        int left_index = lk.LeftPhase(), right_index = lk.RightPhase();
        int nsL = optpro_->Phases_[left_index].nstates, nsR = optpro_->Phases_[right_index].nstates;
        int nnL = optpro_->Phases_[left_index].GetTotalNodes(), nnR = optpro_->Phases_[right_index].GetTotalNodes();
        int sL = Data_->phase_indices[left_index].state[0] - 1, sR = Data_->phase_indices[right_index].state[0] - 1;
        LI.clear(); LJ.clear();
        for (int i = 0; i < nsL; ++i)
            for (int j = 0; j <= i; ++j) { LI.push_back(sL + (nnL + 1) * (i + 1) - 1); LJ.push_back(sL + (nnL + 1) * (j + 1) - 1); }
        for (int i = 0; i < nsR; ++i) {
            int rowstart = sR + (nnR + 1) * i;
            for (int j = 0; j < nsL; ++j) { LI.push_back(rowstart); LJ.push_back(sL + (nnL + 1) * (j + 1) - 1); }
            for (int j = 0; j <= i; ++j) { LI.push_back(rowstart); LJ.push_back(sR + (nnL + 1) * j); } // quirk Q7: nnodesLeft
        }
    }

    // ---- GetHessian :878-1018 ------------------------------------------------------------------
    void GetHessian(double sigma, const Vec& x, const Vec& lambada, Vec& hessain_V)
    {
        hessain_V.clear();
        for (int i = 0; i < optpro_->GetPhaseNum(); ++i) {
            Mat depH = HessDependencies(Data_->allPhaseDependencies[i]);
            Vec pv;
            GetPhaseHessian(i, sigma, lambada, x, depH, pv);
            hessain_V.insert(hessain_V.end(), pv.begin(), pv.end());
        }
        std::vector<NLPWrapper::PhaseVars> pvars;
        for (int ip = 0; ip < optpro_->GetPhaseNum(); ++ip) pvars.push_back(nlp_.Unpack(ip, x));
        for (int ipair = 0; ipair < optpro_->GetLinkageNum(); ++ipair) {
            Vec lv;
            GetLinkHessian(ipair, pvars, lambada, lv);
            hessain_V.insert(hessain_V.end(), lv.begin(), lv.end());
        }
    }

    // ---- GetHessianSparsity :2510-2599 ---------------------------------------------------------
    void GetHessianSparsity(Vec& hessain_I, Vec& hessain_J)
    {
        hessain_I.clear(); hessain_J.clear();
        size_t rowshift = 0, colshift = 0;
        for (int i = 0; i < optpro_->GetPhaseNum(); ++i) {
            const Phase& ph = optpro_->Phases_[i];
            Mat depH = HessDependencies(Data_->allPhaseDependencies[i]);
            Vec pI, pJ;
            GetPhaseHessianSparsity(i, depH, pI, pJ);
            for (size_t e = 0; e < pI.size(); ++e) { hessain_I.push_back(pI[e] + rowshift); hessain_J.push_back(pJ[e] + colshift); }
            size_t numvars = (size_t)ph.nstates * (ph.GetTotalNodes() + 1) + (size_t)ph.ncontrols * ph.GetTotalNodes() + 2;
            rowshift += numvars; colshift += numvars; // :2583-2586
        }
        for (int ipair = 0; ipair < optpro_->GetLinkageNum(); ++ipair) {
            Vec lI, lJ;
            GetLinkHessianSparsity(ipair, lI, lJ);
            hessain_I.insert(hessain_I.end(), lI.begin(), lI.end());
            hessain_J.insert(hessain_J.end(), lJ.begin(), lJ.end());
        }
    }

private:
    std::shared_ptr<FunctionWrapper> fun_;
    LpCalculateData* Data_;
    OptimalProblem* optpro_;
    std::shared_ptr<OptDerive> derive_fun_;
    double tol;
    NLPWrapper nlp_;
};

// LpDerivDependciesChecker.cpp:10-94: set one state/control to NaN at node index 1
// of the guess, call DaeFunction on that single node, mark non-finite outputs.
inline void GetDependiciesForJacobiInEveryPhase(NLPWrapper& nlp, const Vec& guess, std::vector<Mat>& out)
{
    LpCalculateData* Data_ = nlp.calculateData_;
    out.assign(Data_->numphases_, Mat());
    const double nan = std::numeric_limits<double>::quiet_NaN();
    for (int iphase = 0; iphase < Data_->numphases_; ++iphase) {
        NLPWrapper::PhaseVars v = nlp.Unpack(iphase, guess);
        int ns = v.nstates, nc = v.ncontrols, np = v.npaths;
        if (v.sumnodes < 2) throw LpoError("dependency probe needs at least 2 nodes (quirk Q13)");
        Mat dep(ns + np, nc + ns, 0.0);
        SolDae mydae;
        mydae.time_ = Vec(1, v.t_radau[1]);
        mydae.state_ = Mat(1, ns); mydae.contol_ = Mat(1, nc);
        for (int j = 0; j < ns; ++j) mydae.state_(0, j) = v.state_radau(1, j);
        for (int j = 0; j < nc; ++j) mydae.contol_(0, j) = v.control_radau(1, j);
        mydae.phase_num_ = iphase + 1;
        Mat dae, path;
        auto mark = [&](int col) {
            for (int r = 0; r < ns; ++r) if (!std::isfinite(dae(0, r))) dep(r, col) = 1;
            for (int r = 0; r < np; ++r) if (!std::isfinite(path(0, r))) dep(ns + r, col) = 1;
        };
        for (int is = 0; is < ns; ++is) {
            mydae.state_(0, is) = nan;
            nlp.optimalFunction_->DaeFunction(mydae, dae, path);
            mark(is);
            mydae.state_(0, is) = v.state_radau(1, is);
        }
        for (int ic = 0; ic < nc; ++ic) {
            mydae.contol_(0, ic) = nan;
            nlp.optimalFunction_->DaeFunction(mydae, dae, path);
            mark(ns + ic);
            mydae.contol_(0, ic) = v.control_radau(1, ic);
        }
        out[iphase] = dep;
    }
}

} // namespace lpo
