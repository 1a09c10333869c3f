// oracle/lp_nlp.hpp -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of Lpopc::NLPWrapper: the NLP that IPOPT sees.
// Follows Lpopc/src/Core/LpNLPWrapper.cpp:
//   :34-53    GetAllCons          :55-229    GetConsFun
//   :230-259  GetConsJacbi        :260-523   GetWholeJacbi     :524-862   GetPhaseJacbi
//   :863-939  GetObjFun           :940-1104  GetObjGrad
//   :1106-1312 GetPhaseSparsity   :1314-1548 GetWholeSparsity  :1550-1578 GetConsSparsity
// Reference quirks kept on purpose: Q1 (Jacobian mask forced
// all-ones), Q4 (sign of the d/dtf time term), Q5 (gradient: Mayer d/dx0 overwritten,
// tf time term takes element (0) of an outer product), Q9 (link phase numbers).
// autoscale = no (reference default, LpNLPWrapper.hpp:74); nq = 0 (Q3 fenced).
// Armadillo dot products ((1xN)*(Nx1)) are restated as sequential sums.
#pragma once
#include "lp_derive.hpp"
#include "lp_problem.hpp"

namespace lpo {

class LpHessianCalculator; // lp_hessian.hpp

class NLPWrapper {
public:
    NLPWrapper(std::shared_ptr<FunctionWrapper> funs, LpCalculateData* data, OptimalProblem* optpro, std::shared_ptr<OptDerive> deriver)
        : optimalFunction_(funs), calculateData_(data), optpro_(optpro), derive_(deriver) {}

    static int nnz(const Mat& m)
    {
        int c = 0;
        for (double v : m.a) if (v != 0) ++c;
        return c;
    }

    // common unpacking used by every method (:69-96, :533-560, :880-907, :958-975)
    struct PhaseVars {
        int nstates, ncontrols, npaths, nevents, sumnodes;
        double t0, tf, tspan;
        Vec t_radau, x0, xf;
        Mat state_matrix, state_radau, control_radau;
    };
    PhaseVars Unpack(int i, const Vec& x) const
    {
        PhaseVars v;
        v.nstates = calculateData_->SIZES_[i][0];
        v.ncontrols = calculateData_->SIZES_[i][1];
        v.npaths = calculateData_->SIZES_[i][3];
        v.nevents = calculateData_->SIZES_[i][4];
        v.sumnodes = optpro_->Phases_[i].GetTotalNodes();
        const indices& pi = calculateData_->phase_indices[i];
        int s0 = pi.state.front() - 1, s1 = pi.state.back() - 1;
        Vec state_vector(x.begin() + s0, x.begin() + s1 + 1);
        Vec control_vector;
        if (!pi.control.empty()) control_vector.assign(x.begin() + (pi.control.front() - 1), x.begin() + pi.control.back());
        v.t0 = x[pi.time[0] - 1];
        v.tf = x[pi.time[1] - 1];
        v.tspan = v.tf - v.t0;
        const Vec& pts = calculateData_->PS[i].Points;
        v.t_radau.resize(pts.size());
        for (size_t k = 0; k < pts.size(); ++k) v.t_radau[k] = (pts[k] + 1) * (v.tspan / 2.0) + v.t0;
        v.state_matrix = reshape(state_vector, v.sumnodes + 1, v.nstates);
        v.state_radau = Mat(v.sumnodes, v.nstates);
        for (int j = 0; j < v.nstates; ++j)
            for (int k = 0; k < v.sumnodes; ++k) v.state_radau(k, j) = v.state_matrix(k, j);
        v.x0.resize(v.nstates); v.xf.resize(v.nstates);
        for (int j = 0; j < v.nstates; ++j) { v.x0[j] = v.state_matrix(0, j); v.xf[j] = v.state_matrix(v.sumnodes, j); }
        v.control_radau = reshape(control_vector, v.sumnodes, v.ncontrols);
        return v;
    }
    static SolCost MakeSolCost(const PhaseVars& v, int phase_num)
    {
        SolCost c;
        c.initial_time_ = v.t0; c.initial_state_ = v.x0;
        c.terminal_time_ = v.tf; c.terminal_state_ = v.xf;
        c.time_ = v.t_radau; c.state_ = v.state_radau; c.control_ = v.control_radau;
        c.phase_num_ = phase_num;
        return c;
    }
    static SolDae MakeSolDae(const PhaseVars& v, int phase_num)
    {
        SolDae d;
        d.time_ = v.t_radau; d.state_ = v.state_radau; d.contol_ = v.control_radau; d.phase_num_ = phase_num;
        return d;
    }

    // :34-53
    void GetAllCons(const Vec& x, Vec& Cons)
    {
        Vec nonLinearCons;
        GetConsFun(x, nonLinearCons);
        Mat y((int)x.size(), 1);
        y.a = x;
        Mat lin = calculateData_->AlinearMatrix * y;
        Cons = nonLinearCons;
        Cons.insert(Cons.end(), lin.a.begin(), lin.a.end());
    }

    // :55-229
    void GetConsFun(const Vec& x, Vec& Cons)
    {
        int P = calculateData_->numphases_;
        std::vector<Vec> AllCons(P);
        std::vector<SolCost> solTotal(P);
        for (int i = 0; i < P; ++i) {
            PhaseVars v = Unpack(i, x);
            SolDae mySolDae = MakeSolDae(v, i + 1);
            Mat stateout, pathout;
            optimalFunction_->DaeFunction(mySolDae, stateout, pathout);
            Mat odeleft = calculateData_->PS[i].D * v.state_matrix; // :111
            Mat defects(v.sumnodes, v.nstates);
            for (size_t e = 0; e < defects.a.size(); ++e) defects.a[e] = odeleft.a[e] - stateout.a[e] * (v.tspan / 2.0); // :113,:122
            Vec events;
            if (v.nevents > 0) {
                SolEvent ev;
                ev.initial_time_ = v.t0; ev.initial_state_ = v.x0; ev.terminal_time_ = v.tf; ev.terminal_state_ = v.xf; ev.phase_num_ = i + 1;
                events.assign(v.nevents, 0.0);
                optimalFunction_->EventFunction(ev, events);
            }
            Vec consi = defects.a; // :138-164: [defects(:) | pathout(:) | events]
            if (v.npaths > 0) consi.insert(consi.end(), pathout.a.begin(), pathout.a.end());
            consi.insert(consi.end(), events.begin(), events.end());
            AllCons[i] = consi;
            solTotal[i] = MakeSolCost(v, i + 1);
        }
        std::vector<Vec> AllLink(optpro_->GetLinkageNum());
        if (calculateData_->numlinks_ > 0) {
            for (int ipair = 0; ipair < optpro_->GetLinkageNum(); ++ipair) { // :184-209
                int left_index = optpro_->Linkage_[ipair].LeftPhase();
                int right_index = optpro_->Linkage_[ipair].RightPhase();
                SolLink l;
                l.left_state_ = solTotal[left_index].terminal_state_;
                l.left_phase_num_ = left_index; // quirk Q9: 0-based here
                l.right_state_ = solTotal[right_index].initial_state_;
                l.right_phase_num_ = right_index;
                Vec linkout(l.right_state_.size());
                optimalFunction_->LinkFunction(l, linkout);
                AllLink[ipair] = linkout;
            }
        }
        Cons.clear();
        for (auto& c : AllCons) Cons.insert(Cons.end(), c.begin(), c.end());
        for (auto& c : AllLink) Cons.insert(Cons.end(), c.begin(), c.end());
    }

    // :863-939
    double GetObjFun(const Vec& x)
    {
        double cost = 0.0;
        for (int i = 0; i < calculateData_->numphases_; ++i) {
            PhaseVars v = Unpack(i, x);
            SolCost c = MakeSolCost(v, i + 1);
            double temMayer = 0.0;
            Vec temLagrange;
            optimalFunction_->MayerCost(c, temMayer);
            cost += temMayer;
            optimalFunction_->LagrangeCost(c, temLagrange);
            const Vec& w = calculateData_->PS[i].Weights;
            double dot = 0.0;
            for (size_t k = 0; k < w.size(); ++k) dot += w[k] * temLagrange[k];
            cost += dot * (v.tspan / 2.0); // :931-932
        }
        return cost;
    }

    // :940-1104
    void GetObjGrad(const Vec& x, Vec& grad_f)
    {
        grad_f.assign(calculateData_->varbounds_max.size(), 0.0);
        int grad_shift = 0;
        for (int iphase = 0; iphase < optpro_->GetPhaseNum(); ++iphase) {
            PhaseVars v = Unpack(iphase, x);
            int sumnodes = v.sumnodes, nstates = v.nstates, ncontrols = v.ncontrols;
            SolCost c = MakeSolCost(v, iphase + 1);
            Vec dmayerOut, LagrangeOut;
            Mat dLagrangeOut;
            optimalFunction_->LagrangeCost(c, LagrangeOut);
            derive_->DerivMayer(c, dmayerOut);
            derive_->DerivLagrange(c, dLagrangeOut);
            double dMayer_t0 = 0.0, dMayer_tf = 0.0;
            Vec dMayer_x0(nstates, 0.0), dMayer_xf(nstates, 0.0);
            Mat dLagrange_state(sumnodes, nstates, 0.0), dLagrange_control(sumnodes, ncontrols, 0.0);
            Vec dLagrange_time(sumnodes, 0.0);
            if (!dmayerOut.empty()) { // :1007-1021
                dMayer_t0 = dmayerOut[nstates];
                dMayer_tf = dmayerOut[2 * nstates + 1];
                for (int is = 0; is < nstates; ++is) { dMayer_x0[is] = dmayerOut[is]; dMayer_xf[is] = dmayerOut[is + nstates + 1]; }
            }
            if (dLagrangeOut.n_elem() > 0) { // :1023-1039
                dLagrange_time = dLagrangeOut.col(dLagrangeOut.n_cols - 1);
                for (int is = 0; is < nstates; ++is) dLagrange_state.set_col(is, dLagrangeOut.col(is));
                for (int ic = 0; ic < ncontrols; ++ic) dLagrange_control.set_col(ic, dLagrangeOut.col(ic + nstates));
            }
            const Vec& W = calculateData_->PS[iphase].Weights;
            const Vec& Pt = calculateData_->PS[iphase].Points;
            double tspan = v.tspan;
            Vec Jcost(nstates * (sumnodes + 1) + ncontrols * sumnodes + 2, 0.0);
            for (int j = 0; j < nstates; ++j) { // :1045-1055
                int col0 = sumnodes * j + j, colf = sumnodes * (j + 1) + j;
                Jcost[col0] = dMayer_x0[j]; // overwritten just below (quirk Q5)
                for (int k = 0; k < sumnodes; ++k) Jcost[col0 + k] = (W[k] * tspan / 2.0) * dLagrange_state(k, j);
                Jcost[colf] = dMayer_xf[j];
            }
            int colshift = nstates * (sumnodes + 1);
            for (int j = 0; j < ncontrols; ++j) { // :1058-1064
                int colstart = colshift + j * sumnodes;
                for (int k = 0; k < sumnodes; ++k) Jcost[colstart + k] = (W[k] * tspan / 2.0) * dLagrange_control(k, j);
            }
            colshift += ncontrols * sumnodes;
            int t0_col = colshift;
            { // dCost/dt0 :1069-1078
                double ret = 0.0; // trans(W*(-0.5))*LagrangeOut
                for (int k = 0; k < sumnodes; ++k) ret += (W[k] * (-0.5)) * LagrangeOut[k];
                double dot = 0.0; // (trans(W*(tspan/2))*diagmat(dL_time)) * (Points*(-0.5)+0.5)
                for (int k = 0; k < sumnodes; ++k) dot += ((W[k] * (tspan / 2.0)) * dLagrange_time[k]) * (Pt[k] * (-0.5) + 0.5);
                Jcost[t0_col] = dot + dMayer_t0 + ret;
            }
            int tf_col = colshift + 1;
            { // dCost/dtf :1081-1087; `ret3 *= ret2` is (Nx1)*(1xN): element (0,0) only (quirk Q5)
                double ret = 0.0;
                for (int k = 0; k < sumnodes; ++k) ret += (W[k] * 0.5) * LagrangeOut[k];
                double ret3_0 = (Pt[0] * 0.5 + 0.5) * ((W[0] * (tspan / 2.0)) * dLagrange_time[0]);
                Jcost[tf_col] = dMayer_tf + ret + ret3_0;
            }
            for (size_t e = 0; e < Jcost.size(); ++e) grad_f[grad_shift + e] = Jcost[e];
            grad_shift += (int)Jcost.size();
        }
    }

    // :524-862
    void GetPhaseJacbi(int iphase, const Vec& x_all, Mat& idependencies, Vec& Sjac_V, Vec& Sconstant_V)
    {
        PhaseVars v = Unpack(iphase, x_all);
        int sumnodes = v.sumnodes, nstates = v.nstates, ncontrols = v.ncontrols, npaths = v.npaths, nevents = v.nevents;
        double t0 = v.t0, tf = v.tf;
        SolDae mySolDae = MakeSolDae(v, iphase + 1);
        Mat dDaeOut, dPathOut;
        derive_->DerivDae(mySolDae, dDaeOut, dPathOut); // :569
        Mat daeOut, pathOut;
        optimalFunction_->DaeFunction(mySolDae, daeOut, pathOut); // :572
        std::vector<Mat> dDae_state(nstates), dDae_control(ncontrols), dPath_state(nstates), dPath_control(ncontrols);
        for (int is = 0; is < nstates; ++is) { // :584-593
            dDae_state[is] = reshape(dDaeOut.col(is), sumnodes, nstates);
            if (npaths > 0) dPath_state[is] = reshape(dPathOut.col(is), sumnodes, npaths);
        }
        for (int ic = 0; ic < ncontrols; ++ic) { // :595-604
            dDae_control[ic] = reshape(dDaeOut.col(nstates + ic), sumnodes, nstates);
            if (npaths > 0) dPath_control[ic] = reshape(dPathOut.col(nstates + ic), sumnodes, npaths);
        }
        Mat dDae_time = reshape(dDaeOut.col(nstates + ncontrols), sumnodes, nstates); // :606
        Mat dPath_time;
        if (npaths > 0) dPath_time = reshape(dPathOut.col(nstates + ncontrols), sumnodes, npaths);

        Vec dEvent_t0, dEvent_tf;
        std::vector<Vec> dEvent_x0(nstates), dEvent_xf(nstates);
        if (nevents > 0) { // :638-669
            SolEvent ev;
            ev.initial_time_ = t0; ev.terminal_time_ = tf; ev.initial_state_ = v.x0; ev.terminal_state_ = v.xf; ev.phase_num_ = iphase + 1;
            Mat dEventOut;
            derive_->DerivEvent(ev, dEventOut);
            dEvent_t0 = dEventOut.col(nstates);
            dEvent_tf = dEventOut.col(2 * nstates + 1);
            for (int is = 0; is < nstates; ++is) { dEvent_x0[is] = dEventOut.col(is); dEvent_xf[is] = dEventOut.col(is + nstates + 1); }
        }
        int ndiffeqs = nstates;
        for (int i = 0; i < ndiffeqs; ++i) idependencies(i, i) = 1.0;
        int nDependancies = nnz(idependencies);
        Sjac_V.assign((size_t)nDependancies * sumnodes + 2 * (ndiffeqs + npaths) * sumnodes + nevents * (2 * nstates + 2), 0.0);
        Vec DI, DJ, DV;
        dsmatrix::Find(calculateData_->PS[iphase].Doffdiag, DI, DJ, DV); // :686
        int nonZerosDiffMat = (int)DI.size();
        Sconstant_V.assign((size_t)ndiffeqs * nonZerosDiffMat, 0.0);
        const Vec& Points = calculateData_->PS[iphase].Points;
        const std::vector<double>& PsDiag = calculateData_->PS[iphase].Diag.vals; // first N raw COO values (:704-711)
        int S = 0;
        for (int i = 0; i < ndiffeqs; ++i) {
            for (int j = 0; j < nstates; ++j) {
                if (i == j) { // :701-719
                    for (int k = 0; k < sumnodes; ++k) Sjac_V[S + k] = PsDiag[k] - dDae_state[j](k, i) * (tf - t0) / 2.0;
                    S += sumnodes;
                    for (int e = 0; e < nonZerosDiffMat; ++e) Sconstant_V[(size_t)i * nonZerosDiffMat + e] = DV[e];
                } else if (idependencies(i, j) == 1.0) { // :722-728
                    for (int k = 0; k < sumnodes; ++k) Sjac_V[S + k] = -(dDae_state[j](k, i) * (tf - t0) / 2.0);
                    S += sumnodes;
                }
            }
            for (int j = 0; j < ncontrols; ++j) { // :733-743
                if (idependencies(i, j + nstates) == 1.0) {
                    for (int k = 0; k < sumnodes; ++k) Sjac_V[S + k] = -(dDae_control[j](k, i) * (tf - t0) / 2.0);
                    S += sumnodes;
                }
            }
            for (int k = 0; k < sumnodes; ++k) { // d/dt0 :748-752
                double ret = daeOut(k, i) * (0.5);
                double ret2 = -(Points[k] * 0.5) + 0.5;
                ret -= ret2 * (dDae_time(k, i) * (tf - t0) / 2.0);
                Sjac_V[S + k] = ret;
            }
            S += sumnodes;
            for (int k = 0; k < sumnodes; ++k) { // d/dtf :756-760 (sign quirk Q4)
                double ret = -daeOut(k, i) * (0.5);
                double ret2 = (Points[k] * 0.5) + 0.5;
                ret = ret + ret2 * (dDae_time(k, i) * (tf - t0) / 2.0);
                Sjac_V[S + k] = ret;
            }
            S += sumnodes;
        }
        for (int i = 0; i < npaths; ++i) { // :773-820
            for (int j = 0; j < nstates; ++j)
                if (idependencies(i + ndiffeqs, j) == 1.0) {
                    for (int k = 0; k < sumnodes; ++k) Sjac_V[S + k] = dPath_state[j](k, i);
                    S += sumnodes;
                }
            for (int j = 0; j < ncontrols; ++j)
                if (idependencies(i + ndiffeqs, j + nstates)) {
                    for (int k = 0; k < sumnodes; ++k) Sjac_V[S + k] = dPath_control[j](k, i);
                    S += sumnodes;
                }
            for (int k = 0; k < sumnodes; ++k) Sjac_V[S + k] = (-(Points[k] * 0.5) + 0.5) * dPath_time(k, i);
            S += sumnodes;
            for (int k = 0; k < sumnodes; ++k) Sjac_V[S + k] = ((Points[k] * 0.5) + 0.5) * dPath_time(k, i);
            S += sumnodes;
        }
        for (int i = 0; i < nevents; ++i) { // :833-861
            for (int j = 0; j < nstates; ++j) {
                Sjac_V[S++] = dEvent_x0[j][i];
                Sjac_V[S++] = dEvent_xf[j][i];
            }
            Sjac_V[S++] = dEvent_t0[i];
            Sjac_V[S++] = dEvent_tf[i];
        }
    }

    // sizes shared by GetWholeJacbi (:263-323) and GetWholeSparsity (:1317-1379)
    void CountWhole(std::vector<Mat>& dependencies, int& nonZerosSjac, int& nonZerosSconstant)
    {
        int P = optpro_->GetPhaseNum();
        dependencies.assign(P, Mat());
        nonZerosSjac = 0; nonZerosSconstant = 0;
        for (int i = 0; i < P; ++i) {
            const Phase& ph = optpro_->Phases_[i];
            int ns = (int)ph.statemin.size(), nc = (int)ph.controlmin.size(), np = (int)ph.pathmin.size(), ne = (int)ph.eventmin.size();
            int nn = ph.GetTotalNodes();
            dependencies[i] = Mat(ns + np, ns + nc, 0.0);
            for (double& d : dependencies[i].a) d = 1.0; // quirk Q1: "only when use Analytical Derives" but unconditional (:291,:1345)
            int nDependancies = nnz(dependencies[i]);
            nonZerosSjac += nDependancies * nn + 2 * (ns + np) * nn + ne * (2 * ns + 2);
            Vec I, J, V;
            dsmatrix::Find(calculateData_->PS[i].Doffdiag, I, J, V);
            nonZerosSconstant += (int)I.size() * ns;
        }
        for (int ipair = 0; ipair < optpro_->GetLinkageNum(); ++ipair) { // :308-320 (right sizes read from the LEFT phase, quirk Q7)
            const Linkage& lk = optpro_->Linkage_[ipair];
            int nsl = optpro_->Phases_[lk.LeftPhase()].nstates;
            int numlinks = (int)lk.linkmin.size();
            nonZerosSjac += numlinks * (nsl + nsl);
        }
    }

    // :260-523
    void GetWholeJacbi(const Vec& x, Vec& Sjac_V, Vec& Sconstant_V)
    {
        std::vector<Mat> dependencies;
        int nonZerosSjac, nonZerosSconstant;
        CountWhole(dependencies, nonZerosSjac, nonZerosSconstant);
        Sjac_V.assign(nonZerosSjac, 0.0);
        Sconstant_V.assign(nonZerosSconstant, 0.0);
        int Sjac_rowsShift = 0, Sconstant_rowsShift = 0;
        for (int i = 0; i < optpro_->GetPhaseNum(); ++i) {
            Vec pv, pc;
            GetPhaseJacbi(i, x, dependencies[i], pv, pc);
            for (size_t e = 0; e < pv.size(); ++e) Sjac_V[Sjac_rowsShift + e] = pv[e];
            Sjac_rowsShift += (int)pv.size();
            for (size_t e = 0; e < pc.size(); ++e) Sconstant_V[Sconstant_rowsShift + e] = pc[e];
            Sconstant_rowsShift += (int)pc.size();
        }
        std::vector<PhaseVars> pvars;
        for (int ip = 0; ip < optpro_->GetPhaseNum(); ++ip) pvars.push_back(Unpack(ip, x));
        for (int ipair = 0; ipair < optpro_->GetLinkageNum(); ++ipair) { // :406-522
            int left_index = optpro_->Linkage_[ipair].LeftPhase();
            int right_index = optpro_->Linkage_[ipair].RightPhase();
            SolLink l;
            l.left_state_ = pvars[left_index].xf;
            l.left_phase_num_ = left_index + 1;
            l.right_state_ = pvars[right_index].x0;
            l.right_phase_num_ = right_index + 1;
            l.ipair = ipair + 1;
            Mat dLinkOut;
            derive_->DerivLink(l, dLinkOut);
            int nsl = optpro_->Phases_[left_index].nstates, nsr = optpro_->Phases_[right_index].nstates;
            for (int jcol = 0; jcol < nsl; ++jcol) // DLink_xf_left, column-major (:461-471)
                for (int irow = 0; irow < dLinkOut.n_rows; ++irow) Sjac_V[Sjac_rowsShift++] = dLinkOut(irow, jcol);
            for (int jcol = 0; jcol < nsr; ++jcol) // DLink_x0_Right (:492-501)
                for (int irow = 0; irow < dLinkOut.n_rows; ++irow) Sjac_V[Sjac_rowsShift++] = dLinkOut(irow, nsl + jcol);
        }
    }

    // :230-259: [NL | Alinear values | constant D part]
    void GetConsJacbi(const Vec& x, Vec& Sjac_V)
    {
        Vec NL_V, L_I, L_J, L_V, C_V;
        GetWholeJacbi(x, NL_V, C_V);
        dsmatrix::Find(calculateData_->AlinearMatrix, L_I, L_J, L_V);
        Sjac_V = NL_V;
        Sjac_V.insert(Sjac_V.end(), L_V.begin(), L_V.end());
        Sjac_V.insert(Sjac_V.end(), C_V.begin(), C_V.end());
    }

    // :1106-1312
    void GetPhaseSparsity(int iphase, Mat& idependencies, Vec& Sjac_I, Vec& Sjac_J, Vec& Sconstant_I, Vec& Sconstant_J)
    {
        const Phase& ph = optpro_->Phases_[iphase];
        int sumnodes = ph.GetTotalNodes();
        int nstates = (int)ph.statemin.size(), ncontrols = (int)ph.controlmin.size();
        int npaths = (int)ph.pathmin.size(), nevents = (int)ph.eventmin.size();
        int ndiffeqs = nstates, disc_pts = sumnodes + 1;
        for (int i = 0; i < ndiffeqs; ++i) idependencies(i, i) = 1.0;
        int nDependancies = nnz(idependencies);
        size_t total = (size_t)nDependancies * sumnodes + 2 * (ndiffeqs + npaths) * sumnodes + nevents * (2 * nstates + 2);
        Sjac_I.assign(total, 0.0); Sjac_J.assign(total, 0.0);
        Vec DI, DJ, DV;
        dsmatrix::Find(calculateData_->PS[iphase].Doffdiag, DI, DJ, DV);
        int nonZerosDiffMat = (int)DI.size();
        Sconstant_I.assign((size_t)ndiffeqs * nonZerosDiffMat, 0.0);
        Sconstant_J.assign((size_t)ndiffeqs * nonZerosDiffMat, 0.0);
        int S = 0, rowstart, colstart;
        auto run = [&](int rs, int cs, bool fill) { // indexvector + rowstart / + colstart | fill(colstart)
            for (int k = 0; k < sumnodes; ++k) { Sjac_I[S + k] = k + rs; Sjac_J[S + k] = fill ? cs : k + cs; }
            S += sumnodes;
        };
        for (int i = 0; i < ndiffeqs; ++i) {
            rowstart = i * sumnodes;
            for (int j = 0; j < nstates; ++j) {
                colstart = j * disc_pts;
                if (i == j) {
                    run(rowstart, colstart, false);
                    for (int e = 0; e < nonZerosDiffMat; ++e) {
                        Sconstant_I[(size_t)i * nonZerosDiffMat + e] = DI[e] + rowstart;
                        Sconstant_J[(size_t)i * nonZerosDiffMat + e] = DJ[e] + colstart;
                    }
                } else if (idependencies(i, j) == 1.0) {
                    run(rowstart, colstart, false);
                }
            }
            int colshift = nstates * disc_pts;
            for (int j = 0; j < ncontrols; ++j) {
                colstart = colshift + j * sumnodes;
                if (idependencies(i, j + nstates) == 1.0) run(rowstart, colstart, false);
            }
            colshift += ncontrols * sumnodes;
            run(rowstart, colshift, true); // t0 :1196-1198
            colshift++;
            run(rowstart, colshift, true); // tf :1202-1204
        }
        int rowshift = ndiffeqs * sumnodes;
        for (int i = 0; i < npaths; ++i) {
            rowstart = rowshift + i * sumnodes;
            for (int j = 0; j < nstates; ++j) {
                colstart = j * disc_pts;
                if (idependencies(i + ndiffeqs, j) == 1.0) run(rowstart, colstart, false);
            }
            int colshift = nstates * disc_pts;
            for (int j = 0; j < ncontrols; ++j) {
                colstart = colshift + j * sumnodes;
                if (idependencies(i + ndiffeqs, j + nstates)) run(rowstart, colstart, false);
            }
            colshift += ncontrols * sumnodes;
            run(rowstart, colshift, true);
            colshift++;
            run(rowstart, colshift, true);
        }
        rowshift = ndiffeqs * sumnodes + npaths * sumnodes;
        for (int i = 0; i < nevents; ++i) { // :1278-1311
            int row = rowshift + i;
            for (int j = 0; j < nstates; ++j) {
                int col0 = sumnodes * j + j, colf = sumnodes * (j + 1) + j;
                Sjac_I[S] = row; Sjac_J[S] = col0; S++;
                Sjac_I[S] = row; Sjac_J[S] = colf; S++;
            }
            int cols = nstates * (sumnodes + 1) + ncontrols * sumnodes;
            Sjac_I[S] = row; Sjac_J[S] = cols; S++;
            cols++;
            Sjac_I[S] = row; Sjac_J[S] = cols; S++;
        }
    }

    // :1314-1548
    void GetWholeSparsity(Vec& Sjac_I, Vec& Sjac_J, Vec& Sconstant_I, Vec& Sconstant_J)
    {
        std::vector<Mat> dependencies;
        int nonZerosSjac, nonZerosSconstant;
        CountWhole(dependencies, nonZerosSjac, nonZerosSconstant);
        Sjac_I.assign(nonZerosSjac, 0.0); Sjac_J.assign(nonZerosSjac, 0.0);
        Sconstant_I.assign(nonZerosSconstant, 0.0); Sconstant_J.assign(nonZerosSconstant, 0.0);
        int Sjac_rowsShift = 0, Sconstant_rowsShift = 0, rowshift = 0, colshift = 0;
        for (int i = 0; i < optpro_->GetPhaseNum(); ++i) {
            const Phase& ph = optpro_->Phases_[i];
            Vec pI, pJ, cI, cJ;
            GetPhaseSparsity(i, dependencies[i], pI, pJ, cI, cJ);
            for (size_t e = 0; e < pI.size(); ++e) { Sjac_I[Sjac_rowsShift + e] = pI[e] + rowshift; Sjac_J[Sjac_rowsShift + e] = pJ[e] + colshift; }
            Sjac_rowsShift += (int)pI.size();
            for (size_t e = 0; e < cI.size(); ++e) { Sconstant_I[Sconstant_rowsShift + e] = cI[e] + rowshift; Sconstant_J[Sconstant_rowsShift + e] = cJ[e] + colshift; }
            Sconstant_rowsShift += (int)cI.size();
            int nn = ph.GetTotalNodes();
            int numcons = ph.nstates * nn + ph.npaths * nn + ph.nevents;
            int numvars = ph.nstates * (nn + 1) + ph.ncontrols * nn + 2;
            rowshift += numcons;
            colshift += numvars;
        }
        int linkrow = rowshift;
        for (int ipair = 0; ipair < optpro_->GetLinkageNum(); ++ipair) { // :1431-1547
            const Linkage& lk = optpro_->Linkage_[ipair];
            int nlinks = (int)lk.linkmin.size();
            int left_index = lk.LeftPhase(), right_index = lk.RightPhase();
            int nstatesLeft = optpro_->Phases_[left_index].nstates, nstatesRight = optpro_->Phases_[right_index].nstates;
            int nnodesLeft = optpro_->Phases_[left_index].GetTotalNodes(), nnodesRight = optpro_->Phases_[right_index].GetTotalNodes();
            int stateindexstart = calculateData_->phase_indices[left_index].state[0] - 1;
            for (int jcol = 0; jcol < nstatesLeft; ++jcol)
                for (int irow = 0; irow < nlinks; ++irow) {
                    Sjac_I[Sjac_rowsShift] = irow + linkrow;
                    Sjac_J[Sjac_rowsShift] = (jcol + 1) * nnodesLeft + jcol + stateindexstart;
                    Sjac_rowsShift++;
                }
            stateindexstart = calculateData_->phase_indices[right_index].state[0] - 1;
            for (int jcol = 0; jcol < nstatesRight; ++jcol)
                for (int irow = 0; irow < nlinks; ++irow) {
                    Sjac_I[Sjac_rowsShift] = irow + linkrow;
                    Sjac_J[Sjac_rowsShift] = jcol * (nnodesRight + 1) + stateindexstart;
                    Sjac_rowsShift++;
                }
            linkrow += nlinks;
        }
    }

    // :1550-1578 (the function-static cache is dropped: the oracle is rebuilt per mesh)
    void GetConsSparsity(Vec& Sjac_I, Vec& Sjac_J)
    {
        Vec NL_I, NL_J, L_I, L_J, L_V, C_I, C_J;
        GetWholeSparsity(NL_I, NL_J, C_I, C_J);
        dsmatrix::Find(calculateData_->AlinearMatrix, L_I, L_J, L_V);
        int linearConsRowStart = (int)calculateData_->conbounds_min.size();
        Sjac_I = NL_I; Sjac_J = NL_J;
        for (size_t e = 0; e < L_I.size(); ++e) { Sjac_I.push_back(L_I[e] + linearConsRowStart); Sjac_J.push_back(L_J[e]); }
        Sjac_I.insert(Sjac_I.end(), C_I.begin(), C_I.end());
        Sjac_J.insert(Sjac_J.end(), C_J.begin(), C_J.end());
    }

    std::shared_ptr<FunctionWrapper> optimalFunction_;
    LpCalculateData* calculateData_;
    OptimalProblem* optpro_;
    std::shared_ptr<OptDerive> derive_;
};

} // namespace lpo
