// oracle/lp_derive.hpp -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of lpopc's first-derivative providers.
// Follows:
//   Lpopc/src/Core/LpOptDerive.hpp:13-44                    (OptDerive interface)
//   Lpopc/src/Core/LpFiniteDifferenceDerive.cpp:11-98       (DerivMayer)
//                                              :100-192     (DerivLagrange)
//                                              :194-324     (DerivDae)
//                                              :326-409     (DerivEvent)
//                                              :411-506     (DerivLink)
//   Lpopc/src/Core/LpAnalyticDerive.hpp:15-52               (pass-through to user Deriv*)
// Scheme: forward differences, h = tol*(1+|v|) per element, ONE whole variable
// column perturbed at all nodes per user call.  nq = 0 (quirk Q3 fenced).
#pragma once
#include "lp_types.hpp"
#include <memory>

namespace lpo {

class OptDerive {
public:
    virtual ~OptDerive() {}
    virtual void DerivMayer(SolCost&, Vec&) = 0;
    virtual void DerivLagrange(SolCost&, Mat&) = 0;
    virtual void DerivDae(SolDae&, Mat&, Mat&) = 0;
    virtual void DerivEvent(SolEvent&, Mat&) = 0;
    virtual void DerivLink(SolLink&, Mat&) = 0;
};

class LpAnalyticDerive : public OptDerive {
public:
    explicit LpAnalyticDerive(std::shared_ptr<FunctionWrapper> f) : fun_(f) {}
    void DerivMayer(SolCost& s, Vec& d) override { fun_->DerivMayer(s, d); }
    void DerivLagrange(SolCost& s, Mat& d) override { fun_->DerivLagrange(s, d); }
    void DerivDae(SolDae& s, Mat& a, Mat& b) override { fun_->DerivDae(s, a, b); }
    void DerivEvent(SolEvent& s, Mat& d) override { fun_->DerivEvent(s, d); }
    void DerivLink(SolLink& s, Mat& d) override { fun_->DerivLink(s, d); }
private:
    std::shared_ptr<FunctionWrapper> fun_;
};

class LpFDderive : public OptDerive {
public:
    LpFDderive(std::shared_ptr<FunctionWrapper> f, double tol) : fun_(f), tol_(tol) {}

    // LpFiniteDifferenceDerive.cpp:11-98; layout [x0 | t0 | xf | tf]
    void DerivMayer(SolCost& s, Vec& deriv_mayer) override
    {
        size_t nstate = s.state_.n_cols;
        double pert0 = tol_ * (1 + std::fabs(s.initial_time_));
        double pertf = tol_ * (1 + std::fabs(s.terminal_time_));
        Vec pertx0(nstate), pertxf(nstate), x0Pert(nstate), xfPert(nstate);
        for (size_t i = 0; i < nstate; ++i) {
            pertx0[i] = tol_ * (1 + std::fabs(s.initial_state_[i]));
            pertxf[i] = tol_ * (1 + std::fabs(s.terminal_state_[i]));
            x0Pert[i] = s.initial_state_[i] + pertx0[i];
            xfPert[i] = s.terminal_state_[i] + pertxf[i];
        }
        double t0Pert = s.initial_time_ + pert0, tfPert = s.terminal_time_ + pertf;
        double mayerout = 0, perMayerout = 0;
        fun_->MayerCost(s, mayerout);
        double t0 = s.initial_time_;
        s.initial_time_ = t0Pert;
        fun_->MayerCost(s, perMayerout);
        double DMayer_t0 = (perMayerout - mayerout) / pert0;
        s.initial_time_ = t0;
        double tf = s.terminal_time_;
        s.terminal_time_ = tfPert;
        fun_->MayerCost(s, perMayerout);
        double DMayer_tf = (perMayerout - mayerout) / pertf;
        s.terminal_time_ = tf;
        Vec x0 = s.initial_state_, xf = s.terminal_state_;
        Vec DMayer_x0(nstate, 0.0), DMayer_xf(nstate, 0.0);
        for (size_t i = 0; i < nstate; ++i) {
            s.initial_state_[i] = x0Pert[i];
            fun_->MayerCost(s, perMayerout);
            DMayer_x0[i] = (perMayerout - mayerout) / pertx0[i];
            s.initial_state_[i] = x0[i];
            s.terminal_state_[i] = xfPert[i];
            fun_->MayerCost(s, perMayerout);
            DMayer_xf[i] = (perMayerout - mayerout) / pertxf[i];
            s.terminal_state_[i] = xf[i];
        }
        deriv_mayer.assign(nstate + 1 + nstate + 1, 0.0);
        for (size_t i = 0; i < nstate; ++i) deriv_mayer[i] = DMayer_x0[i];
        deriv_mayer[nstate] = DMayer_t0;
        for (size_t i = 0; i < nstate; ++i) deriv_mayer[nstate + 1 + i] = DMayer_xf[i];
        deriv_mayer[2 * nstate + 1] = DMayer_tf;
    }

    // LpFiniteDifferenceDerive.cpp:100-192; columns [dL/dx | dL/du | dL/dt]
    void DerivLagrange(SolCost& s, Mat& deriv_langrange) override
    {
        size_t nstate = s.state_.n_cols, ncontrols = s.control_.n_cols, nnodes = s.state_.n_rows;
        Vec pertTime(nnodes), tRadauPert(nnodes);
        for (size_t k = 0; k < nnodes; ++k) { pertTime[k] = tol_ * (1 + std::fabs(s.time_[k])); tRadauPert[k] = s.time_[k] + pertTime[k]; }
        Mat pertState((int)nnodes, (int)nstate), stateRadauPert((int)nnodes, (int)nstate);
        for (size_t i = 0; i < pertState.a.size(); ++i) { pertState.a[i] = tol_ * (1 + std::fabs(s.state_.a[i])); stateRadauPert.a[i] = s.state_.a[i] + pertState.a[i]; }
        Mat pertControl((int)nnodes, (int)ncontrols), controlRadauPert((int)nnodes, (int)ncontrols);
        for (size_t i = 0; i < pertControl.a.size(); ++i) { pertControl.a[i] = tol_ * (1 + std::fabs(s.control_.a[i])); controlRadauPert.a[i] = s.control_.a[i] + pertControl.a[i]; }
        Vec lagrangeOut, perlagrangeOut;
        fun_->LagrangeCost(s, lagrangeOut);
        deriv_langrange = Mat((int)nnodes, (int)(nstate + ncontrols + 1), 0.0);
        Vec t_radau = s.time_;
        s.time_ = tRadauPert;
        fun_->LagrangeCost(s, perlagrangeOut);
        for (size_t k = 0; k < nnodes; ++k) deriv_langrange((int)k, (int)(nstate + ncontrols)) = (perlagrangeOut[k] - lagrangeOut[k]) / pertTime[k];
        s.time_ = t_radau;
        Mat state_radau = s.state_;
        for (size_t is = 0; is < nstate; ++is) {
            s.state_.set_col((int)is, stateRadauPert.col((int)is));
            fun_->LagrangeCost(s, perlagrangeOut);
            for (size_t k = 0; k < nnodes; ++k) deriv_langrange((int)k, (int)is) = (perlagrangeOut[k] - lagrangeOut[k]) / pertState((int)k, (int)is);
            s.state_.set_col((int)is, state_radau.col((int)is));
        }
        Mat control_radau = s.control_;
        for (size_t ic = 0; ic < ncontrols; ++ic) {
            s.control_.set_col((int)ic, controlRadauPert.col((int)ic));
            fun_->LagrangeCost(s, perlagrangeOut);
            for (size_t k = 0; k < nnodes; ++k) deriv_langrange((int)k, (int)(nstate + ic)) = (perlagrangeOut[k] - lagrangeOut[k]) / pertControl((int)k, (int)ic);
            s.control_.set_col((int)ic, control_radau.col((int)ic));
        }
    }

    // LpFiniteDifferenceDerive.cpp:194-324.
    // deriv_state: (nnodes*nstate) x (nstate+ncontrols+1), column c = reshape of the
    // N x ns quotient matrix for perturbed variable c; deriv_path likewise.
    void DerivDae(SolDae& s, Mat& deriv_state, Mat& deriv_path) override
    {
        int nstate = s.state_.n_cols, ncontrols = s.contol_.n_cols, nnodes = s.state_.n_rows;
        Mat daeout, pathout;
        fun_->DaeFunction(s, daeout, pathout);
        int npaths = pathout.n_cols;
        Vec pertTime(nnodes), tRadauPert(nnodes);
        for (int k = 0; k < nnodes; ++k) { pertTime[k] = tol_ * (1 + std::fabs(s.time_[k])); tRadauPert[k] = s.time_[k] + pertTime[k]; }
        Mat pertState(nnodes, nstate), stateRadauPert(nnodes, nstate);
        for (size_t i = 0; i < pertState.a.size(); ++i) { pertState.a[i] = tol_ * (1 + std::fabs(s.state_.a[i])); stateRadauPert.a[i] = s.state_.a[i] + pertState.a[i]; }
        Mat pertControl(nnodes, ncontrols), controlRadauPert(nnodes, ncontrols);
        for (size_t i = 0; i < pertControl.a.size(); ++i) { pertControl.a[i] = tol_ * (1 + std::fabs(s.contol_.a[i])); controlRadauPert.a[i] = s.contol_.a[i] + pertControl.a[i]; }

        Mat derive_dae(nnodes * (nstate + npaths), nstate + ncontrols + 1);
        Mat perstateout, perpathout;
        // (join_horiz(perstateout,perpathout) - join_horiz(daeout,pathout)) / denominator, reshaped into column `col`
        auto quotient = [&](int col, const double* denom_col) {
            for (int j = 0; j < nstate; ++j)
                for (int k = 0; k < nnodes; ++k) derive_dae(j * nnodes + k, col) = (perstateout(k, j) - daeout(k, j)) / denom_col[k];
            for (int j = 0; j < npaths; ++j)
                for (int k = 0; k < nnodes; ++k) derive_dae((nstate + j) * nnodes + k, col) = (perpathout(k, j) - pathout(k, j)) / denom_col[k];
        };
        Vec t_radau = s.time_; // time :227-243
        s.time_ = tRadauPert;
        fun_->DaeFunction(s, perstateout, perpathout);
        s.time_ = t_radau;
        quotient(nstate + ncontrols, pertTime.data());
        Mat state_radau = s.state_; // states :245-261
        for (int is = 0; is < nstate; ++is) {
            s.state_.set_col(is, stateRadauPert.col(is));
            fun_->DaeFunction(s, perstateout, perpathout);
            quotient(is, &pertState.a[(size_t)is * nnodes]);
            s.state_.set_col(is, state_radau.col(is));
        }
        Mat control_radau = s.contol_; // controls :263-280
        for (int ic = 0; ic < ncontrols; ++ic) {
            s.contol_.set_col(ic, controlRadauPert.col(ic));
            fun_->DaeFunction(s, perstateout, perpathout);
            quotient(nstate + ic, &pertControl.a[(size_t)ic * nnodes]);
            s.contol_.set_col(ic, control_radau.col(ic));
        }
        deriv_state = Mat(nnodes * nstate, derive_dae.n_cols); // :318-323
        for (int c = 0; c < derive_dae.n_cols; ++c)
            for (int r = 0; r < nnodes * nstate; ++r) deriv_state(r, c) = derive_dae(r, c);
        deriv_path = Mat();
        if (npaths > 0) {
            deriv_path = Mat(nnodes * npaths, derive_dae.n_cols);
            for (int c = 0; c < derive_dae.n_cols; ++c)
                for (int r = 0; r < nnodes * npaths; ++r) deriv_path(r, c) = derive_dae(nnodes * nstate + r, c);
        }
    }

    // LpFiniteDifferenceDerive.cpp:326-409; columns [x0 | t0 | xf | tf]
    void DerivEvent(SolEvent& s, Mat& deriv_event) override
    {
        size_t nstate = s.initial_state_.size();
        double pert0 = tol_ * (1 + std::fabs(s.initial_time_));
        double pertf = tol_ * (1 + std::fabs(s.terminal_time_));
        Vec pertx0(nstate), pertxf(nstate), x0Pert(nstate), xfPert(nstate);
        for (size_t i = 0; i < nstate; ++i) {
            pertx0[i] = tol_ * (std::fabs(s.initial_state_[i]) + 1);
            pertxf[i] = tol_ * (std::fabs(s.terminal_state_[i]) + 1);
        }
        Vec eventout;
        fun_->EventFunction(s, eventout);
        size_t nevents = eventout.size();
        double t0Pert = s.initial_time_ + pert0, tfPert = s.terminal_time_ + pertf;
        for (size_t i = 0; i < nstate; ++i) { x0Pert[i] = s.initial_state_[i] + pertx0[i]; xfPert[i] = s.terminal_state_[i] + pertxf[i]; }
        deriv_event = Mat((int)nevents, (int)(2 * nstate + 2), 0.0);
        Vec per;
        double t0 = s.initial_time_;
        s.initial_time_ = t0Pert;
        fun_->EventFunction(s, per);
        for (size_t e = 0; e < nevents; ++e) deriv_event((int)e, (int)nstate) = (per[e] - eventout[e]) / pert0;
        s.initial_time_ = t0;
        double tf = s.terminal_time_;
        s.terminal_time_ = tfPert;
        fun_->EventFunction(s, per);
        for (size_t e = 0; e < nevents; ++e) deriv_event((int)e, (int)(2 * nstate + 1)) = (per[e] - eventout[e]) / pertf;
        s.terminal_time_ = tf;
        Vec x0 = s.initial_state_, xf = s.terminal_state_;
        for (size_t i = 0; i < nstate; ++i) {
            s.initial_state_[i] = x0Pert[i];
            fun_->EventFunction(s, per);
            for (size_t e = 0; e < nevents; ++e) deriv_event((int)e, (int)i) = (per[e] - eventout[e]) / (pertx0[i] * 1.0);
            s.initial_state_[i] = x0[i];
            s.terminal_state_[i] = xfPert[i];
            fun_->EventFunction(s, per);
            for (size_t e = 0; e < nevents; ++e) deriv_event((int)e, (int)(nstate + 1 + i)) = (per[e] - eventout[e]) / (pertxf[i] * 1.0);
            s.terminal_state_[i] = xf[i];
        }
    }

    // LpFiniteDifferenceDerive.cpp:411-506; columns [xf_left | x0_right]
    void DerivLink(SolLink& s, Mat& derive_link) override
    {
        Vec Linkout;
        fun_->LinkFunction(s, Linkout);
        size_t nL = s.left_state_.size(), nR = s.right_state_.size(), nlinks = Linkout.size();
        Vec xf_left = s.left_state_, x0_right = s.right_state_;
        Vec pertL(nL), pertR(nR), xfLeftPert(nL), x0RightPert(nR);
        for (size_t i = 0; i < nL; ++i) { pertL[i] = tol_ * (1 + std::fabs(xf_left[i])); xfLeftPert[i] = xf_left[i] + pertL[i]; }
        for (size_t i = 0; i < nR; ++i) { pertR[i] = tol_ * (1 + std::fabs(x0_right[i])); x0RightPert[i] = x0_right[i] + pertR[i]; }
        derive_link = Mat((int)nlinks, (int)(nL + nR), 0.0);
        Vec per;
        for (size_t i = 0; i < nL; ++i) {
            s.left_state_[i] = xfLeftPert[i];
            fun_->LinkFunction(s, per);
            for (size_t l = 0; l < nlinks; ++l) derive_link((int)l, (int)i) = (per[l] - Linkout[l]) / (1.0 * pertL[i]);
            s.left_state_[i] = xf_left[i];
        }
        for (size_t i = 0; i < nR; ++i) {
            s.right_state_[i] = x0RightPert[i];
            fun_->LinkFunction(s, per);
            for (size_t l = 0; l < nlinks; ++l) derive_link((int)l, (int)(nL + i)) = (per[l] - Linkout[l]) / (1.0 * pertR[i]);
            s.right_state_[i] = x0_right[i];
        }
    }
    double tol() const { return tol_; }
private:
    std::shared_ptr<FunctionWrapper> fun_;
    double tol_;
};

} // namespace lpo
