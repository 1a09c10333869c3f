// oracle/lp_adapter.hpp -- TEST INFRASTRUCTURE ONLY.
//
// Host adapter: presents a pointwise functor set (include/problems/*.h) through
// the reference's whole-mesh matrix interface Lpopc::FunctionWrapper
// (Lpopc/src/Core/LpFunctionWrapper.h:50-69), so the restated transcription code
// calls user functions exactly the way the reference does (one call = all N nodes).
// Analytic Deriv* layouts follow Lpopc/doc/LpopcDoc.tex:727-760,826-860 and the
// consumers in LpNLPWrapper.cpp:584-612,1007-1039.
#pragma once
#include "lp_types.hpp"
#include <cstring>

namespace lpo {

template <class P>
class FunctorAdapter : public FunctionWrapper {
public:
    typename P::Consts C;
    std::vector<int> nevents_; // per phase (0-based)
    int nlinks_ = 0;
    static constexpr int NS = P::NS, NC = P::NC, NPATH = P::NPATH;
    static constexpr int NSa = NS > 0 ? NS : 1, NCa = NC > 0 ? NC : 1, NPa = NPATH > 0 ? NPATH : 1;
    static constexpr int NEa = P::NE_MAX > 0 ? P::NE_MAX : 1, NLa = P::NL_MAX > 0 ? P::NL_MAX : 1;

    FunctorAdapter(const double* consts, int nconsts, const std::vector<int>& nevents, int nlinks) : nevents_(nevents), nlinks_(nlinks)
    {
        std::memset(&C, 0, sizeof C);
        size_t nb = (size_t)nconsts * sizeof(double);
        if (nb > sizeof C) nb = sizeof C;
        if (consts) std::memcpy(&C, consts, nb);
    }
    bool HasAnalytic() const override { return P::HAS_ANALYTIC; }

    void MayerCost(SolCost& s, double& mayer) override
    {
        mayer = P::mayer(C, s.phase_num_, s.initial_time_, s.initial_state_.data(), s.terminal_time_, s.terminal_state_.data());
    }
    void LagrangeCost(SolCost& s, Vec& L) override
    {
        int N = (int)s.time_.size();
        L.assign(N, 0.0);
        double x[NSa], u[NCa];
        for (int k = 0; k < N; ++k) {
            for (int j = 0; j < NS; ++j) x[j] = s.state_(k, j);
            for (int j = 0; j < NC; ++j) u[j] = s.control_(k, j);
            L[k] = P::lagrange(C, s.phase_num_, s.time_[k], x, u);
        }
    }
    void DaeFunction(SolDae& s, Mat& stateout, Mat& pathout) override
    {
        int N = (int)s.time_.size();
        stateout = Mat(N, NS);
        pathout = Mat(N, NPATH);
        double x[NSa], u[NCa], f[NSa], c[NPa];
        for (int k = 0; k < N; ++k) {
            for (int j = 0; j < NS; ++j) x[j] = s.state_(k, j);
            for (int j = 0; j < NC; ++j) u[j] = s.contol_(k, j);
            P::dae(C, s.phase_num_, s.time_[k], x, u, f, c);
            for (int j = 0; j < NS; ++j) stateout(k, j) = f[j];
            for (int j = 0; j < NPATH; ++j) pathout(k, j) = c[j];
        }
    }
    void EventFunction(SolEvent& s, Vec& eventout) override
    {
        int ne = nevents_[s.phase_num_ - 1];
        double e[NEa];
        for (int i = 0; i < NEa; ++i) e[i] = 0.0;
        P::event(C, s.phase_num_, s.initial_time_, s.initial_state_.data(), s.terminal_time_, s.terminal_state_.data(), e);
        eventout.assign(e, e + ne);
    }
    void LinkFunction(SolLink& s, Vec& linkageout) override
    {
        double o[NLa];
        for (int i = 0; i < NLa; ++i) o[i] = 0.0;
        P::link(C, s.left_state_.data(), s.right_state_.data(), o);
        linkageout.assign(o, o + nlinks_);
    }

    // ---- analytic derivatives (only for functor sets with HAS_ANALYTIC) ----
    void DerivDae(SolDae& s, Mat& deriv_state, Mat& deriv_path) override { DerivDaeImpl(s, deriv_state, deriv_path, std::integral_constant<bool, P::HAS_ANALYTIC>()); }
    void DerivLagrange(SolCost& s, Mat& d) override { DerivLagrangeImpl(s, d, std::integral_constant<bool, P::HAS_ANALYTIC>()); }
    void DerivMayer(SolCost& s, Vec& d) override { DerivMayerImpl(s, d, std::integral_constant<bool, P::HAS_ANALYTIC>()); }
    // user derivatives of events and linkages (LpFunctionWrapper.h:64,67): deriv_event is ne x (2 ns + 2) with columns
    // [x0 | t0 | xf | tf], derive_link is nl x 2 ns with columns [xf_left | x0_right]; the functor fills them row-major
    void DerivEvent(SolEvent& s, Mat& d) override { DerivEventImpl(s, d, std::integral_constant<bool, P::HAS_ANALYTIC && (P::NE_MAX > 0)>()); }
    void DerivLink(SolLink& s, Mat& d) override { DerivLinkImpl(s, d, std::integral_constant<bool, P::HAS_ANALYTIC && (P::NL_MAX > 0)>()); }

private:
    void DerivEventImpl(SolEvent&, Mat&, std::false_type) { throw LpoError("functor set has no analytic event derivatives"); }
    void DerivLinkImpl(SolLink&, Mat&, std::false_type) { throw LpoError("functor set has no analytic linkage derivatives"); }
    void DerivEventImpl(SolEvent& s, Mat& de, std::true_type)
    {
        const int ne = nevents_[s.phase_num_ - 1], W = 2 * NS + 2;
        double d[NEa * (2 * NS + 2)];
        P::devent(C, s.phase_num_, s.initial_time_, s.initial_state_.data(), s.terminal_time_, s.terminal_state_.data(), d);
        de = Mat(ne, W, 0.0);
        for (int q = 0; q < ne; ++q)
            for (int c = 0; c < W; ++c) de(q, c) = d[q * W + c];
    }
    void DerivLinkImpl(SolLink& s, Mat& dl, std::true_type)
    {
        const int W = 2 * NS;
        double d[NLa * 2 * NS];
        P::dlink(C, s.left_state_.data(), s.right_state_.data(), d);
        dl = Mat(nlinks_, W, 0.0);
        for (int q = 0; q < nlinks_; ++q)
            for (int c = 0; c < W; ++c) dl(q, c) = d[q * W + c];
    }
    void DerivDaeImpl(SolDae&, Mat&, Mat&, std::false_type) { throw LpoError("functor set has no analytic derivatives"); }
    void DerivLagrangeImpl(SolCost&, Mat&, std::false_type) { throw LpoError("functor set has no analytic derivatives"); }
    void DerivMayerImpl(SolCost&, Vec&, std::false_type) { throw LpoError("functor set has no analytic derivatives"); }
    void DerivDaeImpl(SolDae& s, Mat& deriv_state, Mat& deriv_path, std::true_type)
    {
        int N = (int)s.time_.size();
        const int NV = NS + NC + 1;
        deriv_state = Mat(N * NS, NV);
        deriv_path = NPATH > 0 ? Mat(N * NPATH, NV) : Mat();
        double x[NSa], u[NCa], d[(NS + NPATH) * NV];
        for (int k = 0; k < N; ++k) {
            for (int j = 0; j < NS; ++j) x[j] = s.state_(k, j);
            for (int j = 0; j < NC; ++j) u[j] = s.contol_(k, j);
            P::ddae(C, s.phase_num_, s.time_[k], x, u, d);
            for (int i = 0; i < NS; ++i)
                for (int c = 0; c < NV; ++c) deriv_state(i * N + k, c) = d[i * NV + c];
            for (int i = 0; i < NPATH; ++i)
                for (int c = 0; c < NV; ++c) deriv_path(i * N + k, c) = d[(NS + i) * NV + c];
        }
    }
    void DerivLagrangeImpl(SolCost& s, Mat& dl, std::true_type)
    {
        int N = (int)s.time_.size();
        const int NV = NS + NC + 1;
        dl = Mat(N, NV);
        double x[NSa], u[NCa], d[NV];
        for (int k = 0; k < N; ++k) {
            for (int j = 0; j < NS; ++j) x[j] = s.state_(k, j);
            for (int j = 0; j < NC; ++j) u[j] = s.control_(k, j);
            P::dlagrange(C, s.phase_num_, s.time_[k], x, u, d);
            for (int c = 0; c < NV; ++c) dl(k, c) = d[c];
        }
    }
    void DerivMayerImpl(SolCost& s, Vec& dm, std::true_type)
    {
        double d[2 * NS + 2];
        P::dmayer(C, s.phase_num_, s.initial_time_, s.initial_state_.data(), s.terminal_time_, s.terminal_state_.data(), d);
        dm.assign(d, d + 2 * NS + 2);
    }
};

} // namespace lpo
