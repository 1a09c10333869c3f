// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Drives the REFERENCE'S OWN transcription code -- the unmodified translation units under
// /root/reference/Lpopc/src compiled against oracle/ref_shim/ (see oracle/ref_build.mk) --
// through the same `lpo_*` C API as the restatement (oracle/oracle_capi.cpp), so that
// tests/test_reference_pin.py and oracle/make_golden.py can run both on the same inputs.
// Output: oracle/_ref/liblpopc_ref.so.  Nothing in the product loads it.
//
// Wiring follows the reference:
//   object construction      LpopcAlgorithm::Initialized             Core/LpLpopcAlgorithm.cpp:157-246
//   per-mesh sequence        SetFirstMesh/GetSizes/GetBounds/GetGuess Core/LpLpopcAlgorithm.cpp:17-45,:131-155
//   marshalling              LpopcIpopt::eval_*                       Core/LpopcIpopt.cpp:11-218
// The user functions are the SAME functor headers the device kernels compile
// (include/problems/*.h), presented through the reference's whole-mesh FunctionWrapper
// interface (Core/LpFunctionWrapper.h:50-69) -- the reference's own examples call libm through
// Armadillo, which would differ from the device's deterministic math in the last ulp and so
// blur the comparison of the transcription itself.
#include "LpAnalyticDerive.hpp"
#include "LpBoundsChecker.hpp"
#include "LpCalculateData.hpp"
#include "LpDerivDependciesChecker.h"
#include "LpFiniteDifferenceDerive.hpp"
#include "LpFunctionWrapper.h"
#include "LpGuessChecker.h"
#include "LpHessian.h"
#include "LpNLPWrapper.hpp"
#include "LpOptimalProblem.hpp"
#include "LpPhMeshRefineAlg.hpp"
#include "LpLiuHpMeshRefineAlg.hpp"
#include "LpSizeChecker.h"
#include "LpSolutionError.h"
#include "Nlp2OPConverter.h"
#include "spdlog/sinks/null_sink.h"
#include "RPMGenerator.hpp"

#include "../include/lpopc_b200.h"
#include "../include/problems/all_problems.h"

#include <cstring>
#include <memory>
#include <string>
#include <vector>

using namespace Lpopc;

namespace {

struct RefError : std::runtime_error {
    explicit RefError(const std::string& s) : std::runtime_error(s) {}
};

// pointwise functor set -> the reference's whole-mesh FunctionWrapper
template <class P>
class RefAdapter : public FunctionWrapper {
public:
    typename P::Consts C;
    std::vector<int> nevents_;
    int nlinks_ = 0;
    static constexpr int NS = P::NS, NC = P::NC, NPATH = P::NPATH;
    static constexpr int NSa = NS > 0 ? NS : 1, NCa = NC > 0 ? NC : 1, NPa = NPATH > 0 ? NPATH : 1;
    static constexpr int NEa = P::NE_MAX > 0 ? P::NE_MAX : 1, NLa = P::NL_MAX > 0 ? P::NL_MAX : 1;
    RefAdapter(const double* consts, int nconsts, const std::vector<int>& ne, int nl) : nevents_(ne), nlinks_(nl)
    {
        std::memset(&C, 0, sizeof C);
        size_t nb = (size_t)nconsts * sizeof(double);
        if (nb > sizeof C) nb = sizeof C;
        if (consts) std::memcpy(&C, consts, nb);
    }
    void MayerCost(SolCost& s, double& mayer) override
    {
        mayer = P::mayer(C, s.phase_num_, s.initial_time_, s.initial_state_.memptr(), s.terminal_time_, s.terminal_state_.memptr());
    }
    void LagrangeCost(SolCost& s, vec& L) override
    {
        const int N = (int)s.time_.n_elem;
        L = zeros<vec>(N);
        double x[NSa], u[NCa];
        for (int k = 0; k < N; ++k) {
            for (int j = 0; j < NS; ++j) x[j] = s.state_(k, j);
            for (int j = 0; j < NC; ++j) u[j] = s.control_(k, j);
            L(k) = P::lagrange(C, s.phase_num_, s.time_(k), x, u);
        }
    }
    void DaeFunction(SolDae& s, mat& stateout, mat& pathout) override
    {
        const int N = (int)s.time_.n_elem;
        stateout = zeros<mat>(N, NS);
        pathout = zeros<mat>(N, NPATH);
        double x[NSa], u[NCa], f[NSa], c[NPa];
        for (int k = 0; k < N; ++k) {
            for (int j = 0; j < NS; ++j) x[j] = s.state_(k, j);
            for (int j = 0; j < NC; ++j) u[j] = s.contol_(k, j);
            P::dae(C, s.phase_num_, s.time_(k), x, u, f, c);
            for (int j = 0; j < NS; ++j) stateout(k, j) = f[j];
            for (int j = 0; j < NPATH; ++j) pathout(k, j) = c[j];
        }
    }
    void EventFunction(SolEvent& s, vec& eventout) override
    {
        const int ne = nevents_[s.phase_num_ - 1];
        double e[NEa];
        for (int i = 0; i < NEa; ++i) e[i] = 0.0;
        P::event(C, s.phase_num_, s.initial_time_, s.initial_state_.memptr(), s.terminal_time_, s.terminal_state_.memptr(), e);
        eventout = zeros<vec>(ne);
        for (int i = 0; i < ne; ++i) eventout(i) = e[i];
    }
    void LinkFunction(SolLink& s, vec& linkageout) override
    {
        double o[NLa];
        for (int i = 0; i < NLa; ++i) o[i] = 0.0;
        P::link(C, s.left_state_.memptr(), s.right_state_.memptr(), o);
        linkageout = zeros<vec>(nlinks_);
        for (int i = 0; i < nlinks_; ++i) linkageout(i) = o[i];
    }
    // analytic derivatives: layouts of Lpopc/doc/LpopcDoc.tex:727-760,826-860
    void DerivDae(SolDae& s, mat& deriv_state, mat& deriv_path) override { DerivDaeImpl(s, deriv_state, deriv_path, std::integral_constant<bool, P::HAS_ANALYTIC>()); }
    void DerivLagrange(SolCost& s, mat& d) override { DerivLagrangeImpl(s, d, std::integral_constant<bool, P::HAS_ANALYTIC>()); }
    void DerivMayer(SolCost& s, rowvec& d) override { DerivMayerImpl(s, d, std::integral_constant<bool, P::HAS_ANALYTIC>()); }
    // user derivatives of events and linkages (LpFunctionWrapper.h:64,67; consumed at LpNLPWrapper.cpp:651-668, :441-519)
    void DerivEvent(SolEvent& s, mat& d) override { DerivEventImpl(s, d, std::integral_constant<bool, P::HAS_ANALYTIC && (P::NE_MAX > 0)>()); }
    void DerivLink(SolLink& s, mat& d) override { DerivLinkImpl(s, d, std::integral_constant<bool, P::HAS_ANALYTIC && (P::NL_MAX > 0)>()); }

private:
    void DerivEventImpl(SolEvent&, mat&, std::false_type) { throw RefError("functor set has no analytic event derivatives"); }
    void DerivLinkImpl(SolLink&, mat&, std::false_type) { throw RefError("functor set has no analytic linkage derivatives"); }
    void DerivEventImpl(SolEvent& s, mat& de, std::true_type)
    {
        const int ne = nevents_[s.phase_num_ - 1], W = 2 * NS + 2;
        double d[NEa * (2 * NS + 2)];
        P::devent(C, s.phase_num_, s.initial_time_, s.initial_state_.memptr(), s.terminal_time_, s.terminal_state_.memptr(), d);
        de = zeros<mat>(ne, W);
        for (int q = 0; q < ne; ++q)
            for (int c = 0; c < W; ++c) de(q, c) = d[q * W + c];
    }
    void DerivLinkImpl(SolLink& s, mat& dl, std::true_type)
    {
        const int W = 2 * NS;
        double d[NLa * 2 * NS];
        P::dlink(C, s.left_state_.memptr(), s.right_state_.memptr(), d);
        dl = zeros<mat>(nlinks_, W);
        for (int q = 0; q < nlinks_; ++q)
            for (int c = 0; c < W; ++c) dl(q, c) = d[q * W + c];
    }
    void DerivDaeImpl(SolDae&, mat&, mat&, std::false_type) { throw RefError("functor set has no analytic derivatives"); }
    void DerivLagrangeImpl(SolCost&, mat&, std::false_type) { throw RefError("functor set has no analytic derivatives"); }
    void DerivMayerImpl(SolCost&, rowvec&, std::false_type) { throw RefError("functor set has no analytic derivatives"); }
    void DerivDaeImpl(SolDae& s, mat& deriv_state, mat& deriv_path, std::true_type)
    {
        const int N = (int)s.time_.n_elem;
        const int NV = NS + NC + 1;
        deriv_state = zeros<mat>(N * NS, NV);
        deriv_path = NPATH > 0 ? zeros<mat>(N * NPATH, NV) : mat();
        double x[NSa], u[NCa], d[(NS + NPATH) * NV];
        for (int k = 0; k < N; ++k) {
            for (int j = 0; j < NS; ++j) x[j] = s.state_(k, j);
            for (int j = 0; j < NC; ++j) u[j] = s.contol_(k, j);
            P::ddae(C, s.phase_num_, s.time_(k), x, u, d);
            for (int i = 0; i < NS; ++i)
                for (int c = 0; c < NV; ++c) deriv_state(i * N + k, c) = d[i * NV + c];
            for (int i = 0; i < NPATH; ++i)
                for (int c = 0; c < NV; ++c) deriv_path(i * N + k, c) = d[(NS + i) * NV + c];
        }
    }
    void DerivLagrangeImpl(SolCost& s, mat& dl, std::true_type)
    {
        const int N = (int)s.time_.n_elem;
        const int NV = NS + NC + 1;
        dl = zeros<mat>(N, NV);
        double x[NSa], u[NCa], d[NV];
        for (int k = 0; k < N; ++k) {
            for (int j = 0; j < NS; ++j) x[j] = s.state_(k, j);
            for (int j = 0; j < NC; ++j) u[j] = s.control_(k, j);
            P::dlagrange(C, s.phase_num_, s.time_(k), x, u, d);
            for (int c = 0; c < NV; ++c) dl(k, c) = d[c];
        }
    }
    void DerivMayerImpl(SolCost& s, rowvec& dm, std::true_type)
    {
        double d[2 * NS + 2];
        P::dmayer(C, s.phase_num_, s.initial_time_, s.initial_state_.memptr(), s.terminal_time_, s.terminal_state_.memptr(), d);
        dm = zeros<rowvec>(2 * NS + 2);
        for (int i = 0; i < 2 * NS + 2; ++i) dm(i) = d[i];
    }
};

struct PhaseIn {
    lpb_phase_desc d;
    std::vector<double> smin0, smin, sminf, smax0, smax, smaxf, cmin, cmax, pmin, pmax, emin, emax;
    std::vector<double> mesh;
    std::vector<int> nodes;
    std::vector<double> tguess;
    std::vector<std::vector<double>> xguess, uguess;
};
struct LinkIn {
    int left, right;
    std::vector<double> lmin, lmax;
};

struct Ref {
    std::string functor, err;
    std::vector<double> consts;
    double tol = 1e-6;
    int first_derive = 0;
    std::vector<PhaseIn> ph;
    std::vector<LinkIn> lk;
    bool fresh = false;
    std::vector<umat> dep; // kept across refreshes once probed
    // reference objects (rebuilt per mesh, as the reference's mesh loop re-runs GetSizes..GetGuess)
    shared_ptr<FunctionWrapper> fun;
    shared_ptr<OptimalProblem> op;
    shared_ptr<LpCalculateData> cd;
    shared_ptr<OptDerive> derive;
    shared_ptr<LpHessianCalculator> hess;
    shared_ptr<RPMGenerator> rpm;
    shared_ptr<NLPWrapper> nlp;
    vec jI, jJ, hI, hJ;
    // hp-Liu refinement object: it keeps the history of meshes / states across calls and a pointer to the
    // LpCalculateData it was built with, so `cd` is kept (not re-created) from its construction on
    shared_ptr<LiuHpMeshRefineAlg> liu;
};

template <class P>
shared_ptr<FunctionWrapper> make_adapter(const Ref& r)
{
    std::vector<int> ne;
    for (auto& p : r.ph) {
        if (p.d.nstates != P::NS || p.d.ncontrols != P::NC || p.d.npaths != P::NPATH || p.d.nevents > P::NE_MAX)
            throw RefError("phase sizes do not match functor set " + std::string(P::name()));
        ne.push_back(p.d.nevents);
    }
    int nl = 0;
    for (auto& l : r.lk) nl = (int)l.lmin.size();
    return shared_ptr<FunctionWrapper>(new RefAdapter<P>(r.consts.data(), (int)r.consts.size(), ne, nl));
}

} // namespace
// the reference's own example classes (oracle/ref_examples.cpp)
std::shared_ptr<Lpopc::FunctionWrapper> lpopc_ref_example(const std::string& name);
namespace {

shared_ptr<FunctionWrapper> make_fun(const Ref& r)
{
    if (r.functor.compare(0, 4, "ref:") == 0) { // "ref:<example>": the reference's own user functions
        shared_ptr<FunctionWrapper> f = lpopc_ref_example(r.functor.substr(4));
        if (!f) throw RefError("unknown reference example '" + r.functor + "'");
        return f;
    }
#define REF_TRY(P) \
    if (r.functor == P::name()) return make_adapter<P>(r);
    LPB_FOR_EACH_PROBLEM(REF_TRY)
#undef REF_TRY
    throw RefError("unknown functor set '" + r.functor + "'");
}

void refresh(Ref& r)
{
    r.fun = make_fun(r);
    const int P = (int)r.ph.size(), Lp = (int)r.lk.size();
    r.op.reset(new OptimalProblem(P, Lp, r.fun));
    for (int ip = 0; ip < P; ++ip) {
        const PhaseIn& in = r.ph[ip];
        shared_ptr<Phase> ph(new Phase(ip + 1, in.d.nstates, in.d.ncontrols, in.d.nparameters, in.d.npaths, in.d.nevents));
        ph->SetTimeMin(in.d.t0_min, in.d.tf_min);
        ph->SetTimeMax(in.d.t0_max, in.d.tf_max);
        for (int j = 0; j < in.d.nstates; ++j) {
            ph->SetStateMin(in.smin0[j], in.smin[j], in.sminf[j]);
            ph->SetStateMax(in.smax0[j], in.smax[j], in.smaxf[j]);
        }
        for (int j = 0; j < in.d.ncontrols; ++j) { ph->SetcontrolMin(in.cmin[j]); ph->SetcontrolMax(in.cmax[j]); }
        for (int j = 0; j < in.d.npaths; ++j) { ph->SetpathMin(in.pmin[j]); ph->SetpathMax(in.pmax[j]); }
        for (int j = 0; j < in.d.nevents; ++j) { ph->SeteventMin(in.emin[j]); ph->SeteventMax(in.emax[j]); }
        if (in.d.has_duration) ph->SetDuration(in.d.duration_min, in.d.duration_max);
        // guess (LpGuessChecker needs one to run; default: the two corner points of the bounds)
        std::vector<double> tg = in.tguess;
        std::vector<std::vector<double>> xg = in.xguess, ug = in.uguess;
        if (tg.empty()) {
            tg = {in.d.t0_min, in.d.tf_max > in.d.t0_min ? in.d.tf_max : in.d.t0_min + 1.0};
            xg.assign(in.d.nstates, std::vector<double>(2, 0.0));
            ug.assign(in.d.ncontrols, std::vector<double>(2, 0.0));
        }
        for (double t : tg) ph->SetTimeGuess(t);
        for (int j = 0; j < in.d.nstates; ++j)
            for (double v : xg[j]) ph->SetStateGuess(j + 1, v);
        for (int j = 0; j < in.d.ncontrols; ++j)
            for (double v : ug[j]) ph->SetControlGuess(j + 1, v);
        for (double m : in.mesh) ph->SetMeshPoints(m);
        for (int n : in.nodes) ph->SetNodesPerInterval(n);
        r.op->AddPhase(ph);
    }
    for (int l = 0; l < Lp; ++l) {
        shared_ptr<Linkage> lk(new Linkage(l + 1, r.lk[l].left, r.lk[l].right));
        for (double v : r.lk[l].lmin) lk->SetLinkMin(v);
        for (double v : r.lk[l].lmax) lk->SetLinkMax(v);
        r.op->AddLinkage(lk);
    }
    if (!r.liu) r.cd.reset(new LpCalculateData()); // the reference keeps ONE LpCalculateData per problem across grids
    r.cd->autoscale = false;
    r.cd->current_grid = 1;
    // LpopcAlgorithm::Initialized (LpLpopcAlgorithm.cpp:157-246)
    if (r.first_derive == LPB_DERIVE_ANALYTIC) r.derive = shared_ptr<LpAnalyticDerive>(new LpAnalyticDerive(r.fun));
    else r.derive = shared_ptr<LpFDderive>(new LpFDderive(r.fun, r.tol));
    r.hess.reset(new LpHessianCalculator(r.fun, r.derive, r.cd, r.op, r.tol));
    r.rpm.reset(new RPMGenerator());
    shared_ptr<LpScaleOCP> noscale;
    r.nlp.reset(new NLPWrapper(r.fun, r.cd, r.op, r.derive, r.hess, noscale, r.rpm));
    // per-mesh sequence (LpLpopcAlgorithm.cpp:21-24,:36-40)
    r.nlp->RefreshSparsity();
    LpSizeChecker sc;
    sc.GetSize(r.op, r.cd);
    LpBoundsChecker bc;
    bc.GetBounds(r.op, r.cd);
    LpGuessChecker gc;
    gc.GetGuess(r.op, r.rpm, r.cd);
    if (r.dep.empty()) { // dependency probe not run yet: dense mask
        for (int ip = 0; ip < P; ++ip) {
            umat d(r.ph[ip].d.nstates + r.ph[ip].d.npaths, r.ph[ip].d.nstates + r.ph[ip].d.ncontrols);
            d.fill(1);
            r.dep.push_back(d);
        }
    }
    r.cd->allPhaseDependencies = r.dep;
    r.nlp->GetConsSparsity(r.jI, r.jJ);
    r.nlp->GetHessainSparsity(r.hI, r.hJ);
    r.fresh = true;
}

// Data_->result[iphase] as Nlp2OpConverter::Nlp2OpControl fills it (Core/Nlp2OPConverter.cpp:33-75):
// time on [t0, tf] at the LGR nodes + end point, state (N+1) x ns, control N x nc (+ final row)
void fill_result(Ref& r, const double* x)
{
    const int P = (int)r.ph.size();
    r.cd->result.assign(P, shared_ptr<SolutionData>());
    for (int ip = 0; ip < P; ++ip) {
        const int ns = r.ph[ip].d.nstates, nc = r.ph[ip].d.ncontrols;
        const int N = (int)r.cd->PS[ip]->Points.n_elem;
        const size_t s0 = r.cd->phase_indices[ip]->state[0] - 1;
        shared_ptr<SolutionData> sd(new SolutionData());
        const size_t tcol = s0 + (size_t)ns * (N + 1) + (size_t)nc * N;
        const double t0 = x[tcol], tf = x[tcol + 1];
        vec tau_all = join_vert(r.cd->PS[ip]->Points, ones(1, 1));
        sd->time = (tf - t0) * (tau_all + 1) / 2 + t0;
        sd->state = zeros<mat>(N + 1, ns);
        for (int j = 0; j < ns; ++j)
            for (int k = 0; k <= N; ++k) sd->state(k, j) = x[s0 + (size_t)j * (N + 1) + k];
        if (nc > 0) {
            sd->control = zeros<mat>(N + 1, nc);
            for (int j = 0; j < nc; ++j) {
                for (int k = 0; k < N; ++k) sd->control(k, j) = x[s0 + (size_t)ns * (N + 1) + (size_t)j * N + k];
                sd->control(N, j) = sd->control(N - 1, j); // final row is never read by the error estimator
            }
        }
        r.cd->result[ip] = sd;
    }
}

void ensure_logger()
{
    if (!spdlog::get("lpopc_main_logger")) {
        auto lg = std::make_shared<spdlog::logger>("lpopc_main_logger", std::make_shared<spdlog::sinks::null_sink_st>());
        spdlog::register_logger(lg);
    }
}

vec xvec(Ref* r, const double* x)
{
    const size_t n = r->cd->varbounds_min.size();
    vec v(n);
    for (size_t i = 0; i < n; ++i) v(i) = x[i];
    return v;
}

} // namespace

#define REF_GUARD(r, ...)                   \
    try {                                   \
        __VA_ARGS__;                        \
        return 0;                           \
    } catch (LpopcException & e) {          \
        (r)->err = e.Message();             \
        return -1;                          \
    } catch (const std::exception& e) {     \
        (r)->err = e.what();                \
        return -1;                          \
    }

extern "C" {

void* lpo_create(const lpb_problem_desc* d, char* errbuf, int errlen)
{
    try {
        std::unique_ptr<Ref> r(new Ref());
        r->functor = d->functor ? d->functor : "";
        r->tol = d->fd_tol > 0 ? d->fd_tol : 1e-6;
        r->first_derive = d->first_derive;
        if (d->consts && d->nconsts > 0) r->consts.assign(d->consts, d->consts + d->nconsts);
        for (int i = 0; i < d->nphases; ++i) {
            const lpb_phase_desc& pd = d->phases[i];
            PhaseIn in;
            in.d = pd;
            in.smin0.assign(pd.state_min0, pd.state_min0 + pd.nstates); in.smin.assign(pd.state_min, pd.state_min + pd.nstates); in.sminf.assign(pd.state_minf, pd.state_minf + pd.nstates);
            in.smax0.assign(pd.state_max0, pd.state_max0 + pd.nstates); in.smax.assign(pd.state_max, pd.state_max + pd.nstates); in.smaxf.assign(pd.state_maxf, pd.state_maxf + pd.nstates);
            if (pd.ncontrols) { in.cmin.assign(pd.control_min, pd.control_min + pd.ncontrols); in.cmax.assign(pd.control_max, pd.control_max + pd.ncontrols); }
            if (pd.npaths) { in.pmin.assign(pd.path_min, pd.path_min + pd.npaths); in.pmax.assign(pd.path_max, pd.path_max + pd.npaths); }
            if (pd.nevents) { in.emin.assign(pd.event_min, pd.event_min + pd.nevents); in.emax.assign(pd.event_max, pd.event_max + pd.nevents); }
            in.mesh = {-1.0, 1.0}; // default first mesh (LpMeshRefiner.cpp:30-31,:50)
            in.nodes = {20};
            r->ph.push_back(in);
        }
        for (int i = 0; i < d->nlinkpairs; ++i) {
            LinkIn l;
            l.left = d->links[i].left_phase; l.right = d->links[i].right_phase;
            l.lmin.assign(d->links[i].link_min, d->links[i].link_min + d->links[i].nlinks);
            l.lmax.assign(d->links[i].link_max, d->links[i].link_max + d->links[i].nlinks);
            r->lk.push_back(l);
        }
        return r.release();
    } catch (const std::exception& e) {
        if (errbuf && errlen > 0) { std::strncpy(errbuf, e.what(), errlen - 1); errbuf[errlen - 1] = 0; }
        return nullptr;
    }
}

void lpo_destroy(void* h) { delete (Ref*)h; }
const char* lpo_last_error(void* h) { return ((Ref*)h)->err.c_str(); }

int lpo_set_mesh(void* h, int phase, int K, const double* mesh, const int* nodes)
{
    Ref* r = (Ref*)h;
    REF_GUARD(r, {
        if (phase < 0 || phase >= (int)r->ph.size()) throw RefError("phase out of range");
        r->ph[phase].mesh.assign(mesh, mesh + K + 1);
        r->ph[phase].nodes.assign(nodes, nodes + K);
        r->fresh = false;
    })
}

// guess of one phase: npts time points, x[ns][npts], u[nc][npts] (Phase::SetTimeGuess/SetStateGuess/SetControlGuess)
int lpo_set_guess(void* h, int phase, int npts, const double* t, const double* x, const double* u)
{
    Ref* r = (Ref*)h;
    REF_GUARD(r, {
        if (phase < 0 || phase >= (int)r->ph.size()) throw RefError("phase out of range");
        PhaseIn& in = r->ph[phase];
        in.tguess.assign(t, t + npts);
        in.xguess.clear(); in.uguess.clear();
        for (int j = 0; j < in.d.nstates; ++j) in.xguess.emplace_back(x + (size_t)j * npts, x + (size_t)(j + 1) * npts);
        for (int j = 0; j < in.d.ncontrols; ++j) in.uguess.emplace_back(u + (size_t)j * npts, u + (size_t)(j + 1) * npts);
        r->fresh = false;
    })
}

// the NLP guess vector the reference interpolates onto the LGR nodes (LpGuessChecker.cpp:130-203)
int lpo_get_guess(void* h, double* xguess)
{
    Ref* r = (Ref*)h;
    REF_GUARD(r, {
        if (!r->fresh) refresh(*r);
        for (size_t i = 0; i < r->cd->nlpGuessVector.n_elem; ++i) xguess[i] = r->cd->nlpGuessVector(i);
    })
}

int lpo_refresh(void* h)
{
    Ref* r = (Ref*)h;
    REF_GUARD(r, refresh(*r))
}

int lpo_get_nlp_info(void* h, int* n, int* m, int* nnz_jac, int* nnz_h)
{
    Ref* r = (Ref*)h;
    REF_GUARD(r, {
        if (!r->fresh) refresh(*r);
        *n = (int)r->cd->varbounds_min.size();                             // LpopcIpopt.cpp:13
        *m = (int)(r->cd->conbounds_min.size() + r->cd->linmin.n_elem);   // :14
        *nnz_jac = (int)r->jI.n_elem;
        *nnz_h = (int)r->hI.n_elem;
    })
}

int lpo_get_bounds_info(void* h, double* xl, double* xu, double* gl, double* gu)
{
    Ref* r = (Ref*)h;
    REF_GUARD(r, {
        if (!r->fresh) refresh(*r);
        const size_t n = r->cd->varbounds_min.size(), mc = r->cd->conbounds_min.size();
        for (size_t i = 0; i < n; ++i) { xl[i] = r->cd->varbounds_min[i]; xu[i] = r->cd->varbounds_max[i]; }
        for (size_t i = 0; i < mc; ++i) { gl[i] = r->cd->conbounds_min[i]; gu[i] = r->cd->conbounds_max[i]; }
        for (size_t i = 0; i < r->cd->linmin.n_elem; ++i) { gl[mc + i] = r->cd->linmin(i); gu[mc + i] = r->cd->linmax(i); } // :72-80
    })
}

int lpo_eval_f(void* h, const double* x, double* f)
{
    Ref* r = (Ref*)h;
    REF_GUARD(r, { if (!r->fresh) refresh(*r); *f = r->nlp->GetObjFun(xvec(r, x)); })
}
int lpo_eval_grad_f(void* h, const double* x, double* g)
{
    Ref* r = (Ref*)h;
    REF_GUARD(r, {
        if (!r->fresh) refresh(*r);
        vec gr;
        r->nlp->GetObjGrad(xvec(r, x), gr);
        for (size_t i = 0; i < gr.n_elem; ++i) g[i] = gr(i);
    })
}
int lpo_eval_g(void* h, const double* x, double* g)
{
    Ref* r = (Ref*)h;
    REF_GUARD(r, {
        if (!r->fresh) refresh(*r);
        vec xv = xvec(r, x), c;
        r->nlp->GetAllCons(xv, c);
        for (size_t i = 0; i < c.n_elem; ++i) g[i] = c(i);
    })
}
int lpo_eval_jac_g(void* h, const double* x, int* iRow, int* jCol, double* values)
{
    Ref* r = (Ref*)h;
    REF_GUARD(r, {
        if (!r->fresh) refresh(*r);
        if (!values) { // LpopcIpopt.cpp:156-164
            for (size_t i = 0; i < r->jI.n_elem; ++i) { iRow[i] = (int)r->jI(i); jCol[i] = (int)r->jJ(i); }
        } else {
            vec V;
            r->nlp->GetConsJacbi(xvec(r, x), V);
            for (size_t i = 0; i < V.n_elem; ++i) values[i] = V(i);
        }
    })
}
int lpo_eval_h(void* h, const double* x, double sigma, const double* lambda, int* iRow, int* jCol, double* values)
{
    Ref* r = (Ref*)h;
    REF_GUARD(r, {
        if (!r->fresh) refresh(*r);
        if (!values) { // LpopcIpopt.cpp:187-195
            for (size_t i = 0; i < r->hI.n_elem; ++i) { iRow[i] = (int)r->hI(i); jCol[i] = (int)r->hJ(i); }
        } else {
            const size_t m = r->cd->conbounds_min.size() + r->cd->linmin.n_elem;
            vec lam(m); // all m multipliers (the reference copies m-1, LpopcIpopt.cpp:205-209, quirk Q8)
            for (size_t i = 0; i < m; ++i) lam(i) = lambda[i];
            vec V;
            r->nlp->GetHessainValue(xvec(r, x), sigma, lam, V);
            for (size_t i = 0; i < V.n_elem; ++i) values[i] = V(i);
        }
    })
}

// LpDerivDependciesChecker.cpp:10-94 run on xguess (the reference probes its own NLP guess)
int lpo_probe_dependencies(void* h, const double* xguess, int* dep_out)
{
    Ref* r = (Ref*)h;
    REF_GUARD(r, {
        if (!r->fresh) refresh(*r);
        r->cd->nlpGuessVector = xvec(r, xguess);
        DeriveDependicieshecker chk(r->fun, r->cd);
        chk.GetDependiciesForJacobi(r->cd->allPhaseDependencies);
        r->dep = r->cd->allPhaseDependencies;
        r->nlp->RefreshSparsity();
        r->nlp->GetHessainSparsity(r->hI, r->hJ);
        if (dep_out) {
            size_t k = 0;
            for (auto& m : r->dep)
                for (size_t i = 0; i < m.n_elem; ++i) dep_out[k++] = (int)m[i];
        }
    })
}

int lpo_get_tables(void* h, int phase, double* points, double* weights, int* nD, int* nDiag, int* nDoff)
{
    Ref* r = (Ref*)h;
    REF_GUARD(r, {
        if (!r->fresh) refresh(*r);
        ps& p = *r->cd->PS[phase];
        if (points) for (size_t i = 0; i < p.Points.n_elem; ++i) points[i] = p.Points(i);
        if (weights) for (size_t i = 0; i < p.Weights.n_elem; ++i) weights[i] = p.Weights(i);
        if (nD) *nD = (int)p.D.GetLength();
        if (nDiag) *nDiag = (int)p.Diag.GetLength();
        if (nDoff) *nDoff = (int)p.Doffdiag.GetLength();
    })
}
int lpo_get_coo(void* h, int phase, int which, int* rows, int* cols, double* vals)
{
    Ref* r = (Ref*)h;
    REF_GUARD(r, {
        if (!r->fresh) refresh(*r);
        ps& p = *r->cd->PS[phase];
        dsmatrix& s = which == 0 ? p.D : (which == 1 ? p.Diag : p.Doffdiag);
        vec I, J, V;
        dsmatrix::Find(s, I, J, V);
        for (size_t i = 0; i < V.n_elem; ++i) { rows[i] = (int)I(i); cols[i] = (int)J(i); vals[i] = V(i); }
    })
}

// SolutionErrorChecker::CheckSolutionDiffError (Core/LpSolutionError.cpp:112-166) for every phase: the
// relative error matrix ((sum_k (N_k+1)) + 1) x ns per phase, column-major, phases concatenated.
// rows_out[p] receives the row count of phase p.
int lpo_mesh_error(void* h, const double* x, double* rel_err, int* rows_out)
{
    Ref* r = (Ref*)h;
    REF_GUARD(r, {
        if (!r->fresh) refresh(*r);
        fill_result(*r, x);
        SolutionErrorChecker chk(r->fun, r->cd, r->op);
        size_t k = 0;
        for (size_t ip = 0; ip < r->ph.size(); ++ip) {
            mat e;
            chk.CheckSolutionDiffError((int)ip, e);
            if (rows_out) rows_out[ip] = (int)e.n_rows;
            if (rel_err) for (size_t i = 0; i < e.n_elem; ++i) rel_err[k + i] = e[i];
            k += e.n_elem;
        }
    })
}

// PhMeshRefineAlg::RefineMesh (Core/LpPhMeshRefineAlg.cpp:12-99) on the current mesh and solution x.
// Returns per phase the new mesh: K_out[p], then meshpoints (K+1) and nodes (K) appended to the flat
// output arrays (caller provides room for the worst case).  *no_more_refine = 1 when every interval
// satisfies tol.
int lpo_refine_ph(void* h, const double* x, double tol, int Nmax, int Nmin, int* no_more_refine, int* K_out, double* mesh_out, int* nodes_out)
{
    Ref* r = (Ref*)h;
    REF_GUARD(r, {
        if (!r->fresh) refresh(*r);
        ensure_logger();
        fill_result(*r, x);
        shared_ptr<LpReporter> rep(new LpReporter());
        PhMeshRefineAlg alg(Nmax, Nmin, tol, r->fun, r->cd, rep);
        std::vector<shared_ptr<LpMesh>> meshes;
        const bool done = alg.RefineMesh(r->op, meshes);
        *no_more_refine = done ? 1 : 0;
        size_t km = 0, kn = 0;
        for (size_t ip = 0; ip < meshes.size(); ++ip) {
            K_out[ip] = (int)meshes[ip]->nodesPerInterval.n_elem;
            for (size_t i = 0; i < meshes[ip]->meshpoints.n_elem; ++i) mesh_out[km++] = meshes[ip]->meshpoints(i);
            for (size_t i = 0; i < meshes[ip]->nodesPerInterval.n_elem; ++i) nodes_out[kn++] = (int)meshes[ip]->nodesPerInterval(i);
        }
        r->fresh = false; // RefineMesh rewrote the mesh inside r->op; the next call rebuilds from r->ph
    })
}

// LiuHpMeshRefineAlg::RefineMesh (Core/LpLiuHpMeshRefineAlg.cpp:12-260) on the NLP solution x.  The algorithm object
// persists in the handle (mesh / state / error history); lpo_refine_reset drops it.  Outputs as lpo_refine_ph.
int lpo_refine_hp_liu(void* h, const double* x, double tol, int Nmax, double R, int* no_more_refine, int* K_out, double* mesh_out, int* nodes_out)
{
    Ref* r = (Ref*)h;
    REF_GUARD(r, {
        if (!r->fresh) refresh(*r);
        ensure_logger();
        if (!r->liu) {
            shared_ptr<LpReporter> rep(new LpReporter());
            r->liu.reset(new LiuHpMeshRefineAlg(Nmax, R, tol, r->fun, r->cd, rep));
        }
        fill_result(*r, x);
        std::vector<shared_ptr<LpMesh>> meshes;
        const bool done = r->liu->RefineMesh(r->op, meshes);
        *no_more_refine = done ? 1 : 0;
        size_t km = 0, kn = 0;
        for (size_t ip = 0; ip < meshes.size(); ++ip) {
            if (!meshes[ip]) { // phases the algorithm skipped keep their mesh (the reference leaves the entry empty)
                K_out[ip] = (int)r->ph[ip].nodes.size();
                for (double m : r->ph[ip].mesh) mesh_out[km++] = m;
                for (int n : r->ph[ip].nodes) nodes_out[kn++] = n;
                continue;
            }
            K_out[ip] = (int)meshes[ip]->nodesPerInterval.n_elem;
            for (size_t i = 0; i < meshes[ip]->meshpoints.n_elem; ++i) mesh_out[km++] = meshes[ip]->meshpoints(i);
            for (size_t i = 0; i < meshes[ip]->nodesPerInterval.n_elem; ++i) nodes_out[kn++] = (int)meshes[ip]->nodesPerInterval(i);
        }
        r->fresh = false; // RefineMesh rewrote the mesh inside r->op; the next call rebuilds from r->ph
    })
}

int lpo_refine_reset(void* h)
{
    Ref* r = (Ref*)h;
    REF_GUARD(r, {
        r->liu.reset();
        r->fresh = false;
    })
}

// Nlp2OpConverter::Nlp2OpControl (Core/Nlp2OPConverter.cpp:13-196) on the NLP solution x and multipliers
// lambda.  Output per phase (M = N + 1 rows, column-major), phases concatenated:
//   time[M] | state[M x ns] | control[M x nc] | costate[M x ns] | pathmult[M x np] | Hamiltonian[M] | mayer, lagrange
int lpo_nlp2op(void* h, const double* x, const double* lambda, double* out, double* total_cost)
{
    Ref* r = (Ref*)h;
    REF_GUARD(r, {
        if (!r->fresh) refresh(*r);
        const size_t m = r->cd->conbounds_min.size() + r->cd->linmin.n_elem;
        r->cd->nlpreturn_x = xvec(r, x);
        vec lam(m);
        for (size_t i = 0; i < m; ++i) lam(i) = lambda[i];
        r->cd->nlpreturn_lambda = lam;
        r->cd->autoscale = false;
        Nlp2OpConverter conv;
        conv.Nlp2OpControl(r->fun, r->cd, r->op);
        size_t k = 0;
        for (size_t ip = 0; ip < r->ph.size(); ++ip) {
            const SolutionData& sd = *r->cd->result[ip];
            const int ns = r->ph[ip].d.nstates, nc = r->ph[ip].d.ncontrols, np = r->ph[ip].d.npaths;
            const size_t M = sd.time.n_elem;
            for (size_t i = 0; i < M; ++i) out[k++] = sd.time(i);
            for (size_t i = 0; i < M * ns; ++i) out[k++] = sd.state[i];
            for (size_t i = 0; i < M * nc; ++i) out[k++] = sd.control[i];
            for (size_t i = 0; i < M * ns; ++i) out[k++] = sd.costate[i];
            for (size_t i = 0; i < M * np; ++i) out[k++] = sd.pathmult[i];
            for (size_t i = 0; i < M; ++i) out[k++] = sd.Hamiltonian[i];
            out[k++] = sd.mayerCost;
            out[k++] = sd.lagrangeCost;
        }
        if (total_cost) *total_cost = r->cd->optcontrol_cost;
        r->fresh = false; // Nlp2OpControl rewrote the guesses inside r->op; the next call rebuilds from r->ph
    })
}

int lpo_eval_g_jac_batch(void* h, int nbatch, const double* x, double* g, double* values, int)
{
    Ref* r = (Ref*)h;
    REF_GUARD(r, {
        if (!r->fresh) refresh(*r);
        const size_t n = r->cd->varbounds_min.size(), m = r->cd->conbounds_min.size() + r->cd->linmin.n_elem, nnz = r->jI.n_elem;
        for (int b = 0; b < nbatch; ++b) { // the reference is single-threaded and not re-entrant (function statics)
            vec xv = xvec(r, x + (size_t)b * n), c, V;
            if (g) { r->nlp->GetAllCons(xv, c); for (size_t i = 0; i < m; ++i) g[(size_t)b * m + i] = c(i); }
            if (values) { r->nlp->GetConsJacbi(xv, V); for (size_t i = 0; i < nnz; ++i) values[(size_t)b * nnz + i] = V(i); }
        }
    })
}

} // extern "C"
