// oracle/oracle_capi.cpp -- TEST INFRASTRUCTURE ONLY.
//
// C ABI over the CPU restatement (for ctypes in tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference leg).  Mirrors the marshalling
// of Lpopc::LpopcIpopt (Lpopc/src/Core/LpopcIpopt.cpp:11-218) and the call
// sequence GetSizes -> GetBounds -> GetGuess(PS fill) of
// LpopcAlgorithm::SolveOptimalControlProblem (LpLpopcAlgorithm.cpp:17-45).
// The shipped product never links or loads this library.
#include "../include/lpopc_b200.h"
#include "../include/problems/all_problems.h"
#include "lp_adapter.hpp"
#include "lp_hessian.hpp"
#include "lp_nlp.hpp"
#include <cstring>
#include <thread>

using namespace lpo;

namespace {

struct Oracle {
    OptimalProblem op;
    LpCalculateData cd;
    std::shared_ptr<FunctionWrapper> fun;
    std::shared_ptr<OptDerive> derive;
    std::unique_ptr<NLPWrapper> nlp;
    std::unique_ptr<LpHessianCalculator> hess;
    double tol = 1e-6;
    int first_derive = 0;
    bool fresh = false;
    std::string functor, err;
    Vec consts;
    Vec jI, jJ, hI, hJ;
};

template <class P>
std::shared_ptr<FunctionWrapper> make_adapter(const Oracle& o)
{
    std::vector<int> ne;
    for (auto& ph : o.op.Phases_) {
        if (ph.nstates != P::NS || ph.ncontrols != P::NC || ph.npaths != P::NPATH || ph.nevents > P::NE_MAX)
            throw LpoError("phase sizes do not match functor set " + std::string(P::name()));
        ne.push_back(ph.nevents);
    }
    int nl = 0;
    for (auto& lk : o.op.Linkage_) {
        if ((int)lk.linkmin.size() > P::NL_MAX) throw LpoError("too many links for functor set");
        if (nl && nl != (int)lk.linkmin.size()) throw LpoError("all link pairs must have the same number of links");
        nl = (int)lk.linkmin.size();
    }
    return std::make_shared<FunctorAdapter<P>>(o.consts.data(), (int)o.consts.size(), ne, nl);
}

std::shared_ptr<FunctionWrapper> make_fun(const Oracle& o)
{
#define LPO_TRY(P) \
    if (o.functor == P::name()) return make_adapter<P>(o);
    LPB_FOR_EACH_PROBLEM(LPO_TRY)
#undef LPO_TRY
    throw LpoError("unknown functor set '" + o.functor + "'");
}

void refresh(Oracle& o)
{
    SetAndCheckMesh(o.op);
    GetSize(o.op, o.cd);
    GetBounds(o.op, o.cd);
    FillPS(o.op, o.cd);
    o.fun = make_fun(o);
    if (o.first_derive == LPB_DERIVE_ANALYTIC) {
        if (!o.fun->HasAnalytic()) throw LpoError("functor set has no analytic derivatives");
        o.derive = std::make_shared<LpAnalyticDerive>(o.fun);
    } else {
        o.derive = std::make_shared<LpFDderive>(o.fun, o.tol);
    }
    o.nlp.reset(new NLPWrapper(o.fun, &o.cd, &o.op, o.derive));
    if (o.cd.allPhaseDependencies.empty()) { // default: dense mask (probe not run yet)
        for (auto& ph : o.op.Phases_) o.cd.allPhaseDependencies.push_back(Mat(ph.nstates + ph.npaths, ph.nstates + ph.ncontrols, 1.0));
    }
    o.hess.reset(new LpHessianCalculator(o.fun, &o.cd, &o.op, o.derive, o.tol));
    o.nlp->GetConsSparsity(o.jI, o.jJ);
    o.hess->GetHessianSparsity(o.hI, o.hJ);
    o.fresh = true;
}

} // namespace

#define LPO_GUARD(o, ...)                        \
    try {                                        \
        __VA_ARGS__;                             \
        return 0;                                \
    } catch (const std::exception& e) {          \
        (o)->err = e.what();                     \
        return -1;                               \
    }

extern "C" {

void* lpo_create(const lpb_problem_desc* d, char* errbuf, int errlen)
{
    try {
        std::unique_ptr<Oracle> o(new Oracle());
        o->functor = d->functor ? d->functor : "";
        o->tol = d->fd_tol > 0 ? d->fd_tol : 1e-6;
        o->first_derive = d->first_derive;
        if (d->consts && d->nconsts > 0) o->consts.assign(d->consts, d->consts + d->nconsts);
        for (int i = 0; i < d->nphases; ++i) {
            const lpb_phase_desc& pd = d->phases[i];
            Phase ph;
            ph.nstates = pd.nstates; ph.ncontrols = pd.ncontrols; ph.nparameters = pd.nparameters; ph.npaths = pd.npaths; ph.nevents = pd.nevents;
            for (int j = 0; j < pd.nstates; ++j) {
                Limit mn = {{pd.state_min0[j], pd.state_min[j], pd.state_minf[j]}};
                Limit mx = {{pd.state_max0[j], pd.state_max[j], pd.state_maxf[j]}};
                ph.statemin.push_back(mn); ph.statemax.push_back(mx);
            }
            for (int j = 0; j < pd.ncontrols; ++j) { ph.controlmin.push_back(pd.control_min[j]); ph.controlmax.push_back(pd.control_max[j]); }
            for (int j = 0; j < pd.npaths; ++j) { ph.pathmin.push_back(pd.path_min[j]); ph.pathmax.push_back(pd.path_max[j]); }
            for (int j = 0; j < pd.nevents; ++j) { ph.eventmin.push_back(pd.event_min[j]); ph.eventmax.push_back(pd.event_max[j]); }
            ph.t0_min = pd.t0_min; ph.t0_max = pd.t0_max; ph.tf_min = pd.tf_min; ph.tf_max = pd.tf_max;
            ph.hasduration = pd.has_duration != 0; ph.duration_min = pd.duration_min; ph.duration_max = pd.duration_max;
            // default first mesh: [-1,1] with 20 nodes (LpMeshRefiner.cpp:30-31,50; quirk Q2)
            ph.meshpoints = {-1.0, 1.0};
            ph.nodesperinterval = {20};
            o->op.Phases_.push_back(ph);
        }
        for (int i = 0; i < d->nlinkpairs; ++i) {
            Linkage lk;
            lk.leftphase = d->links[i].left_phase; lk.rightphase = d->links[i].right_phase;
            lk.linkmin.assign(d->links[i].link_min, d->links[i].link_min + d->links[i].nlinks);
            lk.linkmax.assign(d->links[i].link_max, d->links[i].link_max + d->links[i].nlinks);
            o->op.Linkage_.push_back(lk);
        }
        return o.release();
    } catch (const std::exception& e) {
        if (errbuf && errlen > 0) { std::strncpy(errbuf, e.what(), errlen - 1); errbuf[errlen - 1] = 0; }
        return nullptr;
    }
}

void lpo_destroy(void* h) { delete (Oracle*)h; }
const char* lpo_last_error(void* h) { return ((Oracle*)h)->err.c_str(); }

int lpo_set_mesh(void* h, int phase, int K, const double* mesh, const int* nodes)
{
    Oracle* o = (Oracle*)h;
    LPO_GUARD(o, {
        if (phase < 0 || phase >= o->op.GetPhaseNum()) throw LpoError("phase out of range");
        o->op.Phases_[phase].meshpoints.assign(mesh, mesh + K + 1);
        o->op.Phases_[phase].nodesperinterval.assign(nodes, nodes + K);
        o->fresh = false;
    })
}

int lpo_refresh(void* h)
{
    Oracle* o = (Oracle*)h;
    LPO_GUARD(o, refresh(*o))
}

int lpo_get_nlp_info(void* h, int* n, int* m, int* nnz_jac, int* nnz_h)
{
    Oracle* o = (Oracle*)h;
    LPO_GUARD(o, {
        if (!o->fresh) refresh(*o);
        *n = (int)o->cd.varbounds_min.size();                           // LpopcIpopt.cpp:13
        *m = (int)(o->cd.conbounds_min.size() + o->cd.linmin.size()); // :14
        *nnz_jac = (int)o->jI.size();
        *nnz_h = (int)o->hI.size();
    })
}

int lpo_get_bounds_info(void* h, double* xl, double* xu, double* gl, double* gu)
{
    Oracle* o = (Oracle*)h;
    LPO_GUARD(o, {
        if (!o->fresh) refresh(*o);
        size_t n = o->cd.varbounds_min.size(), mc = o->cd.conbounds_min.size();
        for (size_t i = 0; i < n; ++i) { xl[i] = o->cd.varbounds_min[i]; xu[i] = o->cd.varbounds_max[i]; }
        for (size_t i = 0; i < mc; ++i) { gl[i] = o->cd.conbounds_min[i]; gu[i] = o->cd.conbounds_max[i]; }
        for (size_t i = 0; i < o->cd.linmin.size(); ++i) { gl[mc + i] = o->cd.linmin[i]; gu[mc + i] = o->cd.linmax[i]; } // :72-80
    })
}

static Vec xvec(Oracle* o, const double* x) { return Vec(x, x + o->cd.varbounds_min.size()); }

int lpo_eval_f(void* h, const double* x, double* f)
{
    Oracle* o = (Oracle*)h;
    LPO_GUARD(o, { if (!o->fresh) refresh(*o); *f = o->nlp->GetObjFun(xvec(o, x)); })
}
int lpo_eval_grad_f(void* h, const double* x, double* g)
{
    Oracle* o = (Oracle*)h;
    LPO_GUARD(o, {
        if (!o->fresh) refresh(*o);
        Vec gr;
        o->nlp->GetObjGrad(xvec(o, x), gr);
        std::memcpy(g, gr.data(), gr.size() * sizeof(double));
    })
}
int lpo_eval_g(void* h, const double* x, double* g)
{
    Oracle* o = (Oracle*)h;
    LPO_GUARD(o, {
        if (!o->fresh) refresh(*o);
        Vec c;
        o->nlp->GetAllCons(xvec(o, x), c);
        std::memcpy(g, c.data(), c.size() * sizeof(double));
    })
}
int lpo_eval_jac_g(void* h, const double* x, int* iRow, int* jCol, double* values)
{
    Oracle* o = (Oracle*)h;
    LPO_GUARD(o, {
        if (!o->fresh) refresh(*o);
        if (!values) { // LpopcIpopt.cpp:156-164
            for (size_t i = 0; i < o->jI.size(); ++i) { iRow[i] = (int)o->jI[i]; jCol[i] = (int)o->jJ[i]; }
        } else {
            Vec V;
            o->nlp->GetConsJacbi(xvec(o, x), V);
            std::memcpy(values, V.data(), V.size() * sizeof(double));
        }
    })
}
int lpo_eval_h(void* h, const double* x, double sigma, const double* lambda, int* iRow, int* jCol, double* values)
{
    Oracle* o = (Oracle*)h;
    LPO_GUARD(o, {
        if (!o->fresh) refresh(*o);
        if (!values) { // LpopcIpopt.cpp:187-195
            for (size_t i = 0; i < o->hI.size(); ++i) { iRow[i] = (int)o->hI[i]; jCol[i] = (int)o->hJ[i]; }
        } else {
            size_t m = o->cd.conbounds_min.size() + o->cd.linmin.size();
            Vec lam(lambda, lambda + m); // Q8 (last multiplier uninitialised in the reference) is not replicated
            Vec V;
            o->hess->GetHessian(sigma, xvec(o, x), lam, V);
            std::memcpy(values, V.data(), V.size() * sizeof(double));
        }
    })
}

// LpDerivDependciesChecker.cpp:10-94
int lpo_probe_dependencies(void* h, const double* xguess, int* dep_out)
{
    Oracle* o = (Oracle*)h;
    LPO_GUARD(o, {
        if (!o->fresh) refresh(*o);
        GetDependiciesForJacobiInEveryPhase(*o->nlp, xvec(o, xguess), o->cd.allPhaseDependencies);
        o->hess->GetHessianSparsity(o->hI, o->hJ);
        if (dep_out) {
            size_t k = 0;
            for (auto& m : o->cd.allPhaseDependencies)
                for (double v : m.a) dep_out[k++] = (int)v;
        }
    })
}

// table access for known-answer tests
int lpo_get_tables(void* h, int phase, double* points, double* weights, int* nD, int* nDiag, int* nDoff)
{
    Oracle* o = (Oracle*)h;
    LPO_GUARD(o, {
        if (!o->fresh) refresh(*o);
        const ps& p = o->cd.PS[phase];
        if (points) std::memcpy(points, p.Points.data(), p.Points.size() * sizeof(double));
        if (weights) std::memcpy(weights, p.Weights.data(), p.Weights.size() * sizeof(double));
        if (nD) *nD = p.D.GetLength();
        if (nDiag) *nDiag = p.Diag.GetLength();
        if (nDoff) *nDoff = p.Doffdiag.GetLength();
    })
}
int lpo_get_coo(void* h, int phase, int which, int* rows, int* cols, double* vals)
{
    Oracle* o = (Oracle*)h;
    LPO_GUARD(o, {
        if (!o->fresh) refresh(*o);
        const ps& p = o->cd.PS[phase];
        const dsmatrix& s = which == 0 ? p.D : (which == 1 ? p.Diag : p.Doffdiag);
        for (int i = 0; i < s.GetLength(); ++i) { rows[i] = s.rows[i]; cols[i] = s.cols[i]; vals[i] = s.vals[i]; }
    })
}

// Batched evaluation over independent instances with nthreads host threads: the
// CPU baseline of BASELINE config 4 (the reference itself is single-threaded;
// instances are independent, so this is the most favourable CPU arrangement).
int lpo_eval_g_jac_batch(void* h, int nbatch, const double* x, double* g, double* values, int nthreads)
{
    Oracle* o = (Oracle*)h;
    LPO_GUARD(o, {
        if (!o->fresh) refresh(*o);
        size_t n = o->cd.varbounds_min.size(), m = o->cd.conbounds_min.size() + o->cd.linmin.size(), nnz = o->jI.size();
        if (nthreads < 1) nthreads = 1;
        std::vector<std::thread> th;
        std::vector<std::string> errs(nthreads);
        for (int t = 0; t < nthreads; ++t)
            th.emplace_back([&, t]() {
                try {
                    NLPWrapper w(o->fun, &o->cd, &o->op, o->derive);
                    for (int b = t; b < nbatch; b += nthreads) {
                        Vec xb(x + (size_t)b * n, x + (size_t)(b + 1) * n), c, V;
                        if (g) { w.GetAllCons(xb, c); std::memcpy(g + (size_t)b * m, c.data(), m * sizeof(double)); }
                        if (values) { w.GetConsJacbi(xb, V); std::memcpy(values + (size_t)b * nnz, V.data(), nnz * sizeof(double)); }
                    }
                } catch (const std::exception& e) { errs[t] = e.what(); }
            });
        for (auto& t : th) t.join();
        for (auto& e : errs) if (!e.empty()) throw LpoError(e);
    })
}

int lpo_detmath(int which, int n, const double* x, double* y)
{
    for (int i = 0; i < n; ++i) {
        switch (which) {
        case 0: y[i] = lpb_det_exp(x[i]); break;
        case 1: y[i] = lpb_det_tanh(x[i]); break;
        case 2: y[i] = lpb_det_sin(x[i]); break;
        case 3: y[i] = lpb_det_cos(x[i]); break;
        case 4: y[i] = lpb_det_acos(x[i]); break;
        default: return -1;
        }
    }
    return 0;
}

} // extern "C"
