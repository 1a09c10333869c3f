// oracle/lp_problem.hpp -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the problem description, size/bounds/index-map construction
// and PS-table fill of lpopc.
// Follows:
//   Lpopc/src/Core/LpOptimalProblem.hpp:30-326   (Phase / Linkage / OptimalProblem)
//   Lpopc/src/Core/LpCalculateData.hpp:29-111    (indices, ps, LpCalculateData)
//   Lpopc/src/Core/LpSizeChecker.cpp:13-152      (GetSize)
//   Lpopc/src/Core/LpBoundsChecker.cpp:13-348    (GetBounds, index maps, AlinearMatrix)
//   Lpopc/src/Core/LpGuessChecker.cpp:110-122    (PS fill)
//   Lpopc/src/Core/LpMeshRefiner.cpp:10-61       (mesh validation)
// Parameters (nq > 0) are fenced: the reference is self-inconsistent there
// (SURVEY.md Appendix B, Q3), so the oracle throws like the product does.
#pragma once
#include "lp_rpm.hpp"
#include <limits>
#include <memory>

namespace lpo {

struct Limit { double state[3]; }; // LpOptimalProblem.hpp:18-29

struct Phase { // LpOptimalProblem.hpp:30-240 (data only)
    int nstates = 0, ncontrols = 0, nparameters = 0, npaths = 0, nevents = 0;
    std::vector<Limit> statemin, statemax;
    Vec controlmin, controlmax, pathmin, pathmax, eventmin, eventmax;
    double t0_min = 0, t0_max = 0, tf_min = 0, tf_max = 0;
    bool hasduration = false;
    double duration_min = 0, duration_max = 0;
    std::vector<double> meshpoints;
    std::vector<int> nodesperinterval;
    int nodes = 0; // SetTotalNodes
    int GetTotalNodes() const { return nodes; }
};

struct Linkage { // LpOptimalProblem.hpp:242-281
    int leftphase = 0, rightphase = 0; // 1-based as constructed
    Vec linkmin, linkmax;
    int LeftPhase() const { return leftphase - 1; }
    int RightPhase() const { return rightphase - 1; }
};

struct OptimalProblem {
    std::vector<Phase> Phases_;
    std::vector<Linkage> Linkage_;
    int GetPhaseNum() const { return (int)Phases_.size(); }
    int GetLinkageNum() const { return (int)Linkage_.size(); }
};

struct indices { std::vector<int> state, control, time, parameter; }; // LpCalculateData.hpp:29-34
struct ps { Vec Points, Weights; dsmatrix D, Diag, Doffdiag; };        // LpCalculateData.hpp:35-41

struct LpCalculateData { // LpCalculateData.hpp:43-111 (hot-path members)
    std::vector<std::vector<int>> SIZES_;
    int numphases_ = 0, numlinkpairs_ = 0, numlinks_ = 0;
    std::vector<int> variables, constraints;
    Vec linmin, linmax;
    std::vector<double> varbounds_min, varbounds_max, conbounds_min, conbounds_max;
    std::vector<int> totalnodes_perphase;
    std::vector<indices> phase_indices;
    std::vector<std::vector<int>> variable_indices, constraint_indices, link_indices;
    std::vector<Mat> allPhaseDependencies; // (ns+np) x (ns+nc), 0/1
    dsmatrix AlinearMatrix;
    std::vector<ps> PS;
};

// LpMeshRefiner.cpp:10-61 (validation part; defaults are the caller's business)
inline void SetAndCheckMesh(const OptimalProblem& op)
{
    for (int i = 0; i < op.GetPhaseNum(); ++i) {
        const Phase& ph = op.Phases_[i];
        if (ph.meshpoints.size() < 2) throw LpoError("MeshRefinement need at least two meshPoints in phase " + std::to_string(i + 1));
        if (ph.meshpoints.front() != -1 || ph.meshpoints.back() != 1) throw LpoError("meshPoints must span -1 to +1 in phase " + std::to_string(i + 1));
        if (ph.meshpoints.size() != ph.nodesperinterval.size() + 1)
            throw LpoError("Number of nodesPerInterval must match number of mesh intervals in phase " + std::to_string(i + 1));
    }
}

// LpSizeChecker.cpp:13-152
inline void GetSize(OptimalProblem& op, LpCalculateData& cd)
{
    cd.SIZES_.assign(op.GetPhaseNum(), std::vector<int>(5));
    for (int i = 0; i < op.GetPhaseNum(); ++i) {
        Phase& ph = op.Phases_[i];
        int phasetotalnodes = 0;
        for (int v : ph.nodesperinterval) phasetotalnodes += v;
        ph.nodes = phasetotalnodes;
        if (ph.statemin.size() != ph.statemax.size()) throw LpoError("State upper & lower bound MUST be same size in phase" + std::to_string(i + 1));
        if (ph.controlmin.size() != ph.controlmax.size()) throw LpoError("Control upper & lower bound MUST be same size in phase" + std::to_string(i + 1));
        if (ph.pathmin.size() != ph.pathmax.size()) throw LpoError("Path upper & lower bound MUST be same size in phase" + std::to_string(i + 1));
        if (ph.eventmin.size() != ph.eventmax.size()) throw LpoError("Event upper & lower bound MUST be same size in phase" + std::to_string(i + 1));
        if (ph.nparameters != 0) throw LpoError("parameters (nq>0) are unsupported: reference is self-inconsistent (quirk Q3)");
        cd.SIZES_[i][0] = (int)ph.statemin.size();
        cd.SIZES_[i][1] = (int)ph.controlmin.size();
        cd.SIZES_[i][2] = 0;
        cd.SIZES_[i][3] = (int)ph.pathmin.size();
        cd.SIZES_[i][4] = (int)ph.eventmin.size();
    }
    int numlinks = 0;
    for (int i = 0; i < op.GetLinkageNum(); ++i) {
        if (op.Linkage_[i].linkmin.size() != op.Linkage_[i].linkmax.size())
            throw LpoError("Linkage Upper and Lower Bound Vector Must Be Same Size in linkage" + std::to_string(i + 1));
        numlinks += (int)op.Linkage_[i].linkmin.size();
    }
    cd.numphases_ = op.GetPhaseNum();
    cd.numlinkpairs_ = op.GetLinkageNum();
    cd.numlinks_ = numlinks;
}

// LpBoundsChecker.cpp:13-348
inline void GetBounds(const OptimalProblem& op, LpCalculateData& cd)
{
    int variable_offset = 0, constraint_offset = 0;
    std::vector<int> nodes(cd.numphases_);
    for (size_t i = 0; i < nodes.size(); ++i) nodes[i] = op.Phases_[i].GetTotalNodes();
    cd.totalnodes_perphase = nodes;
    std::vector<double> var_min, var_max, con_min, con_max;
    int P = op.GetPhaseNum();
    std::vector<int> opt_varnum(P), opt_connum(P);
    std::vector<indices> opt_phase_indices(P);
    std::vector<std::vector<int>> opt_variable_indices(P), opt_constraint_indices(P);
    for (int i = 0; i < cd.numphases_; ++i) {
        int varnum = 0, connum = 0;
        const Phase& cur = op.Phases_[i];
        int nn = cur.GetTotalNodes();
        for (size_t j = 0; j < cur.statemin.size(); ++j) { // :51-85
            const double* mn = cur.statemin[j].state;
            const double* mx = cur.statemax[j].state;
            if (!(mn[0] <= mx[0] && mn[1] <= mx[1] && mn[2] <= mx[2]))
                throw LpoError("Bounds on State are Inconsistent (i.e. max < min) in Phase:" + std::to_string(i + 1));
            var_min.push_back(mn[0]); var_max.push_back(mx[0]);
            con_min.push_back(0); con_max.push_back(0);
            connum++; varnum++;
            for (int k = 1; k < nn; ++k) {
                var_min.push_back(mn[1]); var_max.push_back(mx[1]); varnum++;
                con_min.push_back(0); con_max.push_back(0); connum++;
            }
            var_min.push_back(mn[2]); var_max.push_back(mx[2]); varnum++;
        }
        for (size_t j = 0; j < cur.controlmin.size(); ++j) { // :88-110
            if (!(cur.controlmin[j] <= cur.controlmax[j]))
                throw LpoError("Bounds on Control are Inconsistent (i.e. max < min) in Phase:" + std::to_string(i + 1));
            for (int k = 0; k < nn; ++k) { varnum++; var_min.push_back(cur.controlmin[j]); var_max.push_back(cur.controlmax[j]); }
        }
        var_min.push_back(cur.t0_min); var_min.push_back(cur.tf_min); varnum += 2; // :111-116
        var_max.push_back(cur.t0_max); var_max.push_back(cur.tf_max);
        for (size_t j = 0; j < cur.pathmin.size(); ++j) { // :141-163
            if (!(cur.pathmin[j] <= cur.pathmax[j]))
                throw LpoError("Bounds on path are Inconsistent (i.e. max < min) in Phase:" + std::to_string(i + 1));
            for (int k = 0; k < nn; ++k) { con_min.push_back(cur.pathmin[j]); con_max.push_back(cur.pathmax[j]); connum++; }
        }
        for (size_t j = 0; j < cur.eventmin.size(); ++j) { // :165-185
            if (!(cur.eventmin[j] <= cur.eventmax[j]))
                throw LpoError("Bounds on event are Inconsistent (i.e. max < min) in Phase:" + std::to_string(i + 1));
            con_min.push_back(cur.eventmin[j]); con_max.push_back(cur.eventmax[j]); connum++;
        }
        opt_varnum[i] = varnum; opt_connum[i] = connum;
        std::vector<int> var_indices(varnum), con_indices(connum); // :188-197 (1-based)
        for (int it = 0; it < varnum; ++it) var_indices[it] = variable_offset + it + 1;
        for (int it = 0; it < connum; ++it) con_indices[it] = constraint_offset + it + 1;
        opt_variable_indices[i] = var_indices;
        opt_constraint_indices[i] = con_indices;
        indices vi; // :199-222
        for (int it = 0; it < (nn + 1) * (int)cur.statemin.size(); ++it) vi.state.push_back(variable_offset + it + 1);
        int state_index = vi.state.back();
        for (int it = 0; it < nn * (int)cur.controlmin.size(); ++it) vi.control.push_back(state_index + it + 1);
        int t0_index = vi.control.empty() ? vi.state.back() + 1 : vi.control.back() + 1;
        int tf_index = t0_index + 1;
        vi.time.push_back(t0_index); vi.time.push_back(tf_index);
        opt_phase_indices[i] = vi;
        variable_offset += opt_varnum[i];
        constraint_offset += opt_connum[i];
    }
    std::vector<std::vector<int>> link_index(cd.numlinkpairs_); // :226-253
    for (int i = 0; i < cd.numlinkpairs_; ++i) {
        const Linkage& lk = op.Linkage_[i];
        for (size_t j = 0; j < lk.linkmin.size(); ++j) {
            if (!(lk.linkmin[j] <= lk.linkmax[j]))
                throw LpoError("Bounds on link are Inconsistent (i.e. max < min) in pair:" + std::to_string(i + 1));
            con_min.push_back(lk.linkmin[j]); con_max.push_back(lk.linkmax[j]);
            link_index[i].push_back(constraint_offset + (int)j + 1); // quirk Q6: offset never advances per pair
        }
    }
    cd.variables = opt_varnum; cd.constraints = opt_connum;
    cd.variable_indices = opt_variable_indices; cd.constraint_indices = opt_constraint_indices;
    cd.link_indices = link_index; cd.phase_indices = opt_phase_indices;
    cd.varbounds_min = var_min; cd.varbounds_max = var_max;
    cd.conbounds_min = con_min; cd.conbounds_max = con_max;

    // linear constraints :265-346
    int numvars = (int)cd.varbounds_min.size();
    int nlin = cd.numphases_ + cd.numlinkpairs_;
    Vec AI(2 * nlin, 0.0), AJ(2 * nlin, 0.0), AV(2 * nlin, 0.0);
    Vec Alinmin(nlin, 0.0), Alinmax(nlin, 0.0);
    int alinrowshift = 0;
    for (int i = 0; i < cd.numphases_; ++i) {
        const Phase& cur = op.Phases_[i];
        int ishift = 0;
        if (i != 0) ishift = cd.variable_indices[i - 1].back();
        int t0_index = ishift + ((nodes[i] + 1) * (int)cur.statemin.size()) + nodes[i] * (int)cur.controlmin.size() + 1;
        int tf_index = t0_index + 1;
        AI[alinrowshift] = i; AJ[alinrowshift] = t0_index - 1; AV[alinrowshift] = -1; alinrowshift++;
        AI[alinrowshift] = i; AJ[alinrowshift] = tf_index - 1; AV[alinrowshift] = 1; alinrowshift++;
        if (cur.hasduration) {
            if (!(cur.duration_min <= cur.duration_max))
                throw LpoError("Bounds on duration are Inconsistent (i.e. max < min) in Phase:" + std::to_string(i + 1));
            Alinmin[i] = cur.duration_min; Alinmax[i] = cur.duration_max;
        } else {
            Alinmin[i] = 0; Alinmax[i] = std::numeric_limits<double>::infinity();
        }
    }
    int istart = cd.numphases_;
    for (int i = 0; i < cd.numlinkpairs_; ++i) {
        const Linkage& lk = op.Linkage_[i];
        int left_phase = lk.LeftPhase(), right_phase = lk.RightPhase();
        int npl = cd.SIZES_[left_phase][2], npr = cd.SIZES_[right_phase][2];
        const std::vector<int>& vl = cd.variable_indices[left_phase];
        const std::vector<int>& vr = cd.variable_indices[right_phase];
        int tf_index_left = vl[vl.size() - npl - 1];
        int t0_index_right = vr[vr.size() - npr - 2];
        AI[alinrowshift] = istart + i; AJ[alinrowshift] = tf_index_left - 1; AV[alinrowshift] = -1; alinrowshift++;
        AI[alinrowshift] = istart + i; AJ[alinrowshift] = t0_index_right - 1; AV[alinrowshift] = 1; alinrowshift++;
        Alinmin[istart + i] = 0; Alinmax[istart + i] = 0;
    }
    cd.linmin = Alinmin; cd.linmax = Alinmax;
    cd.AlinearMatrix = dsmatrix::Sparse(AI, AJ, AV, nlin, numvars);
}

// LpGuessChecker.cpp:110-122: rpm->initialize per phase, copy into PS[iphase]
inline void FillPS(const OptimalProblem& op, LpCalculateData& cd)
{
    cd.PS.assign(op.GetPhaseNum(), ps());
    for (int ip = 0; ip < op.GetPhaseNum(); ++ip) {
        const Phase& ph = op.Phases_[ip];
        RPMGenerator rpm;
        rpm.initialize((int)ph.nodesperinterval.size(), ph.meshpoints, ph.nodesperinterval);
        cd.PS[ip].Points = rpm.RPM_points_;
        cd.PS[ip].Weights = rpm.RPM_weights_;
        cd.PS[ip].D = rpm.RPM_Differentiation_matrix_;
        cd.PS[ip].Diag = rpm.RPM_Differentiation_matrix_diag_;
        cd.PS[ip].Doffdiag = rpm.RPM_Differentiation_matrix_off_diag_;
    }
}

} // namespace lpo
