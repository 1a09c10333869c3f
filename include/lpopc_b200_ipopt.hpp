/* lpopc_b200_ipopt.hpp -- the Ipopt::TNLP adapter over the C ABI (header-only C++).
 *
 * This is the binding a reference maintainer adds next to Lpopc::LpopcIpopt (Lpopc/src/Core/LpopcIpopt.h:18-102,
 * LpopcIpopt.cpp:11-248): NLPSolver::SolveNlp (Core/LpNLPSolver.cpp:13-54) hands `new LpbIpopt(h, guess)` to
 * IpoptApplication::OptimizeTNLP instead of `new LpopcIpopt(...)`; nothing else in the host code changes.
 * Include it after IPOPT's IpTNLP.hpp is on the include path and link with -llpopc_b200.
 *
 *   method                      reference (LpopcIpopt.cpp)                    here
 *   get_nlp_info                :11-25   sizes from LpCalculateData, C_STYLE   lpb_get_nlp_info
 *   get_bounds_info             :27-83   bounds + linear rows                  lpb_get_bounds_info
 *   get_starting_point          :85-104  nlpGuessVector                        the guess given to the constructor
 *   eval_f / eval_grad_f        :106-133 GetObjFun / GetObjGrad                lpb_eval_f / lpb_eval_grad_f
 *   eval_g                      :135-150 GetAllCons                            lpb_eval_g
 *   eval_jac_g                  :152-181 GetConsSparsity / GetConsJacbi        lpb_eval_jac_g (values == NULL: structure)
 *   eval_h                      :183-218 GetHessainSparsity / GetHessainValue  lpb_eval_h   (values == NULL: structure)
 *   finalize_solution           :220-248 copies x, lambda, objective into Data  keeps them and runs lpb_nlp2op
 *                                        (Nlp2OpConverter::Nlp2OpControl follows in the reference's mesh loop)
 *
 * The new_x flags are not needed: the library compares x with the last one it evaluated and serves f, grad f, g and
 * the Jacobian values of one x from a single launch (lpb_handle fast path, lpopc_b200.h).  The handle is owned by the
 * caller: destroy it with lpb_destroy AFTER the IpoptApplication has released this object.
 */
#ifndef LPOPC_B200_IPOPT_HPP
#define LPOPC_B200_IPOPT_HPP
#include "IpTNLP.hpp"
#include "lpopc_b200.h"

#include <algorithm>
#include <string>
#include <vector>

class LpbIpopt : public Ipopt::TNLP {
public:
    typedef Ipopt::Index Index;
    typedef Ipopt::Number Number;

    LpbIpopt(lpb_handle* h, std::vector<double> guess) : h_(h), x0_(std::move(guess)) {}
    virtual ~LpbIpopt() {}

    bool get_nlp_info(Index& n, Index& m, Index& nnz_jac_g, Index& nnz_h_lag, IndexStyleEnum& index_style) override
    {
        index_style = TNLP::C_STYLE; /* LpopcIpopt.cpp:22 */
        return ok(lpb_get_nlp_info(h_, &n, &m, &nnz_jac_g, &nnz_h_lag));
    }
    bool get_bounds_info(Index, Number* x_l, Number* x_u, Index, Number* g_l, Number* g_u) override
    {
        return ok(lpb_get_bounds_info(h_, x_l, x_u, g_l, g_u));
    }
    bool get_starting_point(Index n, bool init_x, Number* x, bool init_z, Number*, Number*, Index, bool init_lambda, Number*) override
    {
        /* LpopcIpopt.cpp:85-104: only x is initialised */
        if (init_z || init_lambda || !init_x || (size_t)n != x0_.size()) return false;
        std::copy(x0_.begin(), x0_.end(), x);
        return true;
    }
    bool eval_f(Index, const Number* x, bool, Number& obj_value) override { return ok(lpb_eval_f(h_, x, &obj_value)); }
    bool eval_grad_f(Index, const Number* x, bool, Number* grad_f) override { return ok(lpb_eval_grad_f(h_, x, grad_f)); }
    bool eval_g(Index, const Number* x, bool, Index, Number* g) override { return ok(lpb_eval_g(h_, x, g)); }
    bool eval_jac_g(Index, const Number* x, bool, Index, Index, Index* iRow, Index* jCol, Number* values) override
    {
        return ok(lpb_eval_jac_g(h_, x, iRow, jCol, values));
    }
    bool eval_h(Index, const Number* x, bool, Number obj_factor, Index, const Number* lambda, bool, Index, Index* iRow, Index* jCol,
                Number* values) override
    {
        /* LpopcIpopt.cpp:205-209 copies lambda[0 .. m-2] only (quirk Q8: the last row is linear, zero Hessian) */
        return ok(lpb_eval_h(h_, x, obj_factor, lambda, iRow, jCol, values));
    }
    void finalize_solution(Ipopt::SolverReturn status, Index n, const Number* x, const Number*, const Number*, Index m, const Number*,
                           const Number* lambda, Number obj_value, const Ipopt::IpoptData*, Ipopt::IpoptCalculatedQuantities*) override
    {
        status_ = (int)status;
        objective_ = obj_value;
        x_.assign(x, x + n);
        lambda_.assign(lambda, lambda + m);
        /* what the reference does next with these arrays (Nlp2OpConverter::Nlp2OpControl, Nlp2OPConverter.cpp:13-159) */
        phase_offsets_.assign(64, 0);
        const long long len = lpb_nlp2op_length(h_, phase_offsets_.data());
        solution_.assign(len > 0 ? (size_t)len : 0, 0.0);
        if (len > 0 && lpb_nlp2op(h_, x_.data(), lambda_.data(), solution_.data(), &cost_) != LPB_OK) error_ = lpb_last_error(h_);
    }

    /* results for the host code that follows the solve */
    int status() const { return status_; }
    double objective() const { return objective_; }
    double cost() const { return cost_; }
    const std::vector<double>& x() const { return x_; }
    const std::vector<double>& lambda() const { return lambda_; }
    const std::vector<double>& solution() const { return solution_; }           /* layout: lpb_nlp2op */
    const std::vector<long long>& phase_offsets() const { return phase_offsets_; }
    const std::string& error() const { return error_; }

private:
    bool ok(int rc)
    {
        if (rc == LPB_OK) return true;
        error_ = lpb_last_error(h_); /* IPOPT sees an evaluation error (returns false), as with LP_CATCH in the reference */
        return false;
    }
    lpb_handle* h_;
    std::vector<double> x0_, x_, lambda_, solution_;
    std::vector<long long> phase_offsets_;
    std::string error_;
    double objective_ = 0.0, cost_ = 0.0;
    int status_ = -1;
};
#endif
