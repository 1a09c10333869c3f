// oracle/ref_shim/IpTNLP.hpp -- TEST INFRASTRUCTURE ONLY.  Stand-in for the IPOPT header of the same name: the
// abstract Ipopt::TNLP interface with IPOPT's published method signatures (the reference's LpopcIpopt.h:30-90
// overrides exactly these), so that (a) the reference's LpopcIpopt.h, included by LpNLPWrapper.cpp:12, parses, and
// (b) the product's TNLP adapter include/lpopc_b200_ipopt.hpp can be COMPILED and driven by tests/shim_harness.cpp.
// IPOPT itself is an un-vendored dependency of the reference (Lpopc/CMakeLists.txt:6,43) and is not in this image;
// there is no solver behind this interface.
#ifndef LPB_SHIM_IPTNLP
#define LPB_SHIM_IPTNLP
namespace Ipopt {
typedef int Index;
typedef double Number;
enum SolverReturn { SUCCESS = 0, MAXITER_EXCEEDED = 1, LOCAL_INFEASIBILITY = 5 };
class IpoptData;
class IpoptCalculatedQuantities;
class TNLP {
public:
    enum IndexStyleEnum { C_STYLE = 0, FORTRAN_STYLE = 1 };
    virtual ~TNLP() {}
    virtual bool get_nlp_info(Index& n, Index& m, Index& nnz_jac_g, Index& nnz_h_lag, IndexStyleEnum& index_style) = 0;
    virtual bool get_bounds_info(Index n, Number* x_l, Number* x_u, Index m, Number* g_l, Number* g_u) = 0;
    virtual bool get_starting_point(Index n, bool init_x, Number* x, bool init_z, Number* z_L, Number* z_U, Index m, bool init_lambda,
                                    Number* lambda) = 0;
    virtual bool eval_f(Index n, const Number* x, bool new_x, Number& obj_value) = 0;
    virtual bool eval_grad_f(Index n, const Number* x, bool new_x, Number* grad_f) = 0;
    virtual bool eval_g(Index n, const Number* x, bool new_x, Index m, Number* g) = 0;
    virtual bool eval_jac_g(Index n, const Number* x, bool new_x, Index m, Index nele_jac, Index* iRow, Index* jCol, Number* values) = 0;
    virtual bool eval_h(Index n, const Number* x, bool new_x, Number obj_factor, Index m, const Number* lambda, bool new_lambda,
                        Index nele_hess, Index* iRow, Index* jCol, Number* values)
    {
        return false; // quasi-Newton runs do not call it
    }
    virtual void finalize_solution(SolverReturn status, Index n, const Number* x, const Number* z_L, const Number* z_U, Index m,
                                   const Number* g, const Number* lambda, Number obj_value, const IpoptData* ip_data,
                                   IpoptCalculatedQuantities* ip_cq) = 0;
};
} // namespace Ipopt
#endif
