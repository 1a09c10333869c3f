// tests/shim_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the product's Ipopt::TNLP adapter (include/lpopc_b200_ipopt.hpp) against the TNLP interface stand-in
// (oracle/ref_shim/IpTNLP.hpp: IPOPT's published signatures; IPOPT itself is not in this image) and drives it through
// the BASE-CLASS pointer in the order IPOPT's TNLPAdapter / IpoptApplication::OptimizeTNLP make the calls:
//   get_nlp_info -> get_bounds_info -> get_starting_point -> eval_jac_g(values = NULL) -> eval_h(values = NULL)
//   -> per iterate: eval_f(new_x = true), eval_grad_f, eval_g, eval_jac_g, eval_h(new_lambda = true) with new_x = false
//   -> finalize_solution
// with arrays it allocates itself (as IPOPT does), and hands everything back to the test.
#include "IpTNLP.hpp"
#include "../include/lpopc_b200_ipopt.hpp"

#include <cstring>
#include <memory>
#include <vector>

using namespace Ipopt;

extern "C" int shim_drive(void* handle, const double* guess, int nx, const double* xs, const double* lambda, double sigma,
                          int* info, double* xl, double* xu, double* gl, double* gu, double* x0, int* jI, int* jJ, int* hI, int* hJ,
                          double* f, double* grad, double* g, double* jac, double* hess, double* sol, long long sol_cap, double* cost,
                          int* refused_start)
{
    lpb_handle* h = static_cast<lpb_handle*>(handle);
    Index n = 0, m = 0, nnz = 0, nnzh = 0;
    if (lpb_get_nlp_info(h, &n, &m, &nnz, &nnzh) != LPB_OK) return -1;
    LpbIpopt* adapter = new LpbIpopt(h, std::vector<double>(guess, guess + n));
    std::unique_ptr<TNLP> nlp(adapter); // every call below goes through the TNLP vtable
    TNLP::IndexStyleEnum style = TNLP::FORTRAN_STYLE;
    Index n2, m2, a2, b2;
    if (!nlp->get_nlp_info(n2, m2, a2, b2, style)) return -2;
    info[0] = n2; info[1] = m2; info[2] = a2; info[3] = b2; info[4] = (int)style;
    if (!nlp->get_bounds_info(n, xl, xu, m, gl, gu)) return -3;
    // IPOPT asks for x only (LpopcIpopt.cpp:85-104 asserts the same); a request for duals must be refused
    *refused_start = nlp->get_starting_point(n, true, x0, true, nullptr, nullptr, m, false, nullptr) ? 0 : 1;
    if (!nlp->get_starting_point(n, true, x0, false, nullptr, nullptr, m, false, nullptr)) return -4;
    if (!nlp->eval_jac_g(n, nullptr, false, m, nnz, jI, jJ, nullptr)) return -5;
    if (!nlp->eval_h(n, nullptr, false, 1.0, m, nullptr, false, nnzh, hI, hJ, nullptr)) return -6;
    std::vector<double> xi(n), lam(lambda, lambda + m), gi(m), gradi(n), jaci(nnz), hessi(nnzh);
    for (int k = 0; k < nx; ++k) {
        std::memcpy(xi.data(), xs + (size_t)k * n, (size_t)n * sizeof(double));
        Number fv = 0.0;
        if (!nlp->eval_f(n, xi.data(), true, fv)) return -10;
        if (!nlp->eval_grad_f(n, xi.data(), false, gradi.data())) return -11;
        if (!nlp->eval_g(n, xi.data(), false, m, gi.data())) return -12;
        if (!nlp->eval_jac_g(n, xi.data(), false, m, nnz, nullptr, nullptr, jaci.data())) return -13;
        if (!nlp->eval_h(n, xi.data(), false, sigma, m, lam.data(), true, nnzh, nullptr, nullptr, hessi.data())) return -14;
        f[k] = fv;
        std::memcpy(grad + (size_t)k * n, gradi.data(), (size_t)n * sizeof(double));
        std::memcpy(g + (size_t)k * m, gi.data(), (size_t)m * sizeof(double));
        std::memcpy(jac + (size_t)k * nnz, jaci.data(), (size_t)nnz * sizeof(double));
        std::memcpy(hess + (size_t)k * nnzh, hessi.data(), (size_t)nnzh * sizeof(double));
    }
    nlp->finalize_solution(SUCCESS, n, xi.data(), nullptr, nullptr, m, gi.data(), lam.data(), f[nx - 1], nullptr, nullptr);
    if (!adapter->error().empty()) return -20;
    if ((long long)adapter->solution().size() > sol_cap) return -21;
    std::memcpy(sol, adapter->solution().data(), adapter->solution().size() * sizeof(double));
    *cost = adapter->cost();
    return (int)adapter->solution().size();
}
