// lpb_tables.hpp -- host-side Radau pseudospectral tables for one phase mesh.
//
// Produces what Lpopc::RPMGenerator::initialize produces (Lpopc/src/Core/RPMGenerator.cpp:43-105):
// LGR points/weights mapped to the global tau axis and the per-interval differentiation
// blocks D_k (N_k x (N_k+1), last row of the (N_k+1)-point matrix dropped), built on the
// *scaled* support points of interval k like the reference (:67-75, CollocD :107-130).
// The layout is device-oriented: one dense column-major block per interval in a flat
// array (threads of an interval read consecutive addresses), not the reference's COO;
// the COO views (D / Diag / Doffdiag with exact zeros dropped, RPMGenerator.cpp:178-180)
// are derived on the GPU by a flag-scan-scatter pass (lpb_structure.cu).
// Arithmetic follows the reference's formulas step by step so table values agree to the
// last bit with a faithful restatement; LGR nodes are cached per N like :17-41.
#pragma once
#include <vector>

namespace lpb {

struct PhaseTables {
    int K = 0;                    // intervals
    int N = 0;                    // total LGR nodes
    std::vector<int> int_n;       // nodes per interval
    std::vector<int> int_row0;    // first node (== first column) of interval
    std::vector<long long> int_d0;// offset of the dense block in dblocks
    std::vector<int> node_interval;
    std::vector<double> tau, w;   // [N]
    std::vector<double> dblocks;  // per interval: column-major N_k x (N_k+1)
};

// LGR points and weights on [-1,1) for n nodes (RPMGenerator.cpp:253-291).
void lgr_points(int n, std::vector<double>& x, std::vector<double>& w);

// Differentiation block on support points s[0..M) (M = n+1): column-major n x M
// (RPMGenerator.cpp:107-124).
void colloc_block(const std::vector<double>& s, std::vector<double>& D);

// Whole phase (RPMGenerator.cpp:43-105).  Throws std::runtime_error on a bad mesh
// (checks of MeshRefiner::SetAndCheckMesh, LpMeshRefiner.cpp:32-58).
void build_phase_tables(int K, const double* meshpoints, const int* nodes, PhaseTables& out);

// Tables of the mesh-error estimator (SolutionErrorChecker, LpSolutionError.cpp:54-166): the mesh with one
// more LGR point per interval.
struct ErrTables {
    int K = 0, M = 0;               // intervals, total new points (sum of N_k + 1)
    std::vector<int> int_m;         // N_k + 1
    std::vector<int> int_rn0;       // first new row of interval k
    std::vector<long long> int_a0;  // offset of the integration block in ablocks
    std::vector<double> tnew;       // [M] interpolation abscissae on [-1, 1) (LpSolutionError.cpp:77)
    std::vector<double> ablocks;    // per interval: column-major M_k x M_k, A_k = inv(D_k[:, 1:]) (RPMGenerator.cpp:85)
};
void build_error_tables(int K, const double* meshpoints, const int* nodes, const std::vector<double>& tau_old, ErrTables& out);

} // namespace lpb
