// lpb_refine_liu.hpp -- hp mesh refinement decision of Liu et al. as the reference implements it
// (Lpopc/src/Core/LpLiuHpMeshRefineAlg.cpp:12-260 RefineMesh, :262-342 power-series coefficients, :347-460 Dividing_mesh /
// Increasing_N, :462-504 Reducing_N, :640-745 CanWeIncreaseN / calculate2nd_derive), selected by the option
// "mesh-refine-methods" = "hp-Liu" (LpMeshRefiner.h:54-61; the reference's hypersensitive example runs with it).
//
// Host logic, like the reference's: it consumes the mesh-error estimate the GPU produces (k_mesh_error) and the state
// part of the NLP solution and decides, per mesh interval, between keeping / reducing the polynomial degree, merging
// with the neighbour, raising the degree, or dividing the interval.  The method is STATEFUL: it compares the current
// solution with the previous mesh, its error estimates and the previous solution, so one LiuRefiner lives in the
// handle and is reset by lpb_refine_reset.
//
// Reference behaviour that is kept on purpose (the goldens come from the reference's own code):
//  * adjacent intervals that both meet the tolerance and ask for the same node count are merged unconditionally: the
//    verdict of Merging_mesh is computed and then ignored (:187-205), so its arithmetic is not restated;
//  * the smoothness test interpolates with nodes mapped to [-1, 1] but evaluates at abscissae in global tau units
//    (:722-731), and indexes the previous STATE rows with mesh-POINT indices (:706-707);
//  * the coefficient test of Reducing_N compares signed coefficients with the tolerance (:492).
// Fenced: where the reference would convert a non-finite or negative double to an unsigned count (undefined
// behaviour) or index out of range, the interval is divided in two / the degree is kept.
#pragma once
#include <vector>

namespace lpb {

struct LiuPhaseInput {
    int ns = 0;
    std::vector<double> mesh;   // K + 1 mesh points of the current mesh
    std::vector<int> nodes;     // K node counts
    std::vector<double> tau;    // N composite LGR points
    std::vector<double> rel;    // relative error, column-major rows x ns, rows = sum(N_k + 1) + 1 (lpb_mesh_error)
    std::vector<double> state;  // (N + 1) x ns, column-major: the state part of the NLP solution
};

class LiuRefiner {
public:
    void reset();
    // returns true when no interval changes (NoMoreRefine); the new mesh of every phase in mesh_out / nodes_out
    bool refine(const std::vector<LiuPhaseInput>& in, double tol, int Nmax, double ratio_R, std::vector<std::vector<double>>& mesh_out,
                std::vector<std::vector<int>>& nodes_out);
    int calls() const { return mesh_index_; }

private:
    struct MeshInfo {
        std::vector<double> mesh, e_k;
        std::vector<int> nodes;
    };
    struct Dense { // column-major
        int rows = 0, cols = 0;
        std::vector<double> a;
        double at(int i, int j) const { return a[(size_t)i + (size_t)j * rows]; }
    };
    int mesh_index_ = 0;
    std::vector<std::vector<MeshInfo>> mesh_history_;      // per call, per phase
    std::vector<std::vector<Dense>> state_history_;        // per call, per phase: (N + 1) x ns
    std::vector<std::vector<std::vector<double>>> mesh_points_history_;

    static const Dense& power_coefficients(int N);
    int reducing_N(const LiuPhaseInput& ph, int seg, const std::vector<int>& state_index, const std::vector<double>& betai, double tol) const;
    bool growth_exponent(int iphase, const LiuPhaseInput& ph, int seg, double e_k, double& q) const;
    bool can_increase_N(int iphase, const LiuPhaseInput& ph, int seg, const std::vector<int>& state_index, double ratio_R) const;
    static void second_derivative(const std::vector<double>& t, const Dense& x, std::vector<double>& interp_t, Dense& d2);
};

// barycentric Lagrange interpolation as SolutionErrorChecker::BarLagrangeInterp does it (LpSolutionError.cpp:10-52)
void bar_lagrange_interp(const std::vector<double>& data_x, const double* data_y, const std::vector<double>& x, std::vector<double>& y);

} // namespace lpb
