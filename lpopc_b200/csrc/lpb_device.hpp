// lpb_device.hpp -- device-side description of one transcribed problem and the
// functor-set dispatch table.  Shared by lpb_api.cu (handle, structure kernels) and
// the per-problem kernel instantiations (lpb_kernels.cuh).
//
// HBM layout (all fp64 unless noted; see DESIGN.md "Data layout"):
//   x[b][n]       IPOPT variable vector per instance, reference layout:
//                 per phase: state j at c_p + j(N+1) + k (N nodes + endpoint), control j at
//                 c_p + ns(N+1) + jN + k, then t0, tf.   Thread k reads x[.. + k]: coalesced.
//   g[b][m]       constraints, reference layout (A.2): defects state-major, paths, events,
//                 links, linear rows.
//   jac[b][nnz]   IPOPT triplet values in the reference's order [NL | L | C] (A.4); every
//                 (row-block, colour) is N contiguous doubles -> a warp stores 256 B runs.
//   tables        tau[N], w[N], ddiag[N], node->interval map, per-interval dense column-major
//                 D blocks, compacted Doffdiag values (constant Jacobian segment).
// ProblemDev is passed BY VALUE to every kernel (__grid_constant__): it lands in the
// constant bank, so per-phase offsets are uniform loads and no device copy is kept.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include "lpb_structure.hpp"

namespace lpb {

struct PhaseDev {
    int N, K;           // nodes, intervals
    int ne;             // events in this phase
    int node0;          // first global node index of this phase (over all phases)
    int var0, con0;     // c_p, r_p: first variable / constraint of the phase (0-based)
    int ndoff;          // Doffdiag entries of the phase (after exact-zero removal)
    int nblkH;          // number of N-long blocks in the xx/ux/uu part of the Hessian I-part
    long long nl0;      // first NL Jacobian value of the phase
    long long ev0;      // first event-row Jacobian value of the phase (inside NL)
    long long c0;       // first constant (Doffdiag) Jacobian value of the phase
    long long hI0, hE0; // first Hessian value of the I-part / E-part of the phase
    const double* tau;
    const double* w;
    const double* ddiag;      // first N values of the compacted Diag COO (LpNLPWrapper.cpp:704-711)
    const int* node_interval; // [N]
    const int* int_row0;      // [K]
    const int* int_n;         // [K]
    const long long* int_d0;  // [K]
    const double* dblocks;
    const double* doff_vals;  // [ndoff]
    const int* hblk;          // [(ns+nc)^2] block index of (row var, col var) in the Hessian I-part, -1 if absent
    const HessRow* hrow;      // [ns+nc] the same per row as (presence mask, first block): one load per row (ns+nc <= 64)
};

struct LinkDev {
    int left, right; // 0-based phases
    int nl;          // links in the pair
    int con0;        // first constraint row of the pair
    int lam0;        // first multiplier row used by the Hessian (link_indices, quirk Q6)
    int pad;
    long long val0;  // first Jacobian value of the pair
    long long h0;    // first Hessian value of the pair
};

struct ProblemDev {
    int P, Lp;
    int n, m, nnz_jac, nnz_h;
    int total_nodes;
    int ns, nc, np;     // functor-set sizes (shared by all phases)
    int lin_con0;       // first linear constraint row
    int analytic;       // first-derive = analytic
    long long lin_val0; // first L value in the Jacobian
    long long ctot;     // total constant (C) Jacobian values per instance
    double tol;         // finite-difference-tol
    const HessEntry* eent; // E-part entry table (same for all phases: sizes are shared)
    const HessEntry* lent; // link-part entry table
    int n_eent, n_lent;
    PhaseDev ph[kMaxPhases];
    LinkDev lk[kMaxLinks];
};

// mesh-error estimator tables of one phase / all phases (lpb_mesherr.cuh)
struct MeshErrPhase {
    int K, M;                 // intervals, new points in total (sum of N_k + 1)
    const int* int_m;         // [K] N_k + 1
    const int* int_rn0;       // [K] first new row of the interval
    const long long* int_a0;  // [K] offset of A_k in ablocks
    const double* tnew;       // [M]
    const double* ablocks;
    long long out0;           // offset of the phase in the (M+1) x ns output matrices (doubles)
};
struct MeshErrDev {
    MeshErrPhase ph[kMaxPhases];
};

struct LaunchOpts {
    int block;        // threads per block for node kernels
    int colour_split; // colour chunks (gridDim.y) of the Jacobian kernel; 0 = auto
    int pair_split;   // pair chunks of the Hessian kernel; 0 = auto
    int sm_count;
    int skip_const;     // 1: do not write the constant tail [L | C] of the Jacobian values (the host-pointer
                        // entry points fill it in the caller's array from a cached copy instead of moving it
                        // over PCIe on every call)
    int stage_values;   // Jacobian kernel: 1 = shared-memory staged write-out where it applies, 0 = per-thread scatter (option "stage_values")
    int no_rotate;      // 1: do not rotate the thread -> node map of the Jacobian kernel (A/B measurement)
    int unroll_colours; // Jacobian kernel variant: -1 = functor default (P::UNROLL_COLOURS), 0 = colour loop, 1 = unrolled;
                        // Hessian node kernel: 0 = generic pair loops, otherwise the tiled kernel where the functor set has HESS_DEP
    int sweep_mode;     // Jacobian kernel of functor sets with sweep hooks: 0 = row-parallel sweep where available, 1 = per-thread sweep
    int hess_variant;   // tuning builds (-DLPB_HESS_VARIANTS): (tile size, CTAs/SM) instantiation of the tiled Hessian kernel
    // optional CUDA events recorded on the launch stream right before / after the dominant
    // node kernel (bench.py's live roofline measurement); null = no timing
    cudaEvent_t ev_begin, ev_end;
};

// Dispatch table of one functor set (filled by LPB_DEFINE_FUNCTOR in lpb_kernels.cuh).
// All launchers are asynchronous on `st` and return the number of kernels launched
// (negative: cudaError_t).
struct FunctorVTable {
    const char* name;
    int NS, NC, NPATH, NE_MAX, NL_MAX;
    int consts_doubles;
    int has_analytic;
    int (*cons_jac)(const ProblemDev& pd, const void* consts, cudaStream_t st, const LaunchOpts& o,
                    int nbatch, const double* x, double* g, double* vals);
    int (*objective)(const ProblemDev& pd, const void* consts, cudaStream_t st, const LaunchOpts& o,
                     int nbatch, const double* x, double* f, double* scratch);
    int (*gradient)(const ProblemDev& pd, const void* consts, cudaStream_t st, const LaunchOpts& o,
                    int nbatch, const double* x, double* grad, double* scratch);
    int (*hessian)(const ProblemDev& pd, const void* consts, cudaStream_t st, const LaunchOpts& o,
                   int nbatch, const double* x, const double* sigma, const double* lambda, double* vals, double* scratch);
    int (*probe)(const ProblemDev& pd, const void* consts, cudaStream_t st, const double* x, int* dep_out);
    size_t (*scratch_doubles)(const ProblemDev& pd, int nbatch);
    int (*mesh_error)(const ProblemDev& pd, const void* consts, cudaStream_t st, const MeshErrDev& me, int total_intervals, int max_n,
                      const double* x, double* tem, double* abs_err);
    // NLP solution -> optimal-control solution (lpb_convert.cuh); out == nullptr: layout query only
    int (*nlp2op)(const ProblemDev& pd, const void* consts, cudaStream_t st, const double* x, const double* lambda,
                  double* out, double* scratch, long long* phase_offsets);
    // declared dependency masks of the functor set (P::HESS_DEP, NS + NPATH + 1 entries) or null
    const unsigned long long* hess_dep;
};

const FunctorVTable* const* functor_registry(int* count);

} // namespace lpb
