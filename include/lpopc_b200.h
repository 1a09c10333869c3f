/* lpopc_b200.h -- C ABI of the B200-native Radau-pseudospectral transcription.
 *
 * Drop-in boundary for lpopc's NLP evaluation path.  The entry points below are
 * what a binding to the reference's `Ipopt::TNLP` adapter would call; each one
 * cites the reference interface it replaces.  Plain pointers and sizes only: no
 * C++ types, no exceptions across the ABI.  Every function returns LPB_OK (0) or
 * a negative error code; lpb_last_error() gives the message.
 *
 * Conventions copied from the reference boundary:
 *   - the caller owns every array; the callee fills in place;
 *   - triplets are 0-based (TNLP::C_STYLE, LpopcIpopt.cpp:22), duplicates are
 *     part of the contract (IPOPT sums them);
 *   - eval_jac_g / eval_h with values == NULL return the structure
 *     (LpopcIpopt.cpp:156-164, :187-195);
 *   - m includes the linear rows (LpopcIpopt.cpp:14);
 *   - one handle = one host thread (the reference is not re-entrant either).
 * The hot path has no CPU fallback: if no CUDA device is usable, lpb_create
 * fails with LPB_ERR_CUDA.
 */
#ifndef LPOPC_B200_H
#define LPOPC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define LPB_OK 0
#define LPB_ERR_INVALID (-1)     /* bad argument / inconsistent description      */
#define LPB_ERR_UNSUPPORTED (-2) /* e.g. nparameters > 0     */
#define LPB_ERR_CUDA (-3)        /* CUDA runtime error / no device               */
#define LPB_ERR_STATE (-4)       /* call order (mesh not set, ...)               */
#define LPB_ERR_UNKNOWN_FUNCTOR (-5)

#define LPB_DERIVE_FINITE_DIFFERENCE 0 /* "first-derive" option, LpOptDerive.hpp:27-37 */
#define LPB_DERIVE_ANALYTIC 1

/* One phase: mirrors Lpopc::Phase (LpOptimalProblem.hpp:30-240).  State limits
 * carry three values per state like Lpopc::Limit (initial node, interior nodes,
 * terminal point; LpBoundsChecker.cpp:51-75). */
typedef struct lpb_phase_desc {
    int nstates, ncontrols, nparameters, npaths, nevents;
    const double* state_min0; const double* state_min; const double* state_minf; /* [nstates] */
    const double* state_max0; const double* state_max; const double* state_maxf; /* [nstates] */
    const double* control_min; const double* control_max;                         /* [ncontrols] */
    const double* path_min; const double* path_max;                               /* [npaths] */
    const double* event_min; const double* event_max;                             /* [nevents] */
    double t0_min, t0_max, tf_min, tf_max; /* Phase::SetTimeMin/Max */
    int has_duration;                      /* Phase::SetDuration   */
    double duration_min, duration_max;
} lpb_phase_desc;

/* One linkage pair: mirrors Lpopc::Linkage (LpOptimalProblem.hpp:242-281);
 * left_phase/right_phase are 1-based like the Linkage constructor. */
typedef struct lpb_link_desc {
    int left_phase, right_phase;
    int nlinks;
    const double* link_min; const double* link_max; /* [nlinks] */
} lpb_link_desc;

/* Whole problem: mirrors Lpopc::OptimalProblem + the options that reach the hot
 * path ("finite-difference-tol", "first-derive"; LpOptDerive.hpp:27-37). */
typedef struct lpb_problem_desc {
    const char* functor; /* name of a registered functor set (include/problems/) */
    int nphases;
    const lpb_phase_desc* phases;
    int nlinkpairs;
    const lpb_link_desc* links;
    const double* consts; /* functor constants, sizeof(Functor::Consts)/8 doubles */
    int nconsts;
    double fd_tol;    /* 0 -> reference default 1e-6 */
    int first_derive; /* LPB_DERIVE_* */
} lpb_problem_desc;

typedef struct lpb_handle lpb_handle;

/* Replaces: LpopcAlgorithm::Initialized object wiring (LpLpopcAlgorithm.cpp:157-246). */
int lpb_create(const lpb_problem_desc* desc, lpb_handle** out);
int lpb_destroy(lpb_handle* h);
const char* lpb_last_error(const lpb_handle* h); /* h may be NULL: last create error */

/* Replaces: Phase::SetMeshPoints / SetNodesPerInterval + MeshRefiner::SetAndCheckMesh
 * (LpMeshRefiner.cpp:10-61).  phase is 0-based; meshpoints has K+1 entries spanning
 * [-1, 1]; nodes_per_interval has K entries (each >= 2). */
int lpb_set_mesh(lpb_handle* h, int phase, int K, const double* meshpoints, const int* nodes_per_interval);

/* Replaces: GetSizes/GetBounds/GetGuess PS-table fill + RefreshSparsity for a new
 * mesh (LpLpopcAlgorithm.cpp:36-45, LpGuessChecker.cpp:110-122, LpNLPWrapper.hpp:89).
 * Uploads the LGR tables and rebuilds every index map / triplet array on the GPU.
 * Called implicitly by lpb_get_nlp_info when the mesh changed. */
int lpb_refresh(lpb_handle* h);

/* Replaces: LpopcIpopt::get_nlp_info (LpopcIpopt.cpp:11-25). */
int lpb_get_nlp_info(lpb_handle* h, int* n, int* m, int* nnz_jac_g, int* nnz_h_lag);
/* Replaces: LpopcIpopt::get_bounds_info (LpopcIpopt.cpp:27-83). */
int lpb_get_bounds_info(lpb_handle* h, double* x_l, double* x_u, double* g_l, double* g_u);

/* Replaces: LpopcIpopt::eval_f -> NLPWrapper::GetObjFun (LpopcIpopt.cpp:106, LpNLPWrapper.cpp:863). */
int lpb_eval_f(lpb_handle* h, const double* x, double* obj_value);
/* Replaces: eval_grad_f -> GetObjGrad (LpopcIpopt.cpp:118, LpNLPWrapper.cpp:940). */
int lpb_eval_grad_f(lpb_handle* h, const double* x, double* grad_f);
/* Replaces: eval_g -> GetAllCons (LpopcIpopt.cpp:135, LpNLPWrapper.cpp:34). */
int lpb_eval_g(lpb_handle* h, const double* x, double* g);
/* Replaces: eval_jac_g -> GetConsSparsity / GetConsJacbi (LpopcIpopt.cpp:152,
 * LpNLPWrapper.cpp:1550, :230).  values == NULL -> fill iRow/jCol only. */
int lpb_eval_jac_g(lpb_handle* h, const double* x, int* iRow, int* jCol, double* values);
/* Replaces: eval_h -> GetHessainSparsity / GetHessainValue (LpopcIpopt.cpp:183,
 * LpNLPWrapper.cpp:2354-2376, LpHessian.cpp:878, :2510). */
int lpb_eval_h(lpb_handle* h, const double* x, double obj_factor, const double* lambda,
               int* iRow, int* jCol, double* values);
/* Fused eval_g + eval_jac_g(values) in one pass (the headline "defect+Jacobian"
 * evaluation; one H2D of x, one D2H of g and values). */
int lpb_eval_g_jac(lpb_handle* h, const double* x, double* g, double* values);

/* Replaces: reading LpCalculateData::PS[phase]->Points / Weights (LpCalculateData.hpp:35-41),
 * the composite LGR nodes on [-1,1) and quadrature weights of the current mesh (what
 * LpGuessChecker.cpp:130-190 interpolates the user's guess onto); N doubles each, either may be NULL. */
int lpb_get_lgr_tables(lpb_handle* h, int phase, double* points, double* weights);

/* Replaces: DeriveDependicieshecker::GetDependiciesForJacobiInEveryPhase
 * (LpDerivDependciesChecker.cpp:10-94): NaN-probe of the dae functor at node 1 of
 * x_guess; the mask feeds the Hessian pattern only.  If it
 * is never called the mask is all-ones.  dep_out (optional) receives, per phase,
 * (ns+np) x (ns+nc) 0/1 values, column-major, phases concatenated. */
int lpb_probe_dependencies(lpb_handle* h, const double* x_guess, int* dep_out);

/* ---- mesh-error estimate and ph refinement (the step between two NLP solves) ----
 * Replaces: SolutionErrorChecker::CheckSolutionDiffError (LpSolutionError.cpp:112-166) -- interpolation of
 * the NLP solution x onto one more LGR point per interval, dae() there and the integration defect run on the
 * GPU.  rows_out[p] = sum_k (N_k + 1) + 1; rel_err: per phase a rows x nstates matrix, column-major,
 * phases concatenated; interval_max: per phase K values (the quantity PhMeshRefineAlg compares with
 * "desired-relative-error").  Any output may be NULL. */
int lpb_mesh_error(lpb_handle* h, const double* x, int* rows_out, double* rel_err, double* interval_max);
/* The same with x and the results on the device (either output may be null); asynchronous on the handle's stream. */
int lpb_mesh_error_dev(lpb_handle* h, const double* d_x, double* d_rel_err, double* d_interval_max);
/* Replaces: PhMeshRefineAlg::RefineMesh / ModifySegment (LpPhMeshRefineAlg.cpp:12-99; options
 * "desired-relative-error", "Nmax", "Nmin" of LpMeshRefiner.h:67-80).  Returns the refined mesh of every phase:
 * K_out[p], then K+1 mesh points and K node counts per phase appended to mesh_out / nodes_out, whose capacities
 * (in elements) are mesh_cap / nodes_cap.  The number of sub-intervals grows with the error (ceil((N_k +
 * log(e/tol)/log(N_k)) / Nmin) per interval), so a call whose result does not fit fails with LPB_ERR_INVALID after
 * filling K_out: size the arrays as sum(K) + nphases and sum(K) and call again.  A non-finite error estimate
 * (diverged solution) is rejected.  Pass the result to lpb_set_mesh. */
int lpb_refine_mesh_ph(lpb_handle* h, const double* x, double tol, int Nmax, int Nmin, int* no_more_refine,
                       int* K_out, double* mesh_out, int mesh_cap, int* nodes_out, int nodes_cap);
/* Replaces: LiuHpMeshRefineAlg::RefineMesh (LpLiuHpMeshRefineAlg.cpp:12-260; "mesh-refine-methods" = "hp-Liu", options
 * "desired-relative-error", "Nmax", "R" of LpMeshRefiner.h:54-61,67-80).  Per interval: keep / reduce the degree or merge
 * with the neighbour where the error estimate meets tol, otherwise raise the degree where the solution is smooth (ratio
 * of second derivatives against the previous grid <= ratio_R) and divide the interval where it is not.  The method
 * compares with the previous grid, its error estimates and its solution: that history lives in the handle, the first
 * call after lpb_create / lpb_refine_reset is "grid 0" (three more nodes where the error is too large), and the mesh
 * passed to lpb_set_mesh between two calls must be the one the previous call returned.  Outputs and capacities as in
 * lpb_refine_mesh_ph. */
int lpb_refine_mesh_hp_liu(lpb_handle* h, const double* x, double tol, int Nmax, double ratio_R, int* no_more_refine,
                           int* K_out, double* mesh_out, int mesh_cap, int* nodes_out, int nodes_cap);
int lpb_refine_reset(lpb_handle* h); /* forget the hp-Liu history (a new solve sequence on this handle) */

/* ---- NLP solution -> optimal-control solution (the step right after the solve) ----
 * Replaces: Nlp2OpConverter::Nlp2OpControl (Nlp2OPConverter.cpp:13-196).  From the NLP solution x and the
 * constraint multipliers lambda (IPOPT's sign, as in finalize_solution, LpopcIpopt.cpp:220) the GPU builds,
 * per phase with N nodes (M = N + 1 rows, column-major), phases concatenated in `out`:
 *   time[M] | state[M x ns] | control[M x nc] | costate[M x ns] | pathmult[M x np] | Hamiltonian[M] | mayer, lagrange
 * (control / pathmult end rows by the reference's natural cubic spline, terminal costate through the last
 * column of D).  lpb_nlp2op_length returns the number of doubles `out` must hold and, if phase_offsets is
 * given, the P + 1 phase offsets into it.  total_cost (optional) = Data_->optcontrol_cost. */
long long lpb_nlp2op_length(lpb_handle* h, long long* phase_offsets);
int lpb_nlp2op(lpb_handle* h, const double* x, const double* lambda, double* out, double* total_cost);
/* The same with x, lambda and the converted solution (lpb_nlp2op_length doubles) on the device; asynchronous. */
int lpb_nlp2op_dev(lpb_handle* h, const double* d_x, const double* d_lambda, double* d_out);

/* ---- batched independent instances (MPC-style; BASELINE config 4) ----------
 * nbatch instances share problem, mesh, tables and pattern; instance b uses
 * x[b*n .. b*n+n).  Host-pointer versions copy H2D/D2H around the kernels. */
int lpb_eval_f_batch(lpb_handle* h, int nbatch, const double* x, double* obj_values);
int lpb_eval_grad_f_batch(lpb_handle* h, int nbatch, const double* x, double* grad_f);
int lpb_eval_g_jac_batch(lpb_handle* h, int nbatch, const double* x, double* g, double* values);
/* lpb_eval_g_jac_batch moves as few bytes over PCIe as the result allows: the mesh-constant tail [L | C]
 * of the values is written by host threads from a cached copy, and -- when `values` is pinned host memory
 * and the batch is large -- only the (row block, column block) segments of [NL] that have ever been
 * non-zero are sent by the GPU (zero-copy stores); the all-zero segments of the reference's forced-dense
 * pattern are verified on the device on every call and written as zeros by the host threads (a segment
 * that turns non-zero is fetched afterwards, so the caller always gets the exact device values).
 * Options "sparse_return" (default 1), "host_fill_const" (1), "host_threads" (0 = auto).
 * Option "auto_pin" (default 0): page-lock large PAGEABLE caller arrays (x, g, values, lambda, grad) that come
 * back with the same address on a second call, as IPOPT's do -- a pageable array is copied through a staging
 * buffer at a few GB/s, a page-locked one moves at PCIe speed and is eligible for the sparse return.  The caller
 * must keep such arrays alive until lpb_destroy or until the option is set back to 0 (which releases them). */
int lpb_eval_h_batch(lpb_handle* h, int nbatch, const double* x, const double* obj_factor,
                     const double* lambda, double* values);

/* ---- device-resident entry points (pointers are CUDA device pointers) ------
 * Asynchronous on the handle's stream; no host synchronisation.  g / values /
 * grad may be NULL to skip that output. */
int lpb_set_stream(lpb_handle* h, void* cuda_stream);
int lpb_eval_f_dev(lpb_handle* h, int nbatch, const double* d_x, double* d_obj);
int lpb_eval_grad_f_dev(lpb_handle* h, int nbatch, const double* d_x, double* d_grad);
int lpb_eval_g_jac_dev(lpb_handle* h, int nbatch, const double* d_x, double* d_g, double* d_values);
int lpb_eval_h_dev(lpb_handle* h, int nbatch, const double* d_x, const double* d_obj_factor,
                   const double* d_lambda, double* d_values);
int lpb_structure_dev(lpb_handle* h, const int** d_jac_iRow, const int** d_jac_jCol,
                      const int** d_h_iRow, const int** d_h_jCol);

/* ---- linear algebra of the batched outer solver (lpopc_b200/solver.py) ----
 * Stands where the reference has IPOPT's sparse linear solver behind NLPSolver::SolveNlp (LpNLPSolver.cpp:13-54,
 * IpoptApplication::OptimizeTNLP :45): for a BATCH of instances the KKT step is one structured factorisation per
 * instance on the GPU instead of one host solve at a time.
 * Batched block-tridiagonal positive definite systems over the mesh intervals: B instances, K diagonal blocks of
 * nb x nb, off-diagonal coupling only through the nbd boundary slots bnd[] (ascending) of the next block.
 * lpb_blocktri_factor: Dp [B][K][nb][nb] (symmetric, lower triangle read), Ep [B][K-1][nbd][nb] ->
 *   L (like Dp, lower triangles), C = E L^-T (like Ep), info[B] (0, or 1 + index of the first non-positive pivot:
 *   the caller's inertia test).  One launch, one CTA per instance.
 * lpb_blocktri_solve: one launch for the whole forward/backward substitution; L and C are HOST arrays of K and
 *   K-1 device pointers with instance strides strideL / strideC (doubles); rhs/out [B][K][nb].
 * Everything else is a device pointer; asynchronous on `cuda_stream`.  Return 0, -1 if the shape is not
 * supported (caller falls back to library routines), or a CUDA error code - 1000. */
int lpb_blocktri_factor(int B, int K, int nb, int nbd, const double* Dp, const double* Ep, const int* bnd, double* L, double* C,
                        int* info, void* cuda_stream);
/* Assembly + factorisation of the block-tridiagonal KKT matrix of the batched interior-point step in one launch
 * (k_kkt_factor, lpb_blocktri.cu): A_i = D_i + diag(diag_i + dw) + gamma Jb_i^T Jb_i with the boundary coupling of the
 * neighbouring intervals, L_i L_i^T = A_i, C_i = E_i' L_i^-T, and the inertia-correction retry (larger dw until every
 * pivot is positive) per instance.  Device pointers; D [B][K][nb][nb], E [B][max(K-1,1)][nbd][nb], Jb [B][K][mr][nb+nbd],
 * diag [B][K][nb], active [B] bytes (0: converged instance, identity factors are written and its info is 0), dw [B]
 * in/out, L / C like D / E, info [B].  Returns 0, -1 (shape not supported) or a negative cudaError_t - 1000.
 * No counterpart in the reference (IPOPT + MUMPS do this on the host). */
int lpb_kkt_factor(int B, int K, int nb, int nbd, int mr, double gamma, const double* D, const double* E, const double* Jb, const double* diag,
                   const int* bnd, const unsigned char* active, double* dw, double* L, double* C, int* info, void* stream);
int lpb_blocktri_solve(int B, int K, int nb, int nbd, const double* const* L, const double* const* C, long long strideL,
                       long long strideC, const int* bnd, const double* rhs, double* out, void* cuda_stream);
/* lpb_blocktri_solve with a per-instance mask: active [B] bytes on the device (or null); instances with 0 are skipped
 * and get a zero solution. */
int lpb_blocktri_solve_masked(int B, int K, int nb, int nbd, const double* const* L, const double* const* C, long long strideL,
                              long long strideC, const int* bnd, const unsigned char* active, const double* rhs, double* out,
                              void* cuda_stream);

/* Batched sparse matrix-vector product y_b = A_b x_b (+ diag_b .* x_b) for matrices that share one CSR structure and
 * take their values in place from per-instance triplet arrays: vals [B][val_stride], entry e of the CSR structure is
 * vals[b][perm[e]].  The batched interior-point step multiplies with the Jacobian, its transpose and the Hessian this
 * way (lpopc_b200/solver.py).  Device pointers; asynchronous on `stream`; returns 0, -1 (bad argument) or a negative
 * cudaError_t - 1000.  No counterpart in the reference (IPOPT does this on the host). */
int lpb_batched_spmv(int B, int nrows, int ncols, const int* rowptr, const int* col, const int* perm, const double* vals,
                     long long val_stride, const double* x, const double* diag, double* y, void* stream);

/* Tuning / introspection (not part of the reference boundary).  Options of lpb_set_option_int (default in brackets;
 * setting any option drops the captured graphs of the single-problem fast path, which are rebuilt on the next call):
 *   host-pointer batch calls (lpb_eval_g_jac_batch):
 *     sparse_return [1]      only (row block, column block) segments that are not all-zero cross PCIe; the rest is
 *                            verified on the device and written by host threads
 *     persistent_values [0]  the caller hands the SAME values array to consecutive calls and leaves it alone in between:
 *                            the host writes nothing after the first call (two sentinels per instance are re-checked)
 *     auto_pin [0]           page-lock large caller arrays that come back with the same address; the caller keeps them
 *                            alive until lpb_destroy
 *     return_mode [0]        0: zero-copy stores of the on-segments, 1: one strided copy-engine transfer per on-run (slower)
 *     e2e_chunks [8]         pipeline depth of the call;  host_threads [0 = half the hardware threads, at most 16]
 *     host_fill_const [1]    constant tail [L | C] written by the host from a cached copy instead of crossing PCIe
 *     sparse_forget          (write-only) forget the learned segment mask
 *   single-problem TNLP calls:
 *     fast_path [-1]         -1: on for n + m + nnz_jac <= 262144, 0: off, 1: on -- one captured graph per new x
 *   kernels:
 *     unroll_colours [-1], colour_split [0 = auto], pair_split [0 = auto], block [128], rotate_nodes [1], stage_values [0],
 *     sweep_mode [0: row kernel where the functor set has the hooks, 1: per-thread sweep, 2/3: row-kernel shapes],
 *     hess_variant (development builds), time_kernels [0] (CUDA events around the node kernels, read by lpb_kernel_time),
 *     debug_skip (bench A/B switches) */
int lpb_set_option_int(lpb_handle* h, const char* name, int value);
long long lpb_kernel_launch_count(const lpb_handle* h); /* kernels launched so far */
/* Counters: "sparse_calls" (host-pointer calls that used the sparse return), "sparse_fixups" (segments fetched
 * after a wrong all-zero prediction), "sparse_on_doubles" (values per instance that cross PCIe on that path),
 * "head_doubles" (size of [NL] per instance), "pinned_buffers" (caller arrays page-locked by "auto_pin"). */
int lpb_get_stat(lpb_handle* h, const char* name, long long* value);
/* With option "time_kernels" = 1 every evaluation brackets its dominant node kernel
 * ("cons_jac": k_cons_jac, "hess_nodes": k_hess_nodes) with CUDA events on the handle's
 * stream; this returns and resets the accumulated device time and launch count. */
int lpb_kernel_time(lpb_handle* h, const char* kernel, double* total_ms, int* count);
/* Device self-test: n pseudo-random (numerator, divisor) pairs through the shared-reciprocal
 * forward-difference quotient of the Jacobian kernels vs IEEE division; *mismatches must be 0. */
int lpb_selftest_fd_division(long long n, unsigned long long seed, long long* mismatches);
int lpb_num_functors(void);
const char* lpb_functor_name(int i);

#ifdef __cplusplus
}
#endif
#endif /* LPOPC_B200_H */
