// lpb_blocktri.cu -- batched block-tridiagonal triangular solves for the batched outer solver
// (lpopc_b200/solver.py::BlockTridiagKKT).
//
// The KKT step of every interior-point iteration factors, per instance, a block-tridiagonal positive definite
// matrix over the mesh intervals: diagonal Cholesky factors L_i (nb x nb, lower) and boundary-row couplings
// C_i = E_i L_{i-1}^-T (nbd x nb; only the first node of interval i+1 couples to interval i).  Each of the
// 1 + refine solves per iteration is then
//     forward   y_i = L_i^-1 (r_i - scatter_bnd(C_{i-1} y_{i-1}))          i = 0 .. K-1
//     backward  x_i = L_i^-T (y_i - C_i^T x_{i+1}[bnd])                     i = K-1 .. 0
// i.e. 2K dependent single-vector triangular solves per instance.  Through the library that is 2K launches whose
// cost does not shrink with the batch (and grows per matrix for small batches: the straggler tail of the
// lockstep iteration).  Here ONE launch does the whole solve: one CTA per instance, thread i owns row i of the
// right-hand side, 32-row panels eliminated with register shuffles and one barrier per panel.
#include <cuda_runtime.h>
#include <mutex>
#include <vector>
#include <cstddef>
#include <cstdint>

namespace {

constexpr int kMaxBlocks = 64;

struct BlockTriArgs {
    const double* L[kMaxBlocks]; // K factors, each [B][nb][nb] row-major with instance stride sL (lower triangle read)
    const double* C[kMaxBlocks]; // K-1 couplings, each [B][nbd][nb] row-major with instance stride sC
    const int* bnd;              // [nbd] boundary slots of a block
    const unsigned char* active; // [B] or null: instances with 0 are skipped (zeros written)
    const double* rhs;           // [B][K][nb]
    double* out;                 // [B][K][nb]
    long long sL, sC;            // doubles between consecutive instances of one factor / coupling
    int B, K, nb, nbd;
};

// Substitution by 32-row panels, factors read straight from global memory (L2): thread r owns row r of the
// right-hand side and, per panel, holds the 32 entries of ITS row (forward) or column (backward) of that panel
// in registers.  The warp that owns the panel's rows eliminates the 32 x 32 diagonal block with register
// shuffles (no barrier inside), publishes the panel's solution in shared memory, and after ONE barrier every
// remaining row folds the panel in with 32 multiply-adds (four independent chains).  Shared memory holds only the
// solution vectors ((K + 1) nb doubles), so occupancy is set by registers, not by a staged factor.
__global__ void __launch_bounds__(256, 2)
k_blocktri_solve(const __grid_constant__ BlockTriArgs a)
{
    extern __shared__ double sm[];
    const int nb = a.nb, nbd = a.nbd, K = a.K;
    double* ys = sm;                          // [K][nb] forward results
    double* piv = ys + (size_t)K * nb;        // [nb, padded to whole panels] right-hand side scratch / backward solution of the current interval
    const int b = blockIdx.x, tid = threadIdx.x;
    if (a.active != nullptr && a.active[b] == 0) { // block-uniform
        for (int e = tid; e < K * nb; e += blockDim.x) a.out[(size_t)b * K * nb + e] = 0.0;
        return;
    }
    const int warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
    const int p0 = warp * 32;                 // first row of this warp's panel (thread = row)
    const unsigned full = 0xffffffffu;
    const bool live = tid < nb;
    if (!live) piv[tid] = 0.0; // padding of the last panel: read (times a zero factor entry) by the backward updates

    // ---- forward: L y = r ----
    for (int i = 0; i < K; ++i) {
        const double* __restrict__ Lg = a.L[i] + (size_t)b * a.sL;
        const double* __restrict__ Lrow = Lg + (size_t)(live ? tid : 0) * nb;
        double t = live ? a.rhs[((size_t)b * K + i) * nb + tid] : 0.0;
        if (i > 0) { // r_i[bnd] -= C_{i-1} y_{i-1}: warp per boundary row, coalesced dot product
            if (live) piv[tid] = t;
            __syncthreads();
            const double* __restrict__ Cg = a.C[i - 1] + (size_t)b * a.sC;
            const double* __restrict__ yp = ys + (size_t)(i - 1) * nb;
            for (int q = warp; q < nbd; q += nwarp) {
                double acc = 0.0;
                for (int c = lane; c < nb; c += 32) acc += Cg[(size_t)q * nb + c] * yp[c];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(full, acc, off);
                if (lane == 0) piv[a.bnd[q]] -= acc;
            }
            __syncthreads();
            if (live) t = piv[tid];
        }
        double* __restrict__ y = ys + (size_t)i * nb;
        for (int q0 = 0; q0 <= p0 && q0 < nb; q0 += 32) { // panels up to and including this warp's own
            double seg[32]; // L[row][q0 .. q0+31] (entries right of the diagonal are never used)
#pragma unroll
            for (int j = 0; j < 32; ++j) seg[j] = (live && q0 + j <= tid) ? Lrow[q0 + j] : 0.0;
            if (p0 == q0) { // owner warp: 32 x 32 diagonal block in registers
                double dg = 1.0; // own diagonal entry L[row][row] = seg[lane]
#pragma unroll
                for (int j = 0; j < 32; ++j) dg = (lane == j) ? seg[j] : dg;
                const double rdg = live ? 1.0 / dg : 0.0;
                double mine = 0.0;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const double yj = __shfl_sync(full, t, j) * __shfl_sync(full, rdg, j);
                    mine = (lane == j) ? yj : mine;
                    if (lane > j) t -= seg[j] * yj;
                }
                if (live) y[tid] = mine;
                __syncthreads(); // matches the barrier the lower warps wait at for this panel
            } else {
                __syncthreads(); // panel q0 published by its owner
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    s0 += seg[j] * y[q0 + j];
                    s1 += seg[j + 1] * y[q0 + j + 1];
                    s2 += seg[j + 2] * y[q0 + j + 2];
                    s3 += seg[j + 3] * y[q0 + j + 3];
                }
                t -= (s0 + s1) + (s2 + s3);
            }
        }
        // warps above the last panels still owe the barriers of the panels after their own
        for (int q0 = p0 + 32; q0 < nb; q0 += 32) __syncthreads();
        __syncthreads(); // y_i complete before the next interval's coupling reads it
    }
    // ---- backward: L^T x = y ----
    for (int i = K - 1; i >= 0; --i) {
        const double* __restrict__ Lg = a.L[i] + (size_t)b * a.sL;
        double t = live ? ys[(size_t)i * nb + tid] : 0.0;
        if (i + 1 < K && live) { // y_i -= C_i^T x_{i+1}[bnd]: thread = column, coalesced over columns
            const double* __restrict__ Cg = a.C[i] + (size_t)b * a.sC;
            const double* __restrict__ xn = a.out + ((size_t)b * K + i + 1) * nb;
            double acc = 0.0;
            for (int q = 0; q < nbd; ++q) acc += Cg[(size_t)q * nb + tid] * xn[a.bnd[q]];
            t -= acc;
        }
        const int qlast = (nb - 1) / 32 * 32;
        for (int q0 = qlast; q0 > p0; q0 -= 32) { // panels after this warp's own: fold them in
            double seg[32]; // L[q0+j][tid], j = 0..31: column tid of the panel's rows, coalesced over tid
#pragma unroll
            for (int j = 0; j < 32; ++j) seg[j] = (live && q0 + j < nb) ? Lg[(size_t)(q0 + j) * nb + tid] : 0.0;
            __syncthreads(); // panel q0 published by its owner
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                s0 += seg[j] * piv[q0 + j];
                s1 += seg[j + 1] * piv[q0 + j + 1];
                s2 += seg[j + 2] * piv[q0 + j + 2];
                s3 += seg[j + 3] * piv[q0 + j + 3];
            }
            t -= (s0 + s1) + (s2 + s3);
        }
        if (p0 < nb) { // own panel: diagonal block of L^T, backward, in registers
            double seg[32]; // L[p0+j][tid] for j >= lane
#pragma unroll
            for (int j = 0; j < 32; ++j) seg[j] = (live && p0 + j < nb && j >= lane) ? Lg[(size_t)(p0 + j) * nb + tid] : 0.0;
            double dg = 1.0;
#pragma unroll
            for (int j = 0; j < 32; ++j) dg = (lane == j) ? seg[j] : dg;
            const double rdg = live ? 1.0 / dg : 0.0;
            double mine = 0.0;
#pragma unroll
            for (int j = 31; j >= 0; --j) {
                const double xj = __shfl_sync(full, t, j) * __shfl_sync(full, rdg, j);
                mine = (lane == j) ? xj : mine;
                if (lane < j) t -= seg[j] * xj;
            }
            if (live) piv[tid] = mine;
        }
        __syncthreads(); // own panel published
        for (int q0 = p0 - 32; q0 >= 0; q0 -= 32) __syncthreads(); // the barriers of the panels before this warp's own
        if (live) a.out[((size_t)b * K + i) * nb + tid] = piv[tid];
        __threadfence_block();
        __syncthreads(); // x_i visible to the block (next interval's coupling) and piv free for reuse
    }
}


// ---- factorisation -----------------------------------------------------------------------------------------
// Block-tridiagonal Cholesky with boundary-row couplings, one CTA per instance, intervals in sequence:
//     A_i  = Dp_i - scatter_bnd(C_{i-1} C_{i-1}^T)      (packed lower triangle in shared memory)
//     L_i  = chol(A_i)                                   (right-looking, thread = row, two barriers per column)
//     C_i  = E_i L_i^-T                                  (nbd right-hand sides, forward substitution in shared memory)
// info[b] = 0, or 1 + the global index of the first non-positive pivot (the inertia test of the caller: the
// factor exists iff the regularised matrix is positive definite); after a failure the remaining factors of the
// instance are set to the identity and its couplings to zero so that the solve stays finite.
struct FactorArgs {
    const double* Dp;   // [B][K][nb][nb] symmetric blocks (lower triangle read)
    const double* Ep;   // [B][K-1][nbd][nb]
    double* L;          // [B][K][nb][nb] (lower triangle written)
    double* C;          // [B][K-1][nbd][nb]
    int* info;          // [B]
    const int* bnd;     // [nbd], ascending
    int B, K, nb, nbd;
};

__device__ __forceinline__ size_t tri(int i) { return (size_t)i * (i + 1) / 2; }

// T threads per matrix row (blockDim = T * rows, thread = (row, sub)): the trailing update of a column and the
// coupling substitution are latency-bound on the longest row, so each row's work is split over T lanes.
template <int T>
__global__ void __launch_bounds__(1024)
k_blocktri_factor(const __grid_constant__ FactorArgs a)
{
    extern __shared__ double sm[];
    const int nb = a.nb, nbd = a.nbd, K = a.K;
    double* A = sm;                        // packed lower triangle
    double* Cs = A + tri(nb);              // [nbd][nb] coupling of the previous / current interval
    double* z = Cs + (size_t)nbd * nb;     // [nbd] pivots of the coupling substitution
    double* col = z + nbd;                 // [nb] the scaled column being eliminated (own array: lets the update loops pipeline)
    const int b = blockIdx.x, tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
    const int row = tid / T, sub = tid - row * T;
    const bool live = row < nb;
    int fail = 0;
    for (int i = 0; i < K; ++i) {
        const double* __restrict__ Dg = a.Dp + ((size_t)b * K + i) * nb * nb;
        double* __restrict__ Lg = a.L + ((size_t)b * K + i) * nb * nb;
        if (fail) { // identity factor, zero coupling
            for (int r = warp; r < nb; r += nwarp)
                for (int c = lane; c <= r; c += 32) Lg[(size_t)r * nb + c] = (r == c) ? 1.0 : 0.0;
            if (i + 1 < K) {
                double* __restrict__ Cg = a.C + ((size_t)b * (K - 1) + i) * nbd * nb;
                for (int e = tid; e < nbd * nb; e += blockDim.x) Cg[e] = 0.0;
            }
            continue;
        }
        for (int r = warp; r < nb; r += nwarp)
            for (int c = lane; c <= r; c += 32) A[tri(r) + c] = Dg[(size_t)r * nb + c];
        __syncthreads();
        if (i > 0) { // A[bnd, bnd] -= C C^T (lower part; bnd ascending)
            for (int e = tid; e < nbd * (nbd + 1) / 2; e += blockDim.x) {
                int q1 = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
                while (tri(q1 + 1) <= (size_t)e) ++q1;
                while (tri(q1) > (size_t)e) --q1;
                const int q2 = e - (int)tri(q1);
                const double* c1 = Cs + (size_t)q1 * nb;
                const double* c2 = Cs + (size_t)q2 * nb;
                double s0 = 0.0, s1 = 0.0;
                int c = 0;
                for (; c + 1 < nb; c += 2) { s0 += c1[c] * c2[c]; s1 += c1[c + 1] * c2[c + 1]; }
                if (c < nb) s0 += c1[c] * c2[c];
                A[tri(a.bnd[q1]) + a.bnd[q2]] -= s0 + s1;
            }
            __syncthreads();
        }
        // right-looking Cholesky
        for (int j = 0; j < nb; ++j) {
            const double ajj = A[tri(j) + j];
            if (!(ajj > 0.0)) { fail = i * nb + j + 1; break; } // uniform: every thread reads the same pivot
            const double ljj = sqrt(ajj);
            double lrj = 0.0;
            if (live && row > j) {
                lrj = A[tri(row) + j] / ljj; // every lane of the row computes it, lane 0 publishes it after the barrier
                if (sub == 0) col[row] = lrj;
            }
            __syncthreads(); // every lane has read A(row, j) and the pivot; col[] is published
            if (sub == 0 && live && row > j) A[tri(row) + j] = lrj;
            if (tid == 0) A[tri(j) + j] = ljj;
            if (live && row > j) {
                double* __restrict__ rp = A + tri(row);
                const double* __restrict__ cj = col;
                for (int c = j + 1 + sub; c <= row; c += T) rp[c] -= lrj * cj[c];
            }
            __syncthreads();
        }
        if (fail) { // this interval failed: hand out the identity for it and everything after
            if (tid == 0) a.info[b] = fail;
            __syncthreads();
            --i; // redo interval i through the identity branch
            continue;
        }
        for (int r = warp; r < nb; r += nwarp)
            for (int c = lane; c <= r; c += 32) Lg[(size_t)r * nb + c] = A[tri(r) + c];
        if (i + 1 < K) { // C_i = E_i L_i^-T: solve L z_q = e_q for the nbd rows e_q of E_i, in place in Cs
            const double* __restrict__ Eg = a.Ep + ((size_t)b * (K - 1) + i) * nbd * nb;
            for (int e = tid; e < nbd * nb; e += blockDim.x) Cs[e] = Eg[e];
            __syncthreads();
            for (int j = 0; j < nb; ++j) {
                if (tid < nbd) {
                    const double v = Cs[(size_t)tid * nb + j] / A[tri(j) + j];
                    Cs[(size_t)tid * nb + j] = v;
                    z[tid] = v;
                }
                __syncthreads();
                if (live && row > j) {
                    const double lrj = A[tri(row) + j];
                    double* __restrict__ cr = Cs + row;
                    const double* __restrict__ zq = z;
                    for (int q = sub; q < nbd; q += T) cr[(size_t)q * nb] -= lrj * zq[q];
                }
                __syncthreads();
            }
            double* __restrict__ Cg = a.C + ((size_t)b * (K - 1) + i) * nbd * nb;
            for (int e = tid; e < nbd * nb; e += blockDim.x) Cg[e] = Cs[e];
        }
        __syncthreads();
    }
    if (!fail && tid == 0) a.info[b] = 0;
}


// ------------------------------------------------------------------------------------------------------------------
// k_kkt_factor: assembly + block-tridiagonal Cholesky of the dual-regularised KKT matrix in ONE launch
// ------------------------------------------------------------------------------------------------------------------
// Per instance and mesh interval i the interior-point step needs (lpopc_b200/solver.py::BlockTridiagKKT.step)
//     A_i = D_i + diag(Sigma_i + dw) + gamma G_i[own, own] (+ on the boundary slots: gamma G_{i-1}[next, next] - C_{i-1} C_{i-1}^T)
//     E_i' = E_i + gamma G_i[next, own],     G_i = Jb_i^T Jb_i     (Jb_i: the interval's rows of the Jacobian, [own block | boundary of block i+1])
//     L_i L_i^T = A_i,   C_i = E_i' L_i^-T
// which round 1 did with library calls: an einsum for G (27 % of a solve's GPU time), elementwise assembly, cuSOLVER's
// batched Cholesky (26 %), triangular solves and tril -- and one more round of all of it per inertia-correction retry.
//
// Here one CTA owns one instance and walks its K intervals.  The matrix that is factorised per interval is the
// AUGMENTED one: the nbd boundary rows of E_i' are appended below A_i as extra block rows.  Eliminating the nb columns
// of A_i then leaves C_i in those rows (the "L" of the appended rows is E_i' L_i^-T) and gamma G_i[next, next] - C_i C_i^T in
// their trailing block -- exactly what the next interval adds to its boundary slots.  One uniform right-looking tile
// algorithm does everything:
//   * the lower triangle of the augmented matrix is cut into T x T tiles (T = 8, or 6 for larger blocks), ONE TILE PER
//     THREAD, held in registers for the whole interval (nb = 140, nbd = 22: 231 tiles of 8 x 8 on 8 warps).  Whole
//     block columns are packed into warps (first fit), so a column never straddles two warps;
//   * SYRK: Jb_i is streamed through shared memory eight rows at a time (scaled by sqrt(gamma), columns permuted to the
//     augmented order); a thread adds x_I x_J^T to its tile, T^2 FMAs per 2 T shared loads; the next chunk's global
//     loads are in flight while the current one is multiplied (two chunk buffers, one barrier per chunk);
//   * factorisation, per block column kb: every tile right of kb subtracts L_I L_J^T from the published panel; the warp
//     holding column kb + 1 then factors its diagonal tile in registers (rsqrt, no division), __syncwarp, solves the
//     tiles below it and publishes the column -- while the other warps are still updating.  One block barrier per block
//     column (18 per interval) instead of one per matrix column;
//   * finished columns stay in an 8-slot ring of panels and leave for global memory four at a time, row by row
//     (a 6- or 8-wide column alone would be written in 48/64-byte pieces: that cost 40 % of the kernel);
//   * the inertia-correction retry (a non-positive pivot -> larger dw, start over) runs inside the kernel per
//     instance, so the Python loop with its synchronisations is gone.
// Bound: the FP64 pipe.  Cycle stamps per warp and step show one DFMA per scheduler every ~4.5 cycles in the SYRK and
// update phases (a B200 scheduler issues one per 4), so the time is the executed multiply-adds (about 2.3 M per
// interval, 1.15x the useful count) over the pipe rate, plus the serial column phases (~20 %) and the tile loads.
// nb = 140, nbd = 22, mr = 96, K = 8, 4096 instances: 23.1 ms against 47.6 ms for the library path (cuSOLVER's batched
// Cholesky alone: 29 ms).
template <int T>
struct KktFactorArgs {
    const double* D;     // [B][K][nb][nb] Hessian blocks (lower triangle read)
    const double* E;     // [B][K-1][nbd][nb]
    const double* Jb;    // [B][K][mr][nb + nbd]
    const double* diag;  // [B][K][nb] added to the diagonal (Sigma + 1 on padded slots)
    const int* bnd;      // [nbd] boundary slots, ascending
    const unsigned char* active; // [B] 0: converged instance, one attempt only
    double* dw;          // [B] in: first regularisation to try; out: the one that worked
    double* L;           // [B][K][nb][nb] (lower triangles written)
    double* C;           // [B][K-1][nbd][nb]
    int* info;           // [B] 0, or 1 when no attempt gave a positive definite matrix (identity factors written)
    double gamma;
    int B, K, nb, nbd, mr;
    int NBA, NBG; // block rows of A / of the appended boundary rows
    unsigned char tI[512], tJ[512]; // thread -> tile of the lower triangle (255: none)
};

// registers per thread for tile edge T (the accumulator tile alone takes 2 T^2) and the warps that budget leaves room
// for: a scheduler's 16384 registers are shared by the warps resident on it
template <int T> struct KktTile;
template <> struct KktTile<6> { static constexpr int regs = 128, max_warps = 16, nslot = 4; };
template <> struct KktTile<8> { static constexpr int regs = 232, max_warps = 8, nslot = 6; };

template <int T>
__global__ void __maxnreg__(KktTile<T>::regs)
k_kkt_factor(const __grid_constant__ KktFactorArgs<T> a)
{
    extern __shared__ double smf[];
    constexpr int RC = 8;          // Jacobian rows per streamed chunk
    constexpr int NSLOT = KktTile<T>::nslot; // chunk elements a thread carries from global to shared memory
    constexpr int TS = (T * T) | 1; // tile stride in the panel (odd: consecutive tiles hit different banks)
    constexpr int JS = T | 1;      // tile stride in a streamed Jacobian row
    constexpr int NSL = 8, WG = 4; // panel slots kept; block columns written out together (WG * T * 8 = 192 contiguous bytes per row)
    const int nb = a.nb, nbd = a.nbd, K = a.K, mr = a.mr, ncol = nb + nbd;
    const int NBA = a.NBA, NBR = a.NBA + a.NBG;
    const int WJ = NBR * JS;
    double* jr = smf;                                // [2][RC][NBR][JS]
    double* Lcol = jr + (size_t)2 * RC * WJ;         // [NSL][NBR][TS]: the last NSL block columns of the factor
    double* Sprev = Lcol + (size_t)NSL * NBR * TS;   // [nbd][nbd]
    double* rdiag = Sprev + (size_t)nbd * nbd;       // [2][T] reciprocals of the diagonal tile's pivots
    int* bidx = reinterpret_cast<int*>(rdiag + 2 * T); // [nb] boundary index of a slot, -1 otherwise
    int* flag = bidx + nb;
    const int b = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
    // thread -> tile (I, J), I >= J: the host packs whole block columns into warps (a column never straddles two), so
    // the diagonal factorisation and the triangular solves of a column synchronise inside ONE warp
    const bool has_tile = a.tI[tid] != 255;
    const int I = has_tile ? a.tI[tid] : 0, J = has_tile ? a.tJ[tid] : 0;
    unsigned colmask; // lanes of this thread's block column
    {
        const int len = NBR - J, lane0 = (tid & 31) - (I - J);
        colmask = len >= 32 ? 0xffffffffu : (((1u << len) - 1u) << lane0);
    }
    for (int i = tid; i < nb; i += nthr) bidx[i] = -1;
    for (int e = tid; e < 2 * RC * WJ; e += nthr) jr[e] = 0.0; // padding lanes of the streamed rows stay zero
    __syncthreads();
    for (int q = tid; q < nbd; q += nthr) bidx[a.bnd[q]] = q;
    __syncthreads();
    const double sg = sqrt(a.gamma);
    double dwv = a.dw[b];
    const bool act = a.active[b] != 0;
    const double* __restrict__ Db = a.D + (size_t)b * K * nb * nb;
    const double* __restrict__ Eb = a.E + (size_t)b * (K > 1 ? K - 1 : 1) * nbd * nb;
    const double* __restrict__ Jbb = a.Jb + (size_t)b * K * mr * ncol;
    const double* __restrict__ dgb = a.diag + (size_t)b * K * nb;
    double* __restrict__ Lb = a.L + (size_t)b * K * nb * nb;
    double* __restrict__ Cb = a.C + (size_t)b * (K > 1 ? K - 1 : 1) * nbd * nb;
    const int r0I = I * T, r0J = J * T;   // first row / column of the tile in the augmented index space
    const bool aug_row = I >= NBA;        // tile in the appended boundary rows
    const bool interior = has_tile && !aug_row && I > J && r0I + T <= nb; // all 36 entries are plain entries of D / L
    // boundary slots among this tile's rows / columns (they receive the previous interval's Schur block)
    unsigned rowm = 0, colm = 0;
    if (has_tile && !aug_row) {
#pragma unroll
        for (int q = 0; q < T; ++q) {
            if (r0I + q < nb && bidx[r0I + q] >= 0) rowm |= 1u << q;
            if (r0J + q < nb && bidx[r0J + q] >= 0) colm |= 1u << q;
        }
    }
    const bool hasb = rowm != 0 && colm != 0;
    // chunk loader: element e = tid + s * nthr of a chunk (RC rows of Jb, contiguous in memory) goes to a fixed place
    int dst[NSLOT], srow[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
        const int e = tid + s * nthr;
        dst[s] = -1; srow[s] = 0;
        if (e < RC * ncol) {
            const int r = e / ncol, c = e - r * ncol;
            const int p = c < nb ? c : NBA * T + (c - nb); // own slots, then the next interval's boundary slots
            dst[s] = r * WJ + (p / T) * JS + (p % T);
            srow[s] = r;
        }
    }
    // in-register Cholesky of a diagonal tile, published (lower part) with the reciprocals of its pivots
    auto factor_diag = [&](double (&t)[T][T], double* __restrict__ dstL, double* __restrict__ rd) {
        bool bad = false;
#pragma unroll
        for (int c = 0; c < T; ++c) {
            const double d = t[c][c];
            if (!(d > 0.0)) bad = true;
            const double rl = rsqrt(d);
            t[c][c] = d * rl;
            rd[c] = rl;
#pragma unroll
            for (int r = c + 1; r < T; ++r) t[r][c] *= rl;
#pragma unroll
            for (int c2 = c + 1; c2 < T; ++c2)
#pragma unroll
                for (int r = c2; r < T; ++r) t[r][c2] -= t[r][c] * t[c2][c];
        }
        if (bad) *flag = 1;
#pragma unroll
        for (int ii = 0; ii < T; ++ii)
#pragma unroll
            for (int jj = 0; jj <= ii; ++jj) dstL[ii * T + jj] = t[ii][jj];
    };
    // block column c of the factor: its warp factors the diagonal tile, solves the tiles below it and publishes them
    auto column_phase = [&](double (&t)[T][T], int c) {
        double* __restrict__ panel = Lcol + (size_t)(c & (NSL - 1)) * NBR * TS;
        double* __restrict__ rd = rdiag + (c & 1) * T;
        if (I == c) factor_diag(t, panel + (size_t)c * TS, rd);
        __syncwarp(colmask);
        if (I > c) { // X L_cc^T = tile
            const double* __restrict__ Lk = panel + (size_t)c * TS;
            double* __restrict__ mine = panel + (size_t)I * TS;
            double lk[T][T], rdc[T];
#pragma unroll
            for (int q = 0; q < T; ++q) {
                rdc[q] = rd[q];
#pragma unroll
                for (int c2 = 0; c2 < q; ++c2) lk[q][c2] = Lk[q * T + c2];
            }
#pragma unroll
            for (int ii = 0; ii < T; ++ii) {
#pragma unroll
                for (int q = 0; q < T; ++q) {
                    double v = t[ii][q];
#pragma unroll
                    for (int c2 = 0; c2 < q; ++c2) v -= t[ii][c2] * lk[q][c2];
                    t[ii][q] = v * rdc[q];
                }
#pragma unroll
                for (int q = 0; q < T; ++q) mine[ii * T + q] = t[ii][q];
            }
        }
    };
    // finished block columns c0 .. c0 + ncb - 1 of L (rows below nb: C) go from the panel slots to global memory, row by
    // row: a thread per element, consecutive threads along a row (a 6-wide block column alone would write 48-byte pieces)
    auto write_group = [&](double* __restrict__ Li, double* __restrict__ Ci, bool haveC, int c0, int ncb) {
        const int W = ncb * T, r_first = c0 * T, nrows = NBR * T - r_first;
        const int rstep = nthr / W, cc = tid % W, pc = r_first + cc, Jc = c0 + cc / T;
        if (tid >= rstep * W || pc >= nb) return;
        const double* __restrict__ src = Lcol + (size_t)(Jc & (NSL - 1)) * NBR * TS + (cc % T);
        for (int rq = tid / W; rq < nrows; rq += rstep) {
            const int r = r_first + rq, Ir = r / T;
            if (Ir < Jc) continue;
            const double v = src[Ir * TS + (r - Ir * T) * T];
            if (r < nb) {
                if (pc <= r) Li[r * nb + pc] = v;
            } else {
                const int q = r - NBA * T;
                if (haveC && q >= 0 && q < nbd) Ci[q * nb + pc] = v;
            }
        }
    };
    bool ok = false;
    for (int attempt = 0; act && attempt < 40 && !ok; ++attempt) { // a converged instance gets identity factors: its step is discarded
        if (tid == 0) *flag = 0;
        __syncthreads();
        bool failed = false;
        for (int i = 0; i < K && !failed; ++i) {
            double t[T][T];
            const double* __restrict__ Ji = Jbb + (size_t)i * mr * ncol;
            double pre[NSLOT];
            auto gload = [&](int rr) {
#pragma unroll
                for (int s = 0; s < NSLOT; ++s)
                    pre[s] = (dst[s] >= 0 && rr + srow[s] < mr) ? Ji[rr * ncol + tid + s * nthr] : 0.0;
            };
            auto sstore = [&](double* __restrict__ buf) {
#pragma unroll
                for (int s = 0; s < NSLOT; ++s)
                    if (dst[s] >= 0) buf[dst[s]] = sg * pre[s]; // scaled here, not at the load: the load stays in flight over the multiply-adds
            };
            gload(0);
            // ---- tile of the augmented matrix before the rank update ----
            if (interior) {
                const double* __restrict__ p = Db + (size_t)i * nb * nb + r0I * nb + r0J;
                if ((nb & 1) == 0 && (T & 1) == 0) { // rows of a tile start on 16-byte boundaries
#pragma unroll
                    for (int ii = 0; ii < T; ++ii)
#pragma unroll
                        for (int jj = 0; jj < T; jj += 2) {
                            const double2 v = *reinterpret_cast<const double2*>(p + ii * nb + jj);
                            t[ii][jj] = v.x; t[ii][jj + 1] = v.y;
                        }
                } else {
#pragma unroll
                    for (int ii = 0; ii < T; ++ii)
#pragma unroll
                        for (int jj = 0; jj < T; ++jj) t[ii][jj] = p[ii * nb + jj];
                }
            } else {
#pragma unroll
                for (int ii = 0; ii < T; ++ii)
#pragma unroll
                    for (int jj = 0; jj < T; ++jj) {
                        const int pr = r0I + ii, pc = r0J + jj;
                        double v = 0.0;
                        if (has_tile) {
                            if (!aug_row) {
                                if (pr < nb && pc < nb) {
                                    v = (pc <= pr) ? Db[((size_t)i * nb + pr) * nb + pc] : Db[((size_t)i * nb + pc) * nb + pr];
                                    if (pr == pc) v += dgb[(size_t)i * nb + pr] + dwv;
                                } else {
                                    v = (pr == pc) ? 1.0 : 0.0; // padding rows of the last A tile
                                }
                            } else if (J < NBA) {
                                const int q = pr - NBA * T;
                                if (i + 1 < K && q < nbd && pc < nb) v = Eb[((size_t)i * nbd + q) * nb + pc];
                            }
                        }
                        t[ii][jj] = v;
                    }
            }
            if (hasb && i > 0) { // + the Schur block the previous interval left on the boundary slots
#pragma unroll
                for (int ii = 0; ii < T; ++ii)
#pragma unroll
                    for (int jj = 0; jj < T; ++jj)
                        if (((rowm >> ii) & 1u) && ((colm >> jj) & 1u)) t[ii][jj] += Sprev[bidx[r0I + ii] * nbd + bidx[r0J + jj]];
            }
            // ---- + gamma Jb^T Jb, streamed: the next chunk's global loads are in flight while this one is multiplied ----
            sstore(jr);
            __syncthreads();
            for (int rr = 0, cb = 0; rr < mr; rr += RC, cb ^= 1) {
                const bool more = rr + RC < mr;
                if (more) gload(rr + RC);
                if (has_tile) {
                    const double* __restrict__ ji = jr + cb * RC * WJ + I * JS;
                    const double* __restrict__ jj_ = jr + cb * RC * WJ + J * JS;
#pragma unroll
                    for (int r = 0; r < RC; ++r) {
                        double xi[T], xj[T];
#pragma unroll
                        for (int c = 0; c < T; ++c) { xi[c] = ji[r * WJ + c]; xj[c] = jj_[r * WJ + c]; }
#pragma unroll
                        for (int ii = 0; ii < T; ++ii)
#pragma unroll
                            for (int jj = 0; jj < T; ++jj) t[ii][jj] += xi[ii] * xj[jj];
                    }
                }
                if (more) sstore(jr + (cb ^ 1) * RC * WJ);
                __syncthreads();
            }
            // ---- right-looking tile Cholesky over the NBA block columns of A; column kb + 1 is finished by its warp
            //      while the other warps are still applying column kb: one block barrier per column ----
            double* __restrict__ Li = Lb + (size_t)i * nb * nb;
            double* __restrict__ Ci = Cb + (size_t)i * nbd * nb;
            if (has_tile && J == 0) column_phase(t, 0);
            __syncthreads();
            if (*flag) failed = true;
            for (int kb = 0; kb < NBA && !failed; ++kb) {
                const double* __restrict__ panel = Lcol + (size_t)(kb & (NSL - 1)) * NBR * TS;
                if (has_tile && J > kb) {
                    const double* __restrict__ LI = panel + (size_t)I * TS;
                    const double* __restrict__ LJ = panel + (size_t)J * TS;
#pragma unroll
                    for (int c = 0; c < T; ++c) {
                        double li[T], lj[T];
#pragma unroll
                        for (int q = 0; q < T; ++q) { li[q] = LI[q * T + c]; lj[q] = LJ[q * T + c]; }
#pragma unroll
                        for (int ii = 0; ii < T; ++ii)
#pragma unroll
                            for (int jj = 0; jj < T; ++jj) t[ii][jj] -= li[ii] * lj[jj];
                    }
                }
                if (has_tile && J == kb + 1 && kb + 1 < NBA) column_phase(t, kb + 1);
                if ((kb % WG) == WG - 1) write_group(Li, Ci, i + 1 < K, kb - (WG - 1), WG); // published one step ago or earlier
                __syncthreads(); // column kb + 1 is complete in its panel slot
                if (*flag) failed = true;
            }
            if (failed) break;
            if (NBA % WG) write_group(Li, Ci, i + 1 < K, NBA - NBA % WG, NBA % WG);
            // ---- trailing block of the appended rows: gamma G[next, next] - C C^T, handed to the next interval ----
            if (has_tile && J >= NBA) {
#pragma unroll
                for (int ii = 0; ii < T; ++ii)
#pragma unroll
                    for (int jj = 0; jj < T; ++jj) {
                        const int q1 = r0I + ii - NBA * T, q2 = r0J + jj - NBA * T;
                        if (q1 < nbd && q2 < nbd && q2 <= q1) {
                            Sprev[q1 * nbd + q2] = t[ii][jj];
                            Sprev[q2 * nbd + q1] = t[ii][jj];
                        }
                    }
            }
            __syncthreads();
        }
        if (!failed) { ok = true; break; }
        dwv = (dwv == 0.0) ? 1e-4 : dwv * (attempt < 2 ? 100.0 : 8.0);
        __syncthreads();
    }
    if (!ok) { // keep the outputs finite: identity factors, zero couplings
        for (int e = tid; e < K * nb * nb; e += nthr) {
            const int r = (e / nb) % nb, c = e % nb;
            if (c <= r) Lb[e] = (r == c) ? 1.0 : 0.0;
        }
        for (int e = tid; e < (K - 1) * nbd * nb; e += nthr) Cb[e] = 0.0;
    }
    if (tid == 0) {
        a.info[b] = (ok || !act) ? 0 : 1;
        a.dw[b] = dwv;
    }
}

// Packing of the block columns (NBR - J tiles each) into warps of 32 lanes.  A warp executes the trailing update of
// step kb as long as ANY of its columns has J > kb, so what the packing costs is sum over warps of min(max J, NBA)
// warp-steps.  First fit by decreasing length pairs the longest column with a short, LATE one and keeps that warp busy
// at a third of its lanes (91 warp-steps for 21 columns on 8 warps); steepest descent over single moves and swaps
// (ties broken towards columns of similar J sharing a warp) brings it to 73, the update phase's instruction count with it.
// Results are cached per shape.  false: more than max_warps warps needed.
static bool pack_columns(int NBR, int NBA, int max_warps, unsigned char* tI, unsigned char* tJ, int* nwarps_out)
{
    struct Packing { int NBR, NBA, max_warps, nwarps; unsigned char bin[32]; };
    static std::mutex mu;
    static std::vector<Packing> cache;
    Packing pk{NBR, NBA, max_warps, 0, {0}};
    bool found = false;
    {
        std::lock_guard<std::mutex> lock(mu);
        for (const Packing& c : cache)
            if (c.NBR == NBR && c.NBA == NBA && c.max_warps == max_warps) { pk = c; found = true; break; }
    }
    if (!found) {
        int fill[32] = {0}, nb = 0;
        for (int J = 0; J < NBR; ++J) { // first fit, longest first
            const int len = NBR - J;
            int w = 0;
            while (w < nb && fill[w] + len > 32) ++w;
            if (w == nb) {
                if (nb == max_warps) return false;
                ++nb;
            }
            pk.bin[J] = (unsigned char)w;
            fill[w] += len;
        }
        auto key = [&](const unsigned char* bin, long long* second) {
            int mx[32];
            for (int w = 0; w < nb; ++w) mx[w] = -1;
            for (int J = 0; J < NBR; ++J) mx[bin[J]] = J > mx[bin[J]] ? J : mx[bin[J]];
            long long c = 0, s2 = 0;
            for (int w = 0; w < nb; ++w) c += mx[w] < 0 ? 0 : (mx[w] < NBA ? mx[w] : NBA);
            for (int J = 0; J < NBR; ++J) {
                const int m = mx[bin[J]] < NBA ? mx[bin[J]] : NBA;
                s2 += m - (J < NBA ? J : NBA);
            }
            *second = s2;
            return c;
        };
        long long cur2 = 0, cur = key(pk.bin, &cur2);
        for (;;) {
            long long best = cur, best2 = cur2;
            int bJ = -1, bw = -1, bJ2 = -1;
            for (int J = 0; J < NBR; ++J) {
                const int wa = pk.bin[J], lenJ = NBR - J;
                for (int w = 0; w < nb; ++w) {
                    if (w == wa) continue;
                    if (fill[w] + lenJ <= 32) { // move J to w
                        pk.bin[J] = (unsigned char)w;
                        long long k2 = 0, k = key(pk.bin, &k2);
                        pk.bin[J] = (unsigned char)wa;
                        if (k < best || (k == best && k2 < best2)) { best = k; best2 = k2; bJ = J; bw = w; bJ2 = -1; }
                    }
                    for (int J2 = 0; J2 < NBR; ++J2) { // swap J and J2
                        if (pk.bin[J2] != w) continue;
                        const int len2 = NBR - J2;
                        if (fill[wa] - lenJ + len2 > 32 || fill[w] - len2 + lenJ > 32) continue;
                        pk.bin[J] = (unsigned char)w; pk.bin[J2] = (unsigned char)wa;
                        long long k2 = 0, k = key(pk.bin, &k2);
                        pk.bin[J] = (unsigned char)wa; pk.bin[J2] = (unsigned char)w;
                        if (k < best || (k == best && k2 < best2)) { best = k; best2 = k2; bJ = J; bw = w; bJ2 = J2; }
                    }
                }
            }
            if (bJ < 0) break;
            const int wa = pk.bin[bJ];
            fill[wa] -= NBR - bJ; fill[bw] += NBR - bJ;
            pk.bin[bJ] = (unsigned char)bw;
            if (bJ2 >= 0) { fill[bw] -= NBR - bJ2; fill[wa] += NBR - bJ2; pk.bin[bJ2] = (unsigned char)wa; }
            cur = best; cur2 = best2;
        }
        pk.nwarps = nb;
        std::lock_guard<std::mutex> lock(mu);
        cache.push_back(pk);
    }
    for (int t = 0; t < 512; ++t) tI[t] = tJ[t] = 255;
    int fill[32] = {0};
    for (int J = 0; J < NBR; ++J) {
        const int w = pk.bin[J], len = NBR - J;
        for (int e = 0; e < len; ++e) {
            tI[w * 32 + fill[w] + e] = (unsigned char)(J + e);
            tJ[w * 32 + fill[w] + e] = (unsigned char)J;
        }
        fill[w] += len;
    }
    *nwarps_out = pk.nwarps;
    return true;
}

struct KktProblem {
    const double *D, *E, *Jb, *diag;
    const int* bnd;
    const unsigned char* active;
    double *dw, *L, *C;
    int* info;
    double gamma;
    int B, K, nb, nbd, mr;
};

template <int T>
int launch_kkt_factor(const KktProblem& q, void* stream)
{
    KktFactorArgs<T> a;
    a.D = q.D; a.E = q.E; a.Jb = q.Jb; a.diag = q.diag; a.bnd = q.bnd; a.active = q.active; a.dw = q.dw; a.L = q.L; a.C = q.C; a.info = q.info;
    a.gamma = q.gamma; a.B = q.B; a.K = q.K; a.nb = q.nb; a.nbd = q.nbd; a.mr = q.mr;
    a.NBA = (q.nb + T - 1) / T; a.NBG = (q.nbd + T - 1) / T;
    const int NBR = a.NBA + a.NBG;
    if (NBR > 32) return -1; // a block column has to fit one warp
    // block columns -> warps (a column never straddles two): pack_columns below
    constexpr int MAXW = KktTile<T>::max_warps;
    int nwarps = 0;
    if (!pack_columns(NBR, a.NBA, MAXW, a.tI, a.tJ, &nwarps)) return -1; // the register budget of one tile per thread
    int threads = nwarps * 32;
    while (8 * (q.nb + q.nbd) > KktTile<T>::nslot * threads) threads += 32; // chunk loader: nslot elements per thread
    if (threads > MAXW * 32) return -1;
    const size_t shm = ((size_t)2 * 8 * NBR * (T | 1) + (size_t)8 * NBR * ((T * T) | 1) + (size_t)q.nbd * q.nbd + 2 * T) * sizeof(double) + ((size_t)q.nb + 4) * sizeof(int);
    if (shm > 200 * 1024) return -1;
    cudaError_t e = cudaFuncSetAttribute(k_kkt_factor<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
    if (e != cudaSuccess) return -(int)e - 1000;
    k_kkt_factor<T><<<q.B, threads, shm, (cudaStream_t)stream>>>(a);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e - 1000;
}


// ------------------------------------------------------------------------------------------------------------------
// k_batched_spmv: y_b = A_b x_b for a batch of sparse matrices that share ONE structure (CSR: rowptr, col) and differ
// in their values, which are read IN PLACE from the instance's triplet value array through perm (CSR position ->
// triplet index): the Jacobian and Hessian values the transcription kernels wrote are used as they lie, no copy, no
// dense block.  The KKT refinement of the batched interior-point step multiplies with J, J^T and H two to four times
// per iteration; through dense blocks (cuBLAS gemv over [B][K][mr][nb + nbd]) that read 4 GB per product, the
// triplets are 0.65 GB.  One CTA per instance, x in shared memory, one thread per row, entries added in CSR order:
// deterministic, and consecutive rows (consecutive nodes of one state) read consecutive triplets.
__global__ void __launch_bounds__(256)
k_batched_spmv(int nrows, int ncols, const int* __restrict__ rowptr, const int* __restrict__ col, const int* __restrict__ perm,
               const double* __restrict__ vals, long long val_stride, const double* __restrict__ x, const double* __restrict__ diag,
               double* __restrict__ y, int x_in_smem)
{
    extern __shared__ double xs[];
    const int b = blockIdx.x;
    const double* __restrict__ xb = x + (size_t)b * ncols;
    if (x_in_smem) {
        for (int i = threadIdx.x; i < ncols; i += blockDim.x) xs[i] = xb[i];
        __syncthreads();
    }
    const double* __restrict__ xv = x_in_smem ? xs : xb;
    const double* __restrict__ vb = vals + (size_t)b * val_stride;
    for (int r = threadIdx.x; r < nrows; r += blockDim.x) {
        double acc = (diag != nullptr) ? diag[(size_t)b * nrows + r] * xv[r] : 0.0; // optional diagonal term (square matrices)
        const int e1 = rowptr[r + 1];
        for (int e = rowptr[r]; e < e1; ++e) acc += vb[perm[e]] * xv[col[e]];
        y[(size_t)b * nrows + r] = acc;
    }
}

} // namespace

extern "C" {

// Device pointers throughout; asynchronous on `stream`.  Returns 0, -1 (shape not supported: K > 64 or
// nb > 256 -- the caller falls back to library solves) or a negative cudaError_t - 1000.
// strideL / strideC: doubles between consecutive instances of one factor / coupling (nb*nb and nbd*nb for
// separately allocated [B][nb][nb] tensors, K*nb*nb and (K-1)*nbd*nb for slices of one [B][K][..] tensor).
int lpb_blocktri_solve_masked(int B, int K, int nb, int nbd, const double* const* L, const double* const* C, long long strideL, long long strideC,
                              const int* bnd, const unsigned char* active, const double* rhs, double* out, void* stream);

int lpb_blocktri_solve(int B, int K, int nb, int nbd, const double* const* L, const double* const* C, long long strideL, long long strideC,
                       const int* bnd, const double* rhs, double* out, void* stream)
{
    return lpb_blocktri_solve_masked(B, K, nb, nbd, L, C, strideL, strideC, bnd, nullptr, rhs, out, stream);
}

// The same with a per-instance mask (device, [B], may be null): instances whose byte is 0 are not solved, zeros are written.
int lpb_blocktri_solve_masked(int B, int K, int nb, int nbd, const double* const* L, const double* const* C, long long strideL, long long strideC,
                              const int* bnd, const unsigned char* active, const double* rhs, double* out, void* stream)
{
    if (B < 1 || K < 1 || K > kMaxBlocks || nb < 1 || nbd < 0) return -1;
    const size_t shm = ((size_t)K * nb + (size_t)(nb + 31) / 32 * 32) * sizeof(double); // piv padded to whole panels
    if (shm > 200 * 1024 || nb > 256) return -1;
    BlockTriArgs a;
    for (int i = 0; i < K; ++i) a.L[i] = L[i];
    for (int i = 0; i + 1 < K; ++i) a.C[i] = C[i];
    a.bnd = bnd; a.active = active; a.rhs = rhs; a.out = out; a.sL = strideL; a.sC = strideC; a.B = B; a.K = K; a.nb = nb; a.nbd = nbd;
    {   // per device and context, and cheap: set on every launch rather than caching it in a process-wide static
        cudaError_t e = cudaFuncSetAttribute(k_blocktri_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
        if (e != cudaSuccess) return -(int)e - 1000;
    }
    const int threads = (nb + 31) / 32 * 32;
    k_blocktri_solve<<<B, threads, shm, (cudaStream_t)stream>>>(a);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e - 1000;
}

// Block-tridiagonal Cholesky (k_blocktri_factor above).  Dp [B][K][nb][nb], Ep [B][K-1][nbd][nb] in; L (same
// shape as Dp, lower triangles), C (same shape as Ep) and info [B] out; bnd ascending.  Same return convention.
int lpb_blocktri_factor(int B, int K, int nb, int nbd, const double* Dp, const double* Ep, const int* bnd, double* L, double* C, int* info,
                        void* stream)
{
    if (B < 1 || K < 1 || nb < 1 || nbd < 0 || nb > 256) return -1;
    const size_t shm = ((size_t)nb * (nb + 1) / 2 + (size_t)nbd * nb + (size_t)nbd + (size_t)nb + 2) * sizeof(double);
    if (shm > 220 * 1024) return -1;
    FactorArgs a;
    a.Dp = Dp; a.Ep = Ep; a.L = L; a.C = C; a.info = info; a.bnd = bnd; a.B = B; a.K = K; a.nb = nb; a.nbd = nbd;
    // lanes per row: 4 for large blocks (the column update is latency-bound on the longest row: 140 x 140 at 64
    // instances 1.96 vs 3.76 ms), 1 for small ones (44 x 44 at 4096 instances 1.84 vs 2.5 ms)
    const int rows = ((nb > nbd ? nb : nbd) + 31) / 32 * 32;
    cudaError_t e;
    if (nb > 64 && rows * 4 <= 1024) {
        e = cudaFuncSetAttribute(k_blocktri_factor<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
        if (e != cudaSuccess) return -(int)e - 1000;
        k_blocktri_factor<4><<<B, rows * 4, shm, (cudaStream_t)stream>>>(a);
    } else {
        e = cudaFuncSetAttribute(k_blocktri_factor<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
        if (e != cudaSuccess) return -(int)e - 1000;
        k_blocktri_factor<1><<<B, rows, shm, (cudaStream_t)stream>>>(a);
    }
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e - 1000;
}

// Assembly + factorisation of the block-tridiagonal KKT matrix (k_kkt_factor above).  dw: in/out per instance.
// Returns 0, -1 (shape not supported: the caller keeps the library path) or a negative cudaError_t - 1000.
int lpb_kkt_factor(int B, int K, int nb, int nbd, int mr, double gamma, const double* D, const double* E, const double* Jb, const double* diag,
                   const int* bnd, const unsigned char* active, double* dw, double* L, double* C, int* info, void* stream)
{
    if (B < 1 || K < 1 || nb < 1 || nbd < 1 || mr < 1) return -1;
    KktProblem q{D, E, Jb, diag, bnd, active, dw, L, C, info, gamma, B, K, nb, nbd, mr};
    // 8 x 8 tiles when the lower triangle then fits 8 warps (two per scheduler, evenly: the kernel is bound by the FP64
    // pipe, so an uneven 4/3/3/3 split of 13 warps costs a quarter), 6 x 6 tiles otherwise
    int rc = launch_kkt_factor<8>(q, stream);
    if (rc == -1) rc = launch_kkt_factor<6>(q, stream);
    return rc;
}

// Batched sparse matrix-vector product with one shared structure (k_batched_spmv above).  rowptr [nrows + 1], col and
// perm [rowptr[nrows]] on the device; vals [B][val_stride] (triplet values, addressed through perm); x [B][ncols];
// diag [B][nrows] or null (adds diag_r x_r: nrows == ncols then); y [B][nrows].  Asynchronous on `stream`.
int lpb_batched_spmv(int B, int nrows, int ncols, const int* rowptr, const int* col, const int* perm, const double* vals, long long val_stride,
                     const double* x, const double* diag, double* y, void* stream)
{
    if (B < 1 || nrows < 1 || ncols < 1 || !rowptr || !col || !perm || !vals || !x || !y || (diag && nrows != ncols)) return -1;
    const size_t shm = (size_t)ncols * sizeof(double);
    const int in_smem = shm <= 160 * 1024 ? 1 : 0;
    if (in_smem && shm > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_batched_spmv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
        if (e != cudaSuccess) return -(int)e - 1000;
    }
    k_batched_spmv<<<B, 256, in_smem ? shm : 0, (cudaStream_t)stream>>>(nrows, ncols, rowptr, col, perm, vals, val_stride, x, diag, y, in_smem);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e - 1000;
}

} // extern "C"
