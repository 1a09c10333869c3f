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
#include <cstddef>
#include <cstdint>

namespace {

constexpr int kMaxBlocks = 64;

struct BlockTriArgs {
    const double* L[kMaxBlocks]; // K factors, each [B][nb][nb] row-major with instance stride sL (lower triangle read)
    const double* C[kMaxBlocks]; // K-1 couplings, each [B][nbd][nb] row-major with instance stride sC
    const int* bnd;              // [nbd] boundary slots of a block
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
__global__ void __launch_bounds__(256)
k_blocktri_solve(const __grid_constant__ BlockTriArgs a)
{
    extern __shared__ double sm[];
    const int nb = a.nb, nbd = a.nbd, K = a.K;
    double* ys = sm;                          // [K][nb] forward results
    double* piv = ys + (size_t)K * nb;        // [nb, padded to whole panels] right-hand side scratch / backward solution of the current interval
    const int b = blockIdx.x, tid = threadIdx.x;
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

} // namespace

extern "C" {

// Device pointers throughout; asynchronous on `stream`.  Returns 0, -1 (shape not supported: K > 64 or
// nb > 256 -- the caller falls back to library solves) or a negative cudaError_t - 1000.
// strideL / strideC: doubles between consecutive instances of one factor / coupling (nb*nb and nbd*nb for
// separately allocated [B][nb][nb] tensors, K*nb*nb and (K-1)*nbd*nb for slices of one [B][K][..] tensor).
int lpb_blocktri_solve(int B, int K, int nb, int nbd, const double* const* L, const double* const* C, long long strideL, long long strideC,
                       const int* bnd, const double* rhs, double* out, void* stream)
{
    if (B < 1 || K < 1 || K > kMaxBlocks || nb < 1 || nbd < 0) return -1;
    const size_t shm = ((size_t)K * nb + (size_t)(nb + 31) / 32 * 32) * sizeof(double); // piv padded to whole panels
    if (shm > 200 * 1024 || nb > 256) return -1;
    BlockTriArgs a;
    for (int i = 0; i < K; ++i) a.L[i] = L[i];
    for (int i = 0; i + 1 < K; ++i) a.C[i] = C[i];
    a.bnd = bnd; a.rhs = rhs; a.out = out; a.sL = strideL; a.sC = strideC; a.B = B; a.K = K; a.nb = nb; a.nbd = nbd;
    static size_t attr = 0;
    if (shm > attr) {
        cudaError_t e = cudaFuncSetAttribute(k_blocktri_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
        if (e != cudaSuccess) return -(int)e - 1000;
        attr = shm;
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
        static size_t attr4 = 0;
        if (shm > attr4) {
            e = cudaFuncSetAttribute(k_blocktri_factor<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
            if (e != cudaSuccess) return -(int)e - 1000;
            attr4 = shm;
        }
        k_blocktri_factor<4><<<B, rows * 4, shm, (cudaStream_t)stream>>>(a);
    } else {
        static size_t attr1 = 0;
        if (shm > attr1) {
            e = cudaFuncSetAttribute(k_blocktri_factor<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
            if (e != cudaSuccess) return -(int)e - 1000;
            attr1 = shm;
        }
        k_blocktri_factor<1><<<B, rows, shm, (cudaStream_t)stream>>>(a);
    }
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -(int)e - 1000;
}

} // extern "C"
