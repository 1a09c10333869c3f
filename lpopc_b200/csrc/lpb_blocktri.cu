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
// lockstep iteration).  Here ONE launch does the whole solve: one CTA per instance, the factor of the current
// interval in shared memory (packed lower triangle, nb (nb+1) / 2 doubles), thread i owns row i of the
// right-hand side, 32-row panels eliminated with register shuffles and one barrier per panel.
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace {

constexpr int kMaxBlocks = 64;

struct BlockTriArgs {
    const double* L[kMaxBlocks]; // K factors, each [B][nb][nb] row-major (lower triangle read)
    const double* C[kMaxBlocks]; // K-1 couplings, each [B][nbd][nb] row-major
    const int* bnd;              // [nbd] boundary slots of a block
    const double* rhs;           // [B][K][nb]
    double* out;                 // [B][K][nb]
    int B, K, nb, nbd;
};

__device__ __forceinline__ size_t tri(int i) { return (size_t)i * (i + 1) / 2; }

// Substitution by 32-row panels: the warp that owns a panel's rows eliminates the 32 x 32 diagonal block with
// register shuffles (no barrier inside), publishes the panel's solution in shared memory, and after ONE barrier
// every remaining row folds the panel in with 32 multiply-adds.  Diagonal reciprocals are taken once per factor.
__global__ void k_blocktri_solve(const __grid_constant__ BlockTriArgs a)
{
    extern __shared__ double sm[];
    const int nb = a.nb, nbd = a.nbd, K = a.K;
    double* Lp = sm;                          // packed lower triangle of the current factor
    double* ys = Lp + tri(nb);                // [K][nb] forward results
    double* piv = ys + (size_t)K * nb;        // [nb] right-hand side scratch / backward solution of the current interval
    double* rd = piv + nb;                    // [nb] reciprocals of the factor's diagonal
    const int b = blockIdx.x, tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
    const int p0 = warp * 32;                 // first row of this warp's panel (thread = row)
    const unsigned full = 0xffffffffu;

    auto load_factor = [&](int i) {
        const double* __restrict__ Lg = a.L[i] + (size_t)b * nb * nb;
        for (int r = warp; r < nb; r += nwarp) // warp per row: coalesced over the row's r + 1 entries
            for (int c = lane; c <= r; c += 32) {
                const double v = Lg[(size_t)r * nb + c];
                Lp[tri(r) + c] = v;
                if (c == r) rd[r] = 1.0 / v;
            }
    };

    // ---- forward ----
    for (int i = 0; i < K; ++i) {
        __syncthreads(); // the previous interval's solve has finished with Lp / rd / piv
        load_factor(i);
        double t = tid < nb ? a.rhs[((size_t)b * K + i) * nb + tid] : 0.0;
        if (i > 0) { // r_i[bnd] -= C_{i-1} y_{i-1}: warp per boundary row, coalesced dot product
            if (tid < nb) piv[tid] = t;
            __syncthreads();
            const double* __restrict__ Cg = a.C[i - 1] + (size_t)b * nbd * nb;
            const double* __restrict__ yp = ys + (size_t)(i - 1) * nb;
            for (int q = warp; q < nbd; q += nwarp) {
                double acc = 0.0;
                for (int c = lane; c < nb; c += 32) acc += Cg[(size_t)q * nb + c] * yp[c];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(full, acc, off);
                if (lane == 0) piv[a.bnd[q]] -= acc;
            }
            __syncthreads();
            if (tid < nb) t = piv[tid];
        } else __syncthreads();
        double* __restrict__ y = ys + (size_t)i * nb;
        for (int q0 = 0; q0 < nb; q0 += 32) { // panel of columns q0 .. q0 + 31
            if (p0 == q0) { // owner warp: 32 x 32 diagonal block, forward, in registers
                double mine = 0.0;
                const int lim = nb - q0 < 32 ? nb - q0 : 32;
                for (int j = 0; j < lim; ++j) {
                    const double yj = __shfl_sync(full, t, j) * rd[q0 + j];
                    if (lane == j) mine = yj;
                    if (lane > j && tid < nb) t -= Lp[tri(tid) + q0 + j] * yj;
                }
                if (tid < nb) y[tid] = mine;
            }
            __syncthreads();
            if (p0 > q0 && tid < nb) { // rows below the panel
                const double* __restrict__ row = Lp + tri(tid) + q0;
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0; // four independent chains: the update is latency-bound
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    s0 += row[j] * y[q0 + j];
                    s1 += row[j + 1] * y[q0 + j + 1];
                    s2 += row[j + 2] * y[q0 + j + 2];
                    s3 += row[j + 3] * y[q0 + j + 3];
                }
                t -= (s0 + s1) + (s2 + s3);
            }
        }
    }
    // ---- backward ----
    for (int i = K - 1; i >= 0; --i) {
        __syncthreads();
        if (i != K - 1) load_factor(i); // the last factor is still resident
        double t = tid < nb ? ys[(size_t)i * nb + tid] : 0.0;
        if (i + 1 < K && tid < nb) { // y_i -= C_i^T x_{i+1}[bnd]: thread = column, coalesced over columns
            const double* __restrict__ Cg = a.C[i] + (size_t)b * nbd * nb;
            const double* __restrict__ xn = a.out + ((size_t)b * K + i + 1) * nb;
            double acc = 0.0;
            for (int q = 0; q < nbd; ++q) acc += Cg[(size_t)q * nb + tid] * xn[a.bnd[q]];
            t -= acc;
        }
        __syncthreads();
        for (int q0 = (nb - 1) / 32 * 32; q0 >= 0; q0 -= 32) { // L^T x = t, panels from the last to the first
            if (p0 == q0) {
                double mine = 0.0;
                const int lim = nb - q0 < 32 ? nb - q0 : 32;
                for (int j = lim - 1; j >= 0; --j) {
                    const double xj = __shfl_sync(full, t, j) * rd[q0 + j];
                    if (lane == j) mine = xj;
                    if (lane < j) t -= Lp[tri(q0 + j) + tid] * xj; // L^T(tid, q0+j) = L(q0+j, tid): contiguous in the packed row
                }
                if (tid < nb) piv[tid] = mine;
            }
            __syncthreads();
            if (p0 < q0) { // rows above the panel
                const int lim = nb - q0 < 32 ? nb - q0 : 32;
                double s0 = 0.0, s1 = 0.0;
                int j = 0;
                for (; j + 1 < lim; j += 2) {
                    s0 += Lp[tri(q0 + j) + tid] * piv[q0 + j];
                    s1 += Lp[tri(q0 + j + 1) + tid] * piv[q0 + j + 1];
                }
                if (j < lim) s0 += Lp[tri(q0 + j) + tid] * piv[q0 + j];
                t -= s0 + s1;
            }
        }
        if (tid < nb) a.out[((size_t)b * K + i) * nb + tid] = piv[tid];
        __threadfence_block();
    }
}

} // namespace

extern "C" {

// Device pointers throughout; asynchronous on `stream`.  Returns 0, -1 (shape not supported: K > 64 or the
// packed factor does not fit in shared memory -- the caller falls back to library solves) or a negative
// cudaError_t - 1000.
int lpb_blocktri_solve(int B, int K, int nb, int nbd, const double* const* L, const double* const* C, const int* bnd,
                       const double* rhs, double* out, void* stream)
{
    if (B < 1 || K < 1 || K > kMaxBlocks || nb < 1 || nbd < 0) return -1;
    const size_t shm = ((size_t)nb * (nb + 1) / 2 + (size_t)K * nb + 2 * (size_t)nb) * sizeof(double);
    if (shm > 220 * 1024 || nb > 1024) return -1;
    BlockTriArgs a;
    for (int i = 0; i < K; ++i) a.L[i] = L[i];
    for (int i = 0; i + 1 < K; ++i) a.C[i] = C[i];
    a.bnd = bnd; a.rhs = rhs; a.out = out; a.B = B; a.K = K; a.nb = nb; a.nbd = nbd;
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

} // extern "C"
