// lpb_mesherr.cuh -- mesh-error estimator on the GPU.
//
//   k_mesh_error   replaces SolutionErrorChecker::SolutionInterpolation + CheckSolutionDiffError
//                  (Lpopc/src/Core/LpSolutionError.cpp:9-166): per mesh interval, barycentric Lagrange
//                  interpolation of the NLP solution onto N_k+1 LGR points, dae() there, integration with
//                  A_k = inv(D_k[:, 1:]) and the absolute defect of the integrated right-hand side.
//
// One block per mesh interval (intervals are independent once the next interval's first node -- an NLP
// variable -- is read as the end value); thread r owns new point r.  Operation order follows the
// reference (weights 1/prod_i(x_i - x_j + delta_ij), H = W / (x - x_j), y = (H y) / sum(H), exact node
// hits return the data value; time uses the reference's `(tf-t0)/2*tau + (tf-t0)/2`, :128).
#pragma once
#include "lpb_kernels.cuh"

namespace lpb {

template <class P>
__global__ void __launch_bounds__(64)
k_mesh_error(const __grid_constant__ ProblemDev pd, const __grid_constant__ typename P::Consts C, const __grid_constant__ MeshErrDev me,
             const double* __restrict__ x, double* __restrict__ tem, double* __restrict__ abs_err)
{
    typedef Dim<P> D;
    extern __shared__ double sm[];
    int p = 0, k = blockIdx.x;
    while (p + 1 < pd.P && k >= me.ph[p].K) { k -= me.ph[p].K; ++p; }
    const PhaseDev& ph = pd.ph[p];
    const MeshErrPhase& mp = me.ph[p];
    const int N = ph.N, M = mp.M;
    const int nk = ph.int_n[k], R = ph.int_row0[k];
    const int mk = mp.int_m[k], Rn = mp.int_rn0[k];
    const double* __restrict__ xb = x + ph.var0;
    const double t0 = xb[(size_t)D::NS * (N + 1) + (size_t)D::NC * N];
    const double tf = xb[(size_t)D::NS * (N + 1) + (size_t)D::NC * N + 1];
    double* sx = sm;                  // [nk + 1] support abscissae (states)
    double* sW = sx + (nk + 1);       // [nk + 1] barycentric weights, state support
    double* sWc = sW + (nk + 1);      // [nk]     barycentric weights, control support
    double* sT = sWc + nk;            // [mk][NS] interpolated states
    double* sF = sT + (size_t)mk * D::NSa; // [mk][NS] (tf-t0)/2 * f
    for (int j = threadIdx.x; j <= nk; j += blockDim.x)
        sx[j] = (j < nk || k + 1 < ph.K) ? ph.tau[R + j] : 1.0; // tau = [Points; 1], :58
    __syncthreads();
    for (int j = threadIdx.x; j <= nk; j += blockDim.x) {
        double pr = 1.0; // prod(X - X' + eye, 0), :24
        for (int i = 0; i <= nk; ++i) pr *= (sx[i] - sx[j]) + (i == j ? 1.0 : 0.0);
        sW[j] = 1 / pr;
        if (j < nk) {
            double pc = 1.0;
            for (int i = 0; i < nk; ++i) pc *= (sx[i] - sx[j]) + (i == j ? 1.0 : 0.0);
            sWc[j] = 1 / pc;
        }
    }
    __syncthreads();
    double* __restrict__ temp = tem + mp.out0;
    double* __restrict__ errp = abs_err + mp.out0;
    for (int r = threadIdx.x; r < mk; r += blockDim.x) {
        const double tq = mp.tnew[Rn + r];
        double xs[D::NSa], us[D::NCa];
        // states: support of nk + 1 points
        {
            int hit = -1;
            double hs = 0.0;
            double num[D::NSa];
#pragma unroll
            for (int s = 0; s < D::NS; ++s) num[s] = 0.0;
            for (int j = 0; j <= nk; ++j) {
                const double dist = tq - sx[j];
                if (dist == 0.0) { hit = j; continue; }
                const double h = sW[j] / dist;
                hs += h;
#pragma unroll
                for (int s = 0; s < D::NS; ++s) num[s] += h * xb[(size_t)s * (N + 1) + R + j];
            }
#pragma unroll
            for (int s = 0; s < D::NS; ++s) xs[s] = hit >= 0 ? xb[(size_t)s * (N + 1) + R + hit] : num[s] / hs;
        }
        // controls: support of nk points (the interval's collocated nodes), :94-102
        {
            int hit = -1;
            double hs = 0.0;
            double num[D::NCa];
#pragma unroll
            for (int s = 0; s < D::NC; ++s) num[s] = 0.0;
            for (int j = 0; j < nk; ++j) {
                const double dist = tq - sx[j];
                if (dist == 0.0) { hit = j; continue; }
                const double h = sWc[j] / dist;
                hs += h;
#pragma unroll
                for (int s = 0; s < D::NC; ++s) num[s] += h * xb[(size_t)D::NS * (N + 1) + (size_t)s * N + R + j];
            }
#pragma unroll
            for (int s = 0; s < D::NC; ++s) us[s] = hit >= 0 ? xb[(size_t)D::NS * (N + 1) + (size_t)s * N + R + hit] : num[s] / hs;
        }
        const double t = (tf - t0) / 2 * tq + (tf - t0) / 2; // :128 (sic: not + (tf+t0)/2)
        double f[D::NSa], c[D::NPa];
        P::dae(C, p + 1, t, xs, us, f, c);
#pragma unroll
        for (int s = 0; s < D::NS; ++s) {
            sT[(size_t)r * D::NSa + s] = xs[s];
            sF[(size_t)r * D::NSa + s] = f[s] * ((tf - t0) / 2.0); // :135
            temp[(size_t)s * (M + 1) + Rn + r] = xs[s];
        }
    }
    __syncthreads();
    const double* __restrict__ A = mp.ablocks + mp.int_a0[k];
    for (int r = threadIdx.x; r < mk; r += blockDim.x) {
#pragma unroll
        for (int s = 0; s < D::NS; ++s) {
            double acc = 0.0; // IntegrationMatrix * daeout: COO order = ascending column per row, zeros dropped
            for (int j = 0; j < mk; ++j) {
                const double a = A[(size_t)j * mk + r];
                if (a != 0.0) acc += a * sF[(size_t)j * D::NSa + s];
            }
            const double integ = 1.0 * sT[s] + acc; // UnityMatrix * temState: the interval's first point, :151
            const double target = (r + 1 < mk) ? sT[(size_t)(r + 1) * D::NSa + s] : xb[(size_t)s * (N + 1) + R + nk];
            errp[(size_t)s * (M + 1) + Rn + r + 1] = fabs(integ - target);
        }
    }
    if (k == 0 && threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < D::NS; ++s) errp[(size_t)s * (M + 1)] = 0.0; // integratedRHS row 0 = temState row 0
    }
    if (k == ph.K - 1 && threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < D::NS; ++s) temp[(size_t)s * (M + 1) + M] = xb[(size_t)s * (N + 1) + N]; // final state row, :109
    }
}

template <class P>
int launch_mesh_error(const ProblemDev& pd, const void* consts, cudaStream_t st, const MeshErrDev& me, int total_intervals, int max_n,
                      const double* x, double* tem, double* abs_err)
{
    typedef Dim<P> D;
    const typename P::Consts& C = *static_cast<const typename P::Consts*>(consts);
    const size_t shm = ((size_t)3 * (max_n + 1) + (size_t)2 * (max_n + 1) * D::NSa) * sizeof(double);
    if (shm > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_mesh_error<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
        if (e != cudaSuccess) return cuda_fail(e);
    }
    k_mesh_error<P><<<total_intervals, 64, shm, st>>>(pd, C, me, x, tem, abs_err);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 1 : cuda_fail(e);
}

} // namespace lpb
