// lpb_convert.cuh -- NLP solution -> optimal-control solution on the GPU.
//
//   k_nlp2op_ends / k_nlp2op_nodes / k_nlp2op_final  replace Nlp2OpConverter::Nlp2OpControl
//   (Lpopc/src/Core/Nlp2OPConverter.cpp:13-196) for the step right after the NLP solve: per phase
//     time        t = (tf - t0) (tau + 1) / 2 + t0 at the N LGR nodes and the end point          (:56-58)
//     state       (N+1) x ns, control (N+1) x nc with the end row extrapolated by the natural cubic
//                 spline of LpGuessChecker::spline_interpolation (LpGuessChecker.cpp:208-262)    (:60-72)
//     costate     -[ lambda / w ; D(:, last)' lambda ]                                           (:80-87)
//     pathmult    2 lambda_path / w / (tf - t0), end row by the same spline                      (:89-121)
//     Hamiltonian L + sum_s costate_s f_s at all N+1 points                                      (:147-151)
//     costs       Mayer, (tf - t0) w' L / 2                                                      (:132-138)
// Output layout of one phase (doubles, column-major matrices with N+1 rows), phases concatenated:
//     [ time | state | control | costate | pathmult | Hamiltonian | mayer, lagrange ]
//
// Reference quirks replicated: the path multipliers are sliced from the UNSCALED multiplier vector at
// offset N*ns of the whole vector, not of the phase (:92, wrong for phases after the first); fenced: the
// reference leaves SolCost::initial_time_ unset before MayerCost (:124 assigns initial_state_ twice), here
// the Mayer functor receives the real t0 (the shipped examples' Mayer costs do not read it).
#pragma once
#include "lpb_kernels.cuh"

namespace lpb {

struct ConvertDev {
    long long out0[kMaxPhases]; // first output double of the phase
};

template <class P>
__host__ __device__ inline long long nlp2op_phase_doubles(int N)
{
    return (long long)(N + 1) * (2 + 2 * P::NS + P::NC + P::NPATH) + 2;
}

// value at x = 1 of the natural cubic spline through (xd[i], y(i)), i < n, xd strictly increasing and < 1
// (LpGuessChecker.cpp:208-262 with kleft = n-1, kright = n: only c[n-2] = 2 z[n-2] and c[n-1] = 0 are read).
//
// The reference runs the forward sweep mu_i = h_i / l_i, z_i = (alpha_i - h_{i-1} z_{i-1}) / l_i over all n knots: a
// serial chain of 2 n divisions (33 ms for one column of 100 000 nodes).  The sweep is a contraction: l_i =
// 2 (h_i + h_{i-1}) - h_{i-1} mu_{i-1} with 0 <= mu_{i-1} < 1/2 gives mu_i < 1/2 and |d z_i / d z_{i-1}| = h_{i-1} / l_i
// < 2/3, |d mu_i / d mu_{i-1}| = mu_i h_{i-1} / l_i < 1/3, so whatever (mu, z) enters knot i is forgotten by a factor
// below (2/3)^k after k knots.  Only z[n-2] and mu[n-2] are read, so the sweep starts kSplineWindow = 512 knots before
// the end with the natural start (mu, z) = (0, 0): the difference to the full sweep is below (2/3)^512 ~ 1e-90 of the
// largest |z| of the column -- no double can tell the two apart unless the column spans more than 70 decades.
constexpr int kSplineWindow = 512;

template <class Y>
__device__ inline double spline_end_value(const double* __restrict__ xd, int n, Y y)
{
    double mu = 0.0, z = 0.0;
    const int first = (n - 1 > kSplineWindow) ? n - 1 - kSplineWindow : 1;
    for (int i = first; i < n - 1; ++i) {
        const double him1 = xd[i] - xd[i - 1];
        const double hi = xd[i + 1] - xd[i];
        const double alphai = 3.0 / hi * (y(i + 1) - y(i)) - 3.0 / him1 * (y(i) - y(i - 1));
        const double li = 2 * (xd[i + 1] - xd[i - 1]) - him1 * mu;
        mu = hi / li;
        z = (alphai - him1 * z) / li;
    }
    const double c_last = 0.0;
    double c_prev = z - mu * c_last; // c[n-2] (z[0] = mu[0] = 0 when n == 2)
    if (n - 2 >= 1) c_prev = 2 * c_prev;
    const double x = 1.0;
    const double h = xd[n - 1] - xd[n - 2];
    const double A = (xd[n - 1] - x) / h;
    const double B = (x - xd[n - 2]) / h;
    const double Cc = (pow(A, 3.0) - A) * (h * h) / 6.0;
    const double Dd = (pow(B, 3.0) - B) * (h * h) / 6.0;
    return A * y(n - 2) + B * y(n - 1) + Cc * c_prev + Dd * c_last;
}

// stage 1: end rows that need a pass over the whole phase -- one block per phase, one thread per column
// (nc control columns, np path-multiplier columns, ns terminal costates)
template <class P>
__global__ void __launch_bounds__(64)
k_nlp2op_ends(const __grid_constant__ ProblemDev pd, const __grid_constant__ ConvertDev cv,
              const double* __restrict__ x, const double* __restrict__ lambda, double* __restrict__ out)
{
    typedef Dim<P> D;
    const int p = blockIdx.x;
    const PhaseDev& ph = pd.ph[p];
    const int N = ph.N, M = N + 1;
    const double* __restrict__ xb = x + ph.var0;
    const double t0 = xb[(size_t)D::NS * M + (size_t)D::NC * N];
    const double tf = xb[(size_t)D::NS * M + (size_t)D::NC * N + 1];
    double* __restrict__ o = out + cv.out0[p];
    double* __restrict__ o_control = o + (size_t)M * (1 + D::NS);
    double* __restrict__ o_costate = o_control + (size_t)M * D::NC;
    double* __restrict__ o_pathmult = o_costate + (size_t)M * D::NS;
    for (int c = threadIdx.x; c < D::NC + D::NP + D::NS; c += blockDim.x) {
        if (c < D::NC) {
            const double* __restrict__ u = xb + (size_t)D::NS * M + (size_t)c * N;
            o_control[(size_t)c * M + N] = spline_end_value(ph.tau, N, [&](int i) { return u[i]; });
        } else if (c < D::NC + D::NP) {
            const int j = c - D::NC;
            // quirk: offset N*ns into the WHOLE multiplier vector (:92)
            const double* __restrict__ lp = lambda + (size_t)N * D::NS + (size_t)j * N;
            const double* __restrict__ w = ph.w;
            o_pathmult[(size_t)j * M + N] =
                spline_end_value(ph.tau, N, [&](int i) { return 2 * (1 / w[i]) * lp[i] / (tf - t0); });
        } else {
            // terminal costate: last column of the composite D (non-zero in the last interval only) times the
            // defect multipliers of the state, rows in ascending order (:85-86)
            const int s = c - D::NC - D::NP;
            const int I = ph.K - 1;
            const int row0 = ph.int_row0[I], nI = ph.int_n[I];
            const double* __restrict__ Db = ph.dblocks + ph.int_d0[I] + (size_t)nI * nI; // column nI of the N_I x (N_I+1) block
            const double* __restrict__ ls = lambda + ph.con0 + (size_t)s * N;
            double acc = 0.0;
            for (int r = 0; r < nI; ++r) acc += Db[r] * ls[row0 + r];
            o_costate[(size_t)s * M + N] = -acc;
        }
    }
}

// stage 2: one thread per point (N collocated nodes + the end point) of every phase
template <class P>
__global__ void __launch_bounds__(128)
k_nlp2op_nodes(const __grid_constant__ ProblemDev pd, const __grid_constant__ typename P::Consts C, const __grid_constant__ ConvertDev cv,
               const double* __restrict__ x, const double* __restrict__ lambda, double* __restrict__ out, double* __restrict__ scr)
{
    typedef Dim<P> D;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= pd.total_nodes + pd.P) return;
    int p = 0;
    while (p + 1 < pd.P && gid >= pd.ph[p + 1].node0 + (p + 1)) ++p;
    const PhaseDev& ph = pd.ph[p];
    const int k = gid - (ph.node0 + p);
    const int N = ph.N, M = N + 1;
    const double* __restrict__ xb = x + ph.var0;
    const double t0 = xb[(size_t)D::NS * M + (size_t)D::NC * N];
    const double tf = xb[(size_t)D::NS * M + (size_t)D::NC * N + 1];
    double* __restrict__ o = out + cv.out0[p];
    double* __restrict__ o_time = o;
    double* __restrict__ o_state = o + M;
    double* __restrict__ o_control = o_state + (size_t)M * D::NS;
    double* __restrict__ o_costate = o_control + (size_t)M * D::NC;
    double* __restrict__ o_pathmult = o_costate + (size_t)M * D::NS;
    double* __restrict__ o_ham = o_pathmult + (size_t)M * D::NP;
    const double tau = k < N ? ph.tau[k] : 1.0;
    const double t = (tf - t0) * (tau + 1) / 2 + t0; // :58
    o_time[k] = t;
    double xs[D::NSa], us[D::NCa], cs[D::NSa];
#pragma unroll
    for (int j = 0; j < D::NS; ++j) {
        xs[j] = xb[(size_t)j * M + k];
        o_state[(size_t)j * M + k] = xs[j];
    }
#pragma unroll
    for (int j = 0; j < D::NC; ++j) {
        if (k < N) {
            us[j] = xb[(size_t)D::NS * M + (size_t)j * N + k];
            o_control[(size_t)j * M + k] = us[j];
        } else us[j] = o_control[(size_t)j * M + N]; // written by k_nlp2op_ends
    }
#pragma unroll
    for (int j = 0; j < D::NS; ++j) {
        if (k < N) {
            cs[j] = -((1 / ph.w[k]) * lambda[ph.con0 + (size_t)j * N + k]); // :83-87
            o_costate[(size_t)j * M + k] = cs[j];
        } else cs[j] = o_costate[(size_t)j * M + N];
    }
    if (k < N) {
#pragma unroll
        for (int j = 0; j < D::NP; ++j)
            o_pathmult[(size_t)j * M + k] = 2 * (1 / ph.w[k]) * lambda[(size_t)N * D::NS + (size_t)j * N + k] / (tf - t0); // :92-97
    }
    double f[D::NSa], c[D::NPa];
    P::dae(C, p + 1, t, xs, us, f, c);
    const double L = P::lagrange(C, p + 1, t, xs, us);
    double hsum = 0.0; // sum(costate % dae, 1): columns left to right (:151)
#pragma unroll
    for (int j = 0; j < D::NS; ++j) hsum += cs[j] * f[j];
    o_ham[k] = L + hsum;
    if (k < N) scr[ph.node0 + k] = ((tf - t0) * ph.w[k]) * L; // terms of (tf - t0) w' L (:137)
}

// stage 3: one block per phase: costs
template <class P>
__global__ void __launch_bounds__(256)
k_nlp2op_final(const __grid_constant__ ProblemDev pd, const __grid_constant__ typename P::Consts C, const __grid_constant__ ConvertDev cv,
               const double* __restrict__ x, const double* __restrict__ scr, double* __restrict__ out)
{
    typedef Dim<P> D;
    __shared__ double sm[256];
    const int p = blockIdx.x;
    const PhaseDev& ph = pd.ph[p];
    const int N = ph.N, M = N + 1;
    const double dot = block_sum<256>(scr + ph.node0, N, sm);
    if (threadIdx.x == 0) {
        const double* xb = x + ph.var0;
        double x0[D::NSa], xf[D::NSa];
#pragma unroll
        for (int j = 0; j < D::NS; ++j) { x0[j] = xb[(size_t)j * M]; xf[j] = xb[(size_t)j * M + N]; }
        const double t0 = xb[(size_t)D::NS * M + (size_t)D::NC * N];
        const double tf = xb[(size_t)D::NS * M + (size_t)D::NC * N + 1];
        double* o = out + cv.out0[p] + (long long)M * (2 + 2 * D::NS + D::NC + D::NP);
        o[0] = P::mayer(C, p + 1, t0, x0, tf, xf);
        o[1] = dot / 2.0;
    }
}

template <class P>
int launch_nlp2op(const ProblemDev& pd, const void* consts, cudaStream_t st, const double* x, const double* lambda,
                  double* out, double* scratch, long long* phase_offsets)
{
    const typename P::Consts& C = *static_cast<const typename P::Consts*>(consts);
    ConvertDev cv;
    long long off = 0;
    for (int p = 0; p < pd.P; ++p) {
        cv.out0[p] = off;
        if (phase_offsets) phase_offsets[p] = off;
        off += nlp2op_phase_doubles<P>(pd.ph[p].N);
    }
    if (phase_offsets) phase_offsets[pd.P] = off;
    if (!out) return 0; // layout query
    const int pts = pd.total_nodes + pd.P;
    k_nlp2op_ends<P><<<pd.P, 64, 0, st>>>(pd, cv, x, lambda, out);
    k_nlp2op_nodes<P><<<(pts + 127) / 128, 128, 0, st>>>(pd, C, cv, x, lambda, out, scratch);
    k_nlp2op_final<P><<<pd.P, 256, 0, st>>>(pd, C, cv, x, scratch, out);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 3 : cuda_fail(e);
}

} // namespace lpb
