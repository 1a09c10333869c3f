// lpb_tables.cpp -- see lpb_tables.hpp.
#include "lpb_tables.hpp"
#include <cmath>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <algorithm>

namespace lpb {

void lgr_points(int n, std::vector<double>& x, std::vector<double>& w)
{
    // Newton iteration on P_{n-1}(x) + P_n(x) with the Legendre three-term recurrence
    // (RPMGenerator.cpp:253-291): start x_k = -cos(2 pi k/(2n-1)), x_0 stays at -1,
    // iterate until max|dx| <= eps.
    static std::map<int, std::pair<std::vector<double>, std::vector<double>>> cache;
    static std::mutex mu;
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(n);
        if (it != cache.end()) { x = it->second.first; w = it->second.second; return; }
    }
    const double pi = 3.14159265358979323846;
    const double eps = 2.220446049250313e-16;
    const int Nm = n - 1; // degree
    x.assign(n, 0.0);
    for (int k = 0; k < n; ++k) x[k] = -1 * std::cos((double)k * ((2 * pi) / (2 * Nm + 1)));
    std::vector<double> P((size_t)n * (n + 1), 0.0); // P[i + n*deg]
    std::vector<double> xold(n, 2.0);
    for (;;) {
        double mx = 0.0;
        for (int i = 0; i < n; ++i) mx = std::fmax(mx, std::fabs(x[i] - xold[i]));
        if (!(mx > eps)) break;
        xold = x;
        for (int i = 0; i < n; ++i) {
            double* Pi = &P[i];
            Pi[0] = 1.0;
            Pi[n] = x[i];
            for (int k = 1; k < n; ++k) Pi[(size_t)n * (k + 1)] = (x[i] * (2 * k + 1) * Pi[(size_t)n * k] - (Pi[(size_t)n * (k - 1)] * k)) / (k + 1);
        }
        for (int i = 1; i < n; ++i) {
            double Pa = P[i + (size_t)n * (n - 1)], Pb = P[i + (size_t)n * n];
            double step = (1.0 - xold[i]) / n;
            step = step * (Pa + Pb);
            x[i] = xold[i] - (step / (Pa - Pb));
        }
    }
    w.assign(n, 0.0);
    w[0] = 2.0 / (n * n);
    for (int i = 1; i < n; ++i) {
        double q = P[i + (size_t)n * Nm] * n;
        w[i] = (1 - x[i]) / (q * q);
    }
    std::lock_guard<std::mutex> lk(mu);
    cache[n] = std::make_pair(x, w);
}

// pairwise two-accumulator sum, the reduction Armadillo's sum() applies per column
static double sum2(const double* s, int n)
{
    double a = 0.0, b = 0.0;
    int i, j;
    for (i = 0, j = 1; j < n; i += 2, j += 2) { a += s[i]; b += s[j]; }
    if (i < n) a += s[i];
    return a + b;
}

void colloc_block(const std::vector<double>& s, std::vector<double>& D)
{
    const int M = (int)s.size(), n = M - 1;
    // Y(i,j) = (delta_ij + s_i) - s_j ; p_i = prod_j Y(i,j) ; G(i,j) = (1/p_i)/((1/p_j) Y(i,j))
    std::vector<double> Y((size_t)M * M), p(M), G((size_t)M * M);
    for (int j = 0; j < M; ++j)
        for (int i = 0; i < M; ++i) Y[i + (size_t)j * M] = ((i == j ? 1.0 : 0.0) + s[i]) - s[j];
    for (int i = 0; i < M; ++i) {
        double acc = 1.0;
        for (int j = 0; j < M; ++j) acc *= Y[i + (size_t)j * M];
        p[i] = acc;
    }
    for (int j = 0; j < M; ++j)
        for (int i = 0; i < M; ++i) G[i + (size_t)j * M] = (1 / p[i]) / ((1 / p[j]) * Y[i + (size_t)j * M]);
    // diagonal from the column sums: G(j,j) = 1 - sum_i G(i,j)   (RPMGenerator.cpp:116-122)
    for (int j = 0; j < M; ++j) G[j + (size_t)j * M] = 1 - sum2(&G[(size_t)j * M], M);
    // D = -G^T with the last row dropped, column-major n x M
    D.assign((size_t)n * M, 0.0);
    for (int j = 0; j < M; ++j)
        for (int i = 0; i < n; ++i) D[i + (size_t)j * n] = -G[j + (size_t)i * M];
}

void build_phase_tables(int K, const double* mesh, const int* nodes, PhaseTables& out)
{
    if (K < 1) throw std::runtime_error("MeshRefinement need at least two meshPoints");
    if (mesh[0] != -1 || mesh[K] != 1) throw std::runtime_error("meshPoints must span -1 to +1");
    out = PhaseTables();
    out.K = K;
    // offsets first, so that the intervals can be filled independently (and in parallel for large meshes: the
    // collocation block of an interval is O(n^3) host arithmetic, 5 ms for 10 000 intervals of 10 nodes on one thread)
    out.int_n.resize(K); out.int_row0.resize(K); out.int_d0.resize(K);
    long long d0 = 0;
    int row0 = 0;
    for (int k = 0; k < K; ++k) {
        const int n = nodes[k];
        if (n < 2) throw std::runtime_error("nodes per interval must be >= 2");
        out.int_n[k] = n; out.int_row0[k] = row0; out.int_d0[k] = d0;
        d0 += (long long)n * (n + 1);
        row0 += n;
    }
    out.N = row0;
    out.tau.resize(row0); out.w.resize(row0); out.node_interval.resize(row0);
    out.dblocks.resize((size_t)d0);
    auto fill = [&](int k0, int k1) {
        std::vector<double> x, w, sall, D;
        for (int k = k0; k < k1; ++k) {
            const int n = nodes[k], r0 = out.int_row0[k];
            lgr_points(n, x, w);
            const double tspan = mesh[k + 1] - mesh[k];
            sall.resize(n + 1);
            for (int i = 0; i < n; ++i) {
                double s = x[i] + 1; // (x+1)*tspan/2 + mesh_k   (RPMGenerator.cpp:67-69)
                s *= tspan / 2.0;
                s += mesh[k];
                sall[i] = s;
                out.tau[r0 + i] = s;
                double ws = w[i] / 2; // w/2*tspan                (:72-73)
                ws *= tspan;
                out.w[r0 + i] = ws;
                out.node_interval[r0 + i] = k;
            }
            sall[n] = mesh[k + 1];
            colloc_block(sall, D);
            std::copy(D.begin(), D.end(), out.dblocks.begin() + out.int_d0[k]);
        }
    };
    unsigned hw = std::thread::hardware_concurrency();
    int nt = K >= 1024 ? (int)(hw > 16 ? 8 : (hw > 1 ? hw / 2 : 1)) : 1;
    if (nt <= 1) {
        fill(0, K);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(fill, (int)((long long)K * t / nt), (int)((long long)K * (t + 1) / nt));
        for (auto& t : th) t.join();
    }
}

// inverse by Gauss-Jordan elimination with partial pivoting, column-major n x n (the reference calls
// Armadillo's inv(), i.e. LAPACK; values agree to rounding)
static void invert(std::vector<double>& a, int n, std::vector<double>& r)
{
    r.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) r[i + (size_t)i * n] = 1.0;
    auto A = [&](int i, int j) -> double& { return a[i + (size_t)j * n]; };
    auto R = [&](int i, int j) -> double& { return r[i + (size_t)j * n]; };
    for (int c = 0; c < n; ++c) {
        int p = c;
        for (int i = c + 1; i < n; ++i)
            if (std::fabs(A(i, c)) > std::fabs(A(p, c))) p = i;
        if (A(p, c) == 0.0) throw std::runtime_error("integration matrix: singular differentiation block");
        if (p != c)
            for (int j = 0; j < n; ++j) { std::swap(A(p, j), A(c, j)); std::swap(R(p, j), R(c, j)); }
        const double d = A(c, c);
        for (int j = 0; j < n; ++j) { A(c, j) /= d; R(c, j) /= d; }
        for (int i = 0; i < n; ++i) {
            if (i == c) continue;
            const double f = A(i, c);
            if (f == 0.0) continue;
            for (int j = 0; j < n; ++j) { A(i, j) -= f * A(c, j); R(i, j) -= f * R(c, j); }
        }
    }
}

void build_error_tables(int K, const double* mesh, const int* nodes, const std::vector<double>& tau_old, ErrTables& out)
{
    out = ErrTables();
    out.K = K;
    std::vector<int> finer(K);
    for (int k = 0; k < K; ++k) finer[k] = nodes[k] + 1;
    PhaseTables ft;
    build_phase_tables(K, mesh, finer.data(), ft); // RPM->initialize(K, meshPoints, nodesPerInterval + 1), :150
    long long a0 = 0;
    int rn0 = 0, r0 = 0;
    for (int k = 0; k < K; ++k) {
        const int m = finer[k];
        std::vector<double> x, w;
        lgr_points(m, x, w); // RPMGenerator::GetLGRPoints(n + 1, ...), :76
        const double time0 = tau_old[r0];
        const double timef = (k + 1 < K) ? tau_old[r0 + nodes[k]] : 1.0; // tau = [Points; 1], :58
        for (int i = 0; i < m; ++i) out.tnew.push_back((x[i] + 1) * (timef - time0) / 2 + time0); // :77
        // A_k = inv(D_k(:, 1:end))
        const double* Dk = ft.dblocks.data() + ft.int_d0[k]; // column-major m x (m+1)
        std::vector<double> sub((size_t)m * m), inv;
        for (int j = 0; j < m; ++j)
            for (int i = 0; i < m; ++i) sub[i + (size_t)j * m] = Dk[i + (size_t)(j + 1) * m];
        invert(sub, m, inv);
        out.int_m.push_back(m);
        out.int_rn0.push_back(rn0);
        out.int_a0.push_back(a0);
        out.ablocks.insert(out.ablocks.end(), inv.begin(), inv.end());
        a0 += (long long)m * m;
        rn0 += m;
        r0 += nodes[k];
    }
    out.M = rn0;
}

} // namespace lpb
