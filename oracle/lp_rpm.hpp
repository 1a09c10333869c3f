// oracle/lp_rpm.hpp -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of Lpopc::RPMGenerator: LGR points/weights and the composite
// Radau differentiation matrices.
// Follows Lpopc/src/Core/RPMGenerator.cpp:17-41 (cache), :43-105 (initialize),
// :107-130 (CollocD), :132-181 (CompositeD), :253-291 (GetLGRPointsImp).
// Armadillo reductions used there: prod(X,1) = sequential product along a row,
// sum(X) = per-column arrayops::accumulate (two interleaved accumulators).
#pragma once
#include "lp_types.hpp"
#include <map>

namespace lpo {

// arma::arrayops::accumulate: two accumulators over even/odd positions
inline double arma_accumulate(const double* src, int n)
{
    double acc1 = 0.0, acc2 = 0.0;
    int i, j;
    for (i = 0, j = 1; j < n; i += 2, j += 2) { acc1 += src[i]; acc2 += src[j]; }
    if (i < n) acc1 += src[i];
    return acc1 + acc2;
}

class RPMGenerator {
public:
    // RPMGenerator.cpp:253-291
    static void GetLGRPointsImp(const int iniN, Vec& x, Vec& w)
    {
        const double pi = 3.14159265358979323846;
        int N = iniN - 1, N1 = N + 1;
        double eps = 2.220446049250313e-16; // datum::eps
        x.assign(N1, 0.0);
        for (int k = 0; k <= N; ++k) x[k] = -1 * std::cos((double)k * ((2 * pi) / (2 * N + 1)));
        Mat P(N1, N1 + 1, 0.0);
        Vec xold(N1, 2.0);
        for (;;) {
            double mx = 0.0;
            for (int i = 0; i < N1; ++i) mx = std::fmax(mx, std::fabs(x[i] - xold[i]));
            if (!(mx > eps)) break;
            xold = x;
            for (int i = 0; i < N1; ++i) { P(i, 0) = 1.0; P(i, 1) = x[i]; }
            for (int k = 1; k < N1; ++k)
                for (int i = 0; i < N1; ++i) {
                    double ret1 = x[i] * (2 * k + 1) * P(i, k) - (P(i, k - 1) * k);
                    P(i, k + 1) = ret1 / (k + 1);
                }
            for (int i = 1; i <= N; ++i) {
                double ret2 = (1.0 - xold[i]) / N1;
                ret2 = ret2 * (P(i, N1 - 1) + P(i, N1));
                ret2 = xold[i] - (ret2 / (P(i, N1 - 1) - P(i, N1)));
                x[i] = ret2;
            }
        }
        w.assign(N1, 0.0);
        w[0] = 2.0 / (N1 * N1);
        for (int i = 1; i <= N; ++i) {
            double ret4 = P(i, N) * N1;
            w[i] = (1 - x[i]) / (ret4 * ret4);
        }
    }

    // RPMGenerator.cpp:17-41: compute once per N
    static void GetLGRPoints(const int iniN, Vec& x, Vec& w)
    {
        static std::map<int, Vec> LGR_points_, LGR_wights_;
        auto xi = LGR_points_.find(iniN);
        auto wi = LGR_wights_.find(iniN);
        if (xi != LGR_points_.end() && wi != LGR_wights_.end()) { x = xi->second; w = wi->second; return; }
        GetLGRPointsImp(iniN, x, w);
        LGR_points_[iniN] = x;
        LGR_wights_[iniN] = w;
    }

    // RPMGenerator.cpp:107-130
    static void CollocD(const Vec& x, Mat& D, Mat& Dd, Mat& Do)
    {
        int N1 = (int)x.size() - 1;
        int M = (int)x.size();
        Mat Ydiff(M, M);
        for (int j = 0; j < M; ++j)
            for (int i = 0; i < M; ++i) Ydiff(i, j) = ((i == j ? 1.0 : 0.0) + x[i]) - x[j]; // eye + Y - trans(Y)
        Vec p(M);
        for (int i = 0; i < M; ++i) { // prod(Ydiff,1)
            double acc = 1.0;
            for (int j = 0; j < M; ++j) acc *= Ydiff(i, j);
            p[i] = acc;
        }
        Mat Dt(M, M); // D = ww / (trans(ww) % Ydiff), ww(i,j) = 1/p_i
        for (int j = 0; j < M; ++j)
            for (int i = 0; i < M; ++i) Dt(i, j) = (1 / p[i]) / ((1 / p[j]) * Ydiff(i, j));
        for (int j = 0; j < M; ++j) Dt(j, j) = 1 - arma_accumulate(&Dt.a[(size_t)j * M], M); // 1 - sum(D)
        D = Mat(N1, M); // -trans(D), last row dropped
        for (int j = 0; j < M; ++j)
            for (int i = 0; i < N1; ++i) D(i, j) = -Dt(j, i);
        Dd = Mat(N1, N1 + 1, 0.0);
        for (int i = 0; i < N1; ++i) Dd(i, i) = D(i, i);
        Do = Mat(N1, N1 + 1);
        for (size_t i = 0; i < Do.a.size(); ++i) Do.a[i] = D.a[i] - Dd.a[i];
    }

    // RPMGenerator.cpp:132-181
    static void CompositeD(const std::vector<Mat>& Dsect, dsmatrix& sparse_matrix)
    {
        int nodes = 0;
        for (auto& d : Dsect) nodes += d.n_rows;
        Vec vec_i, vec_j, vec_v;
        int rowshift = 0, colshift = 0;
        for (auto& d : Dsect) {
            Vec irow, jcol, valu;
            dsmatrix::GeneratRowColValue(d, irow, jcol, valu, rowshift, colshift);
            rowshift += d.n_rows;
            colshift += d.n_rows; // cols[i]-1 with cols = rows+1
            vec_i.insert(vec_i.end(), irow.begin(), irow.end());
            vec_j.insert(vec_j.end(), jcol.begin(), jcol.end());
            vec_v.insert(vec_v.end(), valu.begin(), valu.end());
        }
        sparse_matrix = dsmatrix::Sparse(vec_i, vec_j, vec_v, nodes, nodes + 1);
        dsmatrix::Find(sparse_matrix, vec_i, vec_j, vec_v);
        sparse_matrix = dsmatrix::Sparse(vec_i, vec_j, vec_v, nodes, nodes + 1);
    }

    // RPMGenerator.cpp:43-105 (integration / unity matrices are error-estimator only: out of scope)
    void initialize(int sections, const std::vector<double>& mesh_points, const std::vector<int>& node_per_interval)
    {
        std::vector<Vec> sSeg(sections), wscaled(sections);
        std::vector<Mat> Dsect(sections), Ddsect(sections), Dosect(sections);
        int tau_size_all = 0;
        for (int i = 0; i < sections; ++i) {
            Vec x, w;
            GetLGRPoints(node_per_interval[i], x, w);
            double tspan = mesh_points[i + 1] - mesh_points[i];
            Vec sSegi(x.size()), wscaledi(x.size()), sall(x.size() + 1);
            for (size_t k = 0; k < x.size(); ++k) {
                double s = x[k] + 1;
                s *= tspan / 2.0;
                s += mesh_points[i];
                sSegi[k] = s;
                sall[k] = s;
                double ws = w[k] / 2;
                ws *= tspan;
                wscaledi[k] = ws;
            }
            sall[x.size()] = mesh_points[i + 1];
            CollocD(sall, Dsect[i], Ddsect[i], Dosect[i]);
            sSeg[i] = sSegi;
            wscaled[i] = wscaledi;
            tau_size_all += (int)sSegi.size();
        }
        RPM_points_.clear(); RPM_weights_.clear();
        for (int i = 0; i < sections; ++i) {
            RPM_points_.insert(RPM_points_.end(), sSeg[i].begin(), sSeg[i].end());
            RPM_weights_.insert(RPM_weights_.end(), wscaled[i].begin(), wscaled[i].end());
        }
        (void)tau_size_all;
        CompositeD(Dsect, RPM_Differentiation_matrix_);
        CompositeD(Ddsect, RPM_Differentiation_matrix_diag_);
        CompositeD(Dosect, RPM_Differentiation_matrix_off_diag_);
    }
    Vec RPM_points_, RPM_weights_;
    dsmatrix RPM_Differentiation_matrix_, RPM_Differentiation_matrix_diag_, RPM_Differentiation_matrix_off_diag_;
};

} // namespace lpo
