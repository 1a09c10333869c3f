// oracle/lp_types.hpp -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Minimal dense/sparse containers and the argument structs of the reference's
// problem-definition interface, for the CPU restatement of lpopc's transcription
// path.  Nothing in the shipped product (lpopc_b200/) may include this.
//
// Follows:
//   Lpopc/src/Core/LpFunctionWrapper.h:12-69      (SolCost/SolDae/SolEvent/SolLink, FunctionWrapper)
//   Lpopc/src/SparseMatrix/LpSparseMatrix.cpp:53-155,240-272  (dsmatrix: Sparse, GeneratRowColValue, operator*, Find)
// Armadillo (unpinned, not vendored) is replaced by `Mat`/`Vec` below: column-major
// storage like arma::mat so every reshape/col() in the reference maps 1:1.
#pragma once
#include <cmath>
#include <cstddef>
#include <stdexcept>
#include <string>
#include <vector>

namespace lpo {

typedef std::vector<double> Vec;

struct Mat { // column-major, like arma::mat
    int n_rows = 0, n_cols = 0;
    std::vector<double> a;
    Mat() {}
    Mat(int r, int c, double fill = 0.0) : n_rows(r), n_cols(c), a((size_t)r * c, fill) {}
    double& operator()(int i, int j) { return a[(size_t)i + (size_t)j * n_rows]; }
    double operator()(int i, int j) const { return a[(size_t)i + (size_t)j * n_rows]; }
    size_t n_elem() const { return a.size(); }
    Vec col(int j) const { return Vec(a.begin() + (size_t)j * n_rows, a.begin() + (size_t)(j + 1) * n_rows); }
    void set_col(int j, const Vec& v) { for (int i = 0; i < n_rows; ++i) (*this)(i, j) = v[i]; }
};

// reshape(vec, r, c): column-major reinterpretation (arma::reshape on a column vector)
inline Mat reshape(const Vec& v, int r, int c)
{
    Mat m(r, c);
    for (size_t i = 0; i < m.a.size(); ++i) m.a[i] = v[i];
    return m;
}

// ---- LpFunctionWrapper.h:12-49 ------------------------------------------------
struct SolCost {
    int phase_num_ = 0;
    double initial_time_ = 0;
    Vec initial_state_;
    double terminal_time_ = 0;
    Vec terminal_state_;
    Vec time_;
    Mat state_;
    Mat control_;
    Vec parameter_;
};
struct SolDae {
    int phase_num_ = 0;
    Vec time_;
    Mat state_;
    Mat contol_;
    Vec parameter_;
};
struct SolEvent {
    int phase_num_ = 0;
    double initial_time_ = 0, terminal_time_ = 0;
    Vec initial_state_, terminal_state_;
    Vec parameter_;
};
struct SolLink {
    int left_phase_num_ = 0, right_phase_num_ = 0;
    size_t ipair = 0;
    Vec left_state_, right_state_;
    Vec left_parameter_, right_parameter_;
};

// ---- LpFunctionWrapper.h:50-69 ------------------------------------------------
class FunctionWrapper {
public:
    virtual ~FunctionWrapper() {}
    virtual void MayerCost(SolCost&, double& mayer) { mayer = 0; }
    virtual void DerivMayer(SolCost&, Vec& /*rowvec*/) {}
    virtual void LagrangeCost(SolCost&, Vec& /*N*/) {}
    virtual void DerivLagrange(SolCost&, Mat&) {}
    virtual void DaeFunction(SolDae&, Mat& /*N x ns*/, Mat& /*N x np*/) {}
    virtual void DerivDae(SolDae&, Mat&, Mat&) {}
    virtual void EventFunction(SolEvent&, Vec&) {}
    virtual void DerivEvent(SolEvent&, Mat&) {}
    virtual void LinkFunction(SolLink&, Vec&) {}
    virtual void DerivLink(SolLink&, Mat&) {}
    virtual bool HasAnalytic() const { return false; }
};

// ---- SparseMatrix/LpSparseMatrix.cpp ------------------------------------------
struct dsmatrix { // COO, int indices, double values; storage order is significant
    int m_ = 0, n_ = 0;
    std::vector<int> rows, cols;
    std::vector<double> vals;
    int GetLength() const { return (int)vals.size(); }

    // LpSparseMatrix.cpp:53-78 (indices arrive as doubles and are truncated to int)
    static dsmatrix Sparse(const Vec& r, const Vec& c, const Vec& v, int nrow, int ncol)
    {
        if (r.size() != c.size() || r.size() != v.size()) throw std::runtime_error("dsmatrix::Sparse size mismatch");
        dsmatrix s;
        s.m_ = nrow; s.n_ = ncol;
        s.rows.resize(r.size()); s.cols.resize(r.size()); s.vals.resize(r.size());
        for (size_t i = 0; i < r.size(); ++i) { s.rows[i] = int(r[i]); s.cols[i] = int(c[i]); s.vals[i] = v[i]; }
        return s;
    }
    // LpSparseMatrix.cpp:80-105: column-major flatten of a dense block with shifts
    static void GeneratRowColValue(const Mat& d, Vec& irow, Vec& jcol, Vec& values, int rowShift, int colShift)
    {
        irow.assign(d.n_elem(), 0.0); jcol.assign(d.n_elem(), 0.0); values.assign(d.n_elem(), 0.0);
        int k = 0;
        for (int j = 0; j < d.n_cols; ++j)
            for (int i = 0; i < d.n_rows; ++i) { irow[k] = i + rowShift; jcol[k] = j + colShift; ++k; }
        for (size_t i = 0; i < values.size(); ++i) values[i] = d.a[i];
    }
    // LpSparseMatrix.cpp:240-272: drop exact zeros, keep storage order
    static void Find(const dsmatrix& s, Vec& I, Vec& J, Vec& V)
    {
        I.clear(); J.clear(); V.clear();
        for (int i = 0; i < s.GetLength(); ++i)
            if (s.vals[i] != 0.0) { I.push_back((double)s.rows[i]); J.push_back((double)s.cols[i]); V.push_back(s.vals[i]); }
    }
    // LpSparseMatrix.cpp:127-155: COO x dense, accumulation in storage order
    Mat operator*(const Mat& op) const
    {
        if (n_ != op.n_rows) throw std::runtime_error("dsmatrix*: dimension mismatch");
        Mat res(m_, op.n_cols, 0.0);
        for (int icol = 0; icol < op.n_cols; ++icol) {
            Vec temcol = op.col(icol);
            for (int k = 0; k < GetLength(); ++k) res.a[(size_t)rows[k] + (size_t)icol * m_] += vals[k] * temcol[cols[k]];
        }
        return res;
    }
};

struct LpoError : public std::runtime_error {
    explicit LpoError(const std::string& s) : std::runtime_error(s) {}
};

} // namespace lpo
