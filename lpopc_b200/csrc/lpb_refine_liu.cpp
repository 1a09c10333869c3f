// lpb_refine_liu.cpp -- see lpb_refine_liu.hpp.  Every routine cites the reference lines it follows
// (Lpopc/src/Core/LpLiuHpMeshRefineAlg.cpp unless stated otherwise).  Accumulation orders are Armadillo's documented
// ones (sequential products and row sums, the two-accumulator sum of a vector), because the decisions below compare
// the results with thresholds and pick arg-maxima.
#include "lpb_refine_liu.hpp"

#include "lpb_tables.hpp"

#include <algorithm>
#include <cmath>
#include <limits>
#include <map>
#include <mutex>
#include <stdexcept>

namespace lpb {

namespace {

// sum of a vector: even and odd elements in separate accumulators, added at the end
double pair_sum(const double* p, size_t n)
{
    double a1 = 0.0, a2 = 0.0;
    size_t j;
    for (j = 1; j < n; j += 2) { a1 += *p++; a2 += *p++; }
    if ((j - 1) < n) a1 += *p;
    return a1 + a2;
}

// a + i*delta with the last point exact (linspace)
std::vector<double> linspace(double a, double b, int n)
{
    std::vector<double> v((size_t)n);
    if (n == 1) { v[0] = b; return v; }
    const double delta = (b - a) / double(n - 1);
    for (int i = 0; i + 1 < n; ++i) v[i] = a + double(i) * delta;
    v[n - 1] = b;
    return v;
}

// first index of the maximum, scanning like std::max_element (a later element wins only if it is strictly larger)
int argmax(const double* p, int n)
{
    int best = 0;
    for (int i = 1; i < n; ++i)
        if (p[best] < p[i]) best = i;
    return best;
}

// double -> count as the reference's static_cast<uword>(ceil(.)) where that is defined; false where it is not
bool to_count(double v, long long& out)
{
    if (!std::isfinite(v) || v < 0.0 || v > 1e9) return false;
    out = (long long)v;
    return true;
}

} // namespace

// LpSolutionError.cpp:10-52
void bar_lagrange_interp(const std::vector<double>& dx, const double* dy, const std::vector<double>& x, std::vector<double>& y)
{
    const int M = (int)dx.size(), N = (int)x.size();
    // barycentric weights: 1 / prod_i (x_i - x_j + delta_ij), the product running down the rows (:22-24)
    std::vector<double> w((size_t)M);
    for (int j = 0; j < M; ++j) {
        double p = 1.0;
        for (int i = 0; i < M; ++i) p *= (dx[i] - dx[j]) + (i == j ? 1.0 : 0.0);
        w[j] = 1.0 / p;
    }
    y.assign((size_t)N, 0.0);
    for (int n = 0; n < N; ++n) {
        double num = 0.0, den = 0.0; // (H * data_y)(n) and sum(H, 1)(n): sequential over the nodes (:46-49)
        int on_node = -1;
        for (int j = 0; j < M; ++j) {
            const double dist = x[n] - dx[j];
            if (dist == 0.0) on_node = j; // the reference overwrites with the LAST matching node (:32-51)
            const double hnj = w[j] / (dist == 0.0 ? std::numeric_limits<double>::quiet_NaN() : dist);
            num += hnj * dy[j];
            den += hnj;
        }
        y[n] = on_node >= 0 ? dy[on_node] : num / den;
    }
}

void LiuRefiner::reset()
{
    mesh_index_ = 0;
    mesh_history_.clear();
    state_history_.clear();
    mesh_points_history_.clear();
}

// :262-342: coefficients of the Lagrange basis on the N LGR points + 1 in powers of tau, highest power first;
// (N + 1) x (N + 1), column i belongs to support point i.  Cached per N like alj_map_.
const LiuRefiner::Dense& LiuRefiner::power_coefficients(int N)
{
    static std::map<int, Dense> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(N);
    if (it != cache.end()) return it->second;
    std::vector<double> x, wts;
    lgr_points(N, x, wts);
    x.push_back(1.0);
    Dense alj;
    alj.rows = alj.cols = N + 1;
    alj.a.assign((size_t)(N + 1) * (N + 1), 0.0);
    std::vector<double> t((size_t)N), T((size_t)N * N), Di((size_t)N + 1), powx((size_t)N + 1), prodv((size_t)N + 1);
    for (int i = 0; i <= N; ++i) {
        // the other N support points, negated (:318-322)
        int k = 0;
        for (int j = 0; j <= N; ++j)
            if (j != i) t[k++] = -x[j];
        // CalculateDi (:281-301): elementary symmetric sums by the triangular recurrence, row sums give the coefficients
        std::fill(T.begin(), T.end(), 0.0);
        for (int j = 0; j < N; ++j) T[(size_t)0 + (size_t)j * N] = t[j];
        for (int r = 1; r < N; ++r) {
            for (int j = N - 2; j >= 0; --j) T[(size_t)r + (size_t)j * N] = T[(size_t)r + (size_t)(j + 1) * N] + T[(size_t)(r - 1) + (size_t)(j + 1) * N];
            for (int j = 0; j < N; ++j) T[(size_t)r + (size_t)j * N] *= t[j];
        }
        Di[0] = 1.0;
        for (int r = 0; r < N; ++r) {
            double s = 0.0;
            for (int j = 0; j < N; ++j) s += T[(size_t)r + (size_t)j * N];
            Di[r + 1] = s;
        }
        // value of the numerator polynomial at x_i: powers by repeated multiplication from the top (:323-328)
        powx[N] = 1.0;
        for (int kk = N - 1; kk >= 0; --kk) powx[kk] = powx[kk + 1] * x[i];
        for (int r = 0; r <= N; ++r) prodv[r] = powx[r] * Di[r];
        const double denom = pair_sum(prodv.data(), (size_t)N + 1);
        for (int r = 0; r <= N; ++r) alj.a[(size_t)r + (size_t)i * (N + 1)] = Di[r] / denom;
    }
    return cache.emplace(N, std::move(alj)).first->second;
}

// :462-504: smallest degree whose power-series coefficients (scaled by 1 + max of the state) still exceed the tolerance
int LiuRefiner::reducing_N(const LiuPhaseInput& ph, int seg, const std::vector<int>& sidx, const std::vector<double>& betai, double tol) const
{
    const int Ncur = ph.nodes[seg];
    const Dense& alj = power_coefficients(Ncur);
    const int rows = Ncur + 1, Nall = (int)ph.tau.size() + 1;
    const int istart = sidx[seg];
    int best = 0;
    for (int s = 0; s < ph.ns; ++s) {
        int maxN = 1;
        for (int r = 0; r < rows; ++r) { // bil = alj * segment_state, row r = coefficient of tau^(Ncur - r)
            double acc = 0.0;
            for (int k = 0; k < rows; ++k) acc += alj.at(r, k) * ph.state[(size_t)(istart + k) + (size_t)s * Nall];
            if (acc / betai[s] > tol) { maxN = rows - 1 - r; break; } // first (highest-power) coefficient above the tolerance (:492-500)
        }
        best = std::max(best, maxN);
    }
    return std::max(2, best);
}

// :347-398 / :409-452 ("Find q"): decay exponent of the error between the previous mesh and this one
bool LiuRefiner::growth_exponent(int iphase, const LiuPhaseInput& ph, int seg, double e_k, double& q) const
{
    const MeshInfo& before = mesh_history_[mesh_history_.size() - 2][iphase];
    const double t0 = ph.mesh[seg], tf = ph.mesh[seg + 1];
    const double hcur = tf - t0;
    int itmin = -1, itmax = -1;
    for (int i = 0; i < (int)before.mesh.size(); ++i) {
        if (before.mesh[i] <= t0) itmin = i;                // max(find(meshpointbefore <= t0))
        if (before.mesh[i] >= tf && itmax < 0) itmax = i;   // min(find(meshpointbefore >= tf))
    }
    if (itmin < 0 || itmax < 0 || itmax <= itmin) return false;
    const double h_b = before.mesh[itmax] - before.mesh[itmin];
    const int N = ph.nodes[seg];
    long long N_b = 0;
    double e_k_b = before.e_k[itmin];
    for (int i = itmin; i < itmax; ++i) {
        N_b += before.nodes[i];
        if (e_k_b < before.e_k[i]) e_k_b = before.e_k[i];
    }
    const double fN = (double)N / (double)N_b, fh = hcur / h_b, fe = e_k / e_k_b;
    q = std::ceil(std::log(fe / std::pow((double)N, 5.0 / 2.0)) / std::log(fh / fN));
    return true;
}

// :722-745: second differences of the interpolant on 501 points
void LiuRefiner::second_derivative(const std::vector<double>& t, const Dense& x, std::vector<double>& interp_t, Dense& d2)
{
    const int ntime = (int)t.size();
    const double tf = t[ntime - 1], t0 = t[0];
    std::vector<double> tau((size_t)ntime);
    for (int i = 0; i < ntime; ++i) tau[i] = 2.0 * (t[i] - t0) / (tf - t0) - 1.0;
    const double taustep = 2.0 / 500.0;
    const std::vector<double> tpert = linspace(t0, tf, 501);
    const int np = 501;
    d2.rows = np - 2;
    d2.cols = x.cols;
    d2.a.assign((size_t)(np - 2) * x.cols, 0.0);
    std::vector<double> col;
    for (int s = 0; s < x.cols; ++s) {
        // nodes on [-1, 1], abscissae in the units of t: as the reference does (:731)
        bar_lagrange_interp(tau, &x.a[(size_t)s * x.rows], tpert, col);
        for (int i = 0; i < np - 2; ++i) d2.a[(size_t)i + (size_t)s * (np - 2)] = (col[i + 2] - 2 * col[i + 1] + col[i]) / (taustep * taustep);
    }
    interp_t.assign(tpert.begin(), tpert.begin() + (np - 2));
}

// :640-720: is the solution smooth enough in this interval to raise the degree instead of dividing?
bool LiuRefiner::can_increase_N(int iphase, const LiuPhaseInput& ph, int seg, const std::vector<int>& sidx, double ratio_R) const
{
    const int istart = sidx[seg], iend = sidx[seg + 1], rows = iend - istart + 1;
    const int Nall = (int)ph.tau.size() + 1;
    Dense seg_x;
    seg_x.rows = rows;
    seg_x.cols = ph.ns;
    seg_x.a.resize((size_t)rows * ph.ns);
    std::vector<double> tau_seg((size_t)rows);
    for (int i = 0; i < rows; ++i) {
        tau_seg[i] = (istart + i < (int)ph.tau.size()) ? ph.tau[istart + i] : 1.0;
        for (int s = 0; s < ph.ns; ++s) seg_x.a[(size_t)i + (size_t)s * rows] = ph.state[(size_t)(istart + i) + (size_t)s * Nall];
    }
    std::vector<double> interp_t;
    Dense d2;
    second_derivative(tau_seg, seg_x, interp_t, d2);
    std::vector<double> Pij(ph.ns), tmax(ph.ns), absd((size_t)d2.rows);
    for (int s = 0; s < ph.ns; ++s) {
        for (int i = 0; i < d2.rows; ++i) absd[i] = std::fabs(d2.at(i, s));
        const int im = argmax(absd.data(), d2.rows);
        Pij[s] = absd[im];
        tmax[s] = interp_t[im];
    }
    const double mintime = *std::min_element(tmax.begin(), tmax.end()), maxtime = *std::max_element(tmax.begin(), tmax.end());
    // the mesh that was current when the previous refinement ran, and what it stored (:675-707)
    const MeshInfo& cur = mesh_history_.back()[iphase];
    auto last_where = [&](auto pred) { int r = -1; for (int i = 0; i < (int)cur.mesh.size(); ++i) if (pred(cur.mesh[i])) r = i; return r; };
    auto first_where = [&](auto pred) { for (int i = 0; i < (int)cur.mesh.size(); ++i) if (pred(cur.mesh[i])) return i; return -1; };
    int itmin, itmax;
    if (mintime == maxtime) {
        if (mintime == tau_seg[0]) {
            itmin = last_where([&](double m) { return m <= mintime; });
            itmax = first_where([&](double m) { return m > maxtime; });
        } else if (maxtime == tau_seg[rows - 1]) {
            itmin = last_where([&](double m) { return m < mintime; });
            itmax = first_where([&](double m) { return m >= maxtime; });
        } else {
            itmin = last_where([&](double m) { return m < mintime; });
            itmax = first_where([&](double m) { return m > maxtime; });
        }
    } else {
        itmin = last_where([&](double m) { return m <= mintime; });
        itmax = first_where([&](double m) { return m >= maxtime; });
    }
    const std::vector<double>& time_b = mesh_points_history_.back()[iphase];
    const Dense& state_b = state_history_.back()[iphase];
    if (itmin < 0 || itmax < itmin || itmax >= (int)time_b.size() || itmax >= state_b.rows) return false; // the reference would index out of range
    const int rb = itmax - itmin + 1;
    std::vector<double> seg_t_b(time_b.begin() + itmin, time_b.begin() + itmax + 1);
    Dense seg_x_b;
    seg_x_b.rows = rb;
    seg_x_b.cols = ph.ns;
    seg_x_b.a.resize((size_t)rb * ph.ns);
    for (int i = 0; i < rb; ++i)
        for (int s = 0; s < ph.ns; ++s) seg_x_b.a[(size_t)i + (size_t)s * rb] = state_b.at(itmin + i, s); // state rows by mesh-point index (:706-707)
    std::vector<double> interp_t_b;
    Dense d2b;
    second_derivative(seg_t_b, seg_x_b, interp_t_b, d2b);
    std::vector<double> R(ph.ns);
    for (int s = 0; s < ph.ns; ++s) {
        for (int i = 0; i < d2b.rows; ++i) absd[i] = std::fabs(d2b.at(i, s));
        R[s] = Pij[s] / absd[argmax(absd.data(), d2b.rows)];
    }
    const bool need_dividing = R[argmax(R.data(), ph.ns)] > ratio_R;
    return !need_dividing;
}

// :12-260
bool LiuRefiner::refine(const std::vector<LiuPhaseInput>& in, double tol, int Nmax, double ratio_R, std::vector<std::vector<double>>& mesh_out,
                        std::vector<std::vector<int>>& nodes_out)
{
    const int P = (int)in.size();
    enum Tag { NOT_SATISFIED, SATISFIED, REDUCED, MERGED };
    struct Piece { std::vector<double> mesh; std::vector<int> nodes; };
    bool no_more = true;
    mesh_out.assign(P, std::vector<double>());
    nodes_out.assign(P, std::vector<int>());
    if (mesh_index_ == 0) { // the user's first mesh, no error estimates yet (:23-35)
        std::vector<MeshInfo> first(P);
        for (int p = 0; p < P; ++p) {
            first[p].mesh = in[p].mesh;
            first[p].nodes = in[p].nodes;
            first[p].e_k.assign(in[p].nodes.size(), 0.0);
        }
        mesh_history_.push_back(first);
    }
    std::vector<MeshInfo>& current = mesh_history_.back(); // its e_k is filled below and read by the NEXT call
    std::vector<MeshInfo> produced(P);
    std::vector<Dense> phase_state(P);
    std::vector<std::vector<double>> phase_points(P);
    for (int p = 0; p < P; ++p) {
        const LiuPhaseInput& ph = in[p];
        const int K = (int)ph.nodes.size(), Nall = (int)ph.tau.size() + 1;
        if ((int)current[p].nodes.size() != K) throw std::runtime_error("hp-Liu: the mesh is not the one the previous refinement returned (lpb_refine_reset starts over)");
        std::vector<int> eidx(K + 1, 0), sidx(K + 1, 0); // rows of the error matrix / of the state per interval
        for (int k = 0; k < K; ++k) { eidx[k + 1] = eidx[k] + ph.nodes[k] + 1; sidx[k + 1] = sidx[k] + ph.nodes[k]; }
        const int erows = eidx[K] + 1;
        phase_state[p].rows = Nall;
        phase_state[p].cols = ph.ns;
        phase_state[p].a = ph.state;
        std::vector<double> betai(ph.ns); // 1 + max over the phase of each state (:86)
        for (int s = 0; s < ph.ns; ++s) betai[s] = 1 + ph.state[(size_t)argmax(&ph.state[(size_t)s * Nall], Nall) + (size_t)s * Nall];
        std::vector<Piece> pieces(K);
        std::vector<int> tags(K);
        for (int k = 0; k < K; ++k) {
            // largest relative error of the interval, rows eidx[k] .. eidx[k+1] inclusive (:66-70)
            double emax = -std::numeric_limits<double>::infinity();
            for (int s = 0; s < ph.ns; ++s) {
                const double* col = &ph.rel[(size_t)s * erows];
                const double cm = col[eidx[k] + argmax(col + eidx[k], eidx[k + 1] - eidx[k] + 1)];
                if (emax < cm) emax = cm;
            }
            current[p].e_k[k] = emax;
            const double m0 = ph.mesh[k], mf = ph.mesh[k + 1];
            Piece& pc = pieces[k];
            pc.mesh = {m0, mf};
            if (emax <= tol) { // :73-103
                const int need = reducing_N(ph, k, sidx, betai, tol);
                pc.nodes = {need};
                if (need == ph.nodes[k]) tags[k] = SATISFIED;
                else { tags[k] = REDUCED; no_more = false; }
            } else { // :105-157
                const int Ncur = ph.nodes[k];
                if (mesh_index_ == 0) {
                    pc.nodes = {Ncur + 3}; // second mesh: three more collocation points
                } else {
                    bool divide = !can_increase_N(p, ph, k, sidx, ratio_R);
                    double q = 0.0;
                    const bool have_q = growth_exponent(p, ph, k, emax, q);
                    if (!divide) {
                        long long need = 0;
                        if (have_q && to_count(std::ceil(Ncur * std::pow(emax / tol, 1.0 / (q - 5.0 / 2.0))), need) && need <= Nmax && need >= 1) pc.nodes = {(int)need};
                        else divide = true;
                    }
                    if (divide) {
                        long long H = 0, Hmax = 0, S = 2;
                        if (have_q && to_count(std::ceil(std::pow(emax / tol, 1 / q)), H) && to_count(std::ceil(std::log(emax / tol) / std::log((double)Ncur)), Hmax))
                            S = std::max(std::min(H, Hmax), 2LL);
                        pc.nodes.assign((size_t)S, Ncur);
                        pc.mesh = linspace(m0, mf, (int)S + 1);
                    }
                }
                tags[k] = NOT_SATISFIED;
                no_more = false;
            }
        }
        if (no_more) {
            // the reference leaves such a phase without a new mesh (:162); its problem keeps the current one
            mesh_out[p] = ph.mesh;
            nodes_out[p] = ph.nodes;
            produced[p].mesh = ph.mesh;
            produced[p].nodes = ph.nodes;
            produced[p].e_k.assign(ph.nodes.size(), 0.0);
            phase_points[p] = ph.mesh;
            continue;
        }
        // merge neighbours that both meet the tolerance and ask for the same degree (:172-222; the verdict of
        // Merging_mesh is not used by the reference)
        int r = 0;
        for (int k = 0; k < K; ++k) {
            if (k > 0 && tags[r] != NOT_SATISFIED && tags[r - 1] != NOT_SATISFIED && pieces[r].nodes[0] == pieces[r - 1].nodes[0]) {
                no_more = false;
                pieces[r - 1].mesh[1] = pieces[r].mesh[1];
                pieces.erase(pieces.begin() + r);
                tags.erase(tags.begin() + r);
                tags[r - 1] = MERGED;
            } else {
                ++r;
            }
        }
        std::vector<double>& nm = mesh_out[p];
        std::vector<int>& nn = nodes_out[p];
        nm.push_back(-1.0);
        for (const Piece& pc : pieces) {
            nm.insert(nm.end(), pc.mesh.begin() + 1, pc.mesh.end());
            nn.insert(nn.end(), pc.nodes.begin(), pc.nodes.end());
        }
        produced[p].mesh = nm;
        produced[p].nodes = nn;
        produced[p].e_k.assign(nn.size(), 0.0);
        phase_points[p] = nm;
    }
    mesh_history_.push_back(produced);
    state_history_.push_back(phase_state);
    mesh_points_history_.push_back(phase_points);
    ++mesh_index_;
    return no_more;
}

} // namespace lpb
