// tests/host_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Runs the product's *host-callable* logic on the CPU so that the `-m "not gpu"` suite can
// check it against the oracle without a GPU: the Radau table builder (lpb_tables.cpp), the
// layout computation and the closed-form rank -> (row, col) maps that the GPU structure
// kernels evaluate (lpb_structure.hpp, same functions, one call per entry instead of one
// thread per entry), and the ordered-compaction predicate.  It evaluates no user function
// and is never loaded by the product.
#include "../lpopc_b200/csrc/lpb_structure.hpp"
#include "../lpopc_b200/csrc/lpb_tables.hpp"
#include "../lpopc_b200/csrc/lpb_refine_liu.hpp"
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

using namespace lpb;

namespace {
struct Harness {
    Layout L;
    LayoutTables T;
    std::vector<PhaseTables> tab;
    std::vector<std::vector<int>> doff_a, doff_b, pair_a, pair_b, hblk;
    std::vector<std::vector<double>> doff_v, ddiag;
    std::vector<HessEntry> eent, lent;
    std::vector<int> jI, jJ, hI, hJ;
    std::string err;
};
} // namespace

extern "C" {

void* lpbt_create(int ns, int nc, int np, int P, const int* K, const double* mesh_concat, const int* nodes_concat,
                  const int* ne, int Lp, const int* left, const int* right, const int* nl, const int* dep_concat,
                  char* errbuf, int errlen)
{
    Harness* h = new Harness();
    try {
        std::memset(&h->L, 0, sizeof h->L);
        std::memset(&h->T, 0, sizeof h->T);
        Layout& L = h->L;
        L.P = P; L.Lp = Lp; L.ns = ns; L.nc = nc; L.np = np;
        h->tab.resize(P); h->doff_a.resize(P); h->doff_b.resize(P); h->doff_v.resize(P); h->ddiag.resize(P);
        h->pair_a.resize(P); h->pair_b.resize(P); h->hblk.resize(P);
        const double* mp = mesh_concat;
        const int* nd = nodes_concat;
        const size_t per = (size_t)(ns + np) * (ns + nc);
        for (int p = 0; p < P; ++p) {
            build_phase_tables(K[p], mp, nd, h->tab[p]);
            mp += K[p] + 1; nd += K[p];
            const PhaseTables& t = h->tab[p];
            CompactDev c;
            c.K = t.K; c.N = t.N; c.dblocks = t.dblocks.data(); c.int_d0 = t.int_d0.data();
            c.int_row0 = t.int_row0.data(); c.int_n = t.int_n.data(); c.node_interval = t.node_interval.data();
            c.mode = 1; c.total = t.N;
            h->ddiag[p].assign(t.N, 0.0);
            size_t pos = 0;
            for (long long i = 0; i < c.total; ++i) {
                int a, b; double v;
                if (compact_candidate(c, i, &a, &b, &v)) h->ddiag[p][pos++] = v;
            }
            c.mode = 0; c.total = (long long)t.dblocks.size();
            for (long long i = 0; i < c.total; ++i) {
                int a, b; double v;
                if (compact_candidate(c, i, &a, &b, &v)) { h->doff_a[p].push_back(a); h->doff_b[p].push_back(b); h->doff_v[p].push_back(v); }
            }
            std::vector<int> dep(per, 1);
            if (dep_concat) dep.assign(dep_concat + p * per, dep_concat + (p + 1) * per);
            build_hess_blocks(ns, nc, np, dep, h->pair_a[p], h->pair_b[p], h->hblk[p]);
            L.ph[p].N = t.N; L.ph[p].ne = ne[p]; L.ph[p].ndoff = (int)h->doff_a[p].size(); L.ph[p].nblkH = (int)h->pair_a[p].size();
            h->T.doff_a[p] = h->doff_a[p].data(); h->T.doff_b[p] = h->doff_b[p].data();
            h->T.pair_a[p] = h->pair_a[p].data(); h->T.pair_b[p] = h->pair_b[p].data();
        }
        for (int l = 0; l < Lp; ++l) { L.lk[l].left = left[l]; L.lk[l].right = right[l]; L.lk[l].nl = nl[l]; }
        build_entry_tables(ns, h->eent, h->lent);
        h->T.eent = h->eent.data();
        if (!build_layout(L)) throw std::runtime_error("layout overflow");
        h->jI.resize(L.nnz_jac); h->jJ.resize(L.nnz_jac); h->hI.resize(L.nnz_h); h->hJ.resize(L.nnz_h);
        for (long long e = 0; e < L.nnz_jac; ++e) jac_entry(L, h->T, e, &h->jI[e], &h->jJ[e]);
        for (long long e = 0; e < L.nnz_h; ++e) hess_entry(L, h->T, e, &h->hI[e], &h->hJ[e]);
        return h;
    } catch (const std::exception& e) {
        if (errbuf && errlen > 0) { std::strncpy(errbuf, e.what(), errlen - 1); errbuf[errlen - 1] = 0; }
        delete h;
        return nullptr;
    }
}

void lpbt_destroy(void* p) { delete (Harness*)p; }

void lpbt_info(void* p, int* n, int* m, long long* nnz_jac, long long* nnz_h)
{
    Harness* h = (Harness*)p;
    *n = h->L.n; *m = h->L.m; *nnz_jac = h->L.nnz_jac; *nnz_h = h->L.nnz_h;
}

void lpbt_structure(void* p, int* jI, int* jJ, int* hI, int* hJ)
{
    Harness* h = (Harness*)p;
    std::memcpy(jI, h->jI.data(), h->jI.size() * sizeof(int));
    std::memcpy(jJ, h->jJ.data(), h->jJ.size() * sizeof(int));
    std::memcpy(hI, h->hI.data(), h->hI.size() * sizeof(int));
    std::memcpy(hJ, h->hJ.data(), h->hJ.size() * sizeof(int));
}

// tables of one phase: tau[N], w[N], ddiag[N]; returns ndoff
int lpbt_tables(void* p, int phase, double* tau, double* w, double* ddiag)
{
    Harness* h = (Harness*)p;
    const PhaseTables& t = h->tab[phase];
    std::memcpy(tau, t.tau.data(), t.N * sizeof(double));
    std::memcpy(w, t.w.data(), t.N * sizeof(double));
    std::memcpy(ddiag, h->ddiag[phase].data(), t.N * sizeof(double));
    return (int)h->doff_a[phase].size();
}

void lpbt_doff(void* p, int phase, int* a, int* b, double* v)
{
    Harness* h = (Harness*)p;
    const size_t n = h->doff_a[phase].size();
    std::memcpy(a, h->doff_a[phase].data(), n * sizeof(int));
    std::memcpy(b, h->doff_b[phase].data(), n * sizeof(int));
    std::memcpy(v, h->doff_v[phase].data(), n * sizeof(double));
}

// dense D product order check helper: full composite D as COO in reference order (zeros dropped)
int lpbt_dfull(void* p, int phase, int* a, int* b, double* v)
{
    Harness* h = (Harness*)p;
    const PhaseTables& t = h->tab[phase];
    int cnt = 0;
    for (int I = 0; I < t.K; ++I) {
        const int n = t.int_n[I];
        for (int j = 0; j <= n; ++j)
            for (int i = 0; i < n; ++i) {
                const double d = t.dblocks[t.int_d0[I] + (long long)j * n + i];
                if (d != 0.0) {
                    if (a) { a[cnt] = t.int_row0[I] + i; b[cnt] = t.int_row0[I] + j; v[cnt] = d; }
                    ++cnt;
                }
            }
    }
    return cnt;
}

} // extern "C"

// ---- hp-Liu refinement decision (lpb_refine_liu.cpp) on given error estimates: single phase ----
extern "C" {
void* lpbt_liu_create() { return new lpb::LiuRefiner(); }
void lpbt_liu_destroy(void* p) { delete static_cast<lpb::LiuRefiner*>(p); }
int lpbt_liu_refine(void* p, int ns, int K, const double* mesh, const int* nodes, const double* rel, int rel_rows, const double* state,
                    double tol, int Nmax, double R, int* done, int* K_out, double* mesh_out, int* nodes_out, int cap)
{
    try {
        lpb::PhaseTables tab;
        lpb::build_phase_tables(K, mesh, nodes, tab);
        std::vector<lpb::LiuPhaseInput> in(1);
        in[0].ns = ns;
        in[0].mesh.assign(mesh, mesh + K + 1);
        in[0].nodes.assign(nodes, nodes + K);
        in[0].tau = tab.tau;
        in[0].rel.assign(rel, rel + (size_t)rel_rows * ns);
        in[0].state.assign(state, state + (size_t)(tab.N + 1) * ns);
        std::vector<std::vector<double>> mo;
        std::vector<std::vector<int>> no;
        *done = static_cast<lpb::LiuRefiner*>(p)->refine(in, tol, Nmax, R, mo, no) ? 1 : 0;
        if ((int)no[0].size() > cap) return -2;
        *K_out = (int)no[0].size();
        std::memcpy(mesh_out, mo[0].data(), mo[0].size() * sizeof(double));
        std::memcpy(nodes_out, no[0].data(), no[0].size() * sizeof(int));
        return 0;
    } catch (const std::exception&) {
        return -1;
    }
}
}
