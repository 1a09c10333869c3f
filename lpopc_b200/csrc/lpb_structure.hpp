// lpb_structure.hpp -- closed-form rank -> (row, col) maps of the IPOPT triplet arrays.
//
// The reference builds its sparsity patterns by sequential appends
// (NLPWrapper::GetPhaseSparsity LpNLPWrapper.cpp:1106-1312, GetWholeSparsity :1314-1548,
// GetConsSparsity :1550-1578; LpHessianCalculator::GetPhaseHessianSparsity LpHessian.cpp:601-876,
// GetLinkHessianSparsity :2369-2508, GetHessianSparsity :2510-2599).  Every segment of those
// arrays is a run of N-long blocks or a short fixed list, so entry number e of a segment has
// a closed form; the GPU structure kernels (lpb_api.cu) evaluate it with one thread per entry
// and write int32 iRow/jCol with fully coalesced stores.  The functions are host+device so
// the CPU test-suite can check them against the oracle without a GPU (tests/host_harness.cpp).
//
// All indices 0-based (TNLP::C_STYLE, LpopcIpopt.cpp:22).  nq = 0 (quirk Q3 fenced).
#pragma once

#if defined(__CUDACC__)
#define LPB_SHD __host__ __device__ __forceinline__
#else
#define LPB_SHD inline
#endif

namespace lpb {

constexpr int kMaxPhases = 12;
constexpr int kMaxLinks = 12;

struct HessEntry { int a, b, da, db; }; // endpoint pair + the two perturbations of its denominator (quirk Q10)
// row a of the xx/ux/uu block list of the Hessian I-part: bit b of mask = block (a, b) present; present blocks of a
// row are consecutive, so block index(a, b) = blk0 + popcount(mask below bit b)
struct HessRow { unsigned long long mask; int blk0; int pad; };

// integer description of one phase, enough for every index formula
struct PhaseShape {
    int N;          // LGR nodes
    int ne;         // events
    int var0, con0; // c_p, r_p
    int ndoff;      // compacted Doffdiag entries
    int nblkH;      // N-long blocks in the xx/ux/uu part of the Hessian I-part
};

struct LinkShape {
    int left, right; // 0-based phases
    int nl;
    int con0; // first constraint row of the pair
};

// position of per-node variable v (state v < ns, control v - ns) inside the phase's block of x
LPB_SHD int var_pos(int v, int ns, int N) { return v < ns ? v * (N + 1) : ns * (N + 1) + (v - ns) * N; }
LPB_SHD int t0_pos(int ns, int nc, int N) { return ns * (N + 1) + nc * N; }
// position of endpoint variable e: [0,ns) x0, [ns,2ns) xf, 2ns t0, 2ns+1 tf
LPB_SHD int end_pos(int e, int ns, int nc, int N)
{
    if (e < ns) return e * (N + 1);
    if (e < 2 * ns) return (e - ns) * (N + 1) + N;
    return t0_pos(ns, nc, N) + (e - 2 * ns);
}

// ---- Jacobian -----------------------------------------------------------------------------
// NL entries of one phase (GetPhaseSparsity with the dependency mask forced dense, quirk Q1):
// (ns+np) row blocks x (ns+nc+2) column blocks of N entries, then ne event rows of 2ns+2.
LPB_SHD long long jac_nl_count(int ns, int nc, int np, int N, int ne)
{
    return (long long)(ns + np) * (ns + nc + 2) * N + (long long)ne * (2 * ns + 2);
}
LPB_SHD void jac_nl_entry(const PhaseShape& ph, int ns, int nc, int np, long long e, int* row, int* col)
{
    const int N = ph.N, nblk = ns + nc + 2;
    const long long nodepart = (long long)(ns + np) * nblk * N;
    if (e < nodepart) {
        const int blk = (int)(e / N), k = (int)(e - (long long)blk * N);
        const int i = blk / nblk, cc = blk - i * nblk;
        *row = ph.con0 + i * N + k;                                    // :1152,:1216 rowstart = i*sumnodes
        if (cc < ns + nc) *col = ph.var0 + var_pos(cc, ns, N) + k;     // :1159-1191
        else *col = ph.var0 + t0_pos(ns, nc, N) + (cc - ns - nc);      // :1196-1204 fill(colshift)
    } else {
        const long long ee = e - nodepart;
        const int q = (int)(ee / (2 * ns + 2)), s = (int)(ee - (long long)q * (2 * ns + 2));
        *row = ph.con0 + (ns + np) * N + q;                            // :1278
        if (s < 2 * ns) *col = ph.var0 + (s >> 1) * (N + 1) + ((s & 1) ? N : 0); // x0_j, xf_j interleaved :1283-1292
        else *col = ph.var0 + t0_pos(ns, nc, N) + (s - 2 * ns);        // :1297-1305
    }
}
// NL entries of one link pair (GetWholeSparsity :1431-1547): [xf_left | x0_right] columns,
// column-major over (jcol, irow).
LPB_SHD void jac_link_entry(const LinkShape& lk, const PhaseShape& pl, const PhaseShape& pr, int ns, int idx, int* row, int* col)
{
    const int jcol = idx / lk.nl, irow = idx - jcol * lk.nl;
    *row = lk.con0 + irow;
    if (jcol < ns) *col = (jcol + 1) * pl.N + jcol + pl.var0; // :1487
    else *col = (jcol - ns) * (pr.N + 1) + pr.var0;           // :1521
}
// L entries (Find(AlinearMatrix), LpBoundsChecker.cpp:288-339): two per linear row
LPB_SHD void jac_lin_entry(const PhaseShape* phs, const LinkShape* lks, int P, int ns, int nc, int lin_con0, int idx,
                           int* row, int* col, double* val)
{
    const int r = idx >> 1, second = idx & 1;
    *row = lin_con0 + r;
    *val = second ? 1.0 : -1.0;
    if (r < P) {
        *col = phs[r].var0 + t0_pos(ns, nc, phs[r].N) + second; // (p, t0, -1), (p, tf, +1)
    } else {
        const LinkShape& lk = lks[r - P];
        if (!second) *col = phs[lk.left].var0 + t0_pos(ns, nc, phs[lk.left].N) + 1; // tf of the left phase, -1
        else *col = phs[lk.right].var0 + t0_pos(ns, nc, phs[lk.right].N);         // t0 of the right phase, +1
    }
}
// C entries of one phase (:1162-1166): Doffdiag triplet (a, b) repeated per state
LPB_SHD void jac_const_entry(const PhaseShape& ph, const int* doff_a, const int* doff_b, long long idx, int* row, int* col)
{
    const int i = (int)(idx / ph.ndoff), e = (int)(idx - (long long)i * ph.ndoff);
    *row = ph.con0 + i * ph.N + doff_a[e];
    *col = ph.var0 + i * (ph.N + 1) + doff_b[e];
}

// ---- Hessian ------------------------------------------------------------------------------
// I-part of one phase (GetPhaseHessianSparsity :662-872): nblkH blocks for the (a >= b) pairs
// present in the dependency product, t0 row (ns+nc blocks + 1), tf row (ns+nc blocks + 2).
LPB_SHD long long hess_i_count(int ns, int nc, int N, int nblkH) { return (long long)N * nblkH + 2LL * (ns + nc) * N + 3; }
LPB_SHD void hess_i_entry(const PhaseShape& ph, int ns, int nc, const int* pair_a, const int* pair_b, long long e, int* row, int* col)
{
    const int N = ph.N, NV = ns + nc, R0 = t0_pos(ns, nc, N);
    const long long nb = (long long)ph.nblkH * N;
    if (e < nb) {
        const int blk = (int)(e / N), k = (int)(e - (long long)blk * N);
        *row = ph.var0 + var_pos(pair_a[blk], ns, N) + k;
        *col = ph.var0 + var_pos(pair_b[blk], ns, N) + k;
        return;
    }
    long long r = e - nb;
    if (r < (long long)NV * N) { // t0 row against states/controls
        const int b = (int)(r / N), k = (int)(r - (long long)b * N);
        *row = ph.var0 + R0; *col = ph.var0 + var_pos(b, ns, N) + k;
        return;
    }
    r -= (long long)NV * N;
    if (r == 0) { *row = ph.var0 + R0; *col = ph.var0 + R0; return; }
    r -= 1;
    if (r < (long long)NV * N) { // tf row
        const int b = (int)(r / N), k = (int)(r - (long long)b * N);
        *row = ph.var0 + R0 + 1; *col = ph.var0 + var_pos(b, ns, N) + k;
        return;
    }
    r -= (long long)NV * N;
    *row = ph.var0 + R0 + 1;
    *col = ph.var0 + R0 + (int)r; // (tf,t0) then (tf,tf)
}
// E-part entry count and the (a, b) endpoint pair of entry idx in output order (:675-808)
LPB_SHD int hess_e_count(int ns) { return 2 * ns * ns + 5 * ns + 3; }
LPB_SHD void hess_e_entry(const PhaseShape& ph, int ns, int nc, int a, int b, int* row, int* col)
{
    *row = ph.var0 + end_pos(a, ns, nc, ph.N);
    *col = ph.var0 + end_pos(b, ns, nc, ph.N);
}
// link part (GetLinkHessianSparsity :2403-2507).  Entry order: for i < ns, j <= i: (xfL_i, xfL_j);
// for i < ns: for j < ns: (x0R_i, xfL_j); for j <= i: (x0R_i, "x0R_j" with the LEFT node count, quirk Q7).
LPB_SHD int hess_link_count(int ns) { return ns * (ns + 1) / 2 + ns * ns + ns * (ns + 1) / 2; }
LPB_SHD void hess_link_entry(const PhaseShape& pl, const PhaseShape& pr, int ns, int idx, int* row, int* col)
{
    const int tri = ns * (ns + 1) / 2;
    if (idx < tri) {
        int i = 0;
        while ((i + 1) * (i + 2) / 2 <= idx) ++i;
        const int j = idx - i * (i + 1) / 2;
        *row = pl.var0 + (pl.N + 1) * (i + 1) - 1;
        *col = pl.var0 + (pl.N + 1) * (j + 1) - 1;
        return;
    }
    // rows i = 0..ns-1, each with ns + (i+1) entries
    int r = idx - tri, i = 0;
    while (r >= ns + i + 1) { r -= ns + i + 1; ++i; }
    *row = pr.var0 + (pr.N + 1) * i;
    if (r < ns) *col = pl.var0 + (pl.N + 1) * (r + 1) - 1;
    else *col = pr.var0 + (pl.N + 1) * (r - ns); // quirk Q7: nnodesLeft
}


// ---- ordered compaction of the composite Radau matrices ---------------------------------------
// candidate idx of the Doffdiag (mode 0, dense-entry index space) or Diag (mode 1, node index
// space) COO in the reference's storage order; returns whether Find() keeps it.
struct CompactDev {
    int mode;            // 0 = Doffdiag over the dense-entry index space, 1 = Diag over nodes
    long long total;     // candidates
    int K, N;
    const double* dblocks;
    const long long* int_d0; // [K]
    const int* int_row0;     // [K]
    const int* int_n;        // [K]
    const int* node_interval;// [N]
};

LPB_SHD bool compact_candidate(const CompactDev& c, long long idx, int* a, int* b, double* v)
{
    if (c.mode == 1) {
        const int k = (int)idx;
        const int I = c.node_interval[k];
        const int n = c.int_n[I], r = k - c.int_row0[I];
        *a = k; *b = k;
        *v = c.dblocks[c.int_d0[I] + (long long)r * n + r];
        return *v != 0.0;
    }
    // binary search of the interval owning dense entry idx
    int lo = 0, hi = c.K - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (c.int_d0[mid] <= idx) lo = mid; else hi = mid - 1;
    }
    const int n = c.int_n[lo];
    const long long loc = idx - c.int_d0[lo];
    const int j = (int)(loc / n), i = (int)(loc - (long long)j * n);
    *a = c.int_row0[lo] + i;
    *b = c.int_row0[lo] + j;
    *v = c.dblocks[idx];
    return i != j && *v != 0.0; // Do = D - Dd is exactly 0 on the block diagonal (RPMGenerator.cpp:125-129)
}


// ---- whole-problem layout -------------------------------------------------------------------
// Offsets of every segment of x, g, the Jacobian triplets [NL phases | NL links | L | C]
// (LpNLPWrapper.cpp:246-252,:1563-1573) and the Hessian triplets [per phase I | E, then links]
// (LpHessian.cpp:2569-2598).  POD: passed by value to the structure kernels.
struct Layout {
    int P, Lp, ns, nc, np;
    int n, m, lin_con0, total_nodes;
    PhaseShape ph[kMaxPhases];
    LinkShape lk[kMaxLinks];
    int node0[kMaxPhases];
    int lam0[kMaxLinks];       // first multiplier row the link Hessian reads (quirk Q6)
    long long nl0[kMaxPhases]; // first NL value of the phase
    long long ev0[kMaxPhases]; // first event-row value of the phase
    long long lkv0[kMaxLinks]; // first NL value of the pair
    long long lin_val0;        // first L value
    long long c0[kMaxPhases];  // first C value of the phase
    long long ctot;            // C values in total
    long long nnz_jac;
    long long hI0[kMaxPhases], hE0[kMaxPhases], hL0[kMaxLinks];
    long long nnz_h;
};

// Inputs: P, Lp, ns, nc, np, ph[].{N, ne, ndoff, nblkH}, lk[].{left, right, nl}.  Returns false when
// an index would overflow IPOPT's 32-bit Index.
inline bool build_layout(Layout& L)
{
    const int ns = L.ns, nc = L.nc, np = L.np;
    const long long lim = 2147483647LL;
    long long var0 = 0, con0 = 0, node0 = 0;
    for (int p = 0; p < L.P; ++p) {
        const int N = L.ph[p].N;
        L.ph[p].var0 = (int)var0; L.ph[p].con0 = (int)con0; L.node0[p] = (int)node0;
        var0 += (long long)ns * (N + 1) + (long long)nc * N + 2; // LpBoundsChecker.cpp:51-116
        con0 += (long long)(ns + np) * N + L.ph[p].ne;           // :61-185
        node0 += N;
        if (var0 > lim || con0 > lim) return false;
    }
    L.n = (int)var0;
    L.total_nodes = (int)node0;
    long long linkrow = con0;
    for (int l = 0; l < L.Lp; ++l) {
        L.lk[l].con0 = (int)linkrow;
        L.lam0[l] = (int)con0; // link_indices = constraint_offset + j + 1, never advanced per pair (LpBoundsChecker.cpp:243)
        linkrow += L.lk[l].nl;
    }
    L.lin_con0 = (int)linkrow;
    if (linkrow + L.P + L.Lp > lim) return false;
    L.m = (int)(linkrow + L.P + L.Lp);
    long long off = 0;
    for (int p = 0; p < L.P; ++p) {
        L.nl0[p] = off;
        L.ev0[p] = off + (long long)(ns + np) * (ns + nc + 2) * L.ph[p].N;
        off += jac_nl_count(ns, nc, np, L.ph[p].N, L.ph[p].ne);
    }
    for (int l = 0; l < L.Lp; ++l) {
        L.lkv0[l] = off;
        off += (long long)L.lk[l].nl * (ns + ns); // right sizes read from the LEFT phase (quirk Q7); ns is shared
    }
    L.lin_val0 = off;
    off += 2LL * (L.P + L.Lp);
    L.ctot = 0;
    for (int p = 0; p < L.P; ++p) {
        L.c0[p] = off;
        off += (long long)ns * L.ph[p].ndoff;
        L.ctot += (long long)ns * L.ph[p].ndoff;
    }
    L.nnz_jac = off;
    long long hoff = 0;
    for (int p = 0; p < L.P; ++p) {
        L.hI0[p] = hoff;
        hoff += hess_i_count(ns, nc, L.ph[p].N, L.ph[p].nblkH);
        L.hE0[p] = hoff;
        hoff += hess_e_count(ns);
    }
    for (int l = 0; l < L.Lp; ++l) {
        L.hL0[l] = hoff;
        hoff += hess_link_count(ns);
    }
    L.nnz_h = hoff;
    return off <= lim && hoff <= lim;
}

// per-phase index tables the closed forms read (device pointers in the kernels)
struct LayoutTables {
    const int* doff_a[kMaxPhases];
    const int* doff_b[kMaxPhases];
    const int* pair_a[kMaxPhases];
    const int* pair_b[kMaxPhases];
    const HessEntry* eent;
};

// entry e of the whole Jacobian / Hessian pattern
LPB_SHD void jac_entry(const Layout& L, const LayoutTables& T, long long e, int* row, int* col)
{
    if (e >= L.c0[0]) { // [C]
        int p = 0;
        while (p + 1 < L.P && e >= L.c0[p + 1]) ++p;
        jac_const_entry(L.ph[p], T.doff_a[p], T.doff_b[p], e - L.c0[p], row, col);
    } else if (e >= L.lin_val0) { // [L]
        double v;
        jac_lin_entry(L.ph, L.lk, L.P, L.ns, L.nc, L.lin_con0, (int)(e - L.lin_val0), row, col, &v);
    } else if (L.Lp > 0 && e >= L.lkv0[0]) { // [NL links]
        int l = 0;
        while (l + 1 < L.Lp && e >= L.lkv0[l + 1]) ++l;
        jac_link_entry(L.lk[l], L.ph[L.lk[l].left], L.ph[L.lk[l].right], L.ns, (int)(e - L.lkv0[l]), row, col);
    } else { // [NL phases]
        int p = 0;
        while (p + 1 < L.P && e >= L.nl0[p + 1]) ++p;
        jac_nl_entry(L.ph[p], L.ns, L.nc, L.np, e - L.nl0[p], row, col);
    }
}
LPB_SHD void hess_entry(const Layout& L, const LayoutTables& T, long long e, int* row, int* col)
{
    if (L.Lp > 0 && e >= L.hL0[0]) {
        int l = 0;
        while (l + 1 < L.Lp && e >= L.hL0[l + 1]) ++l;
        hess_link_entry(L.ph[L.lk[l].left], L.ph[L.lk[l].right], L.ns, (int)(e - L.hL0[l]), row, col);
        return;
    }
    int p = 0;
    while (p + 1 < L.P && e >= L.hI0[p + 1]) ++p;
    if (e >= L.hE0[p]) {
        const HessEntry en = T.eent[e - L.hE0[p]];
        hess_e_entry(L.ph[p], L.ns, L.nc, en.a, en.b, row, col);
    } else {
        hess_i_entry(L.ph[p], L.ns, L.nc, T.pair_a[p], T.pair_b[p], e - L.hI0[p], row, col);
    }
}

// ---- Hessian entry tables and block lists (host) ----------------------------------------------
} // namespace lpb
#include <vector>
namespace lpb {

// E-part in output order (GetPhaseHessian, LpHessian.cpp:409-535); (da, db) carry the two
// perturbations whose product is the stencil denominator (quirk Q10, :1588,:1612).
// Link part in output order (GetLinkHessian, :1086-1145): variable index [0,ns) xf_left, [ns,2ns) x0_right.
inline void build_entry_tables(int ns, std::vector<HessEntry>& eent, std::vector<HessEntry>& lent)
{
    eent.clear();
    const int T0 = 2 * ns, TF = 2 * ns + 1;
    for (int i = 0; i < ns; ++i)
        for (int j = 0; j <= i; ++j) {
            eent.push_back({i, j, i, j});                        // x0.x0
            if (i != j) eent.push_back({i, ns + j, i, ns + i}); // x0.xf: pertx0(i)*pertxf(i)
            eent.push_back({ns + i, j, ns + i, j});             // xf.x0
            eent.push_back({ns + i, ns + j, ns + i, ns + i});   // xf.xf: pertxf(i)*pertxf(i)
        }
    for (int i = 0; i < ns; ++i) { eent.push_back({T0, i, T0, i}); eent.push_back({T0, ns + i, T0, ns + i}); }
    eent.push_back({T0, T0, T0, T0});
    for (int i = 0; i < ns; ++i) { eent.push_back({TF, i, TF, i}); eent.push_back({TF, ns + i, TF, ns + i}); }
    eent.push_back({TF, T0, TF, T0});
    eent.push_back({TF, TF, TF, TF});
    lent.clear();
    for (int i = 0; i < ns; ++i)
        for (int j = 0; j <= i; ++j) lent.push_back({i, j, i, j});
    for (int i = 0; i < ns; ++i) {
        for (int j = 0; j < ns; ++j) lent.push_back({j, ns + i, j, ns + i});
        for (int j = 0; j <= i; ++j) lent.push_back({ns + i, ns + j, ns + i, ns + j});
    }
}

// depH = dep' * dep with unit diagonal (LpHessian.cpp:2532-2536); lower-triangle block list in the
// order of GetPhaseHessianSparsity :662-735.  dep is (ns+np) x (ns+nc) column-major 0/1.
inline void build_hess_blocks(int ns, int nc, int np, const std::vector<int>& dep,
                              std::vector<int>& pair_a, std::vector<int>& pair_b, std::vector<int>& hblk, std::vector<HessRow>* hrow = nullptr)
{
    const int NV = ns + nc, NR = ns + np;
    std::vector<int> depH((size_t)NV * NV, 0);
    for (int i = 0; i < NV; ++i)
        for (int j = 0; j < NV; ++j) {
            int acc = 0;
            for (int r = 0; r < NR; ++r) acc += dep[(size_t)i * NR + r] * dep[(size_t)j * NR + r];
            depH[(size_t)i * NV + j] = acc;
        }
    for (int i = 0; i < NV; ++i) depH[(size_t)i * NV + i] = 1;
    pair_a.clear(); pair_b.clear();
    hblk.assign((size_t)NV * NV, -1);
    auto add = [&](int a, int b) {
        hblk[(size_t)a * NV + b] = (int)pair_a.size();
        pair_a.push_back(a); pair_b.push_back(b);
    };
    for (int i = 0; i < ns; ++i)
        for (int j = 0; j <= i; ++j)
            if (depH[(size_t)i * NV + j]) add(i, j);
    for (int i = 0; i < nc; ++i) {
        for (int j = 0; j < ns; ++j)
            if (depH[(size_t)(ns + i) * NV + j]) add(ns + i, j);
        for (int j = 0; j <= i; ++j)
            if (depH[(size_t)(ns + i) * NV + ns + j]) add(ns + i, ns + j);
    }
    if (hrow) {
        hrow->assign((size_t)NV, HessRow{0ull, 0, 0});
        for (int a = 0; a < NV && a < 64; ++a) {
            HessRow r{0ull, -1, 0};
            for (int b = 0; b <= a; ++b)
                if (hblk[(size_t)a * NV + b] >= 0) {
                    if (r.blk0 < 0) r.blk0 = hblk[(size_t)a * NV + b];
                    r.mask |= 1ull << b;
                }
            if (r.blk0 < 0) r.blk0 = 0;
            (*hrow)[a] = r;
        }
    }
}

} // namespace lpb
