"""GPU parity: the CUDA path through the C ABI vs the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): integer triplets bit-exact; fp64 function, Jacobian AND Hessian values within
1e-12 TRUE relative of the restatement (tests/parity.py: |a-b| <= 1e-12 |b| entry by entry, with the documented
floor for entries more than ten orders below their segment's largest), and the bit-equal fraction asserted per
quantity: user functions evaluate bit-identically on both sides, so Jacobian values are expected to be identical
bit for bit and only the quadrature sums of f / grad f and the E-part of the Hessian may differ in the last ulp
(measured table: profiles/r02_parity_report.txt).
"""
import numpy as np
import pytest

import cases
import parity
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu

RTOL = 1e-12


def rel_err(a, b):
    """TRUE relative error max |a-b|/|b| (tests/parity.py); inf when an entry below the documented floor of its
    array is off by more than the floor's share of 1e-12."""
    r = parity.report(a, b)
    if r["max_abs_lo"] > max(parity.ATOL_LO, RTOL * parity.NOISE_FLOOR_REL) * max(r["scale"], 1e-300):
        return float("inf")
    return r["max_rel"]


@pytest.mark.parametrize("name", cases.CASES + ["launch/u5x4", "hypersensitive/u40x3", "orbit_raising/u200x10"])
def test_parity_all_callbacks(nlp_mod, name):
    op = cases.build(name)
    o = Oracle(op)
    g = nlp_mod.TranscribedNLP(op)
    # sizes + integer triplets: bit-exact
    assert g.get_nlp_info() == (o.n, o.m, o.nnz_jac, o.nnz_h)
    gi, gj = g.eval_jac_g(values=False)
    oi, oj = o.jac_structure()
    oi_jac = oi
    assert np.array_equal(gi, oi) and np.array_equal(gj, oj)
    gi, gj = g.eval_h(values=False)
    oi, oj = o.h_structure()
    assert np.array_equal(gi, oi) and np.array_equal(gj, oj)
    for a, b in zip(g.get_bounds_info(), o.bounds()):
        assert np.array_equal(a, b)
    guess, x, sigma, lam = cases.inputs(op, o, 7)
    for xv in (guess, x):
        assert abs(g.eval_f(xv) - o.eval_f(xv)) <= RTOL * abs(o.eval_f(xv))
        parity.assert_parity(g.eval_grad_f(xv), o.eval_grad_f(xv), rtol=RTOL, min_bit_equal=0.9, what="grad f")
        parity.assert_parity(g.eval_g(xv), o.eval_g(xv), rtol=RTOL, min_bit_equal=0.999, what="g")
        jv_g, jv_o = g.eval_jac_g(xv), o.eval_jac_g(xv)
        for sname, idx in parity.jac_segments(op, (o.n, o.m, o.nnz_jac, o.nnz_h), oi_jac).items():
            parity.assert_parity(jv_g[idx], jv_o[idx], rtol=RTOL, min_bit_equal=0.999, what="jac " + sname)
        gg, gv = g.eval_g_jac(xv)
        assert np.array_equal(gg, g.eval_g(xv)) and np.array_equal(gv, jv_g)
    # Hessian: same stencil arithmetic on bit-identical function values
    parity.assert_parity(g.eval_h(x, sigma, lam), o.eval_h(x, sigma, lam), rtol=RTOL, min_bit_equal=0.99, what="hessian")


def test_fd_jacobian_is_bit_exact_for_polynomial_functor(nlp_mod):
    """Bryson-Denham uses + * only: host and device must agree to the last bit."""
    op = cases.build("bryson_denham/ragged")
    o, g = Oracle(op), nlp_mod.TranscribedNLP(op)
    _, x, sigma, lam = cases.inputs(op, o, 3)
    assert np.array_equal(g.eval_g(x), o.eval_g(x))
    assert np.array_equal(g.eval_jac_g(x), o.eval_jac_g(x))
    assert np.array_equal(g.eval_grad_f(x), o.eval_grad_f(x))
    assert np.array_equal(g.eval_h(x, sigma, lam), o.eval_h(x, sigma, lam))


def test_dependency_probe_and_sparse_hessian(nlp_mod):
    op = cases.build("launch")
    o, g = Oracle(op), nlp_mod.TranscribedNLP(op)
    guess, x, sigma, lam = cases.inputs(op, o, 5)
    dep_o, dep_g = o.probe_dependencies(guess), g.probe_dependencies(guess)
    assert np.array_equal(dep_o, dep_g)
    assert g.get_nlp_info() == (o.n, o.m, o.nnz_jac, o.nnz_h)
    gi, gj = g.eval_h(values=False)
    oi, oj = o.h_structure()
    assert np.array_equal(gi, oi) and np.array_equal(gj, oj)
    hv_g, hv_o = g.eval_h(x, sigma, lam), o.eval_h(x, sigma, lam)
    assert rel_err(hv_g, hv_o) <= RTOL


def test_mesh_refresh_rebuilds_index_maps(nlp_mod):
    """Adaptive refinement = new n, m, nnz, new index maps (north_star item 3)."""
    op = cases.build("bryson_denham")
    o, g = Oracle(op), nlp_mod.TranscribedNLP(op)
    for seed, K in ((1, 2), (2, 5), (3, 9)):
        cases.ragged_mesh(op.phases[0], seed, K, 3, 12)
        o.set_mesh(0, op.phases[0].meshpoints, op.phases[0].nodesperinterval); o.refresh()
        g.set_mesh(0, op.phases[0].meshpoints, op.phases[0].nodesperinterval); g.refresh()
        assert g.get_nlp_info() == (o.n, o.m, o.nnz_jac, o.nnz_h)
        for a, b in zip(g.eval_jac_g(values=False), o.jac_structure()):
            assert np.array_equal(a, b)
        for a, b in zip(g.eval_h(values=False), o.h_structure()):
            assert np.array_equal(a, b)
        _, x, _, _ = cases.inputs(op, o, seed)
        assert rel_err(g.eval_jac_g(x), o.eval_jac_g(x)) <= RTOL


def test_batched_instances_match_single(nlp_mod):
    op = cases.build("quadrotor")
    o, g = Oracle(op), nlp_mod.TranscribedNLP(op)
    rng = np.random.Generator(np.random.PCG64(5))
    guess = op.guess(cases.lgr_points_of(o))
    nb = 37
    X = guess[None, :] + 0.05 * rng.uniform(-1, 1, (nb, guess.size))
    G, V = g.eval_g_jac_batch(X)
    F = g.eval_f_batch(X)
    GR = g.eval_grad_f_batch(X)
    lam = rng.uniform(-1, 1, (nb, o.m))
    sg = rng.uniform(0.5, 1.5, nb)
    H = g.eval_h_batch(X, sg, lam)
    og, ov = o.eval_g_jac_batch(X, nthreads=4)
    assert rel_err(G, og) <= RTOL and rel_err(V, ov) <= RTOL
    for b in (0, 17, nb - 1):
        assert abs(F[b] - o.eval_f(X[b])) <= RTOL * abs(F[b])
        assert rel_err(GR[b], o.eval_grad_f(X[b])) <= RTOL
        assert rel_err(H[b], o.eval_h(X[b], sg[b], lam[b])) <= RTOL
        assert np.array_equal(G[b], g.eval_g(X[b])) and np.array_equal(V[b], g.eval_jac_g(X[b]))


def test_error_behaviour(nlp_mod):
    from lpopc_b200 import examples
    op = examples.hypersensitive()
    op.phases[0].parameternum = 1
    with pytest.raises(nlp_mod.LpopcError) as e:
        nlp_mod.TranscribedNLP(op)
    assert e.value.code == -2  # LPB_ERR_UNSUPPORTED (quirk Q3)
    op = examples.hypersensitive()
    op.phases[0].set_mesh([-1.0, 0.9], [4])
    with pytest.raises(nlp_mod.LpopcError, match="span -1 to \\+1"):
        nlp_mod.TranscribedNLP(op)
    op = examples.bryson_denham()
    op.functor = "nope"
    with pytest.raises(nlp_mod.LpopcError) as e:
        nlp_mod.TranscribedNLP(op)
    assert e.value.code == -5


@pytest.mark.gpu
def test_fd_division_is_ieee(nlp_mod):
    """The Jacobian kernels divide by the per-colour step h through one shared reciprocal
    (FdDiv, lpb_kernels.cuh); every quotient must equal IEEE d / h bit for bit."""
    assert nlp_mod.selftest_fd_division(1 << 28, seed=12345) == 0


# ---- BASELINE.json configurations at their full sizes ---------------------------------------------
def _seeded_x(g, seed, nb=1):
    rng = np.random.Generator(np.random.PCG64(seed))
    xl, xu, _, _ = g.get_bounds_info()
    lo, hi = np.maximum(xl, -1.0), np.minimum(xu, 1.5)
    hi = np.where(hi > lo, hi, lo + 1.0)
    return rng.uniform(lo, hi, (nb, g.n))


@pytest.mark.gpu
def test_full_size_config2_orbit_raising_2000_nodes(nlp_mod):
    """BASELINE config 2: single phase, 200 x 10 = 2000 LGR nodes (n = 14007)."""
    from lpopc_b200 import examples
    op = examples.orbit_raising(intervals=200, nodes=10)
    o, g = Oracle(op), nlp_mod.TranscribedNLP(op)
    assert g.get_nlp_info() == (o.n, o.m, o.nnz_jac, o.nnz_h) and o.n == 14007
    for a, b in zip(g.eval_jac_g(values=False) + g.eval_h(values=False), o.jac_structure() + o.h_structure()):
        assert np.array_equal(a, b)
    _, x, sigma, lam = cases.inputs(op, o, 3)
    gg, gv = g.eval_g_jac(x)
    assert rel_err(gg, o.eval_g(x)) <= RTOL and rel_err(gv, o.eval_jac_g(x)) <= RTOL
    assert rel_err(g.eval_grad_f(x), o.eval_grad_f(x)) <= RTOL
    assert rel_err(g.eval_h(x, sigma, lam), o.eval_h(x, sigma, lam)) <= RTOL


@pytest.mark.gpu
def test_full_size_config3_launch_4_phases_10k_variables(nlp_mod):
    """BASELINE config 3: launch vehicle ascent, 4 phases x (25 x 10) nodes, 3 x 7 links (n = 10036)."""
    from lpopc_b200 import examples
    op = examples.launch(intervals=25, nodes=10)
    o, g = Oracle(op), nlp_mod.TranscribedNLP(op)
    assert g.get_nlp_info() == (o.n, o.m, o.nnz_jac, o.nnz_h) and (o.n, o.m, o.nnz_jac) == (10036, 8033, 166388)
    for a, b in zip(g.eval_jac_g(values=False) + g.eval_h(values=False), o.jac_structure() + o.h_structure()):
        assert np.array_equal(a, b)
    _, x, sigma, lam = cases.inputs(op, o, 4)
    gg, gv = g.eval_g_jac(x)
    assert rel_err(gg, o.eval_g(x)) <= RTOL and rel_err(gv, o.eval_jac_g(x)) <= RTOL
    assert abs(g.eval_f(x) - o.eval_f(x)) <= RTOL * abs(o.eval_f(x))
    assert rel_err(g.eval_h(x, sigma, lam), o.eval_h(x, sigma, lam)) <= RTOL


@pytest.mark.gpu
def test_full_size_config4_4096_quadrotor_instances(nlp_mod):
    """BASELINE config 4: 4096 instances on the 8 x 8 mesh.  A sample of instances is compared with the
    oracle; every instance of the batch must equal its own single-instance evaluation (bit for bit),
    and the mesh-constant tail of the values must be identical across instances."""
    from lpopc_b200 import examples
    op = examples.quadrotor(intervals=8, nodes=8)
    o, g = Oracle(op), nlp_mod.TranscribedNLP(op)
    assert (g.n, g.m, g.nnz_jac) == (1038, 769, 19970)
    nb = 4096
    X = _seeded_x(g, 5, nb)
    X[:, -2], X[:, -1] = 0.0, 2.0
    G, V = g.eval_g_jac_batch(X)
    sample = [0, 1, 511, 512, 2047, 4095]
    og, ov = o.eval_g_jac_batch(X[sample], nthreads=4)
    assert rel_err(G[sample], og) <= RTOL and rel_err(V[sample], ov) <= RTOL
    for b in sample:
        g1, v1 = g.eval_g_jac(X[b])
        assert np.array_equal(G[b], g1) and np.array_equal(V[b], v1)
    tail = 6146  # 2 linear-row entries + 12 x 512 Doffdiag entries
    assert np.array_equal(V[:, -tail:], np.broadcast_to(V[0, -tail:], (nb, tail)))
    g.set_option("host_fill_const", 0)  # the all-device path must give the same array
    G2, V2 = g.eval_g_jac_batch(X)
    assert np.array_equal(G, G2) and np.array_equal(V, V2)
    F = g.eval_f_batch(X)
    for b in sample[:3]:
        assert abs(F[b] - o.eval_f(X[b])) <= RTOL * abs(F[b])


@pytest.mark.gpu
def test_full_size_config5_synthetic_100k_nodes(nlp_mod):
    """BASELINE config 5: ns=20, nc=6 synthetic dynamics on 10000 x 10 = 100k LGR nodes
    (n = 2.6 M, nnz_jac = 76 M): structure and values against the oracle, in full."""
    from lpopc_b200 import examples
    op = examples.synthetic20(intervals=10000, nodes=10)
    o, g = Oracle(op), nlp_mod.TranscribedNLP(op)
    assert g.get_nlp_info() == (o.n, o.m, o.nnz_jac, o.nnz_h) and (o.n, o.nnz_jac) == (2600022, 76000002)
    for a, b in zip(g.eval_jac_g(values=False), o.jac_structure()):
        assert np.array_equal(a, b)
    for a, b in zip(g.eval_h(values=False), o.h_structure()):
        assert np.array_equal(a, b)
    x = _seeded_x(g, 7)[0]
    x[-2], x[-1] = 0.0, 10.0
    gg, gv = g.eval_g_jac(x)
    assert rel_err(gg, o.eval_g(x)) <= RTOL
    assert rel_err(gv, o.eval_jac_g(x)) <= RTOL
    assert abs(g.eval_f(x) - o.eval_f(x)) <= RTOL * abs(o.eval_f(x))


@pytest.mark.gpu
def test_full_size_config5_hessian_values_on_a_node_subsample(nlp_mod):
    """BASELINE config 5 Hessian VALUES (76 M-nnz problem, 35.6 M-entry Hessian): the oracle needs minutes for the whole
    mesh, so it evaluates a problem whose first 150 mesh intervals are the first 150 of the 10000-interval mesh (the
    rest of [-1, 1] is one wide interval): on those 1500 nodes tau, w, x, u, lambda, t0, tf and sigma are the same in
    both problems, and every per-node entry of the I-part (one N-long block per variable pair) depends on nothing
    else -- the big problem's values at those nodes must equal the small problem's (TRUE relative 1e-12)."""
    from lpopc_b200 import examples
    K, Ks, Nk = 10000, 150, 10
    op = examples.synthetic20(intervals=K, nodes=Nk)
    g = nlp_mod.TranscribedNLP(op)
    ops = examples.synthetic20(intervals=Ks + 1, nodes=Nk)
    big_mesh = np.asarray(op.phases[0].meshpoints, dtype=np.float64)
    ops.phases[0].set_mesh(np.concatenate([big_mesh[:Ks + 1], [1.0]]), [Nk] * (Ks + 1))
    o = Oracle(ops)
    ns, nc = 20, 6
    N, Ns, M = K * Nk, (Ks + 1) * Nk, Ks * Nk
    x = _seeded_x(g, 11)[0]
    x[-2], x[-1] = 0.0, 10.0
    rng = np.random.Generator(np.random.PCG64(12))
    lam = rng.uniform(-1, 1, g.m)
    xs, lams = np.zeros(o.n), np.zeros(o.m)
    for j in range(ns):
        xs[j * (Ns + 1):j * (Ns + 1) + M] = x[j * (N + 1):j * (N + 1) + M]
        lams[j * Ns:j * Ns + M] = lam[j * N:j * N + M]
    for j in range(nc):
        xs[ns * (Ns + 1) + j * Ns:ns * (Ns + 1) + j * Ns + M] = x[ns * (N + 1) + j * N:ns * (N + 1) + j * N + M]
    xs[-2], xs[-1] = x[-2], x[-1]
    hb, hs = g.eval_h(x, 0.75, lam), o.eval_h(xs, 0.75, lams)
    nblk = (ns + nc) * (ns + nc + 1) // 2  # dense pattern: xx / ux / uu blocks, then the t0 and tf row blocks
    nrow = nblk + 2 * (ns + nc)
    big = np.stack([hb[b * N + (1 if b >= nblk + ns + nc else 0):][:M] for b in range(nrow)])
    small = np.stack([hs[b * Ns + (1 if b >= nblk + ns + nc else 0):][:M] for b in range(nrow)])
    r = parity.assert_parity(big, small, rtol=RTOL, min_bit_equal=0.99, what="config 5 Hessian, first 1500 nodes")
    assert r["n"] == nrow * M and r["scale"] > 0


@pytest.mark.parametrize("name", ["synthetic20", "synthetic20/ragged", "synthetic20/u50x10", "synthetic20/u7x9"])
def test_dae_sweep_hook_is_bit_identical(nlp_mod, name):
    """Functor sets with the optional sweep hooks (include/lpb_functor.h): the one-pass per-thread sweep (dae_sweep) and
    the row-parallel sweep (sweep_pre / sweep_row, one warp per function row) must hand the kernel exactly the values
    dae() returns column by column -- Jacobian values and constraints bit-identical to the plain colour loop on the
    device, single problems and batches, and within 1e-12 of the oracle (which has no sweep)."""
    op = cases.build(name)
    o = Oracle(op)
    g = nlp_mod.TranscribedNLP(op)
    _, x, _, _ = cases.inputs(op, o, 21)
    l0 = g.kernel_launches
    g_r, v_r = g.eval_g_jac(x)           # default: row-parallel sweep
    v_only = g.eval_jac_g(x)
    assert g.kernel_launches > l0
    g.set_option("sweep_mode", 1)        # per-thread sweep (whole colour range in one thread)
    g.set_option("colour_split", 1)
    g_s, v_s = g.eval_g_jac(x)
    g.set_option("unroll_colours", 0)    # forces the plain colour loop
    g_p, v_p = g.eval_g_jac(x)
    for v in (v_r, v_only, v_s):
        assert np.array_equal(v.view(np.int64), v_p.view(np.int64))
    for gg in (g_r, g_s):
        assert np.array_equal(gg.view(np.int64), g_p.view(np.int64))
    assert rel_err(v_r, o.eval_jac_g(x)) <= RTOL
    assert rel_err(g_r, o.eval_g(x)) <= RTOL
    # a batch of instances through the row-parallel kernel (CTAs straddle instances)
    g.set_option("unroll_colours", -1)
    g.set_option("sweep_mode", 0)
    rng = np.random.Generator(np.random.PCG64(22))
    X = x[None, :] + 1e-2 * rng.uniform(-1, 1, (5, x.size))
    GB, VB = g.eval_g_jac_batch(X)
    for b in range(5):
        g1, v1 = g.eval_g_jac(X[b])
        assert np.array_equal(GB[b], g1) and np.array_equal(VB[b], v1)
    g.set_option("unroll_colours", 0)
    GP, VP = g.eval_g_jac_batch(X)
    assert np.array_equal(VB.view(np.int64), VP.view(np.int64)) and np.array_equal(GB.view(np.int64), GP.view(np.int64))
    g.close()


@pytest.mark.parametrize("name,nb", [("quadrotor/u8x8", 64), ("cartpole/u8x8", 128), ("cartpole/u4x8", 64), ("brachistochrone/u2x8", 32)])
def test_staged_jacobian_kernel_is_bit_identical(nlp_mod, name, nb):
    """k_cons_jac_staged (shared-memory staged, contiguous write-out; batches of single-phase instances with
    N | 128) against k_cons_jac (per-thread scatter): same values and constraints bit for bit, and 1e-12 vs the oracle."""
    op = cases.build(name)
    o = Oracle(op)
    g = nlp_mod.TranscribedNLP(op)
    _, x, _, _ = cases.inputs(op, o, 33)
    rng = np.random.Generator(np.random.PCG64(34))
    X = x[None, :] + 1e-3 * rng.uniform(-1, 1, (nb, x.size)) * (np.abs(x) + 0.1)
    g.set_option("colour_split", 1)
    g.set_option("stage_values", 0)
    g_p, v_p = g.eval_g_jac_batch(X)
    g.set_option("stage_values", 1)
    l0 = g.kernel_launches
    g_s, v_s = g.eval_g_jac_batch(X)
    assert g.kernel_launches > l0
    assert np.array_equal(v_s.view(np.int64), v_p.view(np.int64))
    assert np.array_equal(g_s.view(np.int64), g_p.view(np.int64))
    assert rel_err(v_s[3], o.eval_jac_g(X[3])) <= RTOL
    assert rel_err(g_s[3], o.eval_g(X[3])) <= RTOL


TILED_CASES = ["hypersensitive", "hypersensitive_analytic", "bryson_denham", "bryson_denham/ragged", "launch", "launch/ragged",
               "orbit_raising", "brachistochrone", "quadrotor", "quadrotor/u8x8", "cartpole"]


@pytest.mark.parametrize("probe", [False, True])
@pytest.mark.parametrize("name", TILED_CASES)
def test_tiled_hessian_kernel_is_bit_identical(nlp_mod, name, probe):
    """k_hess_tiled (compile-time pair bodies generated from the functor set's HESS_DEP table: stencils only on rows
    that read both variables, branch-free quotients) against k_hess_nodes (run-time pair loops over every row, IEEE
    quotients): every Hessian value bit for bit, on the dense pattern and on the pattern the NaN probe leaves, at a
    generic point and at a point where many function values are tiny next to their increments (the rounding terms of
    rows that read only one variable of a pair are then non-zero)."""
    op = cases.build(name)
    o = Oracle(op)
    g = nlp_mod.TranscribedNLP(op)
    guess, x, sigma, lam = cases.inputs(op, o, 41)
    if probe:
        g.probe_dependencies(guess)
    rng = np.random.Generator(np.random.PCG64(42))
    x_small = x.copy()
    # states and controls within 1e-7 of zero (times untouched): f is tiny where it vanishes at the origin
    off = 0
    for ph in op.phases:
        N = int(np.sum(ph.nodesperinterval))
        nvar = ph.statenum * (N + 1) + ph.controlnum * N
        x_small[off:off + nvar] = 1e-7 * rng.uniform(-1, 1, nvar)
        off += nvar + 2 + ph.parameternum
    for xx in (x, x_small):
        g.set_option("unroll_colours", -1)
        h_t = g.eval_h(xx, sigma, lam)
        g.set_option("unroll_colours", 0)  # generic pair loops
        h_g = g.eval_h(xx, sigma, lam)
        assert np.array_equal(h_t.view(np.int64), h_g.view(np.int64)), np.flatnonzero(h_t.view(np.int64) != h_g.view(np.int64))[:10]
    g.close()


def test_nlp2op_end_rows_on_a_long_mesh(nlp_mod):
    """The spline end rows of lpb_nlp2op on a mesh of 12 000 nodes: k_nlp2op_ends starts the forward sweep 512 knots
    before the end (the sweep forgets its start geometrically); adaptive.natural_spline restates the reference's sweep
    over ALL knots (LpGuessChecker.cpp:208-262, pinned to the reference in test_reference_pin.py) -- same values."""
    from lpopc_b200 import adaptive, examples
    op = examples.synthetic20(intervals=3000, nodes=4)
    g = nlp_mod.TranscribedNLP(op)
    n, m = g.get_nlp_info()[:2]
    rng = np.random.Generator(np.random.PCG64(11))
    x = g.initial_guess() + 1e-2 * rng.standard_normal(n)
    lam = rng.standard_normal(m)
    res, _ = g.nlp2op(x, lam)
    tau = g.lgr_points()[0]
    N, ns, nc = tau.size, len(op.phases[0].statemin), len(op.phases[0].controlmin)
    assert N == 12000 and res[0]["control"].shape == (N + 1, nc)
    for j in range(nc):
        u = x[ns * (N + 1) + j * N: ns * (N + 1) + (j + 1) * N]
        full = float(adaptive.natural_spline(tau, u, np.array([1.0]))[0])
        assert abs(res[0]["control"][N, j] - full) <= RTOL * abs(full), j
        assert np.array_equal(res[0]["control"][:N, j], u)


@pytest.mark.parametrize("name", ["launch/u5x4", "two_stage/ragged", "hypersensitive/u40x3"])
def test_device_resident_nlp2op_and_mesh_error_equal_the_host_calls(nlp_mod, name):
    """lpb_nlp2op_dev / lpb_mesh_error_dev (x, multipliers and results stay on the GPU: what a GPU-resident outer loop
    calls between two solves) deliver bit for bit what the host-pointer calls deliver."""
    import torch
    op = cases.build(name)
    g = nlp_mod.TranscribedNLP(op)
    n, m = g.get_nlp_info()[:2]
    rng = np.random.Generator(np.random.PCG64(3))
    x = g.initial_guess() + 1e-3 * rng.standard_normal(n)
    lam = rng.standard_normal(m)
    dev = torch.device("cuda")
    dx, dl = torch.from_numpy(x).to(dev), torch.from_numpy(lam).to(dev)
    # converted solution
    host, _ = g.nlp2op(x, lam)
    L = g.nlp2op_length()
    d_out = torch.full((L,), float("nan"), dtype=torch.float64, device=dev)
    g.nlp2op_dev(dx.data_ptr(), dl.data_ptr(), d_out.data_ptr())
    torch.cuda.synchronize()
    shapes = [(int(sum(p.nodesperinterval)) + 1, len(p.statemin), len(p.controlmin), len(p.pathmin)) for p in op.phases]
    devres = nlp_mod.unpack_nlp2op(d_out.cpu().numpy(), shapes)
    for a, b in zip(host, devres):
        for key in ("time", "state", "control", "costate", "pathmult", "hamiltonian"):
            assert np.array_equal(a[key], b[key], equal_nan=True), key
        assert a["mayer"] == b["mayer"] and a["lagrange"] == b["lagrange"]
    # mesh error
    rel, imax = g.mesh_error(x)
    nrel, nint = sum(r.size for r in rel), sum(v.size for v in imax)
    d_rel = torch.full((nrel,), float("nan"), dtype=torch.float64, device=dev)
    d_imax = torch.full((nint,), float("nan"), dtype=torch.float64, device=dev)
    g.mesh_error_dev(dx.data_ptr(), d_rel.data_ptr(), d_imax.data_ptr())
    torch.cuda.synchronize()
    flat = np.concatenate([r.T.reshape(-1) for r in rel])  # per phase: column-major rows x ns
    assert np.array_equal(d_rel.cpu().numpy(), flat) and np.array_equal(d_imax.cpu().numpy(), np.concatenate(imax))
    g.mesh_error_dev(dx.data_ptr(), None, d_imax.data_ptr())  # interval maxima only
    torch.cuda.synchronize()
    assert np.array_equal(d_imax.cpu().numpy(), np.concatenate(imax))
