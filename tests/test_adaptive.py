"""Adaptive mesh loop (lpopc_b200/adaptive.py): guess transfer on the CPU; the full loop -- solve, GPU mesh-error
estimate, ph refinement, GPU re-transcription -- on the reference's hypersensitive example (-m gpu)."""
import numpy as np
import pytest

from lpopc_b200 import adaptive, examples


def test_transfer_guess_reproduces_smooth_profiles():
    op = examples.bryson_denham(intervals=3, nodes=5)
    rng = np.random.Generator(np.random.PCG64(1))
    old = [np.sort(rng.uniform(-1, 1, 15))]
    old[0][0] = -1.0
    new = [np.sort(rng.uniform(-1, 1, 22))]
    new[0][0] = -1.0
    f = [lambda t: 1 + 0.5 * t, lambda t: np.sin(t), lambda t: t ** 2, lambda t: np.cos(2 * t)]
    tau_o, tau_n = np.concatenate([old[0], [1.0]]), np.concatenate([new[0], [1.0]])
    x = np.concatenate([f[0](tau_o), f[1](tau_o), f[2](tau_o), f[3](tau_o[:-1]), [0.0, 7.5]])
    y = adaptive.transfer_guess(op, x, old, new, control_end=[[f[3](1.0)]])  # the control's end row comes from lpb_nlp2op in the loop
    assert y.size == 3 * 23 + 22 + 2 and np.array_equal(y[-2:], [0.0, 7.5])
    assert np.allclose(y[:23], f[0](tau_n), atol=1e-12)          # linear profiles are reproduced exactly
    assert np.allclose(y[23:46], f[1](tau_n), atol=2e-2)
    assert np.allclose(y[69:91], f[3](tau_n[:-1]), atol=5e-2)


@pytest.mark.gpu
def test_adaptive_loop_hypersensitive_on_gpu():
    from lpopc_b200 import nlp, solver
    op = examples.hypersensitive(intervals=4, nodes=6)
    for p in op.phases:  # a horizon the ph method resolves in a few grids (the reference example uses 5000)
        p.SetTimeMin(0.0, 50.0); p.SetTimeMax(0.0, 50.0)
        p.timeguess = [0.0, 50.0]
    x, hist = adaptive.solve_adaptive(op, nlp.TranscribedNLP, solver.CudaEvaluator, solver.BatchedIPM, mesh_tol=1e-5, max_grids=12)
    assert len(hist) >= 3 and all(h["status"] == 0 for h in hist)
    # the loop ends on a satisfied mesh, or where the reference's truncated node increment stalls (see adaptive.py)
    assert hist[-1]["mesh_satisfied"] or hist[-1].get("mesh_stalled")
    assert hist[-1]["n"] > hist[0]["n"] and hist[-1]["max_rel_error"] <= 0.05 * hist[0]["max_rel_error"]
    assert hist[-1]["max_rel_error"] <= 16e-5  # within the factor N_k <= Nmax of the tolerance that the truncation leaves
    assert abs(hist[-1]["objective"] - 1.33077) <= 2e-3  # value the refined meshes converge to
    assert x.size == hist[-1]["n"]


@pytest.mark.gpu
def test_adaptive_loop_brachistochrone_with_host_outer_solver_on_gpu():
    """Single phase, free final time, through the whole mesh loop with a HOST outer solver on the CUDA TNLP
    callbacks -- the reference's own arrangement (IPOPT on the host; SciPy SLSQP here): solve -> GPU mesh-error
    estimate -> ph refinement -> GPU re-transcription (new index maps) -> spline warm start, until the mesh
    satisfies the tolerance."""
    from lpopc_b200 import nlp
    op = examples.brachistochrone(intervals=2, nodes=4)
    x, hist = adaptive.solve_adaptive(op, nlp.TranscribedNLP, None, None, mesh_tol=1e-7, max_grids=8,
                                      host_solver=adaptive.slsqp_host_solver(ftol=1e-12))
    assert len(hist) >= 2 and all(h["status"] == 0 for h in hist)
    assert hist[-1]["mesh_satisfied"] and hist[-1]["max_rel_error"] <= 1e-7 < hist[0]["max_rel_error"]
    assert hist[-1]["n"] > hist[0]["n"] and x.size == hist[-1]["n"]
    # the refined mesh reproduces the fine fixed-mesh optimum of the same functor (tests/test_host_outer_loop.py)
    assert abs(hist[-1]["objective"] - 0.824338669) <= 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["ph", "hp-Liu"])
def test_adaptive_loop_reference_hypersensitive_horizon_gpu_resident(method):
    """The reference's hypersensitive example as shipped (t_f = 5000, HyperSensitive.cpp:17; the example itself selects
    hp-Liu, :56) through the whole mesh loop with the GPU-resident outer solver, by both refinement methods: the
    boundary layers are found, the objective settles at the value the refined meshes agree on."""
    from lpopc_b200 import nlp, solver
    op = examples.hypersensitive(intervals=20, nodes=6)
    x, hist = adaptive.solve_adaptive(op, nlp.TranscribedNLP, solver.CudaEvaluator, solver.BatchedIPM, mesh_tol=1e-6, max_grids=10,
                                      method=method, max_iter=300)
    assert all(h["status"] == 0 for h in hist) and len(hist) >= 5
    # hp-Liu resolves the layers with ~100 nodes by grid 7 (error estimate at the tolerance); ph adds nodes more slowly
    assert min(abs(h["objective"] - 1.3308068) for h in hist[3:]) <= (2e-6 if method == "hp-Liu" else 5e-4)
    assert min(h["max_rel_error"] for h in hist) <= 1e-4 * hist[0]["max_rel_error"]
    assert len({tuple(h["nodes"]) for h in hist}) >= 4  # the mesh really changed from grid to grid
    if method == "hp-Liu":
        assert min(h["nodes"][0] for h in hist[5:]) < hist[0]["nodes"][0]  # it also removes nodes where they are not needed


@pytest.mark.gpu
def test_adaptive_loop_brachistochrone_gpu_resident_solver():
    """Free final time with the GPU-RESIDENT outer solver (filter line search): no host solver in the loop."""
    from lpopc_b200 import nlp, solver
    op = examples.brachistochrone(intervals=2, nodes=4)
    # mesh tolerance above the solver's KKT tolerance (1e-6): the error estimate cannot fall below the solve's own accuracy
    x, hist = adaptive.solve_adaptive(op, nlp.TranscribedNLP, solver.CudaEvaluator, solver.BatchedIPM, mesh_tol=1e-5, max_grids=8, max_iter=200)
    assert all(h["status"] == 0 for h in hist) and len(hist) >= 2
    assert hist[-1]["mesh_satisfied"] and hist[-1]["max_rel_error"] <= 1e-5 < hist[0]["max_rel_error"]
    assert abs(hist[-1]["objective"] - 0.824338669) <= 1e-7


@pytest.mark.gpu
def test_adaptive_loop_brachistochrone_to_two_thousand_nodes():
    """BASELINE config 2: brachistochrone through the adaptive loop until the mesh holds ~2k LGR nodes.  The mesh tolerance
    is set below what finite-difference derivatives can certify, so ph refinement keeps dividing (Nmax = 8, Nmin = 4) and
    every grid means new sizes, new index maps on the GPU (lpb_set_mesh + lpb_refresh), a spline-transferred warm start
    and a GPU-resident solve; max_nodes stops the loop at the first mesh beyond 2400 nodes."""
    from lpopc_b200 import nlp, solver
    op = examples.brachistochrone(intervals=40, nodes=6)
    x, hist = adaptive.solve_adaptive(op, nlp.TranscribedNLP, solver.CudaEvaluator, solver.BatchedIPM, mesh_tol=1e-12, nmax=8, nmin=4,
                                      max_grids=24, max_iter=300, max_nodes=2400)
    assert hist[-1].get("node_cap") and not hist[-1]["mesh_satisfied"]
    assert 1600 <= hist[-1]["nodes"][0] <= 2400, [h["nodes"] for h in hist]
    assert all(h["status"] == 0 for h in hist)
    sizes = [h["n"] for h in hist]
    assert all(b > a for a, b in zip(sizes, sizes[1:]))            # every grid was a re-transcription to a larger NLP
    assert all(abs(h["objective"] - 0.824338669) <= 2e-5 for h in hist)  # KKT tolerance 1e-6 per solve
    assert hist[-1]["max_rel_error"] <= 1e-7
    assert x.shape == (hist[-1]["n"],)
