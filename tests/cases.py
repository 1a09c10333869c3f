"""Shared test cases: problem builders at oracle-friendly sizes + seeded inputs."""
import numpy as np

from lpopc_b200 import examples


def ragged_mesh(phase, seed, intervals, nmin=2, nmax=9):
    rng = np.random.Generator(np.random.PCG64(seed))
    cuts = np.sort(rng.uniform(-0.9, 0.9, intervals - 1))
    mesh = np.concatenate([[-1.0], cuts, [1.0]])
    nodes = rng.integers(nmin, nmax + 1, intervals)
    phase.set_mesh(mesh, nodes)


def build(name):
    """name -> OptimalProblem.  Variants after '/' pick the mesh."""
    base, _, var = name.partition("/")
    if base == "hypersensitive":
        op = examples.hypersensitive()
    elif base == "hypersensitive_analytic":
        op = examples.hypersensitive(first_derive="analytic")
    elif base == "bryson_denham":
        op = examples.bryson_denham()
    elif base == "launch":
        op = examples.launch()
    elif base == "orbit_raising":
        op = examples.orbit_raising(intervals=6, nodes=5)
    elif base == "brachistochrone":
        op = examples.brachistochrone(intervals=4, nodes=6)
    elif base == "quadrotor":
        op = examples.quadrotor(intervals=3, nodes=4)
    elif base == "cartpole":
        op = examples.cartpole(intervals=3, nodes=5)
    elif base == "synthetic20":
        op = examples.synthetic20(intervals=3, nodes=4)
    elif base == "two_stage":
        op = examples.two_stage()
    elif base == "two_stage_fd":
        op = examples.two_stage(first_derive="finite-difference")
    else:
        raise KeyError(name)
    if var == "ragged":
        for ip, p in enumerate(op.phases):
            ragged_mesh(p, 100 + ip, 3 + ip)
    elif var == "two":
        for p in op.phases:
            p.set_mesh([-1.0, 1.0], [2])
    elif var.startswith("u"):  # uKxN
        k, n = var[1:].split("x")
        for p in op.phases:
            examples.uniform_mesh(p, int(k), int(n))
    return op


CASES = ["hypersensitive", "hypersensitive/ragged", "hypersensitive/two", "hypersensitive_analytic", "bryson_denham",
         "bryson_denham/ragged", "launch", "launch/ragged", "orbit_raising", "brachistochrone", "quadrotor", "cartpole",
         "synthetic20", "two_stage", "two_stage/ragged", "two_stage_fd"]


def lgr_points_of(oracle):
    return [oracle.tables(ip)["points"] for ip in range(len(oracle.op.phases))]


def inputs(op, oracle, seed):
    """x = reference guess (two-point guess on the LGR nodes) perturbed; sigma; lambda."""
    rng = np.random.Generator(np.random.PCG64(seed))
    guess = op.guess(lgr_points_of(oracle))
    x = guess + 1e-2 * rng.uniform(-1, 1, guess.size) * (np.abs(guess) + 0.1)
    # keep every phase duration positive after the perturbation
    lam = rng.uniform(-1, 1, oracle.m)
    sigma = float(rng.uniform(0.5, 1.5))
    return guess, x, sigma, lam
