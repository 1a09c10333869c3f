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


# ---- manufactured solutions for the mesh-refinement sequences (hp-Liu goldens) ------------------------------------
LIU_CASES = ["hypersensitive/u8x4", "bryson_denham/u7x6"]
LIU_STEPS = 4
LIU_OPTIONS = dict(tol=1e-5, nmax=12, ratio_r=1.2)


def manufactured_x(name, op, points):
    """An NLP vector whose states and controls satisfy the problem's dynamics exactly as functions of time, sampled on
    the composite LGR points of the current mesh (points = [tau per phase]): the mesh-error estimate then measures
    genuine interpolation error -- small where the trajectory is smooth, large in its boundary layers -- which is what
    the refinement decisions react to.  hypersensitive: x(t) with two boundary layers, u = x' + x^3.  Bryson-Denham:
    x1 = a sin(w t), x2 = x1', u = x1'', x3 = int u^2 / 2."""
    base = name.split("/")[0]
    tau = np.concatenate([np.asarray(points[0], dtype=np.float64), [1.0]])
    N = tau.size - 1
    if base == "hypersensitive":
        t0, tf = 0.0, 5000.0
        s = 9.0
        x = 1.5 * np.exp(-(tau + 1.0) * s) + np.exp((tau - 1.0) * s) + 0.05 * np.sin(2.0 * tau)
        dx_dtau = -1.5 * s * np.exp(-(tau + 1.0) * s) + s * np.exp((tau - 1.0) * s) + 0.1 * np.cos(2.0 * tau)
        u = dx_dtau * (2.0 / (tf - t0)) + x ** 3
        return np.concatenate([x, u[:N], [t0, tf]])
    if base == "bryson_denham":
        t0, tf = 0.0, 1.3
        t = (tf - t0) * (tau + 1.0) / 2.0 + t0
        a, w = 0.1, 6.0
        x1, x2, u = a * np.sin(w * t), a * w * np.cos(w * t), -a * w * w * np.sin(w * t)
        x3 = 0.5 * (a * w * w) ** 2 * (t / 2.0 - np.sin(2.0 * w * t) / (4.0 * w))
        return np.concatenate([x1, x2, x3, u[:N], [t0, tf]])
    raise KeyError(name)
