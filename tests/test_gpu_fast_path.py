"""Single-problem fast path (lpb_handle::FastPath): the TNLP callbacks of one x served from ONE captured CUDA graph
launch must return exactly what the regular per-callback path returns, keep doing so when x changes, when only some
callbacks are asked for, across a mesh change, and the cache must never serve a stale x."""
import numpy as np
import pytest

import cases
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.int64)


@pytest.mark.parametrize("name", ["hypersensitive", "bryson_denham/ragged", "launch/u5x4", "orbit_raising", "quadrotor"])
def test_fast_path_matches_regular_path(nlp_mod, name):
    op = cases.build(name)
    o = Oracle(op)
    g = nlp_mod.TranscribedNLP(op)
    _, x, sigma, lam = cases.inputs(op, o, 17)
    xs = [x, x * (1 + 1e-9), x + 1e-3 * np.cos(np.arange(x.size))]
    g.set_option("fast_path", 0)
    ref = [(g.eval_f(v), g.eval_grad_f(v), g.eval_g(v), g.eval_jac_g(v), g.eval_h(v, sigma, lam)) for v in xs]
    g.set_option("fast_path", 1)
    e0, h0 = g.stat("fast_path_evals"), g.stat("fast_path_hits")
    for v, r in zip(xs, ref):
        assert g.eval_f(v) == r[0]
        assert np.array_equal(_bits(g.eval_grad_f(v)), _bits(r[1]))
        assert np.array_equal(_bits(g.eval_g(v)), _bits(r[2]))
        assert np.array_equal(_bits(g.eval_jac_g(v)), _bits(r[3]))
        assert np.array_equal(_bits(g.eval_h(v, sigma, lam)), _bits(r[4]))
    # one graph launch per new x, every other callback (grad_f, g, jac_g, h) served from the stage
    assert g.stat("fast_path_evals") - e0 == len(xs)
    assert g.stat("fast_path_hits") - h0 == 4 * len(xs)
    for v in xs:  # other multipliers on a cached x
        assert np.array_equal(_bits(g.eval_h(v, 0.5 * sigma, -lam)), _bits(g_regular_h(g, v, 0.5 * sigma, -lam)))
    # callbacks in another order, x alternating: never a stale result
    for k in range(6):
        v, r = xs[k % 3], ref[k % 3]
        assert np.array_equal(_bits(g.eval_jac_g(v)), _bits(r[3]))
        assert np.array_equal(_bits(g.eval_g(xs[(k + 1) % 3])), _bits(ref[(k + 1) % 3][2]))
    gg, gv = g.eval_g_jac(xs[1])
    assert np.array_equal(_bits(gg), _bits(ref[1][2])) and np.array_equal(_bits(gv), _bits(ref[1][3]))
    g.close()


def g_regular_h(g, v, sigma, lam):
    g.set_option("fast_path", 0)
    h = g.eval_h(v, sigma, lam)
    g.set_option("fast_path", 1)
    return h


def test_fast_path_survives_a_mesh_change(nlp_mod):
    op = cases.build("bryson_denham")
    o = Oracle(op)
    g = nlp_mod.TranscribedNLP(op)
    g.set_option("fast_path", 1)
    _, x, sigma, lam = cases.inputs(op, o, 5)
    assert np.array_equal(_bits(g.eval_jac_g(x)), _bits(o.eval_jac_g(x)))
    cases.ragged_mesh(op.phases[0], 3, 5, 3, 9)
    o.set_mesh(0, op.phases[0].meshpoints, op.phases[0].nodesperinterval); o.refresh()
    g.set_mesh(0, op.phases[0].meshpoints, op.phases[0].nodesperinterval); g.refresh()
    _, x2, sigma, lam = cases.inputs(op, o, 6)
    assert g.get_nlp_info() == (o.n, o.m, o.nnz_jac, o.nnz_h)
    from test_gpu_parity import rel_err
    assert rel_err(g.eval_g(x2), o.eval_g(x2)) <= 1e-12
    assert rel_err(g.eval_jac_g(x2), o.eval_jac_g(x2)) <= 1e-12
    assert rel_err(g.eval_h(x2, sigma, lam), o.eval_h(x2, sigma, lam)) <= 1e-12
    assert abs(g.eval_f(x2) - o.eval_f(x2)) <= 1e-12 * abs(o.eval_f(x2))
    g.close()
