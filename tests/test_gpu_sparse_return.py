"""Sparse return of the Jacobian head through the host-pointer batch call (lpb_eval_g_jac_batch with
pinned buffers): whatever the GPU predicts to be all-zero, the caller's array must end up holding
exactly the device values -- compared bit for bit with the dense (pageable-buffer) return of the
same call, whose values are pinned to the oracle by test_gpu_parity / test_golden."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a).view(np.int64)


def _pinned(shape, fill=np.nan):
    import torch
    t = torch.full(shape, float(fill), dtype=torch.float64).pin_memory()
    return t, t.numpy()


def _batch_inputs(op, g, nb, seed, scale=0.05):
    base = op.guess(g.lgr_points())
    rng = np.random.Generator(np.random.PCG64(seed))
    return base[None, :] + scale * rng.uniform(-1, 1, (nb, base.size))


@pytest.mark.parametrize("name,nb", [("quadrotor/u8x8", 320), ("cartpole/u8x8", 1500), ("launch/u5x4", 400)])
def test_sparse_return_is_exact(nlp_mod, name, nb):
    import torch
    op = cases.build(name)
    g = nlp_mod.TranscribedNLP(op)
    n, m, nnz, _ = g.get_nlp_info()
    X = [_batch_inputs(op, g, nb, 11 + k) for k in range(3)]
    dense = [g.eval_g_jac_batch(x) for x in X]  # pageable numpy buffers: dense return
    assert g.stat("sparse_calls") == 0
    hx = [torch.from_numpy(x).pin_memory() for x in X]
    tg, hg = _pinned((nb, m))
    tv, hv = _pinned((nb, nnz))
    head = g.stat("head_doubles")
    assert nb * head * 8 >= 4 << 20, "batch too small to take the sparse path"

    # call 1 learns the non-zero segments (dense return + device-side flags)
    g.eval_g_jac_batch_ptr(nb, hx[0].data_ptr(), tg.data_ptr(), tv.data_ptr())
    assert np.array_equal(_bits(hv), _bits(dense[0][1])) and np.array_equal(_bits(hg), _bits(dense[0][0]))
    on = g.stat("sparse_on_doubles")
    assert 0 < on <= head

    # call 2 takes the sparse path into a poisoned buffer: every entry must be rewritten
    hv[:] = np.nan
    hg[:] = np.nan
    g.eval_g_jac_batch_ptr(nb, hx[1].data_ptr(), tg.data_ptr(), tv.data_ptr())
    assert np.array_equal(_bits(hv), _bits(dense[1][1])) and np.array_equal(_bits(hg), _bits(dense[1][0]))
    if on * 4 <= head * 3:
        assert g.stat("sparse_calls") == 1

    # forget everything that was learnt: every segment is predicted zero, the flags must bring all of them back
    g.set_option("sparse_forget", 1)
    hv[:] = np.nan
    before = g.stat("sparse_fixups")
    g.eval_g_jac_batch_ptr(nb, hx[2].data_ptr(), tg.data_ptr(), tv.data_ptr())
    assert np.array_equal(_bits(hv), _bits(dense[2][1])) and np.array_equal(_bits(hg), _bits(dense[2][0]))
    assert g.stat("sparse_fixups") > before
    assert g.stat("sparse_on_doubles") == on  # same set learnt again

    # and the steady state after the fix-up
    hv[:] = np.nan
    g.eval_g_jac_batch_ptr(nb, hx[0].data_ptr(), tg.data_ptr(), tv.data_ptr())
    assert np.array_equal(_bits(hv), _bits(dense[0][1]))


def test_sparse_return_grows_with_the_data(nlp_mod):
    """Learn on inputs whose Jacobian has extra zeros (hover: angles, rates and velocities exactly 0 make
    many trigonometric derivatives vanish), then evaluate generic inputs: newly non-zero segments are fetched."""
    import torch
    op = cases.build("quadrotor/u8x8")
    g = nlp_mod.TranscribedNLP(op)
    n, m, nnz, _ = g.get_nlp_info()
    nb = 320
    X1 = _batch_inputs(op, g, nb, 3)
    X0 = np.zeros_like(X1)
    X0[:, n - 1] = 2.0  # tf
    d0, d1 = g.eval_g_jac_batch(X0), g.eval_g_jac_batch(X1)
    tg, hg = _pinned((nb, m))
    tv, hv = _pinned((nb, nnz))
    h0, h1 = torch.from_numpy(X0).pin_memory(), torch.from_numpy(X1).pin_memory()
    g.eval_g_jac_batch_ptr(nb, h0.data_ptr(), tg.data_ptr(), tv.data_ptr())
    on0 = g.stat("sparse_on_doubles")
    assert np.array_equal(_bits(hv), _bits(d0[1]))
    hv[:] = np.nan
    g.eval_g_jac_batch_ptr(nb, h1.data_ptr(), tg.data_ptr(), tv.data_ptr())
    assert np.array_equal(_bits(hv), _bits(d1[1])) and np.array_equal(_bits(hg), _bits(d1[0]))
    assert g.stat("sparse_on_doubles") >= on0
    if g.stat("sparse_on_doubles") > on0:
        assert g.stat("sparse_fixups") > 0
    # back to the sparser inputs: the larger mask stays (monotone), results exact
    hv[:] = np.nan
    g.eval_g_jac_batch_ptr(nb, h0.data_ptr(), tg.data_ptr(), tv.data_ptr())
    assert np.array_equal(_bits(hv), _bits(d0[1]))


def test_sparse_return_off_and_mesh_change(nlp_mod):
    import torch
    op = cases.build("quadrotor/u8x8")
    g = nlp_mod.TranscribedNLP(op)
    nb = 320
    n, m, nnz, _ = g.get_nlp_info()
    X = _batch_inputs(op, g, nb, 5)
    d = g.eval_g_jac_batch(X)
    hx = torch.from_numpy(X).pin_memory()
    tg, hg = _pinned((nb, m))
    tv, hv = _pinned((nb, nnz))
    g.set_option("sparse_return", 0)
    for _ in range(2):
        g.eval_g_jac_batch_ptr(nb, hx.data_ptr(), tg.data_ptr(), tv.data_ptr())
    assert g.stat("sparse_calls") == 0 and np.array_equal(_bits(hv), _bits(d[1]))
    g.set_option("sparse_return", 1)
    for _ in range(2):
        g.eval_g_jac_batch_ptr(nb, hx.data_ptr(), tg.data_ptr(), tv.data_ptr())
    assert g.stat("sparse_calls") == 1
    # a new mesh forgets the learnt segments (offsets change)
    mesh, nodes = np.linspace(-1, 1, 5), [6, 6, 6, 6]
    op.phases[0].set_mesh(mesh, nodes)
    g.set_mesh(0, mesh, nodes)
    g.refresh()
    assert g.stat("sparse_on_doubles") == -1
    n2, m2, nnz2, _ = g.get_nlp_info()
    X2 = _batch_inputs(op, g, nb, 6)
    assert X2.shape == (nb, n2)
    d2 = g.eval_g_jac_batch(X2)
    tv2, hv2 = _pinned((nb, nnz2))
    tg2, hg2 = _pinned((nb, m2))
    hx2 = torch.from_numpy(X2).pin_memory()
    for _ in range(2):
        hv2[:] = np.nan
        g.eval_g_jac_batch_ptr(nb, hx2.data_ptr(), tg2.data_ptr(), tv2.data_ptr())
        assert np.array_equal(_bits(hv2), _bits(d2[1])) and np.array_equal(_bits(hg2), _bits(d2[0]))
    assert g.stat("sparse_calls") == 2


def test_auto_pin_registers_reused_pageable_buffers(nlp_mod):
    """Option auto_pin: pageable caller arrays that come back with the same address are page-locked on the second
    call (IPOPT reuses its arrays), after which the call takes the DMA / sparse-return path -- results unchanged."""
    op = cases.build("quadrotor/u8x8")
    g = nlp_mod.TranscribedNLP(op)
    n, m, nnz, _ = g.get_nlp_info()
    nb = 320
    X = [np.ascontiguousarray(_batch_inputs(op, g, nb, 40 + k)) for k in range(2)]
    dense = [g.eval_g_jac_batch(x) for x in X]
    xbuf, gbuf, vbuf = np.empty((nb, n)), np.empty((nb, m)), np.empty((nb, nnz))
    g.set_option("auto_pin", 1)
    for k in range(5):
        xbuf[:] = X[k % 2]
        gbuf[:] = np.nan
        vbuf[:] = np.nan
        g.eval_g_jac_batch(xbuf, gbuf, vbuf)
        assert np.array_equal(_bits(vbuf), _bits(dense[k % 2][1])) and np.array_equal(_bits(gbuf), _bits(dense[k % 2][0])), k
        if k == 0:
            assert g.stat("pinned_buffers") == 0
        if k >= 1:
            assert g.stat("pinned_buffers") == 3
    assert g.stat("sparse_calls") >= 2  # pinned from call 2 on: call 2 learns the segments, calls 3.. are sparse
    g.set_option("auto_pin", 0)
    assert g.stat("pinned_buffers") == 0
    g.eval_g_jac_batch(xbuf, gbuf, vbuf)
    assert np.array_equal(_bits(vbuf), _bits(dense[0][1]))


@pytest.mark.parametrize("name,nb", [("quadrotor/u8x8", 320), ("launch/u5x4", 400)])
def test_persistent_values_mode_is_exact(nlp_mod, name, nb):
    """Option persistent_values: the caller reuses ONE values array and leaves it alone between calls, so the constant
    tail and the fill pattern of the off-segments are still in place and the host writes nothing -- only on-segments
    cross PCIe.  The array must still hold exactly the device values after every call; an array that was replaced or
    overwritten (sentinels changed) is refilled; values the GPU rewrites anyway may be poisoned freely."""
    import torch
    op = cases.build(name)
    g = nlp_mod.TranscribedNLP(op)
    n, m, nnz, _ = g.get_nlp_info()
    X = [_batch_inputs(op, g, nb, 31 + k) for k in range(5)]
    dense = [g.eval_g_jac_batch(x) for x in X]
    hx = [torch.from_numpy(x).pin_memory() for x in X]
    tg, hg = _pinned((nb, m))
    tv, hv = _pinned((nb, nnz))
    g.set_option("persistent_values", 1)
    g.eval_g_jac_batch_ptr(nb, hx[0].data_ptr(), tg.data_ptr(), tv.data_ptr())  # learns
    g.eval_g_jac_batch_ptr(nb, hx[1].data_ptr(), tg.data_ptr(), tv.data_ptr())  # sparse, full fill
    assert np.array_equal(_bits(hv), _bits(dense[1][1]))
    assert g.stat("persistent_hits") == 0
    if g.stat("sparse_calls") == 0:
        pytest.skip("head too dense for the sparse path")
    g.eval_g_jac_batch_ptr(nb, hx[2].data_ptr(), tg.data_ptr(), tv.data_ptr())  # in place: no host writes
    assert g.stat("persistent_hits") == 1
    assert np.array_equal(_bits(hv), _bits(dense[2][1])) and np.array_equal(_bits(hg), _bits(dense[2][0]))
    # poison what the device rewrites on every call (every value that differs between two inputs is in an on-segment)
    changed = _bits(dense[2][1]) != _bits(dense[3][1])
    hv[changed] = np.nan
    g.eval_g_jac_batch_ptr(nb, hx[3].data_ptr(), tg.data_ptr(), tv.data_ptr())
    assert g.stat("persistent_hits") == 2
    assert np.array_equal(_bits(hv), _bits(dense[3][1]))
    # an overwritten array (sentinels gone) is detected and refilled completely
    hv[:] = np.nan
    g.eval_g_jac_batch_ptr(nb, hx[4].data_ptr(), tg.data_ptr(), tv.data_ptr())
    assert g.stat("persistent_hits") == 2
    assert np.array_equal(_bits(hv), _bits(dense[4][1]))
    # a different array of the same size: full fill, then in place again
    tv2, hv2 = _pinned((nb, nnz))
    g.eval_g_jac_batch_ptr(nb, hx[0].data_ptr(), tg.data_ptr(), tv2.data_ptr())
    assert g.stat("persistent_hits") == 2 and np.array_equal(_bits(hv2), _bits(dense[0][1]))
    g.eval_g_jac_batch_ptr(nb, hx[1].data_ptr(), tg.data_ptr(), tv2.data_ptr())
    assert g.stat("persistent_hits") == 3 and np.array_equal(_bits(hv2), _bits(dense[1][1]))
    # forgetting the mask changes the plan: the flags bring the segments back and the next call refills
    g.set_option("sparse_forget", 1)
    g.eval_g_jac_batch_ptr(nb, hx[2].data_ptr(), tg.data_ptr(), tv2.data_ptr())
    assert np.array_equal(_bits(hv2), _bits(dense[2][1]))
    g.eval_g_jac_batch_ptr(nb, hx[3].data_ptr(), tg.data_ptr(), tv2.data_ptr())
    assert np.array_equal(_bits(hv2), _bits(dense[3][1]))
    g.close()
