"""hp-Liu mesh refinement (LpLiuHpMeshRefineAlg.cpp:12-260): the product's decision logic (lpb_refine_liu.cpp, behind
lpb_refine_mesh_hp_liu) against sequences of consecutive decisions made by the reference's own code (oracle/_ref,
frozen as tests/golden/liu__*.npz by oracle/make_golden.py) on manufactured solutions re-sampled on every new mesh.
The method is stateful (it compares with the previous grid, its errors and its solution), so a sequence is the test.

CPU: the logic alone, fed with the reference's error matrices.  GPU: the whole step through the C ABI (error estimate
by k_mesh_error, decision on the host), the mesh handed back through lpb_set_mesh / lpb_refresh."""
import os

import numpy as np
import pytest

import cases
import golden_lib
from harness_lib import LiuHarness


def _load(name):
    return np.load(os.path.join(golden_lib.GOLDEN_DIR, "liu__" + name.replace("/", "__") + ".npz"))


def _points(mesh, nodes):
    """composite LGR points of a mesh from the product's host tables (tests/host_harness.cpp)"""
    from lpopc_b200 import examples
    from harness_lib import Harness
    op = examples.hypersensitive()
    op.phases[0].set_mesh(mesh, nodes)
    return Harness(op).tables(0)["points"]


@pytest.mark.parametrize("name", cases.LIU_CASES)
def test_hp_liu_decisions_match_the_reference_sequence_on_cpu(name):
    G = _load(name)
    op = cases.build(name)
    ns = len(op.phases[0].statemin)
    liu = LiuHarness()
    kinds = set()
    for step in range(cases.LIU_STEPS):
        mesh, nodes = G["s%d_in_mesh" % step], G["s%d_in_nodes" % step]
        N = int(nodes.sum())
        x = cases.manufactured_x(name, op, [_points(mesh, nodes)])
        state = x[:ns * (N + 1)].reshape(ns, N + 1).T
        done, mo, no = liu.refine(ns, mesh, nodes, G["s%d_rel" % step], state, **cases.LIU_OPTIONS)
        assert done == bool(G["s%d_done" % step]), step
        assert np.array_equal(no, G["s%d_nodes" % step]), (step, no, G["s%d_nodes" % step])
        assert np.allclose(mo, G["s%d_mesh" % step], rtol=0, atol=1e-14), step
        if step + 1 < cases.LIU_STEPS:
            assert np.array_equal(mo, G["s%d_in_mesh" % (step + 1)]) and np.array_equal(no, G["s%d_in_nodes" % (step + 1)])
        kinds.add("fewer intervals" if no.size < nodes.size else "more intervals" if no.size > nodes.size else "same")
    assert {"fewer intervals", "more intervals"} <= kinds  # the sequences exercise merging and dividing


@pytest.mark.gpu
@pytest.mark.parametrize("name", cases.LIU_CASES)
def test_cuda_hp_liu_refinement_matches_the_reference_sequence(nlp_mod, name):
    G = _load(name)
    op = cases.build(name)
    g = nlp_mod.TranscribedNLP(op)
    for rep in range(2):  # lpb_refine_reset starts the same sequence again
        for step in range(cases.LIU_STEPS):
            mesh, nodes = G["s%d_in_mesh" % step], G["s%d_in_nodes" % step]
            g.set_mesh(0, mesh, nodes)
            g.refresh()
            x = cases.manufactured_x(name, op, g.lgr_points())
            rel, _ = g.mesh_error(x)
            ref = G["s%d_rel" % step]
            assert np.max(np.abs(rel[0] - ref) / np.maximum(1e-30, np.maximum(np.abs(ref), 1e-3 * np.abs(ref).max()))) <= 1e-5  # the estimate itself (LAPACK inverse in the reference vs the product's tables) agrees far better than the decisions need
            done, meshes = g.refine_mesh_hp_liu(x, **cases.LIU_OPTIONS)
            mo, no = meshes[0]
            assert done == bool(G["s%d_done" % step]), step
            assert np.array_equal(no, G["s%d_nodes" % step]), (step, no, G["s%d_nodes" % step])
            assert np.allclose(mo, G["s%d_mesh" % step], rtol=0, atol=1e-14), step
        g.refine_reset()
    g.close()
