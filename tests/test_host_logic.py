"""Product host logic (tables, layout, closed-form index maps) vs the CPU oracle -- no GPU needed."""
import numpy as np
import pytest

import cases
from harness_lib import Harness
from oracle_lib import Oracle


@pytest.mark.parametrize("name", cases.CASES + ["launch/u5x4", "hypersensitive/u40x3"])
def test_structure_and_tables_match_oracle(name):
    op = cases.build(name)
    o = Oracle(op)
    h = Harness(op)
    assert (h.n, h.m, h.nnz_jac, h.nnz_h) == (o.n, o.m, o.nnz_jac, o.nnz_h)
    jI, jJ, hI, hJ = h.structure()
    oI, oJ = o.jac_structure()
    assert np.array_equal(jI, oI) and np.array_equal(jJ, oJ)
    oI, oJ = o.h_structure()
    assert np.array_equal(hI, oI) and np.array_equal(hJ, oJ)
    for ip in range(len(op.phases)):
        th, to = h.tables(ip), o.tables(ip)
        assert np.array_equal(th["points"], to["points"])          # bit-exact
        assert np.array_equal(th["weights"], to["weights"])
        for k in ("D", "Doffdiag"):
            for a, b in zip(th[k], to[k]):
                assert np.array_equal(a, b), k
        N = h.N[ip]
        assert np.array_equal(th["ddiag"], to["Diag"][2][:N])


def test_hessian_structure_with_sparse_dependencies():
    op = cases.build("launch")
    o = Oracle(op)
    guess, x, _, _ = cases.inputs(op, o, 11)
    dep = o.probe_dependencies(guess)
    assert 0 < dep.sum() < dep.size  # launch dynamics are not dense
    h = Harness(op, dep)
    assert h.nnz_h == o.nnz_h
    _, _, hI, hJ = h.structure()
    oI, oJ = o.h_structure()
    assert np.array_equal(hI, oI) and np.array_equal(hJ, oJ)
