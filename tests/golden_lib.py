"""Golden fixtures: outputs of the reference's own code (tests/golden/*.npz, written by
oracle/make_golden.py from oracle/_ref/liblpopc_ref.so)."""
import os

import numpy as np

import cases
import parity

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = cases.CASES + ["launch/u5x4", "hypersensitive/u40x3", "bryson_denham/u7x6"]
RTOL = 1e-12


def load(name):
    return np.load(os.path.join(GOLDEN_DIR, name.replace("/", "__") + ".npz"))


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b))
    return float(np.nanmax(np.abs(a - b) / np.maximum(1.0, np.abs(b)))) if a.size else 0.0


def check_against_golden(name, impl, with_tables=None):
    """impl: object with the TNLP-shaped interface (Oracle or TranscribedNLP adapter below)."""
    G = load(name)
    assert tuple(impl.nlp_info()) == tuple(int(v) for v in G["info"])
    jI, jJ = impl.jac_structure()
    assert np.array_equal(jI, G["jI"]) and np.array_equal(jJ, G["jJ"])
    hI, hJ = impl.h_structure()
    assert np.array_equal(hI, G["hI"]) and np.array_equal(hJ, G["hJ"])
    for a, k in zip(impl.bounds(), ("xl", "xu", "gl", "gu")):
        assert np.array_equal(a, G[k]), k
    # values: TRUE relative 1e-12 entry by entry (tests/parity.py) and the bit-equal fraction per quantity
    op = cases.build(name)
    segs = parity.jac_segments(op, tuple(int(v) for v in G["info"]), G["jI"])
    for xk, sfx in (("guess", "_guess"), ("x", "")):
        x = G[xk]
        f = impl.eval_f(x)
        assert abs(f - float(G["f" + sfx])) <= RTOL * abs(float(G["f" + sfx]))
        parity.assert_parity(impl.eval_grad_f(x), G["grad" + sfx], rtol=RTOL, min_bit_equal=0.9, what="grad f")
        parity.assert_parity(impl.eval_g(x), G["g" + sfx], rtol=RTOL, min_bit_equal=0.999, what="g")
        jv = impl.eval_jac_g(x)
        for sname, idx in segs.items():
            parity.assert_parity(jv[idx], G["jac" + sfx][idx], rtol=RTOL, min_bit_equal=0.999, what="jac " + sname)
    h = impl.eval_h(G["x"], float(G["sigma"]), G["lam"])
    parity.assert_parity(h, G["hess"], rtol=RTOL, min_bit_equal=0.99, what="hessian")
    return G


class CudaAdapter:
    """TranscribedNLP (CUDA, through the C ABI) behind the method names of tests/oracle_lib.Oracle."""

    def __init__(self, nlp):
        self.g = nlp

    def nlp_info(self): return self.g.get_nlp_info()
    def jac_structure(self): return self.g.eval_jac_g(values=False)
    def h_structure(self): return self.g.eval_h(values=False)
    def bounds(self): return self.g.get_bounds_info()
    def probe_dependencies(self, x): return self.g.probe_dependencies(x)
    def eval_f(self, x): return self.g.eval_f(x)
    def eval_grad_f(self, x): return self.g.eval_grad_f(x)
    def eval_g(self, x): return self.g.eval_g(x)
    def eval_jac_g(self, x): return self.g.eval_jac_g(x)
    def eval_h(self, x, s, l): return self.g.eval_h(x, s, l)
