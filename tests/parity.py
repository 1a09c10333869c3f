"""Parity metrics stricter than max|a-b|/max(1,|b|): bit-equal fraction and TRUE relative error per segment.

north_star: fp64 function and Jacobian values within 1e-12 RELATIVE of the reference.  Every implementation here
(CUDA, restatement, reference sources) evaluates the user functions bit-identically (shared functor headers,
deterministic elementary functions, no contraction), so a forward-difference quotient is either bit-equal or differs
only through the few operations that are allowed to round differently (the LGR differentiation-matrix entries, which
each side computes with its own host code, and block sums).  `report` therefore states, per segment:
  bit_equal   fraction of entries with identical bits
  max_rel     max |a-b|/|b| over entries with |b| > floor          (true relative error)
  max_abs_lo  max |a-b| over entries with |b| <= floor              (entries too small to carry 12 digits)
floor = NOISE_FLOOR_REL * max|b| of the segment: a forward difference of f carries an absolute rounding error of
about eps*|f|/h = 2.2e-10*|f| (h = 1e-6), so entries more than ~10 orders below the segment's largest entry are below
the noise of the scheme itself in the REFERENCE too; they are compared absolutely against ATOL_LO * max|b|.
"""
import numpy as np

NOISE_FLOOR_REL = 1e-10
ATOL_LO = 1e-22  # absolute bound below the floor, relative to the segment maximum: 1e-12 * floor


def report(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
    b = np.ascontiguousarray(b, dtype=np.float64).reshape(-1)
    assert a.shape == b.shape
    if a.size == 0:
        return dict(n=0, bit_equal=1.0, max_rel=0.0, max_abs_lo=0.0, scale=0.0)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    fin = ~np.isnan(b)
    a, b = a[fin], b[fin]
    scale = float(np.max(np.abs(b))) if b.size else 0.0
    hi = np.abs(b) > NOISE_FLOOR_REL * scale
    d = np.abs(a - b)
    return dict(n=int(a.size), bit_equal=float(np.mean(a.view(np.int64) == b.view(np.int64))),
                max_rel=float(np.max(d[hi] / np.abs(b[hi]))) if hi.any() else 0.0,
                max_abs_lo=float(np.max(d[~hi])) if (~hi).any() else 0.0, scale=scale)


def assert_parity(a, b, rtol=1e-12, min_bit_equal=0.0, what=""):
    r = report(a, b)
    assert r["max_rel"] <= rtol, (what, r)
    assert r["max_abs_lo"] <= max(ATOL_LO, rtol * NOISE_FLOOR_REL) * max(r["scale"], 1e-300), (what, r)
    assert r["bit_equal"] >= min_bit_equal, (what, r)
    return r


def jac_segments(op, n_info, jI):
    """Index ranges of the Jacobian value vector by segment: NL rows of defects / paths / events, link rows, the
    linear rows L and the constant differentiation-matrix tail C (order [NL | L | C])."""
    n, m, nnz, _ = n_info
    kinds = np.empty(m, dtype=np.int8)  # 0 defect, 1 path, 2 event, 3 link, 4 linear
    r = 0
    for p in op.phases:
        N = int(np.sum(p.nodesperinterval))
        ns, npth, ne = len(p.statemin), len(p.pathmin), len(p.eventmin)
        kinds[r:r + ns * N] = 0
        kinds[r + ns * N:r + (ns + npth) * N] = 1
        kinds[r + (ns + npth) * N:r + (ns + npth) * N + ne] = 2
        r += (ns + npth) * N + ne
    for l in op.links:
        kinds[r:r + len(l.linkmin)] = 3
        r += len(l.linkmin)
    kinds[r:] = 4
    k = kinds[np.asarray(jI)]
    lin = np.flatnonzero(k == 4)
    l0, l1 = (int(lin[0]), int(lin[-1]) + 1) if lin.size else (nnz, nnz)
    seg = {"L": np.arange(l0, l1), "C": np.arange(l1, nnz)}
    head = np.arange(0, l0)
    for name, code in (("NL/defect", 0), ("NL/path", 1), ("NL/event", 2), ("NL/link", 3)):
        seg[name] = head[k[:l0] == code]
    return seg
