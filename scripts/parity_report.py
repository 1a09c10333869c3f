#!/usr/bin/env python
"""Per-segment parity table of the CUDA path against (a) the CPU restatement run here on the same inputs and (b) the
outputs of the reference's own sources frozen in tests/golden: bit-equal fraction and TRUE relative error
(tests/parity.py) of g, grad f, the Jacobian by segment (NL defect/path/event/link rows, L, C) and the Hessian by
segment (I-part, E-part, link part).  Run on a GPU box; the committed copy is profiles/r02_parity_report.txt.

  python scripts/parity_report.py > gpurun_out/parity_report.txt
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
import golden_lib  # noqa: E402
import parity  # noqa: E402
from oracle_lib import Oracle  # noqa: E402 (checker)
from lpopc_b200 import nlp  # noqa: E402


def hess_segments(g, op, nnz_h):
    P, Lp = len(op.phases), len(op.links)
    cuts = []
    for p in range(P):
        cuts.append((g.stat("hess_I0.%d" % p), "I"))
        cuts.append((g.stat("hess_E0.%d" % p), "E"))
    for q in range(Lp):
        cuts.append((g.stat("hess_L0.%d" % q), "link"))
    cuts.sort()
    seg = {}
    for i, (off, kind) in enumerate(cuts):
        end = cuts[i + 1][0] if i + 1 < len(cuts) else nnz_h
        seg.setdefault(kind, []).append(np.arange(off, end))
    return {k: np.concatenate(v) for k, v in seg.items()}


def line(tag, what, r):
    print("%-26s %-14s n=%-8d bit-equal %7.3f%%  max true rel %.2e  max abs below floor %.2e  (segment max %.3g)" %
          (tag, what, r["n"], 100 * r["bit_equal"], r["max_rel"], r["max_abs_lo"], r["scale"]))


def main():
    print("# CUDA path vs CPU restatement (same inputs, seed 7) and vs reference goldens; floor = %.0e * segment max" % parity.NOISE_FLOOR_REL)
    for name in golden_lib.GOLDEN_CASES:
        op = cases.build(name)
        o = Oracle(op)
        g = nlp.TranscribedNLP(op)
        G = golden_lib.load(name)
        x, sigma, lam = G["x"], float(G["sigma"]), G["lam"]
        info = g.get_nlp_info()
        jI, _ = g.eval_jac_g(values=False)
        jseg = parity.jac_segments(op, info, jI)
        hseg = hess_segments(g, op, info[3])
        cg, cv = g.eval_g_jac(x)
        cgrad, ch = g.eval_grad_f(x), g.eval_h(x, sigma, lam)
        for tag, rg, rgrad, rv, rh in (("vs restatement", o.eval_g(x), o.eval_grad_f(x), o.eval_jac_g(x), o.eval_h(x, sigma, lam)),
                                       ("vs reference golden", G["g"], G["grad"], G["jac"], G["hess"])):
            line(name, "g " + tag[3:6], parity.report(cg, rg))
            line(name, "grad " + tag[3:6], parity.report(cgrad, rgrad))
            for k, idx in jseg.items():
                if idx.size:
                    line(name, "jac " + k + " " + tag[3:6], parity.report(cv[idx], rv[idx]))
            for k, idx in hseg.items():
                if idx.size:
                    line(name, "hess " + k + " " + tag[3:6], parity.report(ch[idx], rh[idx]))
        g.close()


if __name__ == "__main__":
    main()
