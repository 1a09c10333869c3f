#!/usr/bin/env python
"""Adaptive ph mesh refinement with a HOST outer solver (SciPy SLSQP in IPOPT's place) on the CUDA TNLP callbacks:
prints the grid history.
    python scripts/adaptive_host_demo.py [brachistochrone|bryson_denham] [mesh_tol]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lpopc_b200 import adaptive, examples, nlp  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "brachistochrone"
tol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-7
op = getattr(examples, name)(intervals=2, nodes=4)
x, hist = adaptive.solve_adaptive(op, nlp.TranscribedNLP, None, None, mesh_tol=tol, max_grids=10, host_solver=adaptive.slsqp_host_solver(ftol=1e-12))
for h in hist:
    print(json.dumps(h))
