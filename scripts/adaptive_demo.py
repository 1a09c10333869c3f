#!/usr/bin/env python
"""Adaptive ph mesh refinement of the reference's hypersensitive example on the GPU: prints the grid history.
    python scripts/adaptive_demo.py [tf] [mesh_tol] [ph|hp-Liu] [intervals] [nodes] [max_grids]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lpopc_b200 import adaptive, examples, nlp, solver  # noqa: E402

tf = float(sys.argv[1]) if len(sys.argv) > 1 else 50.0
tol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-5
method = sys.argv[3] if len(sys.argv) > 3 else "ph"
K0 = int(sys.argv[4]) if len(sys.argv) > 4 else 4
N0 = int(sys.argv[5]) if len(sys.argv) > 5 else 6
grids = int(sys.argv[6]) if len(sys.argv) > 6 else 12
op = examples.hypersensitive(intervals=K0, nodes=N0)
for p in op.phases:  # horizon override (the reference example uses 5000)
    p.SetTimeMin(0.0, tf); p.SetTimeMax(0.0, tf)
    p.timeguess = [0.0, tf]
x, hist = adaptive.solve_adaptive(op, nlp.TranscribedNLP, solver.CudaEvaluator, solver.BatchedIPM, mesh_tol=tol, max_grids=grids, verbose=False, method=method)
for h in hist:
    print(json.dumps(h))
