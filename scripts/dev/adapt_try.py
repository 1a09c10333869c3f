import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lpopc_b200 import adaptive, examples, nlp, solver
prob, K0, N0, tol, method, grids = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), sys.argv[5], int(sys.argv[6])
nmax = int(sys.argv[7]) if len(sys.argv) > 7 else 16
nmin = int(sys.argv[8]) if len(sys.argv) > 8 else 4
op = getattr(examples, prob)(intervals=K0, nodes=N0)
t0 = time.time()
x, hist = adaptive.solve_adaptive(op, nlp.TranscribedNLP, solver.CudaEvaluator, solver.BatchedIPM, mesh_tol=tol, max_grids=grids, method=method, max_iter=300, nmax=nmax, nmin=nmin, verbose=True, max_nodes=int(sys.argv[9]) if len(sys.argv) > 9 else None)
for h in hist:
    print(json.dumps({k: h[k] for k in ("grid", "n", "nodes", "intervals", "objective", "status", "iters", "max_rel_error", "mesh_satisfied")}))
print("seconds", time.time() - t0)
