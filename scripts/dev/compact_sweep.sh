#!/bin/bash
# solves/s of the batched IPM against the batch-compaction threshold (same box, interleaved)
python scripts/solve_bench.py cartpole 64 --chunk 64 >/dev/null 2>&1
for c in 0.5 0.75 0.9 0.5 0.75 0.9; do
  python scripts/solve_bench.py quadrotor 4096 --chunk 4096 --compact-at $c 2>/dev/null | tail -1 > /tmp/sb.json
  python -c "import json; d=json.load(open('/tmp/sb.json')); print('compact_at', $c, d['value'], d['converged'], d['iters_max'], d['seconds'])"
done
