#!/usr/bin/env python
"""Instruction mix + stall summary from `ncu -i X.ncu-rep --page source --csv` output (SASS view)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
ix = {k: i for i, k in enumerate(h)}
data = [r for r in rows[2:] if len(r) == len(h) and r[0].startswith('0x')]
tot = 0; ops = collections.Counter(); samp = collections.Counter(); stalls = collections.Counter()
stall_cols = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
for r in data:
    t = r[ix['Source']].strip().split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    n = int(r[ix['Instructions Executed']]); s = int(r[ix['# Samples']])
    ops[op] += n; samp[op] += s; tot += n
    for k in stall_cols: stalls[k] += int(r[ix[k]])
print('static instr', len(data), 'dyn warp inst', tot)
for op, n in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 16):
    print('%-10s %12d %5.1f%%  samples %d' % (op, n, 100 * n / tot, samp[op]))
print(stalls.most_common(8))
