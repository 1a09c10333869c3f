#!/usr/bin/env python
"""Per CUDA-source-line instruction counts from `ncu --page source --csv --print-source cuda,sass`."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
fil = None; out = []
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': fil = r[1].split('/')[-1]; continue
    if len(r) > 8 and r[0].isdigit():
        try: out.append((int(r[7]), int(r[6]), fil, int(r[0]), r[1].strip()[:110]))
        except ValueError: pass
tot = sum(o[0] for o in out)
print('total warp inst', tot)
for n, s, f, ln, src in sorted(out, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print('%10d %5.1f%% samp %5d  %s:%d  %s' % (n, 100.0 * n / tot, s, f, ln, src))
