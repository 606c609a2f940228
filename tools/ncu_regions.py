#!/usr/bin/env python
"""Stall samples and executed instructions between consecutive barriers of the profiled kernel (SASS order).
usage: python tools/ncu_regions.py rep.ncu-rep n_symbols"""
import csv, io, subprocess, sys
rep, nsym = sys.argv[1], float(sys.argv[2])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(io.StringIO(src)) if len(r) > 6]
h = rows[0]; idx = {n: i for i, n in enumerate(h)}
stall_cols = [i for i, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
seg = []; cur = [0, 0, 0, {}]; tot = 0
for k, r in enumerate(rows[1:]):
    try:
        s = float(r[idx['# Samples']]); e = float(r[idx['Instructions Executed']])
    except ValueError:
        continue
    cur[0] += s; cur[1] += e; cur[2] += 1; tot += s
    for i in stall_cols:
        v = float(r[i] or 0)
        if v: cur[3][h[i][6:]] = cur[3].get(h[i][6:], 0) + v
    srcl = r[idx['Source']]
    if any(t in srcl for t in ('BAR.SYNC', 'SYNCS', 'UBLKCP', 'WARPSYNC.ALL')):
        seg.append((k, srcl.strip()[:34], *cur)); cur = [0, 0, 0, {}]
seg.append((k, 'end', *cur))
for k, srcl, s, e, n, st in seg:
    if s / tot < 0.003: continue
    top = sorted(st.items(), key=lambda x: -x[1])[:4]
    print(f'{k:5d} {srcl:36s} samples {100*s/tot:5.1f}%  warp-inst/symbol {e/nsym:7.1f}  ' + ' '.join(f'{a}:{100*b/tot:.1f}' for a, b in top))
