#!/usr/bin/env python
"""Dynamic per-warp instruction counts of a kernel from an ncu --set full --import-source on report, by
contiguous SASS address range (split wherever the executed count changes by more than --tol, i.e. at the
boundaries of divergent regions), with average active threads per range.

    python tools/ncu_region.py gpurun_out/prof.ncu-rep [--min 5]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
mn = float(sys.argv[sys.argv.index('--min') + 1]) if '--min' in sys.argv else 5.0
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]
ia, isrc, ii, it, isamp = (hdr.index(k) for k in ('Address', 'Source', 'Instructions Executed',
                                                  'Thread Instructions Executed', '# Samples'))
body = [r for r in rows[2:] if len(r) > isamp and r[ia].startswith('0x')]
base = int(body[0][ia], 16)
warps = float(body[0][ii])
recs = [(int(r[ia], 16) - base, r[isrc].strip(), float(r[ii]) / warps, float(r[it]) / max(float(r[ii]), 1), float(r[isamp]))
        for r in body]
tot = sum(r[2] for r in recs)
tots = sum(r[4] for r in recs)
print(f'warps {warps:.0f}  instructions/warp {tot:.1f}  samples {tots:.0f}')
start = 0
for i in range(1, len(recs) + 1):
    if i == len(recs) or abs(recs[i][2] - recs[start][2]) > 0.02 * max(recs[start][2], 0.05):
        seg = recs[start:i]
        n = sum(r[2] for r in seg)
        if n >= mn:
            thr = sum(r[2] * r[3] for r in seg) / n
            print(f'{seg[0][0]:#7x}..{seg[-1][0]:#7x} static {len(seg):4d}  exec/warp each {seg[0][2]:7.2f}  dyn {n:7.1f} '
                  f'({100 * n / tot:4.1f}%)  thr {thr:4.1f}  samples {100 * sum(r[4] for r in seg) / tots:4.1f}%  {seg[0][1][:40]}')
        start = i
