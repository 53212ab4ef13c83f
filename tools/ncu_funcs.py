#!/usr/bin/env python
"""Aggregate an ncu report's per-instruction counts and stall samples by device function (ant_physics.cuh) / file."""
import collections, csv, io, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_attrib as A
rep, kn = sys.argv[1], sys.argv[2]
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
st = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
sec = rows[st[0] + 1:(st[1] if len(st) > 1 else len(rows))]
hdr = sec[0]; ia, ii, isamp = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
base = int(sec[1][ia], 16); W = int(sec[1][ii])
lines = A.sass_lines(kn)
src = open(os.path.join(A.ROOT, 'po_brax_b200/csrc/ant_physics.cuh')).read().splitlines()
def fn_of(f, l):
    if f != 'ant_physics.cuh': return f
    for i in range(l - 1, -1, -1):
        t = src[i]
        if t.startswith('__device__') or t.startswith('template'):
            j = i
            while '(' not in src[j]: j += 1
            return src[j].split('(')[0].split()[-1]
    return '?'
agg, sagg = collections.Counter(), collections.Counter()
for r in sec[1:]:
    (f, l), _ = lines.get(int(r[ia], 16) - base, (('?', 0), ''))
    k = fn_of(f, l); agg[k] += int(r[ii]) / W; sagg[k] += int(r[isamp])
tot, ts = sum(agg.values()), sum(sagg.values())
print(f'instructions per warp {tot:.0f}')
for k, v in agg.most_common(22): print(f'{k:28s} {v:8.1f} instr/warp {100*v/tot:5.1f}%   samples {100*sagg[k]/ts:5.1f}%')
