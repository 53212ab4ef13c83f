#!/usr/bin/env python
"""Opcode mix (instructions per warp) and stall-sample totals of the first kernel in an ncu report."""
import collections, csv, io, re, subprocess, sys
def load(rep):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    st = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
    sec = rows[st[0] + 1:(st[1] if len(st) > 1 else len(rows))]
    hdr = sec[0]; I = hdr.index('Instructions Executed'); W = int(sec[1][I])
    ops = collections.Counter(); stalls = collections.Counter()
    sc = [h for h in hdr if h.startswith('stall_') and 'Not' not in h]
    for r in sec[1:]:
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[1].strip())
        ops[m.group(2).split('.')[0]] += int(r[I]) / W
        for h in sc: stalls[h] += int(r[hdr.index(h)])
    return ops, stalls
if __name__ == '__main__':
    reps = sys.argv[1:]
    data = [load(r) for r in reps]
    keys = sorted(set().union(*[d[0].keys() for d in data]), key=lambda k: -data[0][0].get(k, 0))
    print('op'.ljust(10), *[r.split('/')[-1][:18].rjust(18) for r in reps])
    for k in keys[:36]: print(k.ljust(10), *[f'{d[0].get(k, 0):18.1f}' for d in data])
    print('TOTAL'.ljust(10), *[f'{sum(d[0].values()):18.1f}' for d in data])
    sk = sorted(set().union(*[d[1].keys() for d in data]), key=lambda k: -data[0][1].get(k, 0))
    for k in sk[:10]: print(k.ljust(24), *[f'{d[1].get(k, 0):12d}' for d in data])
