#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
    python tools/launch_summary.py gpurun_out/launches.csv "<command that was profiled>" > profiles/launches_rN_summary.txt
"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[12] == 'gpu__time_duration.sum']
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split('(')[0][:60]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[14].replace(',', '')) / 1e3
tot = sum(v[1] for v in agg.values())
print(f'# ncu --metrics gpu__time_duration.sum --clock-control none: {sys.argv[2] if len(sys.argv) > 2 else ""}')
print('# per-launch times are cold-cache and serialised under ncu: compare shares, not absolutes')
print('kernel, launches, total_us, share_of_captured')
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{k}, {n}, {us:.1f}, {us / tot:.3f}')
st = [(n, us) for k, (n, us) in agg.items() if 'step_kernel' in k]
if st:
    print(f'# step_kernel launches: {st[0][0]}, mean {st[0][1] / st[0][0] / 1e3:.3f} ms; the bench\'s timed region contains step_kernel '
          f'launches only')
