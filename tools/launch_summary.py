#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total time, share.
    python tools/launch_summary.py gpurun_out/launches.csv "header comment" > profiles/launches_rN_summary.txt"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors='replace')) if len(r) > 5]
hdr = next(r for r in rows if 'Kernel Name' in r)
ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    if r is hdr or r[ik] == 'Kernel Name':
        continue
    try:
        v = float(r[iv].replace(',', ''))
    except ValueError:
        continue
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(r[iu], 1.0)
    k = r[ik].split('(')[0][:60]
    tot[k] += v
    cnt[k] += 1
print('# ' + (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print('# per-launch times are cold-cache and serialised under ncu: compare shares, not absolutes')
print('kernel, launches, total_us, share_of_captured, mean_us')
s = sum(tot.values())
for k, v in tot.most_common(12):
    print(f'{k}, {cnt[k]}, {v:.1f}, {v / s:.3f}, {v / cnt[k]:.1f}')
