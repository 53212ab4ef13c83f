"""Static opcode histogram of an address range of a kernel's SASS (no GPU needed).

    cuobjdump -sass -fun <mangled> kernels.cubin > k.sass
    python tools/sass_loop.py k.sass 0x760 0x2cd0
With no range: the widest backward-branch loop.
"""
import re
import sys
from collections import Counter

pat = re.compile(r'/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)(.*?);')
ins = []
for line in open(sys.argv[1]):
    m = pat.search(line)
    if m:
        ins.append((int(m.group(1), 16), m.group(3), m.group(4)))
if len(sys.argv) >= 4:
    lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
else:
    best = (0, 0, 0)
    for a, op, rest in ins:
        if op.startswith('BRA'):
            t = re.search(r'0x([0-9a-f]+)', rest)
            if t and int(t.group(1), 16) < a and a - int(t.group(1), 16) > best[0]:
                best = (a - int(t.group(1), 16), int(t.group(1), 16), a)
    lo, hi = best[1], best[2]
c = Counter(op.split('.')[0] for a, op, _ in ins if lo <= a <= hi)
tot = sum(c.values())
print(f'range {lo:#x}..{hi:#x}: {tot} instructions')
for op, n in c.most_common():
    print(f'  {op:10s} {n:5d}  {100.0 * n / tot:5.1f}%')
