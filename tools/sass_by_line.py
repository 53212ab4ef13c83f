"""Static instruction count of one kernel by source line (nvdisasm -g on the cubin inside a built library; no GPU needed):
where a kernel's instruction-cache footprint comes from.
    python tools/sass_by_line.py po_brax_b200/libpobrax.so 'step_kernelILi2ELb0' [bucket]"""
import os, re, subprocess, sys, tempfile
from collections import Counter
lib, pat = os.path.abspath(sys.argv[1]), sys.argv[2]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 10
d = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', lib], cwd=d, check=True, capture_output=True)
cnt, cur, on = Counter(), None, False
for f in sorted(os.listdir(d)):
    if not f.startswith('kernels.'): continue
    for l in subprocess.run(['nvdisasm', '-g', os.path.join(d, f)], capture_output=True, text=True).stdout.splitlines():
        if l.startswith('\t.section'):
            on = ('.text.' in l) and (pat in l)
            continue
        if not on: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)) // bucket * bucket); continue
        if cur and re.search(r'/\*[0-9a-f]{4,5}\*/\s', l): cnt[cur] += 1
print('instructions', sum(cnt.values()))
for k, n in sorted(cnt.items(), key=lambda x: -x[1])[:40]: print(f'{k[0]:18s} line {k[1]:4d}+  {n}')
