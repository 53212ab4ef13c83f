#!/usr/bin/env python
"""Attribute ncu per-SASS-instruction counts to CUDA source lines.

The ncu source page (SASS view) has executed-instruction counts but, read here without the GPU box's
paths, no source mapping; nvdisasm --print-line-info on the in-tree cubin has the mapping. Both list the
kernel's instructions in the same order, so they are joined by instruction offset.

    python tools/ncu_attrib.py gpurun_out/prof.ncu-rep step_kernelILi1 [--top 40]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_lines(kernel_substr):
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.join(ROOT, 'po_brax_b200', 'libpobrax.so')], cwd=tmp,
                   capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.startswith('kernels') and 'api' not in f][0]
    out = subprocess.run(['nvdisasm', '--print-line-info', '-c', os.path.join(tmp, cubin)], capture_output=True,
                         text=True).stdout.splitlines()
    res, on, cur = {}, False, ('?', 0)
    for ln in out:
        if ln.startswith('.text.'):
            on = kernel_substr in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r'\s*/\*([0-9a-f]+)\*/\s+(.*?);', ln)
        if m:
            res[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return res


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 40
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    # first kernel section only
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
    sec = rows[starts[0] + 1:(starts[1] if len(starts) > 1 else len(rows))]
    hdr = sec[0]
    ia, ii, isamp = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    base = int(sec[1][ia], 16)
    lines = sass_lines(kern)
    per_line, per_file_samples = collections.Counter(), collections.Counter()
    samples = collections.Counter()
    nwarps = int(sec[1][ii])
    tot = 0
    for r in sec[1:]:
        off = int(r[ia], 16) - base
        n, s = int(r[ii]), int(r[isamp])
        key = lines.get(off, (('?', 0), ''))[0]
        per_line[key] += n
        samples[key] += s
        tot += n
    print(f'warps {nwarps}; instructions per warp {tot / nwarps:.0f}; samples {sum(samples.values())}')
    src_cache = {}

    def text(f, l):
        if f not in src_cache:
            for d in ('po_brax_b200/csrc', 'include'):
                p = os.path.join(ROOT, d, f)
                if os.path.exists(p):
                    src_cache[f] = open(p).read().splitlines()
                    break
            else:
                src_cache[f] = []
        s = src_cache[f]
        return s[l - 1].strip()[:90] if 0 < l <= len(s) else ''
    print(f'{"inst/warp":>10} {"%":>5} {"stall%":>6}  location')
    ts = sum(samples.values()) or 1
    for (f, l), n in per_line.most_common(top):
        print(f'{n / nwarps:10.1f} {100 * n / tot:5.1f} {100 * samples[(f, l)] / ts:6.1f}  {f}:{l}  {text(f, l)}')
    byfile = collections.Counter()
    for (f, l), n in per_line.items():
        byfile[f] += n
    print({f: round(n / nwarps) for f, n in byfile.items()})


if __name__ == '__main__':
    main()
