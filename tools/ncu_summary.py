#!/usr/bin/env python
"""Selected raw metrics of the first kernel of an ncu --set full report, one per line (the format of profiles/ncu_step_*.txt).
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "header comment" > profiles/ncu_step_x_r2.txt"""
import re
import subprocess
import sys

KEEP = re.compile(r'^(dram__bytes_(read|write)\.sum|gpu__dram_throughput|gpu__time_duration\.sum|l1tex__t_sector_hit_rate|launch__|'
                  r'lts__t_sector_hit_rate|sm__cycles_elapsed\.max|sm__inst_executed_pipe_\w+\.avg|sm__pipe_\w+_cycles_active\.avg\.pct_of_peak_sustained_active|'
                  r'sm__throughput|sm__warps_active|smsp__average_warps_issue_stalled|smsp__cycles_active\.avg|smsp__inst_executed\.sum |'
                  r'smsp__issue_active|smsp__thread_inst_executed_per_inst_executed|smsp__inst_executed\.sum$)')
rep = sys.argv[1]
print('# ' + (sys.argv[2] if len(sys.argv) > 2 else rep))
txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw'], capture_output=True, text=True).stdout
seen = set()
for ln in txt.splitlines():
    f = ln.split()
    if len(f) >= 2 and KEEP.match(f[0]) and f[0] not in seen:
        seen.add(f[0])
        print(' '.join(f))
