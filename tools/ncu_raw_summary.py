#!/usr/bin/env python
"""The metrics of an ncu --set full report that the design discussion uses (DRAM bytes, pipe utilisation, issue
activity, stall reasons, launch configuration), one per line.

    python tools/ncu_raw_summary.py gpurun_out/prof.ncu-rep "header comment" > profiles/ncu_step_xxx.txt
"""
import csv
import subprocess
import sys

KEEP = ('dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct', 'gpu__time_duration.sum',
        'launch__block_size', 'launch__grid_size', 'launch__occupancy_limit', 'launch__registers_per_thread',
        'launch__shared_mem_per_block', 'launch__waves_per_multiprocessor', 'sm__inst_executed_pipe_',
        'sm__warps_active.avg.pct', 'smsp__issue_active.avg.pct', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__average_warps_issue_stalled_',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'smsp__inst_executed.avg.per_cycle_active')
rep = sys.argv[1]
if len(sys.argv) > 2:
    print('#', sys.argv[2])
txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
for h, u, v in sorted(zip(rows[0], rows[1], rows[2])):
    if any(h.startswith(k) for k in KEEP) and '_not_issued' not in h:
        print(h, u, v)
