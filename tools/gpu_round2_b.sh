#!/bin/bash
# GPU pass B of round 2: the whole parity suite (no -x), episode-phase timings over a full episode length
python -m pytest tests -m gpu -q -s 2>&1 | tail -150 > gpurun_out/r2_pytest_gpu_b.log
python tools/bench_phases.py ant_heavenhell ant_tag > gpurun_out/r2_phases_b.log 2>&1
grep -E "passed|failed|FAILED|parity\]" gpurun_out/r2_pytest_gpu_b.log | tail -40; cat gpurun_out/r2_phases_b.log
