#!/bin/bash
# GPU pass A of round 2: parity suite, bench line, episode-phase timings, sanitizer logs (outputs under gpurun_out/).
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 > gpurun_out/r2_pytest_gpu.log
python bench.py > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
python tools/bench_phases.py > gpurun_out/r2_phases_a.log 2>&1
timeout 400 compute-sanitizer --tool memcheck --log-file gpurun_out/r2_memcheck.log python tools/sanitize_small.py > gpurun_out/r2_memcheck.out 2>&1
timeout 400 compute-sanitizer --tool racecheck --log-file gpurun_out/r2_racecheck.log python tools/sanitize_small.py > gpurun_out/r2_racecheck.out 2>&1
tail -5 gpurun_out/r2_pytest_gpu.log; cat gpurun_out/r2_phases_a.log; tail -3 gpurun_out/r2_memcheck.log gpurun_out/r2_racecheck.log
