#!/bin/bash
# GPU pass C (2 GPUs): whole parity suite incl. the two-device test, then the 2-rank bench line (weak + strong + e2e roof)
python -m pytest tests -m gpu -q -s 2>&1 | tail -80 > gpurun_out/r2_pytest_gpu_c.log
grep -E "passed|failed|FAILED|parity\]" gpurun_out/r2_pytest_gpu_c.log | tail -20
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
tail -c 1500 gpurun_out/r2_bench_2gpu.json; tail -3 gpurun_out/r2_bench_2gpu.err
