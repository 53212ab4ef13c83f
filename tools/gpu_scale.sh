#!/bin/bash
# bench.py at N ranks (torchrun), the way the driver launches it.  usage: bash tools/gpu_scale.sh N [extra bench args]
N=$1; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
tail -c 2200 gpurun_out/r2_bench_${N}gpu.json; tail -2 gpurun_out/r2_bench_${N}gpu.err
