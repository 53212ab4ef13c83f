#!/bin/bash
# ncu --set full capture of ONE step_kernel launch in bench.py's stationary regime (launch #1005 after the reset),
# after the same command has run without ncu.  usage: bash tools/gpu_profile.sh <env> <tag>
ENVN=${1:-ant_heavenhell}; TAG=${2:-r2}
python tools/profile_step.py --env $ENVN --stationary --steps 1010 > gpurun_out/plain_${ENVN}_${TAG}.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:step_kernel --launch-skip 1005 --launch-count 1 \
    -o gpurun_out/prof_${ENVN}_${TAG} -f python tools/profile_step.py --env $ENVN --stationary --steps 1010 > gpurun_out/ncu_${ENVN}_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_${ENVN}_${TAG}.log
