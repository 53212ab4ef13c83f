#!/bin/bash
# ncu --set full capture of ONE step_kernel launch deep into an episode under a periodic action sequence (a gait: the
# ants have walked into the walls).  usage: bash tools/gpu_profile_gait.sh <env> <tag> [launch=445] [period=4]
ENVN=${1:-ant_heavenhell}; TAG=${2:-gait}; L=${3:-445}; P=${4:-4}
python tools/profile_step.py --env $ENVN --period $P --steps $((L + 5)) > gpurun_out/plain_${ENVN}_${TAG}.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:step_kernel --launch-skip $L --launch-count 1 \
    -o gpurun_out/prof_${ENVN}_${TAG} -f python tools/profile_step.py --env $ENVN --period $P --steps $((L + 5)) > gpurun_out/ncu_${ENVN}_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_${ENVN}_${TAG}.log
