#!/bin/bash
# Run on the GPU box (under gpurun): plain bench -> ncu launch list -> ncu full capture of the step kernel.
# usage: tools/gpu_profile.sh <tag>
set -u
TAG=${1:-r1}
B="python bench.py --steps 20 --warmup 30 --no-cpu-baseline --no-e2e"
mkdir -p gpurun_out
$B > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 35 -c 1 -o gpurun_out/prof_$TAG $B > gpurun_out/ncu_full_$TAG.log 2>&1
tail -1 gpurun_out/plain_$TAG.log
tail -2 gpurun_out/ncu_full_$TAG.log
