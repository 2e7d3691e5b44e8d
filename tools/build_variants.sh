#!/bin/bash
# Tuning aid: build libah_b200 variants (block size x blocks/SM) into air_hockey_training_environment_b200/variants/
# usage: tools/build_variants.sh "256:4 256:5 256:6 128:8 128:10 128:12"
set -e
cd "$(dirname "$0")/../air_hockey_training_environment_b200"; mkdir -p variants
for v in $1; do
  th=${v%%:*}; occ=${v##*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -shared -Xcompiler -fPIC -cudart static \
       -DAH_PACKED_THREADS=$th -DAH_PACKED_OCC_F32=$occ $EXTRA -Xptxas -v -o variants/libah_t${th}_o${occ}${TAG}.so csrc/ah_b200.cu 2> variants/ptxas_t${th}_o${occ}${TAG}.txt &
done
wait
grep -A2 "step_packed_kernelIfE" variants/ptxas_*.txt | grep -E "spill|Used" 
