#!/bin/bash
# usage: tools/kbench_build.sh NAME "<extra -D flags for rollout_kernels.cu>"
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
OUT=tools/_kb; mkdir -p $OUT
C=python_motionplanning_b200/csrc
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr"
for f in b200mp_api collision_kernels misc_kernels; do [ -f $OUT/$f.o ] || nvcc $F -c $C/$f.cu -o $OUT/$f.o; done
nvcc $F $@ -Xptxas -v -c $C/rollout_kernels.cu -o $OUT/rollout_$NAME.o 2> $OUT/ptxas_$NAME.log
grep -A2 "rk4_rollout_kernelIdLb1ELb0ELb0" $OUT/ptxas_$NAME.log | grep -E "registers|spill" | tr '\n' ' '; echo
nvcc $F tools/kbench.cu $OUT/b200mp_api.o $OUT/collision_kernels.o $OUT/misc_kernels.o $OUT/rollout_$NAME.o -o $OUT/kbench_$NAME
