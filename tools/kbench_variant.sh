#!/bin/bash
# usage: tools/kbench_variant.sh NAME "<extra flags for rollout_kernels_f64.cu>"   (light: only the FP64 rollout TU is rebuilt;
# run tools/kbench_build.sh base "" first for the other objects)
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift; FLAGS="$*"
OUT=tools/_kb; C=python_motionplanning_b200/csrc; B=python_motionplanning_b200/_build
ARCH=${KB_ARCH:-"-gencode arch=compute_100a,code=sm_100a"}
F="$ARCH -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr"
FL="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr"
nvcc $F $FLAGS -Xptxas -v -c $C/rollout_kernels_f64.cu -o $OUT/rollout_$NAME.o 2> $OUT/ptxas_$NAME.log
nvcc $FL tools/kbench.cu $OUT/api_base.o $B/collision_kernels.o $B/misc_kernels.o $OUT/tracking_base.o $B/lattice_kernels.o $B/rollout_kernels_f32.o $OUT/rollout_$NAME.o -o $OUT/kbench_$NAME
echo "$NAME: $(grep -A2 'rk4_rollout_kernelIdLb1ELb0ELb0ELb1ELb1ELb0ELb0E' $OUT/ptxas_$NAME.log | grep -oE 'Used [0-9]+ registers|[0-9]+ bytes spill stores' | tr '\n' ' ')"
