#!/bin/bash
# usage: tools/kbench_build.sh NAME "<extra -D flags for rollout_kernels_f64.cu>"
# Links tools/kbench.cu against the library's in-tree objects, with rollout_kernels_f64.cu recompiled under the extra flags.
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift; FLAGS="$*"
python -m python_motionplanning_b200.build > /dev/null
OUT=tools/_kb; mkdir -p $OUT
C=python_motionplanning_b200/csrc
B=python_motionplanning_b200/_build
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr"
nvcc $F $FLAGS -Xptxas -v -c $C/rollout_kernels_f64.cu -o $OUT/rollout_$NAME.o 2> $OUT/ptxas_$NAME.log &
# the friction table layout lives in vehicle_rhs.cuh: its builder (b200mp_api.cu) and the other consumer (tracking) follow the flags
nvcc $F $FLAGS -Xcompiler -fPIC -c $C/b200mp_api.cu -o $OUT/api_$NAME.o &
nvcc $F $FLAGS -c $C/tracking_kernels.cu -o $OUT/tracking_$NAME.o &
wait
grep -A2 "rk4_rollout_kernelIdLb1ELb0ELb0E" $OUT/ptxas_$NAME.log | grep -E "registers|spill" | tr '\n' ' '; echo
nvcc $F $FLAGS tools/kbench.cu $OUT/api_$NAME.o $B/collision_kernels.o $B/misc_kernels.o $OUT/tracking_$NAME.o $B/lattice_kernels.o $B/rollout_kernels_f32.o $OUT/rollout_$NAME.o -o $OUT/kbench_$NAME
