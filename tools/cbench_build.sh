#!/bin/bash
# usage: tools/cbench_build.sh [NAME [extra -D flags for collision_kernels.cu]]
# Links tools/cbench.cu against the in-tree object files of libb200mp, with collision_kernels.cu recompiled under the flags.
set -e
cd "$(dirname "$0")/.."
NAME=${1:-cur}; shift || true; FLAGS="$*"
python -m python_motionplanning_b200.build > /dev/null
OUT=tools/_kb; mkdir -p $OUT
C=python_motionplanning_b200/csrc
B=python_motionplanning_b200/_build
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr"
nvcc $F -DB200MP_DEV_TUNABLES=1 $FLAGS -Xptxas -v -c $C/collision_kernels.cu -o $OUT/collision_$NAME.o 2> $OUT/ptxas_collision_$NAME.log
nvcc $F tools/cbench.cu $B/b200mp_api.o $OUT/collision_$NAME.o $B/misc_kernels.o $B/rollout_kernels_f64.o $B/rollout_kernels_f32.o $B/tracking_kernels.o $B/lattice_kernels.o -o $OUT/cbench_$NAME
