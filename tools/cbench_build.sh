#!/bin/bash
# usage: tools/cbench_build.sh   (links tools/cbench.cu against the in-tree object files of libb200mp)
set -e
cd "$(dirname "$0")/.."
python -m python_motionplanning_b200.build > /dev/null
OUT=tools/_kb; mkdir -p $OUT
B=python_motionplanning_b200/_build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/cbench.cu $B/b200mp_api.o $B/collision_kernels.o $B/misc_kernels.o $B/rollout_kernels.o $B/tracking_kernels.o $B/lattice_kernels.o -o $OUT/cbench
