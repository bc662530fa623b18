#!/bin/bash
# usage: tools/libvariant_build.sh NAME SOURCE.cu "<extra -D flags>"
# Builds tools/_kb/libb200mp_NAME.so: the in-tree library with ONE translation unit recompiled under extra flags
# (development A/B of compile-time tunables on the GPU box: copy the variant over python_motionplanning_b200/libb200mp.so there).
set -e
cd "$(dirname "$0")/.."
NAME=$1; SRC=$2; shift; shift; FLAGS="$*"
python -m python_motionplanning_b200.build > /dev/null
OUT=tools/_kb; mkdir -p $OUT
C=python_motionplanning_b200/csrc
B=python_motionplanning_b200/_build
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC"
nvcc $F $FLAGS -c $C/$SRC -o $OUT/${SRC%.cu}_$NAME.o
OBJS=""
for o in $B/*.o; do
  if [ "$(basename $o)" == "${SRC%.cu}.o" ]; then OBJS="$OBJS $OUT/${SRC%.cu}_$NAME.o"; else OBJS="$OBJS $o"; fi
done
nvcc -shared -cudart shared -o $OUT/libb200mp_$NAME.so $OBJS -Xlinker -rpath=/usr/local/cuda/lib64 -Xlinker --no-undefined
echo $OUT/libb200mp_$NAME.so
