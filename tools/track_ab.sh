L=python_motionplanning_b200/libb200mp.so
cp $L /tmp/lib_base.so
for v in base "$@"; do
  if [ $v == base ]; then cp /tmp/lib_base.so $L; else cp tools/_kb/libb200mp_$v.so $L; fi
  python tools/track_bench4.py $v
done
cp /tmp/lib_base.so $L
