for v in base s64 b128 s128 s256 base; do K=tools/_kb/kbench_$v
timeout 40 $K 65536 500 1 $v; timeout 40 $K 65536 500 0 $v; timeout 60 $K 303104 500 1 $v; done
