for v in base st1 st2 base st1; do K=tools/_kb/kbench_$v
timeout 60 $K 65536 500 1 $v; timeout 60 $K 303104 500 1 $v; done
