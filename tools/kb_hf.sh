for v in base hf base hf; do K=tools/_kb/kbench_$v
timeout 40 $K 65536 500 1 $v; timeout 60 $K 303104 500 1 $v; done
