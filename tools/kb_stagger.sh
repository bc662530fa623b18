for v in base stg470 stg1175 stg40000; do K=tools/_kb/kbench_$v
timeout 60 $K 37888 2000 0 $v; timeout 60 $K 37888 2000 1 $v; timeout 60 $K 65536 500 1 $v; done
