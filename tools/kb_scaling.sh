K=tools/_kb/kbench_base
for B in 37888 65536; do for N in 50 100 250 500 1000 2000; do for S in 1 0; do timeout 60 $K $B $N $S scale; done; done; done
timeout 120 $K 303104 500 1 scale8
timeout 120 $K 303104 500 0 scale8
timeout 60 $K 18944 500 1 half
timeout 60 $K 18944 2000 0 half
