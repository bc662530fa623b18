# usage: tools/kb_sweep.sh v1 v2 ...   (config 2 with the trajectory, then eight waves; base first and last)
for v in base "$@" base; do K=tools/_kb/kbench_$v
timeout 40 $K 65536 500 1 $v | sed 's/checksum.*err/err/'; timeout 60 $K 303104 500 1 $v | sed 's/checksum.*err/err/'; done
