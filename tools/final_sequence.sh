# the driver's round-end sequence plus the ncu captures the profiles/ summaries are made from (development helper)
T=${1:-r2s2}
timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --impl reference > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err; echo "ref rc=$?"
timeout 400 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:rk4_rollout_kernel -s 1 -c 1 -f -o gpurun_out/${T}_prof_rollout python tools/prof_target.py rollout --launches 2 > gpurun_out/${T}_ncu_rollout.log 2>&1; echo "ncu rollout rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:track_kernel -s 1 -c 1 -f -o gpurun_out/${T}_prof_track python tools/prof_target.py track --launches 2 --steps 500 > gpurun_out/${T}_ncu_track.log 2>&1; echo "ncu track rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_bench_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${T}_ncu_bench.log 2>&1; echo "ncu bench rc=$?"
