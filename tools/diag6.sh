B="python bench.py --steps 5 --warmup 3 --no-cpu --no-sweep --no-graph"
$B > gpurun_out/plain_b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_b.csv $B > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_b.log | cut -c1-300
