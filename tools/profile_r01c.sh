# ncu evidence, round 1, part c (one GPU): launch list of the pipelined bench step and the shard-exchange kernel.
set -x
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-hnsw --no-graph --no-sweep"
P="python tools/peer_probe.py --batch 1024 --k 10 --iters 5"
$B > gpurun_out/plain_bench_c.log 2>&1 && VQ_PEER_TIMEOUT_MS=500 $P > gpurun_out/plain_peer.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01c_launches_bench.csv $B > gpurun_out/ncu_bench_c.log 2>&1
VQ_PEER_TIMEOUT_MS=500 ncu --set full --clock-control none --import-source on -k regex:peer_exchange_merge -s 4 -c 1 -f -o gpurun_out/r01c_peer_exchange_b1024 $P > gpurun_out/ncu_peer.log 2>&1
ls -la gpurun_out/*.ncu-rep
