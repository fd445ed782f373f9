# First measurements of the next round, cheapest first.  Run under gpurun.
set -x
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
# 1. (8 GPUs) does the two-per-lane queue bound remove the 3.8-5.1 M spread?  three runs each
for cap in 6 0x; do
  for i in 1 2 3; do
    VQ_BENCH_INFLIGHT=${cap/0x/100000} $T --nproc-per-node 8 --master-port 2960$i bench.py --gpus 8 --steps 400 --warmup 5 --no-sweep \
      > gpurun_out/n8_cap${cap}_$i.json 2> gpurun_out/n8_cap${cap}_$i.err
  done
done
# 2. (1 GPU) the N = 8 shard alone: floor of the step with 3 lanes (scan + bootstrap)
python bench.py --rows 125000 --steps 400 --warmup 5 --no-hnsw --no-cpu > gpurun_out/r125_n1.json 2> gpurun_out/r125_n1.err
# 3. (1 GPU) ncu of the CURRENT exchange kernels (line protocol), one simulated rank
VQ_PEER_TIMEOUT_MS=500 ncu --set full --clock-control none --import-source on -k regex:peer_exchange_merge -s 4 -c 1 -f \
  -o gpurun_out/r02_peer_exchange_b1024 python tools/peer_probe.py --batch 1024 --k 10 --iters 5 > gpurun_out/ncu_peer2.log 2>&1
