#!/bin/bash
# N = 8: how far the host may run ahead of the GPU in the device-resident loop (VQ_BENCH_INFLIGHT steps)
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 30 --warmup 3 --no-sweep --no-hnsw --no-cpu --no-api --sustain 0; }
p=29700
for cap in 3 6 4; do
p=$((p+1))
VQ_BENCH_INFLIGHT=$cap run $p 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('inflight',$cap,'value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'par',d['parity']['mismatches'])"
done
