#!/bin/bash
# N = 2: VQ_BENCH_INFLIGHT 3 vs 6
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 30 --warmup 3 --no-sweep --no-hnsw --no-cpu --no-api --sustain 0; }
p=29800
for cap in 3 6; do
p=$((p+1))
VQ_BENCH_INFLIGHT=$cap run $p 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('inflight',$cap,'value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'par',d['parity']['mismatches'])"
done
