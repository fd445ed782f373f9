#!/bin/bash
# A/B: ring depth of the bootstrap sample pass (VQ_BOOT_RING groups; 0 = the scan's full ring)
B="--steps 30 --warmup 3 --no-hnsw --no-cpu --no-api --no-sweep"
for rows in 125000 250000 1000000; do
for ring in 0 2 1; do
VQ_BOOT_RING=$ring timeout 100 python bench.py $B --rows $rows 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('rows',$rows,'boot ring',$ring,'ms',round(d['ms_per_step'],4),'kern',round(d['roofline']['kernel_ms'],4),'par',d['parity']['mismatches'],d['parity']['overflowed_queries'])"
done; done
