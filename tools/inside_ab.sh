#!/bin/bash
# A/B of the threshold bootstrap placement (separate sample pass vs inside the scan) on shard shapes.
B="--steps 30 --warmup 3 --no-hnsw --no-cpu --no-api --no-sweep"
for rows in 125000 250000 1000000; do
for inside in 0 1; do
VQ_EXACT_INSIDE=$inside timeout 100 python bench.py $B --rows $rows 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('rows',$rows,'inside',$inside,'ms',round(d['ms_per_step'],4),'kern',round(d['roofline']['kernel_ms'],4),'gath',round(d['exact_search']['rows_gathered_per_query']),'par',d['parity']['mismatches'],d['parity']['overflowed_queries'])"
done; done
timeout 300 python bench.py --config 4 --rows 1250000 --steps 10 --warmup 3 --no-hnsw --no-cpu --no-api --no-sweep 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('c4 shard', round(d['value']), round(d['ms_per_step'],4), d['parity']['mismatches'])"
