#!/bin/bash
B="--steps 30 --warmup 3 --no-hnsw --no-cpu --no-api"
timeout 300 python bench.py $B 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('1M', round(d['value']), round(d['ms_per_step'],4), 'kern', round(d['roofline']['kernel_ms'],4), d['parity']['mismatches'], [(s['batch'], s['data_kind'], round(s['value']), round(s['ms_per_step'],4), round(s['roofline']['kernel_ms'],4), s['exact_search']['rows_gathered_per_query']) for s in d['sweep']])"
for rows in 125000 250000 500000; do
timeout 200 python bench.py $B --no-sweep --rows $rows 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('rows',$rows,'qps',round(d['value']),'ms',round(d['ms_per_step'],4),'kern',round(d['roofline']['kernel_ms'],4),'gath',round(d['exact_search']['rows_gathered_per_query']),'par',d['parity']['mismatches'])"
done
timeout 300 python bench.py --config 5 --rows 12500000 --steps 5 --warmup 3 --no-hnsw --no-cpu --no-api --no-sweep 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('c5 shard b=4096', round(d['value']), round(d['ms_per_step'],4), 'parity', d['parity']['mismatches'], 'gathered', d['exact_search'], 'overflow', d['overflowed_queries_per_batch'], 'kern', d['roofline']['kernel_ms'])" | cut -c1-700
