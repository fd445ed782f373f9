#!/bin/bash
B="--steps 30 --warmup 3 --no-hnsw --no-cpu --no-api"
timeout 300 python bench.py $B 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('1M', round(d['value']), round(d['ms_per_step'],4), d['parity']['mismatches'], [(s['batch'], s['data_kind'], round(s['value']), round(s['ms_per_step'],4), round(s['roofline']['kernel_ms'],4), s['exact_search']) for s in d['sweep']])"
for b in 1 64 4096; do
timeout 300 python bench.py --config 5 --rows 12500000 --batch $b --steps 5 --warmup 3 --no-hnsw --no-cpu --no-api --no-sweep 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('c5 shard b=$b', round(d['value']), round(d['ms_per_step'],4), 'parity', d['parity'], 'gathered', d['exact_search'], 'overflow', d['overflowed_queries_per_batch'])" | cut -c1-700
done
