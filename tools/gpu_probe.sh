#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_exact.py -x -q 2>&1 | tail -3
B="--steps 30 --warmup 3 --no-hnsw --no-cpu --no-api"
timeout 300 python bench.py $B 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('1M', round(d['value']), round(d['ms_per_step'],4), 'kern', round(d['roofline']['kernel_ms'],4), d['parity']['mismatches'], [(s['batch'], s['data_kind'], round(s['value']), round(s['ms_per_step'],4), round(s['roofline']['kernel_ms'],4), s['exact_search']['rows_gathered_per_query']) for s in d['sweep']])"
for rows in 125000 250000; do
timeout 200 python bench.py $B --no-sweep --rows $rows 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('rows',$rows,'qps',round(d['value']),'ms',round(d['ms_per_step'],4),'kern',round(d['roofline']['kernel_ms'],4),'gath',round(d['exact_search']['rows_gathered_per_query']),'par',d['parity']['mismatches'], 'traffic', d['roofline']['traffic'])"
done
