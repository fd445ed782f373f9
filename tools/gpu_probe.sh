#!/bin/bash
python tools/hnsw_recall_at_scale.py --n 100000 --kinds clip,gauss --select hybrid
python tools/hnsw_recall_at_scale.py --n 1000000 --kinds clip --select hybrid
python tools/hnsw_recall_at_scale.py --n 1000000 --kinds clip --select sequential
B="--steps 30 --warmup 3 --no-hnsw --no-cpu --no-api"
timeout 300 python bench.py $B 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('1M', round(d['value']), round(d['ms_per_step'],4), d['parity']['mismatches'], [(s['batch'], s['data_kind'], round(s['value']), round(s['ms_per_step'],4), round(s['roofline']['kernel_ms'],4)) for s in d['sweep']])"
