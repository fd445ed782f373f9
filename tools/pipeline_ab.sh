#!/bin/bash
# steps in flight (--pipeline) vs value and e2e, N = 1
for p in 3 4 6; do
timeout 100 python bench.py --steps 30 --warmup 5 --no-hnsw --no-cpu --no-api --no-sweep --sustain 0 --pipeline $p 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('pipeline',$p,'value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),d['e2e'].get('host_thread_rank0'),'par',d['parity']['mismatches'])"
done
