#!/bin/bash
# A/B: no bootstrap sample pass at all (VQ_EXACT_BOOT=0: the cooperative bound of the scan itself warms up) vs the default
B="--steps 30 --warmup 3 --no-hnsw --no-cpu --no-api --no-sweep"
for rows in 125000 250000 1000000; do
for cfg in "-1 4000" "0 4000" "0 2000" "0 1000"; do
set -- $cfg
VQ_EXACT_BOOT=$1 VQ_EXACT_REFRESH_NS=$2 timeout 100 python bench.py $B --rows $rows 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('rows',$rows,'boot',$1,'refresh',$2,'ms',round(d['ms_per_step'],4),'kern',round(d['roofline']['kernel_ms'],4),'gath',round(d['exact_search']['rows_gathered_per_query']),'par',d['parity']['mismatches'],d['parity']['overflowed_queries'])"
done; done
