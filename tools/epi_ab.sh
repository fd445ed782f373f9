#!/bin/bash
# A/B of 4 vs 8 epilogue warps (VQ_EXACT_EPI) over batch sizes, 1M-row clip store.
B="--steps 30 --warmup 3 --no-hnsw --no-cpu --no-api --no-sweep"
for b in 1 32 64 128 256 1024; do
for epi in 4 8; do
VQ_EXACT_EPI=$epi timeout 200 python bench.py $B --batch $b 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('b',$b,'epi',$epi,'ms',round(d['ms_per_step'],4),'kern',round(d['roofline']['kernel_ms'],4),'par',d['parity']['mismatches'])"
done; done
