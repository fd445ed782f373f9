#!/bin/bash
# Developer build with -DVQ_SCAN_TRACE: clock stamps of CTA 0's pipeline events (see scan_mma.cu, TR()).
# slots: issuer 0 before/after tmem_empty wait, 2 after turn, 3/4 first group ready / issued, 5/6 last group; epilogue warp 4:
# 8 before / 9 after tmem_full wait, 10 release; epilogue warp 8: 12/13/14
B="--steps 4 --warmup 3 --no-hnsw --no-cpu --no-api --no-sweep --sustain 0 --data gauss"
for cfg in "$@"; do
set -- $cfg
echo "== VQ_MMA_DEBUG=$1 VQ_MMA_CG2=$2"
VQ_MMA_DEBUG=$1 VQ_MMA_CG2=$2 timeout 60 python bench.py $B 2>&1 | grep "scan trace" | tail -4
done
