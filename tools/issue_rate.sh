#!/bin/bash
# MMA issue loop of CTA 0 (cycles per MMA, SM clock) printed by the kernel itself (VQ_MMA_DEBUG=128).
# +2: the producer skips the TMA loads (issue structure alone), +1: the issuers skip the MMAs (store stream alone)
# usage: issue_rate.sh "<dbg> <VQ_MMA_CG2>" ...
B="--steps 12 --warmup 3 --no-hnsw --no-cpu --no-api --no-sweep --sustain 0 --data ${DATA:-gauss}"
for cfg in "$@"; do
set -- $cfg
VQ_MMA_STAGES=${3:-12} VQ_MMA_DEBUG=$1 VQ_MMA_CG2=$2 timeout 60 python bench.py $B 2>&1 | grep "scan_mma dbg" > gpurun_out/issue_$1_$2.log
python - <<PY
import re,collections
d=collections.defaultdict(list)
for l in open('gpurun_out/issue_$1_$2.log'):
    m=re.search(r'(\d+) tiles, (\d+) cycles, (\d+) ns -> ([\d.]+) cycles/MMA, (\d+) MHz',l)
    if m: d[int(m.group(1))].append((float(m.group(4)),int(m.group(5)),int(m.group(3))))
for t,v in sorted(d.items()):
    if t > 100: print("dbg $1 cg2 $2 stages ${3:-12} tiles",t,'n',len(v),'cyc/MMA %.1f'%(sum(x[0] for x in v)/len(v)),'MHz %.0f'%(sum(x[1] for x in v)/len(v)),'us %.1f'%(sum(x[2] for x in v)/len(v)/1e3))
PY
done
