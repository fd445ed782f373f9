#!/bin/bash
# 8-GPU session: config 2 (contract, with parity + breakdown + sharded HNSW), config 4 and config 5.
tag=${1:-r2}
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n "$@"; }
run 8 --steps 30 --warmup 3 > gpurun_out/${tag}_scale_n8.json 2> gpurun_out/${tag}_scale_n8.err; echo "n8 rc=$?"
run 8 --config 5 --steps 5 --warmup 3 --no-sweep > gpurun_out/${tag}_c5_n8.json 2> gpurun_out/${tag}_c5_n8.err; echo "c5 rc=$?"
run 8 --config 4 --steps 10 --warmup 3 --no-sweep > gpurun_out/${tag}_c4_n8.json 2> gpurun_out/${tag}_c4_n8.err; echo "c4 rc=$?"
for f in scale_n8 c5_n8 c4_n8; do echo "== $f"; tail -1 gpurun_out/${tag}_$f.err | cut -c1-300; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_$f.json"))
    print(round(d["value"]), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "parity", d["parity"]["mismatches"], d["parity"]["overflowed_queries"], "kern", d["roofline"] and round(d["roofline"]["kernel_ms"], 4), d.get("step_breakdown"), d.get("hnsw_sharded"))
except Exception as e:
    print("ERR", e)
PY
done
