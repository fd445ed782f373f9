#!/bin/bash
# Multi-GPU session (run under `gpurun --gpus 8`): the contract bench at N = 8 / 4 / 2 (config 2, with the N > 1 parity
# check and the sharded HNSW leg), then BASELINE configs 4 and 5 at N = 8.  One JSON line per run under gpurun_out/.
tag=${1:-r2}
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n "$@"; }
run 8 --steps 30 --warmup 3 > gpurun_out/${tag}_scale_n8.json 2> gpurun_out/${tag}_scale_n8.err; echo "n8 rc=$?"
run 4 --steps 30 --warmup 3 --no-sweep > gpurun_out/${tag}_scale_n4.json 2> gpurun_out/${tag}_scale_n4.err; echo "n4 rc=$?"
run 2 --steps 30 --warmup 3 --no-sweep > gpurun_out/${tag}_scale_n2.json 2> gpurun_out/${tag}_scale_n2.err; echo "n2 rc=$?"
python bench.py --steps 30 --warmup 3 --no-sweep --no-hnsw --no-cpu --no-api > gpurun_out/${tag}_scale_n1.json 2> gpurun_out/${tag}_scale_n1.err; echo "n1 rc=$?"
run 8 --config 4 --steps 10 --warmup 3 --no-sweep > gpurun_out/${tag}_c4_n8.json 2> gpurun_out/${tag}_c4_n8.err; echo "c4 n8 rc=$?"
run 8 --config 5 --steps 5 --warmup 3 --no-sweep > gpurun_out/${tag}_c5_n8.json 2> gpurun_out/${tag}_c5_n8.err; echo "c5 n8 rc=$?"
for f in scale_n8 scale_n4 scale_n2 scale_n1 c4_n8 c5_n8; do echo "== $f"; tail -2 gpurun_out/${tag}_$f.err | cut -c1-300; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_$f.json"))
    print(round(d["value"]), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "parity", d["parity"]["mismatches"], d["parity"]["overflowed_queries"], "kern", d["roofline"] and round(d["roofline"]["kernel_ms"], 4), d["config"]["scan_path"], d["config"]["launch"], d.get("hnsw_sharded"))
except Exception as e:
    print("ERR", e)
PY
done
