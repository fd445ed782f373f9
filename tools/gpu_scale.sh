#!/bin/bash
# Multi-GPU session (run under `gpurun --gpus G`): the contract bench at N = G .. 2 (config 2, with the N > 1 parity
# check and the sharded HNSW leg), then BASELINE configs 4 and 5 at N = G.  One JSON line per run under gpurun_out/.
tag=${1:-r2}; G=${2:-8}
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n "$@"; }
names=""
for n in 8 4 2; do
  if [ $n -le $G ]; then
    extra="--no-sweep"; [ $n -eq $G ] && extra=""
    run $n --steps 30 --warmup 3 $extra > gpurun_out/${tag}_scale_n$n.json 2> gpurun_out/${tag}_scale_n$n.err; echo "n$n rc=$?"; names="$names scale_n$n"
  fi
done
timeout 300 python bench.py --steps 30 --warmup 3 --no-sweep --no-hnsw --no-cpu --no-api > gpurun_out/${tag}_scale_n1.json 2> gpurun_out/${tag}_scale_n1.err; echo "n1 rc=$?"
run $G --config 4 --steps 10 --warmup 3 --no-sweep > gpurun_out/${tag}_c4_n$G.json 2> gpurun_out/${tag}_c4_n$G.err; echo "c4 rc=$?"
run $G --config 5 --steps 5 --warmup 3 --no-sweep > gpurun_out/${tag}_c5_n$G.json 2> gpurun_out/${tag}_c5_n$G.err; echo "c5 rc=$?"
for f in $names scale_n1 c4_n$G c5_n$G; do echo "== $f"; tail -2 gpurun_out/${tag}_$f.err | cut -c1-300; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_$f.json"))
    print(round(d["value"]), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "parity", d["parity"]["mismatches"], d["parity"]["overflowed_queries"], "kern", d["roofline"] and round(d["roofline"]["kernel_ms"], 4), d["config"]["scan_path"], d["config"]["launch"], d.get("hnsw_sharded"))
except Exception as e:
    print("ERR", e)
PY
done
