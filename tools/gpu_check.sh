#!/bin/bash
# Developer helper run under gpurun: exact-path GPU tests, the contract bench at 1M, 125k / 250k-row shards (the per-GPU
# shapes of 8- and 4-way sharded config 2) and config 4's per-GPU shape.
tag=${1:-x}
timeout 900 python -m pytest tests/test_gpu_exact.py tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py -x -q -k "not config4 and not config5" > gpurun_out/${tag}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_tests.log; tail -4 gpurun_out/${tag}_tests.log
B="--steps 30 --warmup 3 --no-hnsw --no-cpu --no-api"
timeout 300 python bench.py $B > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${tag}_bench.err
for rows in 125000 250000; do
  timeout 200 python bench.py $B --no-sweep --rows $rows > gpurun_out/${tag}_bench_$rows.json 2> gpurun_out/${tag}_bench_$rows.err; echo "bench $rows rc=$?"
  VQ_EXACT_REFRESH_NS=0 timeout 200 python bench.py $B --no-sweep --rows $rows > gpurun_out/${tag}_bench_${rows}_norefresh.json 2> gpurun_out/${tag}_bench_${rows}_norefresh.err; echo "bench $rows norefresh rc=$?"
done
timeout 200 python bench.py $B --no-sweep --rows 125000 --data gauss > gpurun_out/${tag}_bench_125000_gauss.json 2> gpurun_out/${tag}_bench_125000_gauss.err; echo "bench125 gauss rc=$?"
timeout 300 python bench.py --config 4 --rows 1250000 --steps 10 --warmup 3 --no-hnsw --no-cpu --no-api --no-sweep > gpurun_out/${tag}_bench_c4shard.json 2> gpurun_out/${tag}_bench_c4shard.err; echo "bench c4 shard rc=$?"; tail -2 gpurun_out/${tag}_bench_c4shard.err
timeout 600 python tools/hnsw_hybrid_1m.py diverse > gpurun_out/${tag}_hnsw_hybrid.log 2>&1; cat gpurun_out/${tag}_hnsw_hybrid.log | tail -6
