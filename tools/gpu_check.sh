#!/bin/bash
# Developer helper run under gpurun: exact-path GPU tests, the contract bench at 1M, a 125k-row shard (the per-GPU
# shape of an 8-way sharded config 2) and config 4's per-GPU shape.
tag=${1:-x}
timeout 900 python -m pytest tests/test_gpu_exact.py tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py -x -q -k "not config4 and not config5" > gpurun_out/${tag}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_tests.log; tail -4 gpurun_out/${tag}_tests.log
B="--steps 30 --warmup 3 --no-hnsw --no-cpu --no-api"
timeout 300 python bench.py $B > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${tag}_bench.err
timeout 200 python bench.py $B --no-sweep --rows 125000 > gpurun_out/${tag}_bench_125k.json 2> gpurun_out/${tag}_bench_125k.err; echo "bench125 rc=$?"
timeout 200 python bench.py $B --no-sweep --rows 125000 --data gauss > gpurun_out/${tag}_bench_125k_gauss.json 2> gpurun_out/${tag}_bench_125k_gauss.err; echo "bench125 gauss rc=$?"
timeout 300 python bench.py --config 4 --rows 1250000 --steps 10 --warmup 3 --no-hnsw --no-cpu --no-api --no-sweep > gpurun_out/${tag}_bench_c4shard.json 2> gpurun_out/${tag}_bench_c4shard.err; echo "bench c4 shard rc=$?"; tail -2 gpurun_out/${tag}_bench_c4shard.err
