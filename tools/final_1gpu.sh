#!/bin/bash
# End-of-round record on one GPU: the full GPU test suite, the driver's command line (both arms) and the default run.
tag=${1:-final}
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/${tag}_gputests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_gputests.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "reference rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_driver.json 2> gpurun_out/${tag}_bench_driver.err; echo "driver-like rc=$?"
timeout 600 python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err; echo "default rc=$?"
timeout 300 python bench.py --config 4 --rows 1250000 --steps 10 --warmup 3 --no-hnsw --no-cpu --no-api --no-sweep > gpurun_out/${tag}_c4shard.json 2>/dev/null; echo "c4 shard rc=$?"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
