set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; tail -c 3000 gpurun_out/bench_r01.json; tail -5 gpurun_out/bench_r01.err
