timeout 900 python -m pytest tests -x -q -m gpu -s 2>&1 | grep -E "recall|passed|failed|Error" | tail -12
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu 2>&1 | tail -c 2600
timeout 600 python tools/bench_hnsw.py --n 1000000 --kind clip 2>&1 | tail -2
