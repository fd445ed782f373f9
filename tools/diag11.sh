timeout 600 python -m pytest tests/test_gpu_hnsw.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python tools/bench_hnsw.py --n 1000000 --kind clip 2>&1 | tail -1
VQ_HNSW_REGLIST=0 timeout 600 python tools/bench_hnsw.py --n 1000000 --kind clip 2>&1 | tail -1
