set -x
timeout 120 python tools/mma_check.py 2>&1 | grep -v "torch fp32" | tail -9
timeout 300 python tools/quick_bench.py --dtypes bf16 --paths mma,fma --batches 1,32,128,1024 2>&1 | tail -12
timeout 600 python -m pytest tests/test_gpu_hnsw.py -q -m gpu -s 2>&1 | grep -E "recall|passed|failed|Error|assert" | head -30
