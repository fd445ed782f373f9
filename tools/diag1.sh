set -x
for dbg in 0 1 2 4 5 6; do
  echo "== VQ_MMA_DEBUG=$dbg"
  VQ_MMA_DEBUG=$dbg timeout 120 python tools/quick_bench.py --dtypes bf16 --paths mma --batches 1,32,1024 --k 10 --iters 10 2>&1 | tail -3
done
echo "== k=32"
timeout 120 python tools/quick_bench.py --dtypes bf16 --paths mma --batches 1,32,128,256,1024 --k 32 --iters 10 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_hnsw.py -q -m gpu -s -k recall_bar 2>&1 | grep -E "recall|passed|failed|Error|assert" | head -30
