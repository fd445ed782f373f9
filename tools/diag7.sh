timeout 120 python tools/mma_check.py 2>&1 | grep -v "torch fp32" | tail -8
timeout 900 python -m pytest tests/test_gpu_exact.py -x -q -m gpu 2>&1 | tail -3
echo "== k=32"; timeout 120 python tools/quick_bench.py --dtypes bf16 --paths mma --batches 1,32,128,256,1024,4096 --k 32 --iters 10 2>&1 | tail -6 | cut -c1-200
for dbg in 128 132; do
echo "== VQ_MMA_DEBUG=$dbg"
VQ_MMA_DEBUG=$dbg timeout 120 python tools/quick_bench.py --dtypes bf16 --paths mma --batches 1024 --k 32 --iters 2 2>&1 | grep -E "dbg\]|dtype" | awk '!seen[$0]++' | cut -c1-150 | tail -3
done
