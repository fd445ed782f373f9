set -x
timeout 120 python tools/mma_check.py 2>&1 | grep -v "torch fp32" | tail -9
timeout 600 python -m pytest tests/test_gpu_exact.py -x -q -m gpu 2>&1 | tail -5
for boot in 1 0; do
echo "== BOOT=$boot k=10"
VQ_MMA_BOOT=$boot timeout 120 python tools/quick_bench.py --dtypes bf16 --paths mma --batches 1,32,128,1024 --k 10 --iters 10 2>&1 | tail -4
echo "== BOOT=$boot k=32"
VQ_MMA_BOOT=$boot timeout 120 python tools/quick_bench.py --dtypes bf16 --paths mma --batches 1,32,128,256,1024 --k 32 --iters 10 2>&1 | tail -5
done
for dbg in 4 6; do
echo "== VQ_MMA_DEBUG=$dbg"
VQ_MMA_DEBUG=$dbg timeout 120 python tools/quick_bench.py --dtypes bf16 --paths mma --batches 1,32,1024 --k 10 --iters 10 2>&1 | tail -3
done
