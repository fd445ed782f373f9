for boot in 1 0; do
echo "== BOOT=$boot n=125000"
VQ_MMA_BOOT=$boot timeout 120 python tools/quick_bench.py --n 125000 --dtypes bf16 --paths mma --batches 32,256,1024 --k 32 --iters 10 2>&1 | tail -3 | cut -c1-200
echo "== BOOT=$boot n=250000"
VQ_MMA_BOOT=$boot timeout 120 python tools/quick_bench.py --n 250000 --dtypes bf16 --paths mma --batches 256,1024 --k 32 --iters 10 2>&1 | tail -2 | cut -c1-200
done
