for dbg in 0 4 1 5 6; do
echo "== VQ_MMA_DEBUG=$dbg"
VQ_MMA_DEBUG=$dbg timeout 120 python tools/quick_bench.py --dtypes bf16 --paths mma --batches 1,32 --k 32 --iters 10 2>&1 | tail -2 | cut -c1-170
done
