for dbg in 128 134 254; do
echo "== VQ_MMA_DEBUG=$dbg"
VQ_MMA_DEBUG=$dbg VQ_MMA_BOOT=0 timeout 120 python tools/quick_bench.py --dtypes bf16 --paths mma --batches 1,1024 --k 10 --iters 2 2>&1 | grep -E "dbg\]|dtype" | awk '!seen[$0]++' | cut -c1-150 | tail -12
done
