VQ_MMA_CLUSTER=2 timeout 120 python tools/mma_check.py 2>&1 | grep -v "torch fp32" | tail -4
for c in 0 2; do
echo "== CLUSTER=$c"
VQ_MMA_CLUSTER=$c timeout 120 python tools/quick_bench.py --dtypes bf16 --paths mma --batches 256,1024,4096 --k 32 --iters 10 2>&1 | tail -3 | cut -c1-200
done
