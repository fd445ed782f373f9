timeout 120 python tools/mma_check.py 2>&1 | grep -v "torch fp32" | tail -1
for g in 0 1 2; do
echo "== GROUP log2=$g"
VQ_MMA_GROUP=$g timeout 120 python tools/quick_bench.py --dtypes bf16 --paths mma --batches 1,32,128,1024 --k 32 --iters 10 2>&1 | tail -4 | cut -c1-170
done
