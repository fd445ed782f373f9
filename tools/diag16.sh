timeout 120 python tools/mma_check.py 2>&1 | grep -v "torch fp32" | tail -1
timeout 120 python tools/quick_bench.py --dtypes bf16 --paths mma --batches 1,32,256,1024,4096 --k 32 --iters 10 2>&1 | tail -5 | cut -c1-200
