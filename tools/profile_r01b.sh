set -x
Q32="python tools/quick_bench.py --dtypes bf16 --paths mma --batches 32 --k 32 --iters 2"
Q1024="python tools/quick_bench.py --dtypes bf16 --paths mma --batches 1024 --k 32 --iters 2"
ncu --set full --clock-control none --import-source on -k regex:scan_mma_bf16_kernel -s 7 -c 1 -o gpurun_out/r01_scan_mma_b32 -f $Q32 > gpurun_out/ncu_q32.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_mma_bf16_kernel -s 7 -c 1 -o gpurun_out/r01_scan_mma_b1024 -f $Q1024 > gpurun_out/ncu_q1024.log 2>&1
