#!/bin/bash
# ncu evidence, round 2, after the CTA-pair (cta_group::2) scan kernel (one GPU).  Every command first exits 0 without ncu; a number printed under ncu is never a bench value.
set -x
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-hnsw --no-api --no-graph --no-sweep --sustain 0 --pipeline 1"
P="python tools/shard_probe.py --old 0"
$B > gpurun_out/r02b_plain_bench.log 2>&1 && $P --n 1000000 > gpurun_out/r02b_plain_probe.log 2>&1 || exit 1
# launch list of the contract step (eager launches, one step in flight so that the list reads in order)
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02b_launches_bench.csv $B > gpurun_out/r02b_ncu_bench.log 2>&1
# full captures: the exact-mode scan (launches of scan_mma_bf16_kernel alternate boot pass / main pass: odd index = main)
ncu --set full --clock-control none --import-source on -k regex:scan_mma_bf16_kernel -s 11 -c 1 -f -o gpurun_out/r02b_scan_exact_b1024_clip $P --n 1000000 > gpurun_out/r02b_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_mma_bf16_kernel -s 11 -c 1 -f -o gpurun_out/r02b_scan_exact_b1024_gauss $P --n 1000000 --kind gauss > gpurun_out/r02b_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_mma_bf16_kernel -s 11 -c 1 -f -o gpurun_out/r02b_scan_exact_b32_clip $P --n 1000000 --batch 32 > gpurun_out/r02b_ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_mma_bf16_kernel -s 11 -c 1 -f -o gpurun_out/r02b_scan_exact_b1024_shard125k $P --n 125000 > gpurun_out/r02b_ncu_d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:exact_finish_kernel -s 6 -c 1 -f -o gpurun_out/r02b_exact_finish_b1024_clip $P --n 1000000 > gpurun_out/r02b_ncu_e.log 2>&1
# DRAM traffic of the per-GPU shard shapes of N = 2 / 4 (roofline.traffic at N > 1)
ncu --set full --clock-control none -k regex:scan_mma_bf16_kernel -s 11 -c 1 -f -o gpurun_out/r02b_scan_exact_b1024_shard250k $P --n 250000 > gpurun_out/r02b_ncu_f.log 2>&1
ncu --set full --clock-control none -k regex:scan_mma_bf16_kernel -s 11 -c 1 -f -o gpurun_out/r02b_scan_exact_b1024_shard500k $P --n 500000 > gpurun_out/r02b_ncu_g.log 2>&1
# smoke() under a kernel-serialising profiler must pass
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02b_launches_smoke.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_ncu_smoke.log 2>&1; echo "smoke under ncu rc=$?" >> gpurun_out/r02b_ncu_smoke.log
ls -la gpurun_out/*.ncu-rep
tail -3 gpurun_out/r02b_ncu_smoke.log
