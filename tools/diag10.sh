P="python tools/bench_hnsw.py --n 200000 --queries 4096 --efs 256"
timeout 300 $P > gpurun_out/hnsw_plain.log 2>&1 && tail -1 gpurun_out/hnsw_plain.log
ncu --set full --clock-control none --import-source on -k regex:hnsw_search -s 1 -c 1 -o gpurun_out/r01_hnsw_ef256 $P > gpurun_out/ncu_hnsw.log 2>&1
tail -2 gpurun_out/ncu_hnsw.log | cut -c1-300
