# First measurements of the next round, cheapest first.  Run under gpurun.
set -x
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="--steps 30 --warmup 3 --no-hnsw --no-cpu --no-api --no-sweep"
# 1. (1 GPU) the N = 8 shard alone: where the 0.157 ms step goes.  Launch list (serialised) + the pipelined step with the
#    sample pass forced onto fewer tiles (does a smaller sample shorten the sample pass -> scan chain more than it costs in gather?)
python bench.py $B --rows 125000 > gpurun_out/r125_default.json 2> gpurun_out/r125_default.err
for sub in 4 8; do VQ_EXACT_SUB=$sub python bench.py $B --rows 125000 > gpurun_out/r125_sub$sub.json 2> gpurun_out/r125_sub$sub.err; done
# 2. (1 GPU) the pipeline events of the pair kernel on a SHORT scan (54 tiles per CTA): how long until the first tile's
#    epilogue runs, how long the tail is (tools/scan_trace.sh needs the tracing build)
VQ_NVCC_EXTRA=-DVQ_SCAN_TRACE python -m video_quierer_b200.build --force && bash tools/scan_trace.sh "0 1" > gpurun_out/trace_1m.log 2>&1
python -m video_quierer_b200.build --force
# 3. (2 GPUs, cheap) host queue depth and the exchange: value vs e2e at 1 / 2 / 3 steps ahead
for cap in 1 2 3; do
  VQ_BENCH_INFLIGHT=$cap $T --nproc-per-node 2 --master-port 2981$cap bench.py --gpus 2 $B --sustain 0 \
    > gpurun_out/n2_cap$cap.json 2> gpurun_out/n2_cap$cap.err
done
# 4. (8 GPUs) three runs of the contract line: the run-to-run spread at N = 8 was 3.85-4.98 M QPS this round
for i in 1 2 3; do
  $T --nproc-per-node 8 --master-port 2960$i bench.py --gpus 8 --steps 30 --warmup 3 --no-sweep --no-hnsw --no-cpu --no-api \
    > gpurun_out/n8_run$i.json 2> gpurun_out/n8_run$i.err
done
# 5. (1 GPU) HNSW: ncu of the search kernel at ef 128 (the bound is per-hop bookkeeping on shared memory)
python tools/bench_hnsw.py --efs 128 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 1 -c 1 -f -o gpurun_out/r03_hnsw_ef128 \
  python tools/bench_hnsw.py --efs 128 > gpurun_out/ncu_hnsw.log 2>&1
