#!/usr/bin/env python3
"""Developer probe: one shard shape (rows x 512, batch 1024), kernel time of the exact-mode scan (vq_profile events)
and of the round-1 list-mode route on the same box; env knobs are set by the caller (VQ_EXACT_BOOT, VQ_MMA_DEBUG, ...)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_quierer_b200 import _lib, engine
from video_quierer_b200.flat_index import exact_search, two_stage_search
from tools.bench_hnsw import device_rows

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=125_000)
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--kind", default="clip")
ap.add_argument("--old", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda", 0)
lib = _lib.load()
stores = []
for c in range(5):                      # rotate copies: the shard must not sit in L2
    st = engine.DeviceStore(512, dev, keep_fp32=True, keep_bf16=True)
    st.append(device_rows(a.kind, a.n, 512, dev, 1))
    stores.append(st)
q = device_rows(a.kind, a.batch, 512, dev, 2)
sc = engine.Scanner(dev)


def timed(fn, iters=20):
    for i in range(5):
        fn(stores[i % 5])
    torch.cuda.synchronize()
    ks, e0, e1 = [], torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lib.vq_profile_enable(1)
    for i in range(iters):
        fn(stores[i % 5])
        ks.append(lib.vq_profile_last_kernel_ms())
    lib.vq_profile_enable(0)
    e0.record()
    for i in range(iters):
        fn(stores[i % 5])
    e1.record()
    torch.cuda.synchronize()
    ks.sort()
    return round(ks[len(ks) // 2], 4), round(e0.elapsed_time(e1) / iters, 4)


res = {"n": a.n, "batch": a.batch, "kind": a.kind, "env": {k: v for k, v in os.environ.items() if k.startswith("VQ_")}}
res["exact_kernel_ms"], res["exact_step_ms"] = timed(lambda st: exact_search(sc, st, q, 10))
if a.old:
    res["list_kernel_ms"], res["list_step_ms"] = timed(lambda st: two_stage_search(sc, st, q, 10))
print(json.dumps(res), flush=True)
