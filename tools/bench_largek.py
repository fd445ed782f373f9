#!/usr/bin/env python3
"""Developer benchmark of BASELINE config 4's per-GPU shape (10M x 768 over 8 GPUs = 1.25M rows per GPU,
k = 100): the tensor-core route (sampled bound + collect + re-score) against the fp32 FMA scan."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_quierer_b200 import _lib, engine
from video_quierer_b200.flat_index import two_stage_search
from tools.bench_hnsw import device_rows
from tools.bench_clip import timed

dev = torch.device("cuda", 0)
n, dim, k = 1_250_000, 768, 100
for kind in ("gauss", "clip"):
    st = engine.DeviceStore(dim, dev, keep_fp32=True, keep_bf16=True)
    st.append(device_rows(kind, n, dim, dev, 1))
    sc = engine.Scanner(dev)
    for b in (1, 32, 1024):
        q = device_rows(kind, b, dim, dev, 2)
        ms, (s, r, _) = timed(lambda: two_stage_search(sc, st, q, k), iters=5)
        res = {"kind": kind, "n": n, "dim": dim, "k": k, "batch": b, "tensor_route_ms": round(ms, 3), "path": sc.last_path}
        if b <= 32:
            mf, (sf, rf) = timed(lambda: sc.scan(st.f32, st.n, dim, q, k, _lib.NORM_EPS, "fma"), iters=3)
            res["fma_ms"] = round(mf, 3)
            res["id_match"] = round((r == rf).float().mean().item(), 5)
        print(json.dumps(res), flush=True)
    del st
