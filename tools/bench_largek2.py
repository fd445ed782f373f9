#!/usr/bin/env python3
"""Stage timing of the large-k route (config 4 per-GPU shape)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_quierer_b200 import _lib, engine
from video_quierer_b200.flat_index import BF16_SCORE_EPS, LARGE_K_CAP
from tools.bench_hnsw import device_rows
from tools.bench_clip import timed

dev = torch.device("cuda", 0)
n, dim, k = 1_250_000, 768, 100
st = engine.DeviceStore(dim, dev, keep_fp32=True, keep_bf16=True)
st.append(device_rows("gauss", n, dim, dev, 1))
sc = engine.Scanner(dev)
stride = min(64, LARGE_K_CAP // (3 * k))
sample = st.sample_f32(stride)
for b in (32, 1024):
    q = device_rows("gauss", b, dim, dev, 2)
    ms1, (s_smp, _) = timed(lambda: sc.scan(sample, sample.shape[0], dim, q, k, _lib.NORM_EPS, "fma"), iters=3)
    thr = s_smp[:, k - 1] - BF16_SCORE_EPS
    ms2, (s, r, over) = timed(lambda: sc.collect(st.bf16, st.f32, st.n, dim, q, k, thr, LARGE_K_CAP), iters=3)
    # candidate counts: scan again with a huge threshold margin? read them from the workspace is not exposed; estimate
    cnt = ((st.bf16[:, :dim].float() @ (q[:4] / q[:4].norm(dim=1, keepdim=True)).T) >= thr[:4]).sum(dim=0)
    print(json.dumps({"batch": b, "sample_rows": int(sample.shape[0]), "sample_fma_ms": round(ms1, 3), "collect_ms": round(ms2, 3),
                      "overflow": int(over.sum()), "rows_above_thr(first 4 queries)": cnt.cpu().tolist()}), flush=True)
