#!/usr/bin/env python3
"""Developer benchmark: the two-stage exact search on CLUSTERED (CLIP-like, near-duplicate heavy) rows,
where the certificate rejects many queries: time of the two-stage pass, of the collect pass that resolves
them, and (for comparison) of the fp32 FMA scan that used to resolve them."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_quierer_b200 import _lib, engine
from video_quierer_b200.flat_index import exact_fallback, resolve_uncertified, two_stage_search
from tools.bench_hnsw import device_rows


def timed(fn, iters=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--batches", default="32,1024")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    x = device_rows("clip", a.n, 512, dev, 1)
    st = engine.DeviceStore(512, dev, keep_fp32=True, keep_bf16=True)
    st.append(x)
    sc = engine.Scanner(dev)
    for b in [int(v) for v in a.batches.split(",")]:
        q = device_rows("clip", b, 512, dev, 2)
        ms2, (s, r, bad) = timed(lambda: two_stage_search(sc, st, q, 10))
        idx = torch.nonzero(bad).flatten()
        res = {"n": a.n, "batch": b, "two_stage_ms": round(ms2, 4), "uncertified": int(len(idx))}
        if len(idx):
            msc, (s2, r2) = timed(lambda: resolve_uncertified(sc, st, q, 10, idx, s))
            res["collect_ms"] = round(msc, 4)
            sub = idx[:64]
            msf, (s3, r3) = timed(lambda: exact_fallback(sc, st, q, 10, sub), iters=3)
            res["fma_ms_for_%d" % len(sub)] = round(msf, 4)
            same = (r2[: len(sub)] == r3).float().mean().item()
            res["collect_vs_fma_id_match"] = round(same, 5)
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
