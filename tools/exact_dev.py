#!/usr/bin/env python3
"""Developer benchmark of the single-pass exact search (vq_search_exact) on S-gauss and S-clip rows:
parity against the fp32 FMA scan on a query subset, step time at batch 1 / 32 / 1024 next to the old
two-stage (+ collect) route, rows gathered / re-scored per query."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_quierer_b200 import _lib, engine
from video_quierer_b200.flat_index import exact_search, resolve_uncertified, two_stage_search
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
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--batches", default="1,32,1024")
    ap.add_argument("--kinds", default="gauss,clip")
    ap.add_argument("--old", type=int, default=1)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    for kind in a.kinds.split(","):
        x = device_rows(kind, a.n, a.dim, dev, 1)
        st = engine.DeviceStore(a.dim, dev, keep_fp32=True, keep_bf16=True)
        st.append(x)
        del x
        sc = engine.Scanner(dev)
        print(json.dumps({"kind": kind, "n": a.n, "bounds": st.bounds.cpu().tolist()}), flush=True)
        for b in [int(v) for v in a.batches.split(",")]:
            q = device_rows(kind, b, a.dim, dev, 2)
            stats = torch.zeros((b, 2), dtype=torch.int32, device=dev)
            s, r, over = sc.exact(st, q, a.k, stats=stats)
            torch.cuda.synchronize()
            ms, _ = timed(lambda: exact_search(sc, st, q, a.k))
            lib.vq_profile_enable(1)
            exact_search(sc, st, q, a.k)
            kms = lib.vq_profile_last_kernel_ms()
            lib.vq_profile_enable(0)
            sub = min(b, 48)
            sf, rf = sc.scan(st.f32, st.n, a.dim, q[:sub].contiguous(), a.k, _lib.NORM_EPS, "fma")
            sth = stats.cpu().float()
            res = {"kind": kind, "batch": b, "exact_ms": round(ms, 4), "scan_kernel_ms": round(kms, 4), "launches": sc.last_launches,
                   "overflow": int(over.sum()), "gathered_mean": round(sth[:, 0].mean().item(), 1), "gathered_max": int(sth[:, 0].max()),
                   "rescored_mean": round(sth[:, 1].mean().item(), 1), "rescored_max": int(sth[:, 1].max()),
                   "ids_equal_fma": bool(torch.equal(r[:sub], rf)), "scores_equal_fma": bool(torch.equal(s[:sub], sf)),
                   "max_score_diff": float((s[:sub] - sf).abs().max())}
            if a.old:
                ms2, (s2, r2, bad) = timed(lambda: two_stage_search(sc, st, q, a.k))
                idx = torch.nonzero(bad).flatten()
                res["old_two_stage_ms"] = round(ms2, 4)
                res["old_uncertified"] = int(len(idx))
                if len(idx):
                    msc, _ = timed(lambda: resolve_uncertified(sc, st, q, a.k, idx, s2))
                    res["old_collect_ms"] = round(msc, 4)
            print(json.dumps(res), flush=True)
        del st, sc
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
