#!/usr/bin/env python3
"""BASELINE config 3: HNSW M=16 ef_construction=200 ef_search 64-256 on N x 512, recall@10 vs the
exact GPU scan, QPS and achieved gather bandwidth (bytes = evals*ld*elem + hops*deg*4, both
counted by the kernel).  Developer benchmark; the contract line is printed by /bench.py."""
import argparse, json, os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from video_quierer_b200 import _lib, engine
from video_quierer_b200.flat_index import exact_fallback, exact_search
from video_quierer_b200.hnsw_index import B200HNSWIndex


def device_rows(kind, n, dim, dev, seed):
    blk = 1 << 16
    centres = None
    if kind == "clip":
        g = torch.Generator(device=dev).manual_seed(77)
        c = max(256, n // 4096)
        centres = torch.randn((c, dim), device=dev, generator=g)
        centres /= centres.norm(dim=1, keepdim=True)
    out = []
    for b0 in range(0, n, blk):
        g = torch.Generator(device=dev).manual_seed(seed * 100003 + b0 // blk)
        m = min(blk, n - b0)
        x = torch.randn((m, dim), device=dev, generator=g)
        if centres is not None:
            ids = torch.randint(0, centres.shape[0], (m,), device=dev, generator=g)
            x = centres[ids] + 0.35 / dim ** 0.5 * x
        out.append(x / x.norm(dim=1, keepdim=True))
    return torch.cat(out)


def measure(a, dev=None):
    """a: namespace with n, dim, queries, kind, efs, search_dtype.  Returns the result dict."""
    dev = dev or torch.device("cuda", 0)
    x = device_rows(a.kind, a.n, a.dim, dev, 1)
    q = device_rows(a.kind, a.queries, a.dim, dev, 2)
    h = B200HNSWIndex(dimension=a.dim, M=16, ef_construction=200, ef_search=64, max_M=16, search_dtype=a.search_dtype)
    # bulk ingest through the public surface: rows stay on the device, levels from the reference's distribution
    t0 = time.time()
    h.add_device_rows(x, level_seed=0)
    t1 = time.time()
    h.build()
    torch.cuda.synchronize()
    build_s = time.time() - t1
    # exact ground truth: the single-pass exact search (queries whose gather overflowed re-run on the fp32 FMA scan)
    sc = engine.Scanner(dev)
    truth = []
    for s0 in range(0, a.queries, 1024):
        qq = q[s0:s0 + 1024].contiguous()
        s, r, over = exact_search(sc, h._store, qq, 10)
        r = r.cpu().numpy()
        bad = np.nonzero(over.cpu().numpy())[0]
        if len(bad):
            _, rf = exact_fallback(sc, h._store, qq, 10, torch.from_numpy(bad).to(dev))
            r[bad] = rf.cpu().numpy()
        truth.append(r)
    truth = np.concatenate(truth)
    lib = _lib.load(); lib.vq_profile_enable(1)
    res = {"n": a.n, "dim": a.dim, "kind": a.kind, "build_s": round(build_s, 2), "max_level": int(max(h._level_list)),
           "search_dtype": a.search_dtype, "runs": []}
    elem = 2 if a.search_dtype == "bf16" else 4
    for ef in [int(v) for v in a.efs.split(",")]:
        h.ef_search = ef
        h.search_arrays(q[:256], 10)
        torch.cuda.synchronize()
        t0 = time.time()
        dist, rows = h.search_arrays(q, 10)
        wall = time.time() - t0
        kms = lib.vq_profile_last_kernel_ms()
        st = h.last_stats.astype(np.float64)
        rec = np.mean([len(set(rows[i]) & set(truth[i])) / 10 for i in range(a.queries)])
        gbytes = (st[:, 0].sum() * h._store.ld * elem + st[:, 1].sum() * 16 * 4) / 1e9
        res["runs"].append({"ef": ef, "recall@10": round(float(rec), 4), "qps_kernel": round(a.queries / (kms * 1e-3)),
                            "qps_wall": round(a.queries / wall), "kernel_ms": round(kms, 3),
                            "evals_per_query": round(st[:, 0].mean(), 1), "hops_per_query": round(st[:, 1].mean(), 1),
                            "gather_GBps": round(gbytes / (kms * 1e-3), 1), "overflow": int(st[:, 2].sum())})
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--queries", type=int, default=10000)
    ap.add_argument("--kind", default="clip")
    ap.add_argument("--efs", default="64,128,256")
    ap.add_argument("--search-dtype", default="fp32")
    print(json.dumps(measure(ap.parse_args())), flush=True)


if __name__ == "__main__":
    main()
