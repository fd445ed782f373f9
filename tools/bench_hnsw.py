#!/usr/bin/env python3
"""BASELINE config 3: HNSW M=16 ef_construction=200 ef_search 64-256 on N x 512, recall@10 vs the
exact GPU scan, QPS and achieved gather bandwidth (bytes = evals*ld*elem + hops*deg*4, both
counted by the kernel).  Developer benchmark; the contract line is printed by /bench.py."""
import argparse, json, os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from video_quierer_b200 import _lib, engine
from video_quierer_b200.flat_index import two_stage_search
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
    # bulk ingest through the public surface: levels from the reference's distribution, rows stay on the device
    random.seed(0)
    t0 = time.time()
    h._ids = list(range(a.n)); h._row_of = {}  # ids == rows for the benchmark (no per-row dict needed)
    u = np.random.default_rng(0).random(a.n)
    h._level_list = list((-np.log(np.maximum(u, 1e-300)) * h.level_generation_factor).astype(np.int32))
    h.element_count = a.n
    h._store.append(x, _lib.NORM_NONE)
    lv = np.asarray(h._level_list); h._entry_row = int(np.argmax(lv == lv.max())); h.entry_point = h._entry_row
    t1 = time.time()
    h.build()
    torch.cuda.synchronize()
    build_s = time.time() - t1
    # exact ground truth with the certified two-stage scan
    sc = engine.Scanner(dev)
    truth = []
    for s0 in range(0, a.queries, 1024):
        s, r, bad = two_stage_search(sc, h._store, q[s0:s0 + 1024].contiguous(), 10)
        truth.append(r.cpu().numpy())
    truth = np.concatenate(truth)
    lib = _lib.load(); lib.vq_profile_enable(1)
    res = {"n": a.n, "dim": a.dim, "kind": a.kind, "build_s": round(build_s, 2), "max_level": int(lv.max()),
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
