#!/usr/bin/env python3
"""The REFERENCE's HNSW recall@10 at scale: builds the graph with the C++
restatement of /root/reference/src/indexes/hnsw.py (oracle/hnsw_ref.cpp — reproduces the reference's 10k graphs edge
for edge) at N rows with the reference's level stream (`random.seed(0)`), searches with ef_search 64/128/256 and
writes recall@10 against the float64 brute-force top-10 to tests/golden/hnsw_ref_recall.json.  CPU only; the GPU
tests / bench compare the B200 index against these numbers on the same synthetic rows.

    python tools/ref_recall_at_scale.py --n 100000 --kind clip
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import hnsw_ref
from video_quierer_b200.utils import synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "hnsw_ref_recall.json")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100_000)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--kind", default="clip", choices=["clip", "gauss"])
    ap.add_argument("--queries", type=int, default=1000)
    ap.add_argument("--efs", default="64,128,256")
    a = ap.parse_args()
    gen = synth.clip_like if a.kind == "clip" else synth.gauss
    store = gen(a.n, a.dim, seed=synth.STORE_SEED)
    queries = synth.clip_like(a.queries, a.dim, seed=synth.QUERY_SEED, n_store=a.n) if a.kind == "clip" else synth.gauss(a.queries, a.dim, seed=synth.QUERY_SEED)
    levels = hnsw_ref.reference_levels(a.n, 0)
    t0 = time.time()
    h = hnsw_ref.RefHNSW(a.dim, 16, 200, 64, 16).build(store, levels)
    build_s = time.time() - t0
    qn = queries / np.linalg.norm(queries, axis=1, keepdims=True)
    truth = np.empty((a.queries, 10), np.int64)
    for s in range(0, a.queries, 64):
        sims = qn[s:s + 64].astype(np.float64) @ h.rows.astype(np.float64).T
        truth[s:s + 64] = np.argsort(-sims, axis=1)[:, :10]
    res = {"n": a.n, "dim": a.dim, "kind": a.kind, "queries": a.queries, "M": 16, "max_M": 16, "ef_construction": 200,
           "level_seed": 0, "store_seed": synth.STORE_SEED, "query_seed": synth.QUERY_SEED, "build_s": round(build_s, 1),
           "store_sha": synth.sha256_of(store), "runs": {}}
    for ef in [int(v) for v in a.efs.split(",")]:
        t1 = time.time()
        ids, _, evals = h.search(queries, 10, ef)
        el = time.time() - t1
        rec = float(np.mean([len(set(ids[i]) & set(truth[i])) / 10 for i in range(a.queries)]))
        res["runs"][str(ef)] = {"recall@10": rec, "evals_per_query": evals / a.queries, "qps_1thread": a.queries / el}
        print(a.kind, a.n, "ef", ef, "recall", round(rec, 4), "evals/q", evals / a.queries, flush=True)
    allr = json.load(open(OUT)) if os.path.exists(OUT) else {}
    allr[f"{a.kind}_{a.n}"] = res
    json.dump(allr, open(OUT, "w"), indent=1, sort_keys=True)
    print("written", OUT)


if __name__ == "__main__":
    main()
