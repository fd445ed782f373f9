#!/usr/bin/env python3
"""GPU side of the HNSW recall bar at scale: B200HNSWIndex on the SAME synthetic rows / queries as
tools/ref_recall_at_scale.py (tests/golden/hnsw_ref_recall.json), recall@10 at ef 64/128/256 next to the reference's."""
import argparse, json, os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from video_quierer_b200.hnsw_index import B200HNSWIndex
from video_quierer_b200.utils import synth

REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "hnsw_ref_recall.json")


def gpu_truth(store, queries, k=10):
    dev = torch.device("cuda", 0)
    x = torch.from_numpy(store).to(dev).double()
    x /= x.norm(dim=1, keepdim=True)
    q = torch.from_numpy(queries).to(dev).double()
    q /= q.norm(dim=1, keepdim=True)
    out = []
    for s in range(0, len(queries), 128):
        out.append(torch.topk(q[s:s + 128] @ x.T, k, dim=1).indices.cpu().numpy())
    return np.concatenate(out)


def measure(kind, n, dim=512, nq=1000, efs=(64, 128, 256), **kw):
    gen = synth.clip_like if kind == "clip" else synth.gauss
    store = gen(n, dim, seed=synth.STORE_SEED)
    queries = synth.clip_like(nq, dim, seed=synth.QUERY_SEED, n_store=n) if kind == "clip" else synth.gauss(nq, dim, seed=synth.QUERY_SEED)
    truth = gpu_truth(store, queries)
    random.seed(0)
    h = B200HNSWIndex(dimension=dim, M=16, ef_construction=200, ef_search=64, max_M=16, **kw)
    t0 = time.time()
    h.add_batch(list(store), list(range(n)))
    h.build()
    build_s = time.time() - t0
    out = {"kind": kind, "n": n, "build_s": round(build_s, 2), "select": h.select, "runs": {}}
    for ef in efs:
        h.ef_search = ef
        _, rows = h.search_arrays(queries, 10)
        rec = float(np.mean([len(set(rows[i]) & set(truth[i])) / 10 for i in range(nq)]))
        out["runs"][str(ef)] = {"recall@10": rec, "evals_per_query": float(h.last_stats[:, 0].mean())}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100_000)
    ap.add_argument("--kinds", default="clip,gauss")
    ap.add_argument("--select", default="hybrid")
    ap.add_argument("--max-candidates", type=int, default=63)
    a = ap.parse_args()
    ref = json.load(open(REF)) if os.path.exists(REF) else {}
    for kind in a.kinds.split(","):
        r = measure(kind, a.n, select=a.select, max_candidates=a.max_candidates)
        rr = ref.get(f"{kind}_{a.n}", {}).get("runs", {})
        r["reference"] = {ef: round(v["recall@10"], 4) for ef, v in rr.items()}
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
