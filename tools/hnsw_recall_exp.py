#!/usr/bin/env python3
"""Developer experiment: recall@10 of GPU-built graphs vs the reference-built graph (golden), both
searched by the GPU kernel on 2000 fresh queries (the 100-query golden set is too noisy to rank them)."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import compare
from video_quierer_b200.utils import synth
from video_quierer_b200.hnsw_index import B200HNSWIndex

NQ = 2000
for name in ("gauss", "clip"):
    g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", f"hnsw_{name}10k.npz"), allow_pickle=True)
    n, d = 10000, 512
    gen = synth.clip_like if name == "clip" else synth.gauss
    store = gen(n, d, seed=synth.STORE_SEED)
    queries = synth.clip_like(NQ, d, seed=77, n_store=n) if name == "clip" else synth.gauss(NQ, d, seed=77)
    stored = np.stack([x / np.linalg.norm(x) for x in store])
    truth = np.argsort(-(queries @ stored.T), axis=1)[:, :10]
    h = B200HNSWIndex(dimension=d)
    h.load_arrays(stored, g["levels"], g["adj0"], g["upper_off"], g["upper_adj"], int(g["entry"]))
    out = []
    for ef in (64, 128, 256):
        h.ef_search = ef
        _, rows = h.search_arrays(queries, 10)
        out.append(round(compare.recall_at_k(rows, truth), 4))
    print(name, "reference graph", out, f"evals/query={h.last_stats[:,0].mean():.0f}")
    for sel, mc in (("diverse", 63), ("diverse", 95), ("closest", 63)):
        random.seed(0)
        b = B200HNSWIndex(dimension=d, M=16, ef_construction=200, ef_search=64, max_M=16, select=sel, max_candidates=mc)
        b.add_batch(list(store), list(range(n)))
        out = []
        for ef in (64, 128, 256):
            b.ef_search = ef
            _, rows = b.search_arrays(queries, 10)
            out.append(round(compare.recall_at_k(rows, truth), 4))
        print(f"  {sel:8s} cand={mc:3d} -> {out}  evals/query={b.last_stats[:,0].mean():.0f}")
