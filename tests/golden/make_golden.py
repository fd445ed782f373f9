#!/usr/bin/env python3
"""Generate the golden fixtures by running the UNMODIFIED reference from /root/reference.

Run in the authoring container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py [--skip-10k]

Nothing from the reference is copied: its two modules are imported in place
(`video_search_overhaul.py::SimpleVideoIndex`, `src/indexes/hnsw.py::OptimizedHNSWIndex`),
fed seeded synthetic inputs, and only their OUTPUTS (ids, scores, graphs) are stored.
Small inputs are stored too (as float16, exactly representable) so those fixtures do not
depend on a numpy RNG stream; the large ones store a sha256 of the regenerated input.
"""

from __future__ import annotations

import argparse
import importlib.util
import os
import random
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from video_quierer_b200.utils import synth  # noqa: E402
from oracle.hnsw import links_to_arrays  # noqa: E402

REF = "/root/reference"


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ref_modules():
    hn = _load("ref_hnsw", os.path.join(REF, "src/indexes/hnsw.py"))
    ov = _load("ref_overhaul", os.path.join(REF, "video_search_overhaul.py"))
    return hn, ov


def run_exact(ov, store, queries, k):
    idx = ov.SimpleVideoIndex()
    for i, x in enumerate(store):
        idx.add_frame(x, f"v{i % 7}.mp4", float(i) * 0.5)
    rows, scores = [], []
    for q in queries:
        res = idx.search(q, k)
        rows.append([r["frame_id"] for r in res])
        scores.append([r["score"] for r in res])
    return np.array(rows, np.int64), np.array(scores, np.float64), idx


def ref_graph_arrays(h, n):
    links = []
    for lv in sorted(h.graph.keys()):
        while len(links) <= lv:
            links.append({})
        links[lv] = {int(u): set(int(v) for v in nb) for u, nb in h.graph[lv].items()}
    level_of = [int(h.levels[i]) for i in range(n)]
    return links_to_arrays(links, level_of, h.entry_point, h.M, h.max_M)


def build_ref_hnsw(hn, store, seed, **kw):
    random.seed(seed)
    h = hn.OptimizedHNSWIndex(dimension=store.shape[1], num_threads=1, **kw)
    t0 = time.time()
    for i, x in enumerate(store):
        h.add(x, i)
    return h, time.time() - t0


def search_ref_hnsw(h, queries, k, ef):
    h.ef_search = ef
    ids = np.full((len(queries), k), -1, np.int64)
    dist = np.full((len(queries), k), np.nan, np.float64)
    t0 = time.time()
    for b, q in enumerate(queries):
        res = h.search(q, k)
        for j, r in enumerate(res):
            ids[b, j] = r["id"]
            dist[b, j] = float(r["distance"])
    return ids, dist, time.time() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-10k", action="store_true")
    args = ap.parse_args()
    hn, ov = ref_modules()

    # ------------------------------------------------------------------ exact, small
    rng = np.random.default_rng(1234)
    store16 = synth.gauss(1024, 512, seed=11).astype(np.float16)
    q16 = rng.standard_normal((32, 512)).astype(np.float16)        # NOT normalised on purpose
    store = store16.astype(np.float32)
    queries = q16.astype(np.float32)
    out = {"store_f16": store16, "queries_f16": q16}
    for k in (1, 10, 50):
        r, s, idx = run_exact(ov, store, queries, k)
        out[f"rows_k{k}"], out[f"scores_k{k}"] = r, s
    # k > N
    r, s, _ = run_exact(ov, store[:7], queries[:4], 50)
    out["rows_kgtn"], out["scores_kgtn"] = r, s
    # zero query, and float64 query (dtype-preserving path)
    res0 = idx.search(np.zeros(512, np.float32), 5)
    out["zero_scores"] = np.array([x["score"] for x in res0])
    res64 = idx.search(queries[0].astype(np.float64), 10)
    out["rows_f64q"] = np.array([x["frame_id"] for x in res64], np.int64)
    out["scores_f64q"] = np.array([x["score"] for x in res64], np.float64)
    # exact ties: duplicate rows
    tie_store = store[:64].copy()
    tie_store[[5, 17, 40]] = tie_store[3]
    r, s, _ = run_exact(ov, tie_store, tie_store[[3]], 6)
    out["tie_rows"], out["tie_scores"] = r, s
    # metadata dict shape
    res = idx.search(queries[1], 3)
    out["dict_keys"] = np.array(sorted(res[0].keys()))
    out["dict_example"] = np.array([[x["frame_id"], x["timestamp"]] for x in res], np.float64)
    out["dict_video"] = np.array([x["video_name"] for x in res])
    # empty index
    out["empty_len"] = np.array(len(ov.SimpleVideoIndex().search(queries[0], 5)))
    np.savez_compressed(os.path.join(HERE, "exact_small.npz"), **out)
    print("exact_small done")

    # ------------------------------------------------------------------ exact, C1 (seeded)
    store = synth.gauss(10000, 512, seed=synth.STORE_SEED)
    queries = np.random.default_rng(synth.QUERY_SEED).standard_normal((100, 512), dtype=np.float32)
    t0 = time.time()
    r, s, _ = run_exact(ov, store, queries, 10)
    dt = time.time() - t0
    np.savez_compressed(os.path.join(HERE, "exact_c1.npz"), rows=r, scores=s,
                        store_sha=synth.sha256_of(store), query_sha=synth.sha256_of(queries),
                        ref_seconds=np.array(dt), n=10000, dim=512, k=10)
    print(f"exact_c1 done: {dt / 100 * 1e3:.2f} ms/query as shipped")

    # ------------------------------------------------------------------ HNSW, small (inputs stored)
    s16 = synth.clip_like(1500, 64, seed=21).astype(np.float16)
    q16 = synth.clip_like(64, 64, seed=22, n_store=1500).astype(np.float16)
    store = s16.astype(np.float32)
    queries = q16.astype(np.float32)
    h, bt = build_ref_hnsw(hn, store, seed=0, M=16, ef_construction=200, ef_search=50, max_M=16)
    g = ref_graph_arrays(h, len(store))
    out = {"store_f16": s16, "queries_f16": q16, "levels": g.levels, "adj0": g.adj0,
           "upper_off": g.upper_off, "upper_adj": g.upper_adj, "entry": np.array(g.entry),
           "max_level": np.array(g.max_level), "build_seconds": np.array(bt),
           "stored_vectors": np.stack([h.data[i] for i in range(len(store))])}
    for ef in (10, 50, 128):
        ids, dist, _ = search_ref_hnsw(h, queries, 10, ef)
        out[f"ids_ef{ef}"], out[f"dist_ef{ef}"] = ids, dist
    ids, dist, _ = search_ref_hnsw(h, queries, 100, 50)        # k > ef_search → ef = k
    out["ids_k100"], out["dist_k100"] = ids, dist
    np.savez_compressed(os.path.join(HERE, "hnsw_small.npz"), **out)
    print(f"hnsw_small done ({bt:.1f}s build)")

    # a second small graph with non-default M / max_M (M != max_M exercises the :191 rule)
    s16 = synth.gauss(800, 32, seed=31).astype(np.float16)
    q16 = synth.gauss(32, 32, seed=32).astype(np.float16)
    store, queries = s16.astype(np.float32), q16.astype(np.float32)
    h, bt = build_ref_hnsw(hn, store, seed=5, M=8, ef_construction=60, ef_search=40, max_M=12)
    g = ref_graph_arrays(h, len(store))
    out = {"store_f16": s16, "queries_f16": q16, "levels": g.levels, "adj0": g.adj0,
           "upper_off": g.upper_off, "upper_adj": g.upper_adj, "entry": np.array(g.entry),
           "max_level": np.array(g.max_level)}
    ids, dist, _ = search_ref_hnsw(h, queries, 5, 40)
    out["ids_ef40"], out["dist_ef40"] = ids, dist
    np.savez_compressed(os.path.join(HERE, "hnsw_m8.npz"), **out)
    print("hnsw_m8 done")

    if args.skip_10k:
        return
    # ------------------------------------------------------------------ HNSW, 10k (seeded inputs)
    for name, gen in (("clip", synth.clip_like), ("gauss", synth.gauss)):
        n, d = 10000, 512
        store = gen(n, d, seed=synth.STORE_SEED)
        if name == "clip":
            queries = synth.clip_like(100, d, seed=synth.QUERY_SEED, n_store=n)
        else:
            queries = synth.gauss(100, d, seed=synth.QUERY_SEED)
        h, bt = build_ref_hnsw(hn, store, seed=0, M=16, ef_construction=200, ef_search=64, max_M=16)
        g = ref_graph_arrays(h, n)
        truth = np.argsort(-(queries @ store.T), axis=1)[:, :10]
        out = {"store_sha": synth.sha256_of(store), "query_sha": synth.sha256_of(queries),
               "levels": g.levels, "adj0": g.adj0, "upper_off": g.upper_off, "upper_adj": g.upper_adj,
               "entry": np.array(g.entry), "max_level": np.array(g.max_level),
               "build_seconds": np.array(bt), "truth": truth}
        for ef in (64, 128, 256):
            ids, dist, st = search_ref_hnsw(h, queries, 10, ef)
            rec = np.mean([len(set(ids[i]) & set(truth[i])) / 10 for i in range(len(queries))])
            out[f"ids_ef{ef}"], out[f"dist_ef{ef}"] = ids, dist
            out[f"recall_ef{ef}"] = np.array(rec)
            out[f"qps_ef{ef}"] = np.array(len(queries) / st)
            print(f"hnsw_{name}10k ef={ef}: recall@10={rec:.3f} qps={len(queries) / st:.0f}")
        np.savez_compressed(os.path.join(HERE, f"hnsw_{name}10k.npz"), **out)
        print(f"hnsw_{name}10k done ({bt:.0f}s build)")


if __name__ == "__main__":
    main()
