"""The HNSW oracle (oracle/hnsw.py) against graphs and results produced by the UNMODIFIED
reference src/indexes/hnsw.py (tests/golden/make_golden.py).  CPU only."""
import random

import numpy as np
import pytest

from oracle import compare
from oracle.hnsw import GraphArrays, OracleHNSW, search_arrays
from video_quierer_b200.utils import synth


def _graph(g) -> GraphArrays:
    return GraphArrays(g["levels"], g["adj0"], g["upper_off"], g["upper_adj"], int(g["entry"]), int(g["max_level"]))


def _same_graph(a: GraphArrays, b: GraphArrays):
    assert np.array_equal(a.levels, b.levels)
    assert a.entry == b.entry and a.max_level == b.max_level
    assert np.array_equal(a.adj0, b.adj0)
    assert np.array_equal(a.upper_off, b.upper_off)
    assert np.array_equal(a.upper_adj, b.upper_adj)


@pytest.mark.parametrize("name,seed,kw", [
    ("hnsw_small.npz", 0, dict(M=16, ef_construction=200, ef_search=50, max_M=16)),
    ("hnsw_m8.npz", 5, dict(M=8, ef_construction=60, ef_search=40, max_M=12)),
])
def test_build_reproduces_reference_graph(golden, name, seed, kw):
    g = golden(name)
    store = g["store_f16"].astype(np.float32)
    random.seed(seed)                      # the reference draws levels from the global stream
    h = OracleHNSW(dimension=store.shape[1], **kw)
    for x in store:
        h.add(x)
    _same_graph(h.to_arrays(), _graph(g))
    if "stored_vectors" in g.files:
        assert np.array_equal(h.store(), g["stored_vectors"])


def test_search_matches_reference_ids_and_distances(golden):
    g = golden("hnsw_small.npz")
    store = g["stored_vectors"]
    queries = g["queries_f16"].astype(np.float32)
    ga = _graph(g)
    for ef in (10, 50, 128):
        for b, q in enumerate(queries):
            found, evals, hops = search_arrays(store, ga, q, 10, ef)
            ids = [v for _, v in found]
            assert ids == [int(x) for x in g[f"ids_ef{ef}"][b] if x >= 0]
            assert np.array_equal(np.array([d for d, _ in found], np.float64), g[f"dist_ef{ef}"][b][: len(found)])
            assert evals > 0 and hops > 0
    # k > ef_search → ef = k (hnsw.py:264)
    for b, q in enumerate(queries[:8]):
        found, _, _ = search_arrays(store, ga, q, 100, 50)
        assert [v for _, v in found] == [int(x) for x in g["ids_k100"][b] if x >= 0]


def test_search_m8_graph(golden):
    g = golden("hnsw_m8.npz")
    h_store = g["store_f16"].astype(np.float32)
    h_store = h_store / np.linalg.norm(h_store, axis=1, keepdims=True)
    ga = _graph(g)
    queries = g["queries_f16"].astype(np.float32)
    same = 0
    for b, q in enumerate(queries):
        found, _, _ = search_arrays(h_store, ga, q, 5, 40)
        same += [v for _, v in found] == [int(x) for x in g["ids_ef40"][b] if x >= 0]
    # rows normalised with a vectorised norm may differ by an ulp from the reference's per-row
    # norm, so allow a stray near-tie
    assert same >= len(queries) - 1


@pytest.mark.parametrize("name", ["clip", "gauss"])
def test_10k_reference_graph_recall_bar(golden, name):
    """Search the reference-built 10k graph with the oracle: same ids as the reference, and
    the recall@10 the reference itself achieved (the bar the CUDA path must meet)."""
    g = golden(f"hnsw_{name}10k.npz")
    n, d = 10000, 512
    gen = synth.clip_like if name == "clip" else synth.gauss
    store = gen(n, d, seed=synth.STORE_SEED)
    queries = synth.clip_like(100, d, seed=synth.QUERY_SEED, n_store=n) if name == "clip" else synth.gauss(100, d, seed=synth.QUERY_SEED)
    assert synth.sha256_of(store) == str(g["store_sha"]), "numpy RNG stream changed; regenerate golden"
    assert synth.sha256_of(queries) == str(g["query_sha"])
    # the reference re-normalises stored rows with a per-row np.linalg.norm (hnsw.py:157)
    stored = np.stack([x / np.linalg.norm(x) for x in store])
    ga = _graph(g)
    ef = 64
    found_rows, match = [], 0
    for b, q in enumerate(queries[:40]):
        found, _, _ = search_arrays(stored, ga, q, 10, ef)
        ids = [v for _, v in found]
        found_rows.append(ids)
        match += ids == [int(x) for x in g[f"ids_ef{ef}"][b] if x >= 0]
    assert match == 40
    rec = compare.recall_at_k(found_rows, g["truth"][:40])
    ref_rec = compare.recall_at_k(g[f"ids_ef{ef}"][:40], g["truth"][:40])
    assert abs(rec - ref_rec) < 1e-9
