"""The C++ restatement of the reference HNSW (oracle/hnsw_ref.cpp) against graphs and results produced by the
UNMODIFIED reference src/indexes/hnsw.py (tests/golden/make_golden.py): edge for edge, id for id.  CPU only."""
import random

import numpy as np
import pytest

from oracle import hnsw_ref
from oracle.hnsw import GraphArrays
from video_quierer_b200.utils import synth


def _graph(g) -> GraphArrays:
    return GraphArrays(g["levels"], g["adj0"], g["upper_off"], g["upper_adj"], int(g["entry"]), int(g["max_level"]))


def _same_graph(a: GraphArrays, b: GraphArrays):
    assert np.array_equal(a.levels, b.levels)
    assert a.entry == b.entry and a.max_level == b.max_level
    assert np.array_equal(a.adj0, b.adj0)
    assert np.array_equal(a.upper_off, b.upper_off)
    assert np.array_equal(a.upper_adj, b.upper_adj)


def test_restated_cpython_set_iterates_like_the_real_one():
    """Neighbour sets are iterated in hash-table order by the reference; the restatement must agree with CPython."""
    rng = random.Random(7)
    for _ in range(800):
        s, ops = set(), []
        universe = rng.choice([20, 100, 1000, 100000, 10_000_000])
        for _ in range(rng.randint(1, 150)):
            if s and rng.random() < 0.3:
                k = rng.choice(list(s))
                s.discard(k)
                ops.append(("d", k))
            else:
                k = rng.randrange(universe)
                s.add(k)
                ops.append(k)
        assert hnsw_ref.pyset_order(ops) == list(s)


@pytest.mark.parametrize("name,seed,kw", [
    ("hnsw_small.npz", 0, dict(M=16, ef_construction=200, ef_search=50, max_M=16)),
    ("hnsw_m8.npz", 5, dict(M=8, ef_construction=60, ef_search=40, max_M=12)),
])
def test_small_reference_graphs_edge_for_edge(golden, name, seed, kw):
    g = golden(name)
    store = g["store_f16"].astype(np.float32)
    levels = hnsw_ref.reference_levels(len(store), seed)
    assert np.array_equal(levels, g["levels"])             # the level stream of random.seed(seed)
    h = hnsw_ref.RefHNSW(dimension=store.shape[1], **kw).build(store, levels)
    _same_graph(h.to_arrays(), _graph(g))
    if "stored_vectors" in g.files:
        assert np.array_equal(h.rows, g["stored_vectors"])
    queries = g["queries_f16"].astype(np.float32)
    if name == "hnsw_small.npz":
        for ef in (10, 50, 128):
            ids, d, _ = h.search(queries, 10, ef)
            assert np.array_equal(ids, g[f"ids_ef{ef}"]) and np.array_equal(d.astype(np.float64), g[f"dist_ef{ef}"])
        ids, _, _ = h.search(queries[:8], 100, 50)          # k > ef_search -> ef = k (hnsw.py:264)
        assert np.array_equal(ids, g["ids_k100"][:8])
    else:
        ids, _, _ = h.search(queries, 5, 40)
        assert np.array_equal(ids, g["ids_ef40"])


@pytest.mark.parametrize("name", ["clip", "gauss"])
def test_10k_reference_graph_edge_for_edge(golden, name):
    """BASELINE config 1 scale: the reference's own 10k x 512 graph (M=16, ef_construction=200, max_M=16,
    random.seed(0)) rebuilt by the restatement in seconds — every edge, every search result, every distance."""
    g = golden(f"hnsw_{name}10k.npz")
    n, d = 10000, 512
    store = (synth.clip_like if name == "clip" else synth.gauss)(n, d, seed=synth.STORE_SEED)
    queries = synth.clip_like(100, d, seed=synth.QUERY_SEED, n_store=n) if name == "clip" else synth.gauss(100, d, seed=synth.QUERY_SEED)
    assert synth.sha256_of(store) == str(g["store_sha"]), "numpy RNG stream changed; regenerate golden"
    h = hnsw_ref.RefHNSW(d, 16, 200, 64, 16).build(store, hnsw_ref.reference_levels(n, 0))
    _same_graph(h.to_arrays(), _graph(g))
    for ef in (64, 128, 256):
        ids, dist, _ = h.search(queries, 10, ef)
        assert np.array_equal(ids, g[f"ids_ef{ef}"])
        assert np.array_equal(dist.astype(np.float64), g[f"dist_ef{ef}"])
        rec = np.mean([len(set(ids[i]) & set(g["truth"][i])) / 10 for i in range(100)])
        assert abs(rec - float(g[f"recall_ef{ef}"])) < 1e-9
