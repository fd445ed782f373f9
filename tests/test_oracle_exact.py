"""The exact-search oracle against the golden outputs of the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np

from oracle import exact, compare
from video_quierer_b200.utils import synth


def _inputs(g):
    return g["store_f16"].astype(np.float32), g["queries_f16"].astype(np.float32)


def test_small_ids_and_scores_bit_identical(golden):
    g = golden("exact_small.npz")
    store, queries = _inputs(g)
    for k in (1, 10, 50):
        rows, scores = exact.exact_search_batch(store, queries, k)
        # same numpy, same arithmetic → the restatement must agree bit for bit
        assert np.array_equal(rows, g[f"rows_k{k}"])
        assert np.array_equal(scores, g[f"scores_k{k}"])


def test_k_greater_than_n(golden):
    g = golden("exact_small.npz")
    store, queries = _inputs(g)
    rows, scores = exact.exact_search_batch(store[:7], queries[:4], 50)
    assert rows.shape == (4, 7)
    assert np.array_equal(rows, g["rows_kgtn"]) and np.array_equal(scores, g["scores_kgtn"])


def test_zero_query_and_f64_query_and_empty(golden):
    g = golden("exact_small.npz")
    store, queries = _inputs(g)
    _, s0 = exact.exact_search(store, np.zeros(512, np.float32), 5)
    assert np.array_equal(s0, g["zero_scores"]) and np.all(s0 == 0.0)
    r64, s64 = exact.exact_search(store, queries[0].astype(np.float64), 10)
    assert np.array_equal(r64, g["rows_f64q"]) and np.array_equal(s64, g["scores_f64q"])
    r, s = exact.exact_search(np.zeros((0, 512), np.float32), queries[0], 5)
    assert len(r) == 0 and int(g["empty_len"]) == 0


def test_ties_reference_order(golden):
    g = golden("exact_small.npz")
    store, _ = _inputs(g)
    tie_store = store[:64].copy()
    tie_store[[5, 17, 40]] = tie_store[3]
    rows, scores = exact.exact_search_batch(tie_store, tie_store[[3]], 6)
    assert np.array_equal(rows, g["tie_rows"])
    # the four duplicates tie exactly; the reference's order among them is whatever the
    # unstable np.argsort yields (here 40,5,3,17) — which is why ties are compared by score
    assert sorted(g["tie_rows"][0][:4]) == [3, 5, 17, 40]
    # and the tie-aware comparator accepts the engine's (row asc) order for the same scores
    eng = rows.copy()
    eng[0, :4] = [3, 5, 17, 40]
    assert compare.check_topk_batch(eng, scores, g["tie_rows"], g["tie_scores"]) == []


def test_dict_shape(golden):
    g = golden("exact_small.npz")
    store, queries = _inputs(g)
    md = [{"video_name": f"v{i % 7}.mp4", "timestamp": float(i) * 0.5, "frame_id": i} for i in range(len(store))]
    res = exact.search_dicts(store, md, queries[1], 3)
    assert sorted(res[0].keys()) == list(g["dict_keys"])
    assert [[r["frame_id"], r["timestamp"]] for r in res] == g["dict_example"].tolist()
    assert [r["video_name"] for r in res] == list(g["dict_video"])
    assert exact.formatted_time(125.7) == "2m5s"


def test_c1_full_config(golden):
    """BASELINE config 1: 10k x 512 store, 100 queries, k=10 — inputs regenerated from the seed."""
    g = golden("exact_c1.npz")
    store = synth.gauss(10000, 512, seed=synth.STORE_SEED)
    queries = np.random.default_rng(synth.QUERY_SEED).standard_normal((100, 512), dtype=np.float32)
    assert synth.sha256_of(store) == str(g["store_sha"]), "numpy RNG stream changed; regenerate golden"
    assert synth.sha256_of(queries) == str(g["query_sha"])
    rows, scores = exact.exact_search_batch(store, queries, 10)
    assert np.array_equal(rows, g["rows"]) and np.array_equal(scores, g["scores"])
    # float64 ground truth agrees with the fp32 reference inside the north-star tolerance
    r64, s64 = exact.topk_f64(store, queries, 10)
    assert compare.check_topk_batch(g["rows"], g["scores"], r64, s64) == []


def test_comparator_rejects_real_differences():
    rows_r = np.array([[5, 3, 9]]); s_r = np.array([[0.9, 0.8, 0.7]])
    assert compare.check_topk_batch(rows_r, s_r, rows_r, s_r) == []
    assert compare.check_topk_batch(np.array([[5, 9, 3]]), np.array([[0.9, 0.7, 0.8]]), rows_r, s_r) != []
    assert compare.check_topk_batch(rows_r, s_r * (1 + 1e-4), rows_r, s_r) != []
    assert compare.check_topk_batch(np.array([[5, 3, 11]]), np.array([[0.9, 0.8, 0.7]]), rows_r, s_r) == []  # boundary tie
    assert compare.check_topk_batch(np.array([[5, 3, 11]]), np.array([[0.9, 0.8, 0.6]]), rows_r, s_r) != []
