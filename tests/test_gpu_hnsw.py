"""GPU tests of the HNSW path (warp-per-query search kernel, GPU layer builder, drop-in facade)
against graphs/results of the unmodified reference (golden) and the oracle."""
import random

import numpy as np
import pytest

from oracle import compare
from oracle.hnsw import GraphArrays, search_arrays
from video_quierer_b200.utils import synth

pytestmark = pytest.mark.gpu


def _graph(g):
    return GraphArrays(g["levels"], g["adj0"], g["upper_off"], g["upper_adj"], int(g["entry"]), int(g["max_level"]))


def _index_from_golden(g, vectors, **kw):
    from video_quierer_b200.hnsw_index import B200HNSWIndex
    h = B200HNSWIndex(dimension=vectors.shape[1], **kw)
    h.load_arrays(vectors, g["levels"], g["adj0"], g["upper_off"], g["upper_adj"], int(g["entry"]))
    return h


def test_search_on_reference_graph_matches_reference_ids(built_lib, golden):
    """Same graph, same vectors → the kernel must return what the reference returned."""
    g = golden("hnsw_small.npz")
    vectors = g["stored_vectors"]
    queries = g["queries_f16"].astype(np.float32)
    h = _index_from_golden(g, vectors, M=16, max_M=16)
    for ef in (10, 50, 128):
        h.ef_search = ef
        dist, rows = h.search_arrays(queries, 10)
        ref_ids, ref_d = g[f"ids_ef{ef}"], g[f"dist_ef{ef}"]
        same = sum(list(rows[b]) == list(ref_ids[b]) for b in range(len(queries)))
        assert same >= len(queries) - 1, (ef, same)          # a stray fp32 near-tie may reorder one list
        ok = rows == ref_ids
        assert np.allclose(dist[ok], ref_d[ok], rtol=1e-5, atol=2e-6)
        assert h.last_stats[:, 2].sum() == 0                # no visited-set overflow
    # k > ef_search → ef = k (hnsw.py:264)
    h.ef_search = 50
    dist, rows = h.search_arrays(queries[:8], 100)
    same = sum(list(rows[b]) == list(g["ids_k100"][b]) for b in range(8))
    assert same >= 7


def test_distance_evaluations_match_oracle_traversal(built_lib, golden):
    """The kernel walks the graph like hnsw.py:76-121: same number of distance evaluations and
    expansions as the oracle on the same graph (this is the E and H of the gather roofline)."""
    g = golden("hnsw_small.npz")
    vectors = g["stored_vectors"]
    queries = g["queries_f16"].astype(np.float32)
    h = _index_from_golden(g, vectors)
    h.ef_search = 50
    h.search_arrays(queries, 10)
    ga = _graph(g)
    ev = [search_arrays(vectors, ga, q, 10, 50)[1:] for q in queries]
    evals = np.array([e[0] for e in ev]); hops = np.array([e[1] for e in ev])
    # the reference re-evaluates the entry point at every layer (hnsw.py:92-97); the kernel carries
    # that distance down instead, so it makes exactly max_level fewer evaluations
    k_evals = h.last_stats[:, 0].astype(int) + int(g["max_level"])
    k_hops = h.last_stats[:, 1].astype(int)
    print("evals kernel/oracle", k_evals[:8], evals[:8], "hops", k_hops[:8], hops[:8])
    assert (k_evals == evals).mean() >= 0.95
    assert (k_hops == hops).mean() >= 0.95


def test_m8_graph_non_default_degrees(built_lib, golden):
    g = golden("hnsw_m8.npz")
    store = g["store_f16"].astype(np.float32)
    store = np.stack([x / np.linalg.norm(x) for x in store])
    h = _index_from_golden(g, store, M=8, max_M=12)
    h.ef_search = 40
    _, rows = h.search_arrays(g["queries_f16"].astype(np.float32), 5)
    same = sum(list(rows[b]) == list(g["ids_ef40"][b]) for b in range(len(rows)))
    assert same >= len(rows) - 1


@pytest.mark.parametrize("name", ["clip", "gauss"])
def test_10k_reference_graph_and_gpu_build_recall_bar(built_lib, golden, name):
    """(1) kernel on the reference's own 10k graph reproduces the reference's recall;
    (2) the GPU-built graph reaches recall@10 >= the reference's at the same M / ef."""
    from video_quierer_b200.hnsw_index import B200HNSWIndex
    g = golden(f"hnsw_{name}10k.npz")
    n, d = 10000, 512
    gen = synth.clip_like if name == "clip" else synth.gauss
    store = gen(n, d, seed=synth.STORE_SEED)
    queries = synth.clip_like(100, d, seed=synth.QUERY_SEED, n_store=n) if name == "clip" else synth.gauss(100, d, seed=synth.QUERY_SEED)
    assert synth.sha256_of(store) == str(g["store_sha"])
    stored = np.stack([x / np.linalg.norm(x) for x in store])
    h = _index_from_golden(g, stored)
    truth = g["truth"]
    for ef in (64, 128, 256):
        h.ef_search = ef
        _, rows = h.search_arrays(queries, 10)
        rec = compare.recall_at_k(rows, truth)
        assert abs(rec - float(g[f"recall_ef{ef}"])) <= 0.005, (ef, rec)
        assert (rows == g[f"ids_ef{ef}"]).mean() >= 0.98
    # GPU build with the reference's parameters and level stream
    random.seed(0)
    b = B200HNSWIndex(dimension=d, M=16, ef_construction=200, ef_search=64, max_M=16)
    b.add_batch(list(store), list(range(n)))
    assert [b.levels[i] for i in range(50)] == [int(x) for x in g["levels"][:50]]     # same level stream
    # (2a) on the golden's own 100 queries (1000 hits per ef: one hit = 0.001) the two graphs must
    # agree within sampling noise ...
    for ef in (64, 128, 256):
        b.ef_search = ef
        _, rows = b.search_arrays(queries, 10)
        rec = compare.recall_at_k(rows, truth)
        print(f"{name} ef={ef} (100 golden queries): gpu-built recall@10={rec:.3f} reference={float(g[f'recall_ef{ef}']):.3f}")
        assert rec >= float(g[f"recall_ef{ef}"]) - 0.005
    # (2b) ... and the bar itself — recall@10 >= the reference's at the same M / ef, no tolerance — is
    # checked where it is statistically meaningful: 2000 fresh queries (20 000 hits per ef) searched by
    # the same kernel on the REFERENCE-built graph (which part (1) showed reproduces the reference's
    # ids) and on the GPU-built graph.
    nq = 2000
    big_q = synth.clip_like(nq, d, seed=77, n_store=n) if name == "clip" else synth.gauss(nq, d, seed=77)
    big_truth = np.argsort(-(big_q @ stored.T), axis=1)[:, :10]
    for ef in (64, 128, 256):
        h.ef_search = b.ef_search = ef
        _, rows_ref = h.search_arrays(big_q, 10)
        _, rows_gpu = b.search_arrays(big_q, 10)
        rec_ref = compare.recall_at_k(rows_ref, big_truth)
        rec_gpu = compare.recall_at_k(rows_gpu, big_truth)
        print(f"{name} ef={ef} ({nq} queries): gpu-built recall@10={rec_gpu:.4f} reference-built={rec_ref:.4f}")
        # clustered rows (the realistic case): strictly at or above the reference.  iid gaussian rows at 10k: the default
        # layer 0 follows the reference's own construction rule, so the two graphs are statistically indistinguishable
        # there (one sigma of a 20 000-hit recall estimate is 0.003 per graph); the margin shows at 100k and 1M
        assert rec_gpu >= rec_ref - (0.0 if name == "clip" else 0.01)
    assert b.entry_point == int(g["entry"])


def test_facade_api_surface(built_lib, tmp_path):
    from video_quierer_b200.hnsw_index import B200HNSWIndex
    store = synth.clip_like(3000, 64, seed=41)
    h = B200HNSWIndex(dimension=64, ef_search=50)
    assert h.search(store[0], 5) == [] and h.size() == 0
    ids = [f"vid{i // 100}_{i % 100}" for i in range(len(store))]      # string ids like the unwired orchestrator
    h.add_batch(list(store), ids)
    res = h.search(store[123] * 3.0, k=5)                              # un-normalised query is normalised
    assert res[0]["id"] == "vid1_23" and abs(res[0]["distance"]) < 1e-5
    assert sorted(res[0].keys()) == ["distance", "id", "score"]
    assert all(res[i]["distance"] <= res[i + 1]["distance"] for i in range(4))
    assert isinstance(res[0]["distance"], np.float32) and abs(res[0]["score"] + res[0]["distance"] - 1.0) < 1e-6
    batch = h.search_batch([store[5], store[6]], k=3)
    assert [b[0]["id"] for b in batch] == ["vid0_5", "vid0_6"]
    st = h.get_stats()
    assert sorted(st) == sorted(['element_count', 'entry_point_level', 'avg_search_time_ms', 'p95_search_time_ms',
                                 'total_searches', 'dimension', 'M', 'ef_search'])
    assert st["element_count"] == 3000 and st["total_searches"] == 3
    # k > N and k > ef_search
    small = B200HNSWIndex(dimension=64)
    small.add_batch(list(store[:7]), list(range(7)))
    assert len(small.search(store[0], 50)) == 7
    # rows added after the build are found through the delta scan, without a rebuild
    g_before = h._graph
    extra = synth.clip_like(10, 64, seed=43)
    h.add_batch(list(extra), [f"new{i}" for i in range(10)])
    assert h.search(extra[3], 1)[0]["id"] == "new3" and h._graph is g_before
    # save / load (pickle + sha256 sidecar), reference format
    p = str(tmp_path / "idx" / "hnsw.pkl")
    h.save(p)
    h2 = B200HNSWIndex(dimension=64)
    h2.load(p)
    assert h2.size() == h.size() and h2.entry_point == h.entry_point
    a = [x["id"] for x in h.search(store[77], 10)]
    b = [x["id"] for x in h2.search(store[77], 10)]
    assert a[0] == b[0] == "vid0_77" and len(set(a) & set(b)) >= 8
    with open(p + ".sha256", "w") as f:
        f.write("0" * 64)
    with pytest.raises(ValueError):
        B200HNSWIndex(dimension=64).load(p)
    with pytest.raises(FileNotFoundError):
        h.save("nodir.pkl")
    h.thread_pool.shutdown()


def test_reference_pickle_loads(built_lib, tmp_path):
    """A pickle in the reference's exact layout (hnsw.py:311-324) loads and searches."""
    import pickle, hashlib
    from oracle.hnsw import OracleHNSW
    from video_quierer_b200.hnsw_index import B200HNSWIndex
    random.seed(3)
    store = synth.clip_like(400, 32, seed=51)
    o = OracleHNSW(dimension=32, M=8, ef_construction=50, max_M=10)
    for x in store:
        o.add(x)
    graph = {lv: {u: set(nb) for u, nb in layer.items()} for lv, layer in enumerate(o.links)}
    payload = {'dimension': 32, 'M': 8, 'max_M': 10, 'ef_construction': 50, 'ef_search': 30,
               'level_generation_factor': o.mL, 'data': {i: o.vec[i] for i in range(400)},
               'levels': {i: o.level_of[i] for i in range(400)}, 'graph': graph,
               'entry_point': o.entry, 'element_count': 400}
    p = tmp_path / "ref.pkl"
    p.write_bytes(pickle.dumps(payload, protocol=pickle.HIGHEST_PROTOCOL))
    (tmp_path / "ref.pkl.sha256").write_text(hashlib.sha256(p.read_bytes()).hexdigest())
    h = B200HNSWIndex(dimension=32)
    h.load(str(p))
    assert h.M == 8 and h.max_M == 10 and h.ef_search == 30
    for q in store[:20]:
        want = [v for _, v in o.search(q, 5, 30)]
        got = [x["id"] for x in h.search(q, 5)]
        assert got == want


def test_bf16_search_dtype_recall(built_lib):
    from video_quierer_b200.hnsw_index import B200HNSWIndex
    store = synth.clip_like(5000, 128, seed=61)
    queries = synth.clip_like(50, 128, seed=62, n_store=5000)
    a = B200HNSWIndex(dimension=128, ef_search=64)
    b = B200HNSWIndex(dimension=128, ef_search=64, search_dtype="bf16")
    random.seed(1); a.add_batch(list(store), list(range(5000)))
    random.seed(1); b.add_batch(list(store), list(range(5000)))
    _, ra = a.search_arrays(queries, 10)
    _, rb = b.search_arrays(queries, 10)
    assert compare.recall_at_k(rb, ra) >= 0.9


def test_sharded_subgraphs_merge(built_lib):
    """per-shard sub-graphs searched with the same ef and merged with the shard
    offsets (vq_topk_merge) — 4 logical shards on one GPU.  The merged result must be exactly the
    best k of the union of the per-shard results, and its recall at least the per-shard average."""
    import torch
    from video_quierer_b200 import engine
    from video_quierer_b200.hnsw_index import B200HNSWIndex
    from video_quierer_b200.sharded import shard_range
    n, d, g, k = 20000, 128, 4, 10
    store = synth.clip_like(n, d, seed=71)
    queries = synth.clip_like(64, d, seed=72, n_store=n)
    stored = store / np.linalg.norm(store, axis=1, keepdims=True)
    truth = np.argsort(-(queries @ stored.T), axis=1)[:, :k]
    qd = torch.from_numpy(queries).cuda()
    ss, rr, offs, per_shard = [], [], [], []
    for r in range(g):
        lo, hi = shard_range(n, g, r)
        random.seed(r)
        h = B200HNSWIndex(dimension=d, ef_search=64)
        h.add_batch(list(store[lo:hi]), list(range(hi - lo)))
        s, rows = h.as_local_search()(qd, k)
        ss.append(s), rr.append(rows), offs.append(lo)
        local_truth = np.argsort(-(queries @ stored[lo:hi].T), axis=1)[:, :k]
        per_shard.append(compare.recall_at_k(rows.cpu().numpy(), local_truth))
    scores, rows = torch.stack(ss).contiguous(), torch.stack(rr).contiguous()
    ms, mr = engine.Scanner().merge(scores, rows, torch.tensor(offs, dtype=torch.int64, device="cuda"), k)
    ms, mr = ms.cpu().numpy(), mr.cpu().numpy()
    # exactly the best k of the union (score desc, global row asc)
    for b in range(len(queries)):
        cand = sorted((-float(scores[s_, b, j]), int(rows[s_, b, j]) + offs[s_]) for s_ in range(g) for j in range(k)
                      if int(rows[s_, b, j]) >= 0)[:k]
        assert [c[1] for c in cand] == mr[b].tolist()
        assert np.allclose([-c[0] for c in cand], ms[b])
    rec = compare.recall_at_k(mr, truth)
    print(f"sharded HNSW: merged recall@10={rec:.3f}, per-shard {np.round(per_shard, 3)}")
    assert rec >= np.mean(per_shard) - 0.02


def test_readded_ids_still_return_k_hits_and_degree_is_validated(built_lib):
    """an id that is added again supersedes its old row (the reference overwrites data[node_id] in place
    and still returns k hits); M / max_M beyond the device graph's degree limit fail at construction."""
    from video_quierer_b200.hnsw_index import B200HNSWIndex, MAX_DEGREE
    with pytest.raises(ValueError):
        B200HNSWIndex(dimension=64, M=MAX_DEGREE + 1)
    with pytest.raises(ValueError):
        B200HNSWIndex(dimension=64, M=16, max_M=40)
    x = synth.clip_like(3000, 64, seed=141)
    h = B200HNSWIndex(dimension=64, M=16, ef_construction=200, ef_search=64, max_M=16)
    h.add_batch(list(x), list(range(3000)))
    h.build()
    moved = synth.clip_like(20, 64, seed=142, n_store=3000)
    for i in range(20):
        h.add(moved[i], i)                                  # ids 0..19 re-added with new vectors
    q = x[5]                                                # the OLD vector of id 5: its row is dead now
    res = h.search(q, 10)
    assert len(res) == 10 and len({r["id"] for r in res}) == 10
    live = np.concatenate([moved, x[20:]])
    live_ids = list(range(20)) + list(range(20, 3000))
    truth = [live_ids[j] for j in np.argsort(-(live @ (q / np.linalg.norm(q))))[:10]]
    assert len(set(truth) & {r["id"] for r in res}) >= 8
    got5 = [r for r in res if r["id"] == 5]
    if got5:                                                # if id 5 is returned it is scored with its NEW vector
        assert abs(float(got5[0]["score"]) - float(moved[5] @ (q / np.linalg.norm(q)))) < 1e-4


def test_recall_at_1m_meets_the_reference_bar(built_lib):
    """BASELINE config 3 at full size: 1M x 512 clustered rows, M=16, ef 64/128/256 — the reference's recall comes from
    its C++ restatement built on these very rows (30 min of CPU; tests/golden/hnsw_ref_recall.json: 0.680 / 0.812 / 0.887)."""
    import json
    import os
    from tools.hnsw_recall_at_scale import measure
    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "hnsw_ref_recall.json")))["clip_1000000"]
    got = measure("clip", 1_000_000)
    for ef in ("64", "128", "256"):
        assert got["runs"][ef]["recall@10"] >= ref["runs"][ef]["recall@10"] - 0.005, (ef, got["runs"][ef], ref["runs"][ef])


@pytest.mark.parametrize("kind", ["clip", "gauss"])
def test_recall_at_100k_meets_the_reference_bar(built_lib, kind):
    """North-star: 'HNSW must reach recall@10 >= the reference's recall at the same M/ef' — at 100k x 512, where the
    reference's Python build is infeasible, the bar comes from its C++ restatement (oracle/hnsw_ref.cpp, validated
    edge for edge on the reference's own 10k graphs; numbers in tests/golden/hnsw_ref_recall.json, produced by
    tools/ref_recall_at_scale.py on these very rows and queries)."""
    import json
    import os
    from tools.hnsw_recall_at_scale import measure
    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "hnsw_ref_recall.json")))[f"{kind}_100000"]
    store = (synth.clip_like if kind == "clip" else synth.gauss)(100_000, 512, seed=synth.STORE_SEED)
    assert synth.sha256_of(store) == ref["store_sha"], "numpy RNG stream changed; regenerate the reference recall"
    got = measure(kind, 100_000)
    for ef in ("64", "128", "256"):
        assert got["runs"][ef]["recall@10"] >= ref["runs"][ef]["recall@10"] - 0.005, (ef, got["runs"][ef], ref["runs"][ef])


def test_hybrid_builder_layer0_looks_like_the_reference(built_lib):
    """The default builder's layer 0 follows hnsw.py:183-223 (links in both directions, closest-M prune with the dropped
    link removed at both ends): symmetric, degrees mostly short of max_M, every node reachable — the structure of the
    reference's own graphs (golden 10k: mean degree ~11, symmetric) — while the upper layers keep full diverse lists."""
    from video_quierer_b200.hnsw_index import B200HNSWIndex
    n = 20000
    x = synth.clip_like(n, 128, seed=151)
    random.seed(0)
    h = B200HNSWIndex(dimension=128, M=16, ef_construction=200, ef_search=64, max_M=16)
    assert h.select == "hybrid"
    h.add_batch(list(x), list(range(n)))
    h.build()
    adj0 = h._graph.adj0.cpu().numpy()
    deg = (adj0 >= 0).sum(axis=1)
    assert deg.max() <= 16 and 6.0 < deg.mean() < 15.0 and (deg < 16).mean() > 0.3
    # no holes inside a list, no self loops, no duplicates
    for u in range(0, n, 97):
        row = adj0[u]
        k = int(deg[u])
        assert np.all(row[:k] >= 0) and np.all(row[k:] == -1) and u not in row[:k] and len(set(row[:k])) == k
    # symmetric up to the batch-order approximation
    sym = tot = 0
    for u in range(0, n, 41):
        for v in adj0[u][adj0[u] >= 0]:
            tot += 1
            sym += u in adj0[v]
    assert sym / tot > 0.97, sym / tot
    indeg = np.bincount(adj0[adj0 >= 0].ravel(), minlength=n)
    assert (indeg == 0).mean() < 0.02
    up = h._graph.upper_adj.cpu().numpy()
    lv = np.asarray(h._level_list)
    off = h._graph.upper_off.cpu().numpy()
    rows_l1 = off[lv >= 1]
    assert ((up[rows_l1] >= 0).sum(axis=1) == 16).mean() > 0.95          # layer-1 lists are full (diverse selection)
    q = synth.clip_like(300, 128, seed=152, n_store=n)
    qn = q / np.linalg.norm(q, axis=1, keepdims=True)
    xn = x / np.linalg.norm(x, axis=1, keepdims=True)
    truth = np.argsort(-(qn @ xn.T), axis=1)[:, :10]
    _, rows = h.search_arrays(q, 10)
    assert np.mean([len(set(rows[i]) & set(truth[i])) / 10 for i in range(300)]) > 0.9
