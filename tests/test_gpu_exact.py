"""GPU parity tests of the exact path, through the C-ABI (ctypes) and the drop-in facade,
against the oracle and the golden outputs of the unmodified reference."""
import numpy as np
import pytest

from oracle import compare, exact
from video_quierer_b200.utils import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(built_lib):
    import torch
    from video_quierer_b200 import _lib, engine
    return engine, _lib, torch


def _scan(eng, store_np, queries_np, k, dtype="fp32", path="fma", norm=None):
    engine, _lib, torch = eng
    st = engine.DeviceStore(store_np.shape[1], keep_fp32=True, keep_bf16=(dtype == "bf16"))
    st.append(store_np)
    sc = engine.Scanner()
    q = engine.as_device_queries(queries_np, store_np.shape[1], st.device)
    s, r = sc.scan(st.view(dtype), st.n, st.dim, q, k, _lib.NORM_EPS if norm is None else norm, path)
    torch.cuda.synchronize()
    return s.cpu().numpy(), r.cpu().numpy(), sc


def test_golden_small_all_k(eng, golden):
    g = golden("exact_small.npz")
    store = g["store_f16"].astype(np.float32)
    queries = g["queries_f16"].astype(np.float32)
    for k in (1, 10, 50):
        s, r, _ = _scan(eng, store, queries, k)
        assert compare.check_topk_batch(r, s, g[f"rows_k{k}"], g[f"scores_k{k}"]) == []
        assert compare.id_match_fraction(r, g[f"rows_k{k}"]) == 1.0   # no ties in this fixture


@pytest.mark.parametrize("b", [1, 2, 3, 5, 8, 13, 16, 17, 32, 40])
def test_batch_sizes_vs_oracle(eng, b):
    store = synth.gauss(5000, 512, seed=3)
    queries = np.random.default_rng(4).standard_normal((b, 512), dtype=np.float32)
    s, r, sc = _scan(eng, store, queries, 10)
    ro, so = exact.exact_search_batch(store, queries, 10)
    assert compare.check_topk_batch(r, s, ro, so) == []
    assert sc.last_path == "scan_fma_f32" and sc.last_launches >= 3


@pytest.mark.parametrize("n,dim", [(1, 512), (3, 64), (127, 128), (128, 768), (129, 96), (4097, 512), (10001, 100)])
def test_ragged_shapes(eng, n, dim):
    store = synth.gauss(n, dim, seed=5)
    queries = np.random.default_rng(6).standard_normal((4, dim), dtype=np.float32)
    k = min(10, n)
    s, r, _ = _scan(eng, store, queries, k)
    ro, so = exact.exact_search_batch(store, queries, k)
    assert compare.check_topk_batch(r, s, ro, so) == []


def test_c1_full_config_golden(eng, golden):
    """BASELINE config 1 against the reference's own outputs."""
    g = golden("exact_c1.npz")
    store = synth.gauss(10000, 512, seed=synth.STORE_SEED)
    queries = np.random.default_rng(synth.QUERY_SEED).standard_normal((100, 512), dtype=np.float32)
    assert synth.sha256_of(store) == str(g["store_sha"])
    s, r, _ = _scan(eng, store, queries, 10)
    assert compare.check_topk_batch(r, s, g["rows"], g["scores"]) == []
    assert compare.id_match_fraction(r, g["rows"]) == 1.0
    assert np.max(np.abs(s - g["scores"]) / np.abs(g["scores"])) < 1e-5


def test_k100_and_large_k(eng):
    store = synth.clip_like(20000, 256, seed=7)
    queries = synth.clip_like(9, 256, seed=8, n_store=20000)
    for k in (100, 257, 1024):
        s, r, _ = _scan(eng, store, queries, k)
        ro, so = exact.exact_search_batch(store, queries, k)
        assert compare.check_topk_batch(r, s, ro, so) == []


def test_ties_deterministic_rule(eng):
    store = synth.with_ties(4000, 128, seed=9)
    queries = store[[10, 200, 3999]] + 0.0
    s, r, _ = _scan(eng, store, queries, 20)
    ro, so = exact.exact_search_batch(store, queries, 20)
    assert compare.check_topk_batch(r, s, ro, so) == []
    # engine rule: equal scores are listed by ascending row
    for b in range(len(r)):
        for i in range(19):
            if s[b, i] == s[b, i + 1]:
                assert r[b, i] < r[b, i + 1]


def test_zero_query_and_k_gt_n(eng):
    store = synth.gauss(7, 64, seed=10)
    s, r, _ = _scan(eng, store, np.zeros((1, 64), np.float32), 5)
    assert np.all(s == 0.0) and list(r[0]) == [0, 1, 2, 3, 4]
    s, r, _ = _scan(eng, store, store[:2], 50)
    assert np.all(r[:, :7] >= 0) and np.all(r[:, 7:] == -1)


def test_bf16_store_scan_matches_bf16_oracle(eng):
    """bf16 store, fp32 accumulate: exactly the oracle run on the bf16-rounded rows."""
    _, _, torch = eng
    store = synth.clip_like(6000, 512, seed=11)
    queries = synth.clip_like(6, 512, seed=12, n_store=6000)
    s, r, sc = _scan(eng, store, queries, 10, dtype="bf16")
    rounded = torch.from_numpy(store).to(torch.bfloat16).to(torch.float32).numpy()
    ro, so = exact.exact_search_batch(rounded, queries, 10)
    assert compare.check_topk_batch(r, s, ro, so) == []
    assert sc.last_path == "scan_fma_bf16"


def test_linearity_property_full_size(eng):
    """Size-independent property at a BASELINE-scale store: score(q1+q2-normalised) ordering is
    consistent with an independent recomputation of the returned rows on the host."""
    n = 200_000
    store = synth.gauss(n, 512, seed=13)
    queries = np.random.default_rng(14).standard_normal((3, 512), dtype=np.float32)
    s, r, _ = _scan(eng, store, queries, 10)
    qn = queries / (np.linalg.norm(queries, axis=1, keepdims=True) + 1e-10)
    for b in range(3):
        host = store[r[b]].astype(np.float64) @ qn[b].astype(np.float64)
        assert np.allclose(host, s[b], rtol=1e-5, atol=1e-7)
        assert np.all(np.diff(s[b]) <= 0)
        # nothing outside the returned set beats the k-th score
        full = store @ qn[b]
        assert np.sum(full > s[b, -1] + 1e-6) <= 9


@pytest.mark.parametrize("store_dtype", ["bf16", "fp32"])
def test_facade_matches_reference_dicts(eng, golden, store_dtype):
    from video_quierer_b200.flat_index import B200FlatIndex
    g = golden("exact_small.npz")
    store = g["store_f16"].astype(np.float32)
    queries = g["queries_f16"].astype(np.float32)
    idx = B200FlatIndex(store_dtype=store_dtype)
    assert idx.search(queries[0], 5) == []
    for i, x in enumerate(store):
        idx.add_frame(x, f"v{i % 7}.mp4", float(i) * 0.5)
    res = idx.search(queries[1], 3)
    assert sorted(res[0].keys()) == list(g["dict_keys"])
    assert [[x["frame_id"], x["timestamp"]] for x in res] == g["dict_example"].tolist()
    assert [x["video_name"] for x in res] == list(g["dict_video"])
    assert isinstance(res[0]["score"], float)
    # batch = the /api/search/batch loop, one launch
    batch = idx.search_batch(queries[:8], 10)
    assert [[h["frame_id"] for h in hits] for hits in batch] == g["rows_k10"][:8].tolist()
    # mutation through the attribute surface the route handlers use
    idx.embeddings.pop(int(g["rows_k10"][0][0])); idx.metadata.pop(int(g["rows_k10"][0][0]))
    res2 = idx.search(queries[0], 1)
    assert res2[0]["frame_id"] == int(g["rows_k10"][0][1])
    idx.embeddings = []; idx.metadata = []
    assert idx.search(queries[0], 5) == []


def test_facade_save_load_roundtrip(eng, tmp_path):
    from video_quierer_b200.flat_index import B200FlatIndex
    store = synth.gauss(300, 64, seed=15)
    a = B200FlatIndex()
    a.add_frames(store, ["a.mp4"] * 300, [float(i) for i in range(300)])
    a.video_hashes["a.mp4"] = "h"
    p = tmp_path / "video_search_cache.pkl"
    assert a.save_to_disk(p) is True
    b = B200FlatIndex()
    assert b.load_from_disk(p) is True and b.video_hashes == {"a.mp4": "h"}
    assert b.load_from_disk(tmp_path / "missing.pkl") is False
    q = store[17]
    assert [h["frame_id"] for h in a.search(q, 5)] == [h["frame_id"] for h in b.search(q, 5)]


def test_bf16_rescore_mode_is_exact(eng):
    from video_quierer_b200.flat_index import B200FlatIndex
    store = synth.clip_like(8000, 512, seed=16)
    queries = synth.clip_like(5, 512, seed=17, n_store=8000)
    idx = B200FlatIndex(store_dtype="bf16", rescore=True)
    idx.add_frames(store, ["a.mp4"] * len(store), np.arange(len(store), dtype=float))
    s, r = idx.search_arrays(queries, 10)
    ro, so = exact.exact_search_batch(store, queries, 10)
    assert compare.check_topk_batch(r, s, ro, so) == []


# ----------------------------------------------------------------------------- tcgen05 path
def _bf16_round(a, torch):
    return torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).to(torch.float32).numpy()


@pytest.mark.parametrize("n,dim,b,k", [(128, 64, 1, 1), (1000, 128, 17, 10), (5000, 512, 32, 10), (20000, 512, 130, 10),
                                       (4097, 256, 5, 32), (60000, 768, 300, 16), (777, 100, 3, 7)])
def test_mma_path_vs_oracle_on_bf16_operands(eng, n, dim, b, k):
    """tcgen05 scan: bf16 x bf16 products are exact and accumulate in fp32, so against the oracle
    fed the same bf16-rounded operands the ids are identical and scores agree to ~1e-6."""
    engine, _lib, torch = eng
    store = _bf16_round(synth.gauss(n, dim, seed=21), torch)
    q = synth.gauss(b, dim, seed=22)
    q = _bf16_round(q, torch)
    st = engine.DeviceStore(dim, keep_fp32=False, keep_bf16=True)
    st.append(store)
    sc = engine.Scanner()
    s, r = sc.scan(st.view("bf16"), st.n, st.dim, engine.as_device_queries(q, dim, st.device), k, _lib.NORM_NONE, "mma")
    torch.cuda.synchronize()
    assert sc.last_path == "scan_mma_bf16"
    ro, so = exact.scan_prenormalised_f64(store, q, k)
    assert compare.check_topk_batch(r.cpu().numpy(), s.cpu().numpy(), ro, so) == []


def test_mma_auto_dispatch_and_fp32_never_uses_tensor_path(eng):
    engine, _lib, torch = eng
    store = synth.gauss(3000, 128, seed=23)
    st = engine.DeviceStore(128, keep_fp32=True, keep_bf16=True)
    st.append(store)
    sc = engine.Scanner()
    q = engine.as_device_queries(store[:4], 128, st.device)
    sc.scan(st.view("bf16"), st.n, 128, q, 10, _lib.NORM_EPS, "auto")
    assert sc.last_path == "scan_mma_bf16"
    sc.scan(st.view("fp32"), st.n, 128, q, 10, _lib.NORM_EPS, "auto")
    assert sc.last_path == "scan_fma_f32"                 # parity: no silent tf32
    sc.scan(st.view("bf16"), st.n, 128, q, 100, _lib.NORM_EPS, "auto")
    assert sc.last_path == "scan_fma_bf16"                # k beyond the register lists falls back to FMA
    with pytest.raises(_lib.VQError):
        sc.scan(st.view("fp32"), st.n, 128, q, 10, _lib.NORM_EPS, "mma")


@pytest.mark.parametrize("gen,n,b", [("gauss", 50000, 40), ("clip", 30000, 200)])
def test_exact_single_pass_facade(eng, gen, n, b):
    """The default facade mode: one tensor-core pass (vq_search_exact) == the oracle, no fallback."""
    from video_quierer_b200.flat_index import B200FlatIndex
    store = synth.gauss(n, 512, seed=31) if gen == "gauss" else synth.clip_like(n, 512, seed=31)
    queries = np.random.default_rng(32).standard_normal((b, 512), dtype=np.float32) if gen == "gauss" \
        else synth.clip_like(b, 512, seed=32, n_store=n)
    idx = B200FlatIndex()
    idx.add_frames(store, ["a.mp4"] * n, np.arange(n, dtype=float))
    s, r = idx.search_arrays(queries, 10)
    ro, so = exact.exact_search_batch(store, queries, 10)
    assert compare.check_topk_batch(r, s, ro, so) == []
    # clustered rows produce scores closer than fp32 summation-order noise: those may swap places (north-star:
    # ties within 1e-5 are excluded from the id comparison, which is what check_topk_batch applies)
    assert compare.id_match_fraction(r, ro) >= (1.0 if gen == "gauss" else 0.99)
    assert idx.last_scan_path == "scan_mma_bf16<exact>+finish"
    assert idx.stats == {"exact_queries": b, "overflow_queries": 0}


def _exact(eng, store, q, k, norm=None):
    engine, _lib, torch = eng
    st = engine.DeviceStore(store.shape[1], keep_fp32=True, keep_bf16=True)
    st.append(store)
    sc = engine.Scanner()
    qd = engine.as_device_queries(q, store.shape[1], st.device)
    if k <= 64:
        s, r, over = sc.exact(st, qd, k, _lib.NORM_EPS if norm is None else norm)
        assert sc.last_path == "scan_mma_bf16<exact>+finish"
    else:           # the router: single pass where the gather cannot overflow, else sample pass + collect pass
        from video_quierer_b200.flat_index import exact_search
        s, r, over = exact_search(sc, st, qd, k)
        assert sc.last_path.startswith("scan_mma_bf16<")
    torch.cuda.synchronize()
    return s.cpu().numpy(), r.cpu().numpy(), over.cpu().numpy(), st


@pytest.mark.parametrize("gen,n,dim,b,k", [
    ("gauss", 1, 512, 1, 1), ("gauss", 7, 64, 3, 5), ("gauss", 127, 128, 4, 10), ("gauss", 129, 96, 4, 10),
    ("gauss", 4097, 512, 5, 10), ("gauss", 10001, 100, 17, 10), ("clip", 20000, 512, 130, 10),
    ("clip", 60000, 768, 300, 16), ("gauss", 40000, 256, 33, 32), ("clip", 50000, 512, 24, 50),
    ("gauss", 30000, 512, 9, 64), ("clip", 300000, 512, 40, 10), ("ties", 8000, 128, 12, 20),
    ("gauss", 30000, 768, 12, 100), ("gauss", 120000, 768, 12, 100), ("clip", 90000, 512, 40, 100), ("gauss", 200000, 512, 300, 128),
    ("clip", 150000, 768, 260, 100)])
def test_vq_search_exact_vs_oracle(eng, gen, n, dim, b, k):
    """vq_search_exact over ragged shapes, every list width (k <= 16 / 32 / 64), bootstrap as a separate pass (one query
    tile) and inside the scan (several), clustered data and exact duplicates; k = 100 / 128 (BASELINE config 4) through
    the router (listless single pass on small stores, sample + collect passes with the exact_finish stage beyond):
    ids and scores of the oracle."""
    store = {"gauss": synth.gauss, "clip": synth.clip_like, "ties": synth.with_ties}[gen](n, dim, seed=101)
    if gen == "ties":
        q = store[np.arange(b) * 7 % n] + 0.0
    else:
        q = synth.gauss(b, dim, seed=102) if gen == "gauss" else synth.clip_like(b, dim, seed=102, n_store=n)
    s, r, over, _ = _exact(eng, store, q, k)
    assert not over.any()
    ro, so = exact.exact_search_batch(store, q, k)
    assert compare.check_topk_batch(r, s, ro, so) == []
    kk = min(k, n)
    assert np.all(r[:, :kk] >= 0) and np.all(r[:, kk:] == -1)
    if gen == "ties":                     # engine rule: equal scores are listed by ascending row
        for bi in range(b):
            for i in range(kk - 1):
                if s[bi, i] == s[bi, i + 1]:
                    assert r[bi, i] < r[bi, i + 1]


def test_vq_search_exact_rows_not_unit_norm(eng):
    """The reference never normalises stored rows (video_search_overhaul.py:33,53): rows of norm 0.2 .. 6 —
    the error bound scales with the tracked row norms and the result stays exact."""
    rng = np.random.default_rng(111)
    store = synth.clip_like(30000, 512, seed=110) * rng.uniform(0.2, 6.0, size=(30000, 1)).astype(np.float32)
    q = synth.clip_like(20, 512, seed=112, n_store=30000) * 3.0
    s, r, over, st = _exact(eng, store, q, 10)
    assert not over.any()
    ro, so = exact.exact_search_batch(store, q, 10)
    assert compare.check_topk_batch(r, s, ro, so) == []
    b0, b1 = st.bounds.cpu().numpy().tolist()
    bf = eng[2].from_numpy(store).to(eng[2].bfloat16).to(eng[2].float32).numpy()
    assert abs(b0 - np.linalg.norm(bf, axis=1).max()) < 1e-3 * b0
    assert abs(b1 - np.linalg.norm(bf - store, axis=1).max()) < 1e-3 * b1
    assert st.max_row_norm() >= np.linalg.norm(store, axis=1).max() * 0.999


def test_vq_search_exact_near_duplicates_and_overflow(eng):
    """500 rows closer together than the bf16 resolution: all of them are gathered and re-scored, the result
    is exact without any fallback.  A zero query ties with EVERY row: the gather overflows, the flag is set
    and the facade answers from the fp32 FMA scan."""
    from video_quierer_b200.flat_index import B200FlatIndex
    rng = np.random.default_rng(41)
    base = synth.gauss(1, 256, seed=40)[0]
    near = base[None, :] + 1e-4 * rng.standard_normal((500, 256)).astype(np.float32)
    near /= np.linalg.norm(near, axis=1, keepdims=True)
    store = np.concatenate([synth.gauss(9000, 256, seed=42), near.astype(np.float32)])
    q = np.stack([base, synth.gauss(1, 256, seed=43)[0], np.zeros(256, np.float32)])
    ro, so = exact.exact_search_batch(store, q, 10)
    s, r, over, _ = _exact(eng, store, q, 10)
    assert over.tolist() == [0, 0, 1]
    assert compare.check_topk_batch(r[:2], s[:2], ro[:2], so[:2]) == []
    idx = B200FlatIndex()
    idx.add_frames(store, ["a.mp4"] * len(store), np.arange(len(store), dtype=float))
    s2, r2 = idx.search_arrays(q, 10)
    assert compare.check_topk_batch(r2, s2, ro, so) == []
    assert np.all(s2[2] == 0.0) and r2[2].tolist() == list(range(10))
    assert idx.stats["overflow_queries"] == 1


def test_two_stage_low_level_still_certifies(eng):
    """vq_search_two_stage (kept for k_cand experiments) with the sound eps: certified queries are exact."""
    engine, _lib, torch = eng
    from video_quierer_b200.flat_index import two_stage_search
    store = synth.gauss(50000, 512, seed=31)
    q = synth.gauss(40, 512, seed=32)
    st = engine.DeviceStore(512, keep_fp32=True, keep_bf16=True)
    st.append(store)
    sc = engine.Scanner()
    s, r, bad = two_stage_search(sc, st, engine.as_device_queries(q, 512, st.device), 10)
    ok = bad.cpu().numpy() == 0
    ro, so = exact.exact_search_batch(store, q, 10)
    assert compare.check_topk_batch(r.cpu().numpy()[ok], s.cpu().numpy()[ok], ro[ok], so[ok]) == []


def test_collect_pass_resolves_uncertified_and_overflows_to_fma(eng):
    """vq_search_collect: (1) with the threshold s_k - eps it returns the exact top-k although hundreds
    of rows sit inside the bf16 resolution; (2) more rows than `cap` above the threshold -> overflow
    flag, and the facade then answers from the fp32 FMA scan."""
    engine, _lib, torch = eng
    from video_quierer_b200.flat_index import BF16_SCORE_EPS, B200FlatIndex
    rng = np.random.default_rng(51)
    base = synth.gauss(1, 256, seed=50)[0]
    near = base[None, :] + 2e-4 * rng.standard_normal((6000, 256)).astype(np.float32)
    near /= np.linalg.norm(near, axis=1, keepdims=True)
    store = np.concatenate([synth.gauss(30000, 256, seed=52), near.astype(np.float32)])
    rng.shuffle(store)
    q = np.stack([base, synth.gauss(1, 256, seed=53)[0]])
    ro, so = exact.exact_search_batch(store, q, 10)
    st = engine.DeviceStore(256, keep_fp32=True, keep_bf16=True)
    st.append(store)
    sc = engine.Scanner()
    qd = engine.as_device_queries(q, 256, st.device)
    thr = torch.from_numpy(so[:, 9].astype(np.float32) - np.float32(BF16_SCORE_EPS)).to(st.device)
    # (1) cap large enough for the 6000 near-duplicates
    s, r, over = sc.collect(st.bf16, st.f32, st.n, 256, qd, 10, thr, cap=8192)
    assert over.cpu().numpy().tolist() == [0, 0] and sc.last_path.startswith("scan_mma_bf16<collect>")
    assert compare.check_topk_batch(r.cpu().numpy(), s.cpu().numpy(), ro, so) == []
    # (2) cap too small for query 0 -> overflow flagged, query 1 still exact
    s, r, over = sc.collect(st.bf16, st.f32, st.n, 256, qd, 10, thr, cap=1024)
    assert over.cpu().numpy().tolist() == [1, 0]
    assert compare.check_topk_batch(r.cpu().numpy()[1:], s.cpu().numpy()[1:], ro[1:], so[1:]) == []
    # the facade: 6000 near-duplicates overflow the 4096-slot gather of query 0 -> fp32 FMA scan, still exact
    idx = B200FlatIndex()
    idx.add_frames(store, ["a.mp4"] * len(store), np.arange(len(store), dtype=float))
    s2, r2 = idx.search_arrays(q, 10)
    assert compare.check_topk_batch(r2, s2, ro, so) == []
    assert idx.stats["overflow_queries"] == 1


@pytest.mark.parametrize("gen", ["gauss", "clip"])
def test_k50_uses_tensor_path(eng, gen):
    """k = 50 (the API maximum, src/api/routes.py:56-59) is served by the single-pass exact search (64-entry
    register lists); the fp32 FMA scan is not needed."""
    engine, _lib, torch = eng
    from video_quierer_b200.flat_index import B200FlatIndex
    n, b, k = 40000, 24, 50
    store = synth.gauss(n, 512, seed=61) if gen == "gauss" else synth.clip_like(n, 512, seed=61)
    queries = synth.gauss(b, 512, seed=62) if gen == "gauss" else synth.clip_like(b, 512, seed=62, n_store=n)
    idx = B200FlatIndex(store_dtype="bf16", rescore=True)
    idx.add_frames(store, ["a.mp4"] * n, np.arange(n, dtype=float))
    s, r = idx.search_arrays(queries, k)
    ro, so = exact.exact_search_batch(store, queries, k)
    assert compare.check_topk_batch(r, s, ro, so) == []
    assert idx.last_scan_path == "scan_mma_bf16<exact>+finish" and idx.stats["overflow_queries"] == 0


def test_query_similarity_cache_probe_matches_reference_semantics(eng):
    """B200QueryResultCache vs a straight restatement of the reference loop (cache.py:447-478)."""
    from video_quierer_b200.query_cache import B200QueryResultCache

    class DictCache:
        def __init__(self): self.d = {}
        def get(self, k): return self.d.get(k)
        def put(self, k, v, ttl=None): self.d[k] = v; return True
        def clear(self): self.d.clear()

    rng = np.random.default_rng(71)
    base = synth.gauss(300, 512, seed=70)
    qc = B200QueryResultCache(DictCache(), similarity_threshold=0.95)
    for i, v in enumerate(base):
        qc.cache_results(v, 5, [{"video_id": f"v{i}"}])
    qc.cache_results(base[7], 10, [{"video_id": "k10"}])                       # another k: separate pool
    assert qc.get_cached_results(base[3], 5) == [{"video_id": "v3"}]           # exact key hit, no probe
    assert qc.probes == 0
    near = base[42] + 0.1 * rng.standard_normal(512).astype(np.float32) / np.sqrt(512)    # cosine ~0.995
    far = base[42] + 0.6 * rng.standard_normal(512).astype(np.float32) / np.sqrt(512)     # cosine ~0.86
    def ref_probe(q, k):
        best, out = 0, None
        for key, v in qc.query_vectors.items():
            if not key.endswith(f":{k}"):
                continue
            sim = np.dot(q, v) / (np.linalg.norm(q) * np.linalg.norm(v))
            if sim > 0.95 and sim > best and qc.cache.get(key) is not None:
                best, out = sim, qc.cache.get(key)
        return out
    assert qc.get_cached_results(near * 3.0, 5) == ref_probe(near * 3.0, 5) == [{"video_id": "v42"}]
    assert qc.get_cached_results(far, 5) is None and ref_probe(far, 5) is None
    assert qc.get_cached_results(base[7] * 1.0001, 10) == [{"video_id": "k10"}]
    assert qc.get_cached_results(near, 7) is None                              # no pool for k = 7
    # best candidate evicted from the backing cache -> the next best above the threshold is used
    twin = base[42] + 0.05 * rng.standard_normal(512).astype(np.float32) / np.sqrt(512)
    qc.cache_results(twin, 5, [{"video_id": "twin"}])
    key42 = [k for k, v in qc.query_vectors.items() if v is base[42] or np.array_equal(v, base[42])][0]
    del qc.cache.d[key42]
    assert qc.get_cached_results(near, 5) == ref_probe(near, 5) == [{"video_id": "twin"}]
    qc.invalidate_results("v1")
    assert qc.get_cached_results(near, 5) is None and qc.query_vectors == {}


@pytest.mark.parametrize("gen,n,dim,k", [("gauss", 120000, 768, 100), ("clip", 90000, 512, 100), ("gauss", 120000, 256, 200)])
def test_large_k_sampled_collect_is_exact(eng, gen, n, dim, k):
    """k beyond the register lists (BASELINE config 4: k = 100, dim 768): sampled fp32 bound + one
    tensor-core collect pass + fp32 re-score of everything gathered == the oracle."""
    engine, _lib, torch = eng
    from video_quierer_b200.flat_index import two_stage_search
    store = synth.gauss(n, dim, seed=91) if gen == "gauss" else synth.clip_like(n, dim, seed=91)
    q = synth.gauss(12, dim, seed=92) if gen == "gauss" else synth.clip_like(12, dim, seed=92, n_store=n)
    st = engine.DeviceStore(dim, keep_fp32=True, keep_bf16=True)
    st.append(store)
    sc = engine.Scanner()
    s, r, bad = two_stage_search(sc, st, engine.as_device_queries(q, dim, st.device), k)
    assert int(bad.sum()) == 0 and sc.last_path.startswith("scan_mma_bf16<collect>")
    ro, so = exact.exact_search_batch(store, q, k)
    assert compare.check_topk_batch(r.cpu().numpy(), s.cpu().numpy(), ro, so) == []


def test_facade_chunks_very_large_batches(eng):
    """More queries than one launch chain takes (MAX_BATCH): the facade chunks, results unchanged."""
    from video_quierer_b200 import flat_index
    from video_quierer_b200.flat_index import B200FlatIndex
    store = synth.gauss(3000, 64, seed=95)
    q = synth.gauss(700, 64, seed=96)
    idx = B200FlatIndex(store_dtype="bf16", rescore=True)
    idx.add_frames(store, ["a.mp4"] * len(store), np.arange(len(store), dtype=float))
    s0, r0 = idx.search_arrays(q, 5)
    old = flat_index.MAX_BATCH
    try:
        flat_index.MAX_BATCH = 256
        s1, r1 = idx.search_arrays(q, 5)
    finally:
        flat_index.MAX_BATCH = old
    assert np.array_equal(r0, r1) and np.array_equal(s0, s1)
    ro, so = exact.exact_search_batch(store, q, 5)
    assert compare.check_topk_batch(r1, s1, ro, so) == []


def test_empty_store_fill_is_minus_inf(eng):
    """n == 0 through the C-ABI: rows -1 and scores -inf as the header promises (not a NaN pattern)."""
    engine, _lib, torch = eng
    sc = engine.Scanner()
    mat = torch.zeros((1, 64), dtype=torch.float32, device=sc.device)
    q = torch.randn((3, 64), device=sc.device)
    s, r = sc.scan(mat, 0, 64, q, 5)
    torch.cuda.synchronize()
    assert bool((r == -1).all()) and bool(torch.isinf(s).all()) and bool((s < 0).all())


def test_add_frames_takes_cuda_tensors(eng):
    """the encoder's output stays on the device (no .cpu().numpy() round trip); host appends
    and device appends interleave, the list surface keeps working."""
    engine, _lib, torch = eng
    from video_quierer_b200.flat_index import B200FlatIndex, DeviceRows
    store = synth.clip_like(9000, 512, seed=121)
    q = synth.clip_like(6, 512, seed=122, n_store=9000)
    idx = B200FlatIndex()
    idx.add_frames(store[:2000], ["a.mp4"] * 2000, np.arange(2000, dtype=float))                  # host rows first
    idx.add_frames(torch.from_numpy(store[2000:7000]).cuda(), ["b.mp4"] * 5000, np.arange(5000, dtype=float))
    idx.add_frame(store[7000], "c.mp4", 0.0)                                                       # reference call
    idx.add_frames(torch.from_numpy(store[7001:]).cuda(), ["d.mp4"] * 1999, np.arange(1999, dtype=float))
    assert isinstance(idx.embeddings, DeviceRows) and len(idx.embeddings) == 9000 == len(idx.metadata)
    assert np.array_equal(idx.embeddings[6999], store[6999]) and np.array_equal(idx.embeddings[-1], store[-1])
    s, r = idx.search_arrays(q, 10)
    ro, so = exact.exact_search_batch(store, q, 10)
    assert compare.check_topk_batch(r, s, ro, so) == []
    hit = idx.search(store[7000], 1)[0]
    assert hit["video_name"] == "c.mp4" and hit["frame_id"] == 7000
    idx.materialise()                                      # back to the reference's list form for handlers that pop rows
    idx.embeddings.pop(7000); idx.metadata.pop(7000)
    assert idx.search(store[7000], 1)[0]["frame_id"] != 7000 or idx.search(store[7000], 1)[0]["video_name"] != "c.mp4"


def test_scheduler_and_microbatcher_over_real_index(eng):
    """SURVEY.md 8(f) rank 1 on the device: BatchSearchScheduler (the body of /api/search/batch) and the MicroBatcher
    over a real B200FlatIndex return what the reference's sequential loop returns."""
    from video_quierer_b200.flat_index import B200FlatIndex
    from video_quierer_b200.scheduler import BatchSearchScheduler, MicroBatcher
    store = synth.clip_like(12000, 512, seed=131)
    idx = B200FlatIndex()
    idx.add_frames(store, [f"v{i // 500}.mp4" for i in range(12000)], [float(i % 500) for i in range(12000)])
    texts = [f"query {i}" for i in range(40)]
    vecs = {t: synth.clip_like(1, 512, seed=200 + i, n_store=12000)[0] for i, t in enumerate(texts)}
    sched = BatchSearchScheduler(encode=lambda t: vecs[t], index=idx)
    body = sched.batch_response(texts, 5)
    assert body["query_count"] == 40 and body["total_results"] == 200
    for t, entry in zip(texts, body["results"]):
        ro, so = exact.exact_search(store, vecs[t], 5)
        assert entry["query"] == t and entry["count"] == 5
        assert compare.check_topk_batch(np.array([[h["frame_id"] for h in entry["results"]]]),
                                        np.array([[h["score"] for h in entry["results"]]], dtype=np.float32), ro[None], so[None]) == []
        assert all(h["formatted_time"] == f"{int(h['timestamp'] // 60)}m{int(h['timestamp'] % 60)}s" for h in entry["results"])
    single = sched.single_response(texts[3], 5)
    assert [h["frame_id"] for h in single["results"]] == [h["frame_id"] for h in body["results"][3]["results"]]
    mb = MicroBatcher(idx, max_batch=16, max_wait_ms=5.0)
    try:
        futs = [mb.submit(vecs[t], 5) for t in texts]
        got = [f.result(timeout=60) for f in futs]
    finally:
        mb.close()
    assert [[h["frame_id"] for h in hits] for hits in got] == [[h["frame_id"] for h in e["results"]] for e in body["results"]]
    assert mb.batches_flushed < 40                          # requests were coalesced
