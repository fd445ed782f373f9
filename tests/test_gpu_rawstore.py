"""GPU round trips of the raw store format through the drop-in facades."""
import random

import numpy as np
import pytest

from oracle import compare, exact
from video_quierer_b200.utils import synth

pytestmark = pytest.mark.gpu


def test_flat_index_raw_roundtrip(built_lib, tmp_path):
    from video_quierer_b200.flat_index import B200FlatIndex, LazyRows
    store = synth.clip_like(6000, 96, seed=81)                   # dim 96 -> ld 128: padding columns travel too
    q = synth.clip_like(7, 96, seed=82, n_store=6000)
    a = B200FlatIndex(store_dtype="bf16", rescore=True)
    a.add_frames(store, [f"v{i // 100}.mp4" for i in range(len(store))], np.arange(len(store), dtype=float))
    a.video_hashes = {"v0.mp4": "h"}
    sa, ra = a.search_arrays(q, 10)
    a.save_raw(tmp_path / "flat")
    b = B200FlatIndex()
    b.load_raw(tmp_path / "flat")
    assert isinstance(b.embeddings, LazyRows) and len(b.embeddings) == 6000 and b.store_dtype == "bf16"
    assert np.array_equal(b.embeddings[17], store[17]) and b.video_hashes == {"v0.mp4": "h"}
    sb, rb = b.search_arrays(q, 10)
    assert np.array_equal(ra, rb) and np.array_equal(sa, sb)     # same device bytes -> same answer
    hits = b.search(q[0], 3)
    assert hits[0]["video_name"] == a.search(q[0], 3)[0]["video_name"] and "score" in hits[0]
    # appends after a raw load go to the tail and are found
    extra = synth.clip_like(3, 96, seed=83, n_store=6000)
    for i, e in enumerate(extra):
        b.add_frame(e, "new.mp4", float(i))
    assert len(b.embeddings) == 6003 and b.search(extra[1], 1)[0]["video_name"] == "new.mp4"
    ro, so = exact.exact_search_batch(np.concatenate([store, extra]), q, 10)
    sb, rb = b.search_arrays(q, 10)
    assert compare.check_topk_batch(rb, sb, ro, so) == []
    # the list form of the reference comes back on demand (route handlers pop rows)
    b.materialise()
    b.embeddings.pop(0); b.metadata.pop(0)
    ro, so = exact.exact_search_batch(np.concatenate([store, extra])[1:], q, 10)
    sb, rb = b.search_arrays(q, 10)
    assert compare.check_topk_batch(rb, sb, ro, so) == []


def test_hnsw_index_raw_roundtrip(built_lib, tmp_path):
    from video_quierer_b200.hnsw_index import B200HNSWIndex
    store = synth.clip_like(4000, 64, seed=91)
    random.seed(9)
    h = B200HNSWIndex(dimension=64, ef_search=48)
    ids = [f"vid{i // 50}_{i % 50}" for i in range(len(store))]
    h.add_batch(list(store), ids)
    q = synth.clip_like(20, 64, seed=92, n_store=4000)
    want = [[x["id"] for x in hits] for hits in h.search_batch(list(q), k=10)]
    h.save_raw(tmp_path / "hnsw")
    g = B200HNSWIndex(dimension=8)                               # parameters come from the store
    g.load_raw(tmp_path / "hnsw")
    assert g.dimension == 64 and g.ef_search == 48 and g.size() == 4000 and g.entry_point == h.entry_point
    got = [[x["id"] for x in hits] for hits in g.search_batch(list(q), k=10)]
    assert got == want                                           # same graph + same rows -> same traversal
    (tmp_path / "hnsw" / "adj0.bin").write_bytes(b"\0" * (tmp_path / "hnsw" / "adj0.bin").stat().st_size)
    with pytest.raises(ValueError):
        B200HNSWIndex(dimension=64).load_raw(tmp_path / "hnsw")
