"""CPU tests of the host-side batch scheduler / micro-batcher (response shapes of the reference's
/api/search and /api/search/batch handlers, src/api/routes.py:589-645) with a stand-in index, and of
the list facade that records the route handlers' direct mutations of `index.embeddings`."""
import threading
import time

import numpy as np
import pytest

from video_quierer_b200.scheduler import BatchSearchScheduler, MicroBatcher


class FakeIndex:
    """search_batch(vectors, k) -> per query: k hits whose score encodes the query's first component."""

    def __init__(self):
        self.calls = []
        self.lock = threading.Lock()

    def search_batch(self, vectors, k):
        vectors = np.asarray(vectors)
        with self.lock:
            self.calls.append(len(vectors))
        return [[{"video_name": "v.mp4", "timestamp": 61.5 + i, "frame_id": i, "score": float(v[0]) - i}
                 for i in range(k)] for v in vectors]


def _encode(text):
    return np.full(4, float(len(text)), dtype=np.float32)


def test_batch_response_shape_and_single_launch():
    idx = FakeIndex()
    s = BatchSearchScheduler(encode=_encode, index=idx)
    out = s.batch_response(["a", "bbb", "cc"], k=2)
    assert idx.calls == [3]                                   # ONE search for the whole batch (routes.py:627-634 loops)
    assert out["query_count"] == 3 and out["total_results"] == 6
    assert [r["query"] for r in out["results"]] == ["a", "bbb", "cc"]
    assert all(r["count"] == 2 and len(r["results"]) == 2 for r in out["results"])
    hit = out["results"][1]["results"][0]
    assert hit["score"] == 3.0 and hit["formatted_time"] == "1m1s"   # video_search_overhaul.py:451-453
    assert s.search_batch([], k=3) == []


def test_single_response_shape():
    s = BatchSearchScheduler(encode=_encode, index=FakeIndex())
    out = s.single_response("  car  ", k=3, use_cache=False)
    assert sorted(out) == ["from_cache", "performance", "query_id", "results", "search_time_ms"]
    assert out["from_cache"] is False and out["performance"] == {"results_count": 3} and len(out["results"]) == 3
    with pytest.raises(ValueError):
        s.single_response("   ")


def test_micro_batcher_coalesces_concurrent_requests():
    idx = FakeIndex()
    mb = MicroBatcher(idx, max_batch=16, max_wait_ms=30.0)
    try:
        futs = [mb.submit(np.full(4, float(i), np.float32), k=1 + i % 3) for i in range(16)]
        res = [f.result(timeout=5) for f in futs]
        assert [len(r) for r in res] == [1 + i % 3 for i in range(16)]        # every caller gets its own k
        assert [r[0]["score"] for r in res] == [float(i) for i in range(16)]  # and its own query's hits
        assert sum(idx.calls) == 16 and len(idx.calls) <= 2                   # coalesced, not 16 launches
        t0 = time.monotonic()
        assert len(mb.submit(np.zeros(4, np.float32), k=2).result(timeout=5)) == 2
        assert time.monotonic() - t0 < 1.0                                     # the timeout flushes a lone request
    finally:
        mb.close()


def test_micro_batcher_propagates_errors():
    class Boom:
        def search_batch(self, v, k):
            raise RuntimeError("device lost")
    mb = MicroBatcher(Boom(), max_batch=4, max_wait_ms=1.0)
    try:
        with pytest.raises(RuntimeError):
            mb.submit(np.zeros(4, np.float32)).result(timeout=5)
    finally:
        mb.close()


def test_embedding_list_tracks_the_valid_prefix():
    """Route handlers mutate `index.embeddings` in place (routes.py:754-762: pop; :979: rebind to []).
    The facade only re-uploads rows at or after the first mutated position."""
    from video_quierer_b200.flat_index import EmbeddingList
    e = EmbeddingList([np.zeros(2)] * 10)
    e.valid_prefix = 10
    e.append(np.ones(2)); e.extend([np.ones(2)] * 2)
    assert e.valid_prefix == 10 and len(e) == 13              # appends never invalidate uploaded rows
    e.pop(7)
    assert e.valid_prefix == 7
    e.valid_prefix = len(e)
    e.pop()                                                   # last row
    assert e.valid_prefix == len(e)
    del e[3]
    assert e.valid_prefix == 3
    e[1] = np.ones(2)
    assert e.valid_prefix == 1
    e.valid_prefix = len(e); e.insert(0, np.ones(2)); assert e.valid_prefix == 0
    e.valid_prefix = len(e); e.clear(); assert e.valid_prefix == 0 and len(e) == 0
    e += [np.ones(2)]
    assert isinstance(e, EmbeddingList) and len(e) == 1
