"""Batch-search scheduler — replaces the sequential per-query loop behind
``POST /api/search/batch`` (reference src/api/routes.py:621-645) and offers a micro-batcher for
concurrent single ``/api/search`` calls (in the spirit of the reference's ``BatchProcessor``,
src/core/feature_extractor.py:261-354: flush on size or on a short timeout).

The response shapes are the reference's:
  single : list of metadata dicts + 'score' + 'formatted_time'   (video_search_overhaul.py:439-456)
  batch  : [{"query": q, "results": [...], "count": n}, ...]      (routes.py:630-634)
"""

from __future__ import annotations

import threading
import time
from concurrent.futures import Future
from typing import Callable, List, Sequence

import numpy as np


def _formatted_time(ts: float) -> str:
    return f"{int(ts // 60)}m{int(ts % 60)}s"          # video_search_overhaul.py:451-453


class BatchSearchScheduler:
    """One scan launch per batch of text queries.

    `system` is the live orchestrator (anything with `.processor.encode_text_query(str)` and
    `.index` being a `B200FlatIndex`); alternatively pass `encode=` and `index=` directly.
    """

    def __init__(self, system=None, *, encode: Callable[[str], np.ndarray] | None = None, index=None):
        self.encode = encode if encode is not None else system.processor.encode_text_query
        self.index = index if index is not None else system.index

    def search_vectors(self, vectors: np.ndarray, k: int) -> List[List[dict]]:
        hits = self.index.search_batch(vectors, k)
        for res in hits:
            for r in res:
                if 'timestamp' in r:
                    r['formatted_time'] = _formatted_time(r['timestamp'])
        return hits

    def search_batch(self, queries: Sequence[str], k: int = 5) -> List[dict]:
        """The body of /api/search/batch: encode every text, ONE batched search, scatter back."""
        if len(queries) == 0:
            return []
        vecs = np.stack([np.asarray(self.encode(q), dtype=np.float32) for q in queries])
        hits = self.search_vectors(vecs, k)
        return [{"query": q, "results": res, "count": len(res)} for q, res in zip(queries, hits)]

    def single_response(self, query: str, k: int = 5, use_cache: bool = True) -> dict:
        """Full JSON body of POST /api/search (routes.py:589-613): blank query -> ValueError (the handler
        answers 400), `from_cache` echoes the request flag like the reference does."""
        import uuid
        query = query.strip()
        if not query:
            raise ValueError("No query provided")
        t0 = time.time()
        results = self.search_batch([query], k)[0]["results"]
        return {"results": results, "search_time_ms": (time.time() - t0) * 1000, "from_cache": use_cache,
                "query_id": str(uuid.uuid4()), "performance": {"results_count": len(results)}}

    def batch_response(self, queries: Sequence[str], k: int = 5) -> dict:
        """Full JSON body of the handler (routes.py:636-640)."""
        results = self.search_batch(queries, k)
        return {"results": results, "query_count": len(queries),
                "total_results": sum(r["count"] for r in results)}


class MicroBatcher:
    """Coalesces concurrent single-vector searches into batched launches.

    `submit(vector, k)` returns a Future of the per-query hit list.  A background thread flushes
    when `max_batch` requests are waiting or the oldest has waited `max_wait_ms`.
    """

    def __init__(self, index, max_batch: int = 64, max_wait_ms: float = 2.0):
        self.index = index
        self.max_batch = int(max_batch)
        self.max_wait = max_wait_ms / 1e3
        self._pending: list = []
        self._cv = threading.Condition()
        self._stop = False
        self.batches_flushed = 0
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def submit(self, vector: np.ndarray, k: int = 5) -> Future:
        fut: Future = Future()
        with self._cv:
            self._pending.append((np.asarray(vector, dtype=np.float32), int(k), fut, time.monotonic()))
            self._cv.notify()
        return fut

    def _take(self):
        with self._cv:
            while not self._stop:
                if self._pending:
                    age = time.monotonic() - self._pending[0][3]
                    if len(self._pending) >= self.max_batch or age >= self.max_wait:
                        batch, self._pending = self._pending[: self.max_batch], self._pending[self.max_batch:]
                        return batch
                    self._cv.wait(timeout=max(self.max_wait - age, 1e-4))
                else:
                    self._cv.wait(timeout=0.1)
            return None

    def _run(self):
        while True:
            batch = self._take()
            if batch is None:
                return
            try:
                kmax = max(b[1] for b in batch)
                hits = self.index.search_batch(np.stack([b[0] for b in batch]), kmax)
                for (_, k, fut, _), res in zip(batch, hits):
                    fut.set_result(res[:k])
            except Exception as e:  # noqa: BLE001 — deliver the failure to every waiter
                for _, _, fut, _ in batch:
                    if not fut.done():
                        fut.set_exception(e)
            self.batches_flushed += 1

    def close(self):
        with self._cv:
            self._stop = True
            self._cv.notify_all()
        self._thread.join(timeout=2)
