"""Device-side query-similarity cache probe.

Drop-in for the reference's ``QueryResultCache`` (reference src/storage/cache.py:384-488): same
constructor, ``get_cached_results`` / ``cache_results`` / ``invalidate_results``, same cache keys
(``text_query:md5(text):k`` / ``vector_query:md5(vector bytes):k``, :405-410, :436-441) and the
same semantics — a vector query that misses its exact key reuses the results of the MOST similar
cached query vector whose cosine similarity exceeds ``similarity_threshold`` and whose results are
still in the backing cache (:447-478).

What changes is the probe: the reference loops over every cached vector in Python with one
``np.dot`` + two ``np.linalg.norm`` each (:458-465).  Here the cached query vectors of each k live
L2-normalised in a small device matrix and the probe is ONE `vq_scan_topk` launch (the same exact
scan + fused top-k kernel as the frame search), returning the few best candidates in similarity
order; the host then walks them exactly like the reference walks its dict.

The backing ``cache`` is whatever the orchestrator already uses (the reference's
``MultiLevelCache`` / ``SimpleCache``: ``get`` / ``put`` / ``clear``) and is not touched.
"""

from __future__ import annotations

import hashlib
import logging
from typing import Any, Dict, List, Optional

import numpy as np

from . import _lib
from .engine import DeviceStore, Scanner, _require_cuda, as_device_queries

logger = logging.getLogger(__name__)

PROBE_K = 8          # candidates per probe (the reference takes the best one whose results still exist)


class B200QueryResultCache:
    def __init__(self, cache, similarity_threshold: float = 0.95, device=None):
        self.cache = cache
        self.similarity_threshold = similarity_threshold
        self.query_vectors: Dict[str, np.ndarray] = {}      # same public attribute as the reference
        self.device = _require_cuda(device)
        self._scanner = Scanner(self.device)
        self._per_k: Dict[int, dict] = {}                   # k -> {"store": DeviceStore, "keys": [cache_key per row]}
        self.probes = 0

    # ------------------------------------------------------------------ keys (cache.py:405-410)
    @staticmethod
    def _key(query_vector, k: int, query_text: Optional[str]) -> str:
        if query_text:
            return f"text_query:{hashlib.md5(query_text.encode()).hexdigest()}:{k}"
        return f"vector_query:{hashlib.md5(query_vector.tobytes()).hexdigest()}:{k}"

    def get_cached_results(self, query_vector: Any, k: int, query_text: Optional[str] = None) -> Optional[List[Dict]]:
        results = self.cache.get(self._key(query_vector, k, query_text))
        if results is not None:
            return results
        if query_text is None:
            return self._find_similar_cached_query(query_vector, k)
        return None

    def cache_results(self, query_vector: Any, k: int, results: List[Dict], query_text: Optional[str] = None,
                      ttl: Optional[int] = None) -> None:
        cache_key = self._key(query_vector, k, query_text)
        if not query_text and cache_key not in self.query_vectors:
            self.query_vectors[cache_key] = query_vector
            v = np.asarray(query_vector, dtype=np.float32).reshape(1, -1)
            slot = self._per_k.get(k)
            if slot is None or slot["store"].dim != v.shape[1]:
                slot = self._per_k[k] = {"store": DeviceStore(v.shape[1], self.device, keep_fp32=True), "keys": []}
            slot["store"].append(v, _lib.NORM_PLAIN)            # stored unit-norm: the scan returns the cosine
            slot["keys"].append(cache_key)
        self.cache.put(cache_key, results, ttl)

    # ------------------------------------------------------------------ the probe (cache.py:447-478)
    def _find_similar_cached_query(self, query_vector: Any, k: int) -> Optional[List[Dict]]:
        try:
            slot = self._per_k.get(k)
            if slot is None or slot["store"].n == 0:
                return None
            st = slot["store"]
            q = as_device_queries(np.asarray(query_vector, dtype=np.float32), st.dim, self.device)
            kk = min(PROBE_K, st.n)
            scores, rows = self._scanner.scan(st.f32, st.n, st.dim, q, kk, _lib.NORM_PLAIN, "fma")
            self.probes += 1
            for s, r in zip(scores[0].cpu().numpy(), rows[0].cpu().numpy()):
                if r < 0 or not (s > self.similarity_threshold):     # best first: nothing further qualifies
                    break
                results = self.cache.get(slot["keys"][int(r)])
                if results is not None:
                    return results
            return None
        except Exception as e:  # noqa: BLE001 — like the reference (:476-478)
            logger.warning(f"Similarity matching failed: {e}")
            return None

    def invalidate_results(self, video_id: str) -> None:
        """cache.py:480-488: clears everything (and, here, the device-side vectors with it)."""
        logger.info(f"Invalidating cache for video: {video_id}")
        self.cache.clear()
        self.query_vectors.clear()
        self._per_k.clear()
