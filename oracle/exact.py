"""ORACLE — test infrastructure only, never imported by the product path.

CPU restatement (numpy) of the reference's *exact* frame search, the live path behind
``/api/search``:  ``SimpleVideoIndex.search``  (reference ``video_search_overhaul.py:40-64``).

Parity status: **pinned** — `tests/golden/make_golden.py` runs the unmodified reference
class from ``/root/reference`` on seeded inputs and commits its outputs under
``tests/golden/``; `tests/test_oracle_exact.py` checks every function here against them.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference
arm may import this module.
"""

from __future__ import annotations

import numpy as np


def normalise_query(q: np.ndarray) -> np.ndarray:
    """``q / (‖q‖ + 1e-10)`` — reference video_search_overhaul.py:49-50.

    dtype-preserving like the reference (float64 in → float64 out); a zero query stays
    all-zero (scores become 0.0), it is *not* rejected.
    """
    return q / (np.linalg.norm(q) + 1e-10)


def stack_store(embeddings) -> np.ndarray:
    """``np.vstack(self.embeddings)`` — reference video_search_overhaul.py:46 (done per query there)."""
    return np.vstack(embeddings)


def exact_search(store: np.ndarray, query: np.ndarray, k: int = 5):
    """Rows and scores exactly as the reference returns them.

    store: [N, D] float32 (rows assumed unit-norm; the reference never re-normalises them)
    Follows video_search_overhaul.py:42-62: empty store → ([], []); normalise the query
    (:49-50); ``np.dot(store, q)`` (:53); ``np.argsort(sim)[::-1][:k]`` (:56); score =
    ``float(sim[idx])`` (:61).  k > N returns N hits.
    """
    if store is None or len(store) == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.float64)
    qn = normalise_query(query)
    sims = np.dot(store, qn)
    top = np.argsort(sims)[::-1][:k]
    return top.astype(np.int64), np.array([float(sims[i]) for i in top], dtype=np.float64)


def exact_search_batch(store: np.ndarray, queries: np.ndarray, k: int):
    """The `/api/search/batch` loop (reference src/api/routes.py:627-634): one
    `exact_search` per query, strictly sequential.  Returns [B,k'] rows and scores."""
    rows, scores = [], []
    for q in queries:
        r, s = exact_search(store, q, k)
        rows.append(r)
        scores.append(s)
    return np.stack(rows), np.stack(scores)


def search_dicts(store: np.ndarray, metadata: list, query: np.ndarray, k: int = 5):
    """The dict form the reference hands to the API layer (video_search_overhaul.py:58-62)."""
    rows, scores = exact_search(store, query, k)
    out = []
    for r, s in zip(rows, scores):
        md = dict(metadata[int(r)])
        md["score"] = float(s)
        out.append(md)
    return out


def formatted_time(ts: float) -> str:
    """reference video_search_overhaul.py:451-453."""
    return f"{int(ts // 60)}m{int(ts % 60)}s"


# --------------------------------------------------------------------------- ground truth

def scores_f64(store: np.ndarray, queries: np.ndarray) -> np.ndarray:
    """float64 ground truth of the normalised-query scores, [B, N] (independent of any
    fp32 summation order — SURVEY.md §4 item 2)."""
    q = queries.astype(np.float64)
    q = q / (np.linalg.norm(q, axis=1, keepdims=True) + 1e-10)
    return q @ store.astype(np.float64).T


def topk_f64(store: np.ndarray, queries: np.ndarray, k: int):
    s = scores_f64(store, queries)
    k = min(k, s.shape[1])
    # deterministic tie rule of the new engine: score desc, row asc
    order = np.lexsort((np.arange(s.shape[1])[None, :].repeat(s.shape[0], 0), -s), axis=1)[:, :k]
    return order.astype(np.int64), np.take_along_axis(s, order, axis=1)


def scan_prenormalised_f64(store: np.ndarray, queries: np.ndarray, k: int):
    """float64 scores of queries used AS GIVEN (no normalisation) and their top-k under the
    engine's (score desc, row asc) rule — ground truth for kernels fed pre-rounded operands."""
    s = queries.astype(np.float64) @ store.astype(np.float64).T
    k = min(k, s.shape[1])
    order = np.lexsort((np.broadcast_to(np.arange(s.shape[1]), s.shape), -s), axis=1)[:, :k]
    return order.astype(np.int64), np.take_along_axis(s, order, axis=1)
