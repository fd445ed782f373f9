"""ORACLE — test infrastructure only.  Comparison rules of BASELINE.json's north_star:

* exact search: identical top-k frame ids **excluding score ties within 1e-5**, scores
  within 1e-5 relative;
* HNSW: recall@k against exact ground truth.

The reference orders exactly-equal scores highest-row-first (np.argsort + [::-1],
video_search_overhaul.py:56, observed not guaranteed) and HNSW lowest-id-first
(hnsw.py:269); the new engine uses (score desc, row asc).  Ties are therefore compared by
score, not by id.
"""

from __future__ import annotations

import numpy as np

TIE_TOL = 1e-5
SCORE_RTOL = 1e-5


def _close(a, b, rtol=SCORE_RTOL, atol=1e-7):
    return abs(float(a) - float(b)) <= rtol * max(abs(float(a)), abs(float(b))) + atol


def check_topk(rows_t, scores_t, rows_r, scores_r, tie_tol=TIE_TOL, rtol=SCORE_RTOL):
    """Return (ok, message) for one query.  `*_t` is the implementation under test,
    `*_r` the reference/oracle.  Both best-first."""
    rows_t = np.asarray(rows_t).astype(np.int64)
    rows_r = np.asarray(rows_r).astype(np.int64)
    scores_t = np.asarray(scores_t, dtype=np.float64)
    scores_r = np.asarray(scores_r, dtype=np.float64)
    if rows_t.shape != rows_r.shape:
        return False, f"length {rows_t.shape} vs {rows_r.shape}"
    ref_score = {int(r): float(s) for r, s in zip(rows_r, scores_r)}
    for i, (rt, st, rr, sr) in enumerate(zip(rows_t, scores_t, rows_r, scores_r)):
        if not _close(st, sr, rtol=max(rtol, tie_tol)):
            return False, f"rank {i}: score {st!r} vs {sr!r}"
        if rt == rr:
            if not _close(st, sr, rtol=rtol):
                return False, f"rank {i} row {rt}: score {st!r} vs {sr!r} (> {rtol} rel)"
            continue
        # different row at this rank: only legal inside a tie group (scores within tie_tol)
        if int(rt) in ref_score:
            if abs(ref_score[int(rt)] - sr) > tie_tol:
                return False, f"rank {i}: row {rt} vs {rr}, not a tie ({ref_score[int(rt)]} vs {sr})"
        else:
            # row fell outside the reference's top-k: must tie with the k-th reference score
            if abs(st - scores_r[-1]) > tie_tol:
                return False, f"rank {i}: row {rt} not in reference top-k and not a boundary tie"
    return True, "ok"


def check_topk_batch(rows_t, scores_t, rows_r, scores_r, **kw):
    bad = []
    for b in range(len(rows_r)):
        ok, msg = check_topk(rows_t[b], scores_t[b], rows_r[b], scores_r[b], **kw)
        if not ok:
            bad.append((b, msg))
    return bad


def id_match_fraction(rows_t, rows_r) -> float:
    rows_t, rows_r = np.asarray(rows_t), np.asarray(rows_r)
    return float((rows_t == rows_r).mean()) if rows_r.size else 1.0


def recall_at_k(found_rows, truth_rows) -> float:
    """Mean |found ∩ truth| / |truth| over queries (rows may be ragged lists; −1 = empty)."""
    hits = total = 0
    for f, t in zip(found_rows, truth_rows):
        t = [int(x) for x in t if int(x) >= 0]
        fs = {int(x) for x in f if int(x) >= 0}
        hits += sum(1 for x in t if x in fs)
        total += len(t)
    return hits / max(total, 1)
