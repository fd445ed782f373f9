"""Deterministic synthetic frame-embedding generators.

Shared by tests, bench.py and the golden-fixture script so that the oracle, the
reference and the CUDA path always see byte-identical inputs.  All generators use
``np.random.default_rng(seed)`` and return float32.

  S-gauss : rows and queries iid N(0,1)^D, L2-normalised.
  S-clip  : clustered (temporally adjacent frames look alike): C unit-norm centres,
            row = centre[uniform id] + 0.35 * N(0,1)^D / sqrt(D), normalised.
  S-ties  : S-gauss with 1 % of rows duplicated exactly (tie rule exerciser).
"""

from __future__ import annotations

import hashlib

import numpy as np

STORE_SEED = 0
QUERY_SEED = 1


def _normalise_rows(x: np.ndarray) -> np.ndarray:
    n = np.linalg.norm(x, axis=1, keepdims=True).astype(np.float32)
    return (x / n).astype(np.float32)


def gauss(n: int, dim: int, seed: int = STORE_SEED, chunk: int = 1 << 16) -> np.ndarray:
    """S-gauss rows, generated in fixed-size chunks so any prefix is reproducible."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, dim), dtype=np.float32)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        out[s:e] = _normalise_rows(rng.standard_normal((e - s, dim), dtype=np.float32))
    return out


def clip_centres(n_store: int, dim: int, seed: int = STORE_SEED) -> np.ndarray:
    c = max(256, n_store // 4096)
    rng = np.random.default_rng(seed + 7919)
    return _normalise_rows(rng.standard_normal((c, dim), dtype=np.float32))


def clip_like(n: int, dim: int, seed: int = STORE_SEED, n_store: int | None = None,
              sigma: float = 0.35, chunk: int = 1 << 16) -> np.ndarray:
    """S-clip rows (or queries: pass the store's size as ``n_store`` and another seed)."""
    centres = clip_centres(n if n_store is None else n_store, dim)
    rng = np.random.default_rng(seed)
    out = np.empty((n, dim), dtype=np.float32)
    scale = np.float32(sigma / np.sqrt(dim))
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        ids = rng.integers(0, centres.shape[0], size=e - s)
        noise = rng.standard_normal((e - s, dim), dtype=np.float32)
        out[s:e] = _normalise_rows(centres[ids] + scale * noise)
    return out


def with_ties(n: int, dim: int, seed: int = STORE_SEED, frac: float = 0.01) -> np.ndarray:
    """S-ties: S-gauss with ``frac`` of the rows overwritten by exact copies of others."""
    x = gauss(n, dim, seed)
    rng = np.random.default_rng(seed + 104729)
    m = max(1, int(n * frac))
    dst = rng.choice(n, size=m, replace=False)
    src = rng.integers(0, n, size=m)
    x[dst] = x[src]
    return x


def sha256_of(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
