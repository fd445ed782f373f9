"""ORACLE — test infrastructure only, never imported by the product path.

ctypes wrapper of oracle/hnsw_ref.cpp: the C++ restatement of the reference's pure-Python HNSW
(/root/reference/src/indexes/hnsw.py:68-74, 76-121, 123-148, 150-229, 238-280) that makes the reference's recall
measurable at 100k+ rows.  Everything the Python code does in NumPy stays in NumPy here (row / query
normalisation `v / np.linalg.norm(v)` :157,250; the level stream of Python's global `random` :68-74) and the C++
side receives the `cblas_sdot` of the OpenBLAS that NumPy itself loaded, so distances are bit-identical to
`1.0 - np.dot(a, b)` (:66).

Parity status: **pinned** — tests/test_oracle_hnsw_ref.py rebuilds the golden graphs of the UNMODIFIED reference
(tests/golden/make_golden.py) edge for edge and reproduces its search results id for id.
"""

from __future__ import annotations

import ctypes as C
import glob
import math
import os
import random
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hnsw_ref.cpp")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libhnsw_ref.so")

_lib = None
_blas = None


def build(force: bool = False) -> str:
    """g++ -O2 of the one source file into oracle/_build/ (git-ignored; travels to the GPU box)."""
    os.makedirs(OUT_DIR, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        r = subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", LIB, SRC], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"g++ failed for {SRC}:\n{r.stderr}")
    return LIB


def _numpy_sdot():
    """Address of cblas_sdot in the OpenBLAS NumPy loaded (scipy-openblas builds prefix / suffix the symbol)."""
    global _blas
    np.dot(np.ones(4, np.float32), np.ones(4, np.float32))      # make sure BLAS is loaded
    cands = glob.glob(os.path.join(os.path.dirname(np.__file__), "..", "numpy.libs", "*openblas*.so*"))
    for path in cands:
        lib = C.CDLL(path)
        for name in ("scipy_cblas_sdot64_", "cblas_sdot64_", "scipy_cblas_sdot", "cblas_sdot"):
            fn = getattr(lib, name, None)
            if fn is not None:
                _blas = lib
                ilp64 = name.endswith("64_")
                return C.cast(fn, C.c_void_p).value, ilp64
    raise ImportError("no cblas_sdot found in numpy.libs: the C++ HNSW restatement needs NumPy's own OpenBLAS for bit-identical distances")


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(LIB)
        lib.href_create.restype = C.c_void_p
        lib.href_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.href_destroy.argtypes = [C.c_void_p]
        lib.href_build.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        lib.href_size.restype = C.c_int64
        lib.href_size.argtypes = [C.c_void_p]
        lib.href_entry.argtypes = [C.c_void_p]
        lib.href_dist_evals.restype = C.c_uint64
        lib.href_dist_evals.argtypes = [C.c_void_p]
        lib.href_neighbours.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        lib.href_search.restype = C.c_uint64
        lib.href_search.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib.href_pyset_order.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        _lib = lib
    return _lib


def reference_levels(n: int, seed: int, level_generation_factor: float = 1.0 / math.log(2.0)):
    """The level stream of n consecutive `add` calls after `random.seed(seed)` (hnsw.py:68-74: one uniform each)."""
    random.seed(seed)
    return np.array([int(-math.log(random.uniform(0, 1)) * level_generation_factor) for _ in range(n)], dtype=np.int32)


def normalise_rows(x: np.ndarray) -> np.ndarray:
    """hnsw.py:157 — per-row `v / np.linalg.norm(v)` (a vectorised norm may differ by an ulp)."""
    return np.stack([v / np.linalg.norm(v) for v in np.asarray(x, dtype=np.float32)]).astype(np.float32)


class RefHNSW:
    """The reference index restated; nodes are dense ints in insertion order."""

    def __init__(self, dimension=512, M=16, ef_construction=200, ef_search=50, max_M=16):
        self.lib = load()
        fn, ilp64 = _numpy_sdot()
        if not ilp64:
            raise ImportError("NumPy's OpenBLAS uses 32-bit integers: rebuild hnsw_ref.cpp's SdotFn accordingly")
        self.dimension, self.M, self.max_M, self.ef_search = dimension, M, max_M, ef_search
        self.h = C.c_void_p(self.lib.href_create(dimension, M, max_M, ef_construction, C.c_void_p(fn)))
        self.rows = None
        self.levels = None

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.href_destroy(self.h)
            self.h = None

    def build(self, vectors: np.ndarray, levels: np.ndarray, already_normalised: bool = False):
        self.rows = np.ascontiguousarray(vectors if already_normalised else normalise_rows(vectors), dtype=np.float32)
        self.levels = np.ascontiguousarray(levels, dtype=np.int32)
        assert self.rows.shape == (len(self.levels), self.dimension)
        self.lib.href_build(self.h, self.rows.ctypes.data, self.levels.ctypes.data, len(self.levels))
        return self

    @property
    def entry(self) -> int:
        return int(self.lib.href_entry(self.h))

    def search(self, queries: np.ndarray, k: int = 5, ef_search: int | None = None):
        """→ (ids [b,k] int32, distances [b,k] float32, distance evaluations)."""
        q = np.ascontiguousarray(np.stack([v / np.linalg.norm(v) for v in np.asarray(queries, dtype=np.float32)]), dtype=np.float32)
        b = q.shape[0]
        d = np.empty((b, k), np.float32)
        ids = np.empty((b, k), np.int32)
        ev = self.lib.href_search(self.h, q.ctypes.data, b, k, self.ef_search if ef_search is None else ef_search,
                                  d.ctypes.data, ids.ctypes.data)
        return ids, d, int(ev)

    def to_arrays(self):
        """Dense graph format shared with oracle/hnsw.py (`GraphArrays`)."""
        from .hnsw import GraphArrays
        n = len(self.levels)
        adj0 = np.full((n, self.max_M), -1, np.int32)
        upper_off = np.full(n, -1, np.int32)
        slots = 0
        for u in range(n):
            if self.levels[u] > 0:
                upper_off[u] = slots
                slots += int(self.levels[u])
        upper_adj = np.full((max(slots, 1), self.M), -1, np.int32)
        buf = np.empty(max(self.M, self.max_M) + 8, np.int32)
        for u in range(n):
            for lv in range(int(self.levels[u]) + 1):
                c = self.lib.href_neighbours(self.h, lv, u, buf.ctypes.data, len(buf))
                assert 0 <= c <= (self.max_M if lv == 0 else self.M), (u, lv, c)
                if lv == 0:
                    adj0[u, :c] = buf[:c]
                else:
                    upper_adj[upper_off[u] + lv - 1, :c] = buf[:c]
        e = self.entry
        return GraphArrays(self.levels.copy(), adj0, upper_off, upper_adj, e, int(self.levels[e]) if e >= 0 else 0)


def pyset_order(ops):
    """Iteration order of the restated CPython set after `ops` (int key = add, ('d', key) = discard) — test hook."""
    lib = load()
    enc = np.array([(-(o[1] + 1)) if isinstance(o, tuple) else o for o in ops], dtype=np.int64)
    out = np.empty(len(ops) + 1, np.int64)
    m = lib.href_pyset_order(enc.ctypes.data, len(enc), out.ctypes.data, len(out))
    return out[:m].tolist()
