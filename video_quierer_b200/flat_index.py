"""B200FlatIndex — drop-in for the reference's live exact index,
``SimpleVideoIndex`` (reference video_search_overhaul.py:23-106).

Same attribute surface (`embeddings`, `metadata`, `video_hashes`), same methods
(`add_frame`, `search`, `save_to_disk`, `load_from_disk`), same result dicts and error
conventions (empty index → ``[]``; k > N → N hits; save/load swallow errors and return a
bool).  What changes is the engine: the per-query ``np.vstack`` + ``np.dot`` + ``np.argsort``
(:46-56) becomes one `vq_scan_topk` launch over a device-resident matrix.

The route handlers of the reference reach *through* the index and mutate
``index.embeddings`` / ``index.metadata`` directly (``pop(i)``, rebind to ``[]``, ``len``;
src/api/routes.py:754-762,979-981,1014-1016), so `embeddings` is a list subclass that
records mutations; the device matrix is re-synchronised lazily before the next search
(append-only changes upload only the new rows).
"""

from __future__ import annotations

import logging
import os
import pickle
import threading
from pathlib import Path
from typing import Dict, List

import numpy as np
import torch

from . import _lib
from .engine import DeviceStore, Scanner, _require_cuda, as_device_queries

logger = logging.getLogger(__name__)


class EmbeddingList(list):
    """`list` of per-frame float32 arrays that remembers how far the device copy is valid."""

    def __init__(self, *a):
        super().__init__(*a)
        self.valid_prefix = 0          # rows [0, valid_prefix) are unchanged since the last sync

    def _touch(self, i: int):
        n = len(self)
        if i < 0:
            i += n
        self.valid_prefix = max(0, min(self.valid_prefix, i))

    def pop(self, i=-1):
        self._touch(i if i >= 0 else len(self) + i)
        return super().pop(i)

    def __delitem__(self, i):
        self._touch(0 if isinstance(i, slice) else i)
        super().__delitem__(i)

    def __setitem__(self, i, v):
        self._touch(0 if isinstance(i, slice) else i)
        super().__setitem__(i, v)

    def insert(self, i, v):
        self._touch(i)
        super().insert(i, v)

    def remove(self, v):
        self.valid_prefix = 0
        super().remove(v)

    def clear(self):
        self.valid_prefix = 0
        super().clear()

    def sort(self, *a, **k):
        self.valid_prefix = 0
        super().sort(*a, **k)

    def reverse(self):
        self.valid_prefix = 0
        super().reverse()

    def __iadd__(self, other):
        super().extend(other)
        return self
    # append / extend only add rows past valid_prefix → nothing to record


class LazyRows:
    """`embeddings` of an index loaded from a raw store: the rows stay in the memmap / on the device
    instead of becoming one Python object each (hopeless at 100M rows).  Reads (`len`, indexing,
    iteration, truthiness) work; `append` / `extend` work (new rows live in `tail` and are uploaded
    incrementally); removing or replacing one of the stored rows needs the list form: call
    `B200FlatIndex.materialise()` first (fine up to a few million rows)."""

    def __init__(self, rows_f32, rows_bf16, dim: int):
        self._f32, self._bf16, self._dim = rows_f32, rows_bf16, int(dim)
        self.base_n = int((rows_f32 if rows_f32 is not None else rows_bf16).shape[0])
        self.tail = EmbeddingList()

    def __len__(self):
        return self.base_n + len(self.tail)

    def __bool__(self):
        return len(self) > 0

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if i >= self.base_n:
            return self.tail[i - self.base_n]
        if self._f32 is not None:
            return np.array(self._f32[i, : self._dim])
        return (self._bf16[i, : self._dim].astype(np.uint32) << 16).view(np.float32)      # widen bf16 exactly

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def append(self, v):
        self.tail.append(v)

    def extend(self, vs):
        self.tail.extend(vs)

    def _frozen(self, *a, **k):
        raise NotImplementedError("rows of a raw-loaded store are immutable; call index.materialise() to get the "
                                  "reference's mutable list form")
    pop = __delitem__ = __setitem__ = insert = remove = clear = sort = reverse = _frozen


class DeviceRows(LazyRows):
    """`embeddings` of an index whose rows were handed over on the DEVICE (`add_frames` with a CUDA tensor,
    `adopt_store`): they live only in the device matrix; reads copy single rows back on demand.  Same
    contract as `LazyRows` (append / extend work, removing stored rows needs `materialise()`)."""

    def __init__(self, store: DeviceStore):
        self._store_ref, self._dim = store, store.dim
        self._f32 = self._bf16 = None
        self.base_n = store.n
        self.tail = EmbeddingList()

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if i >= self.base_n:
            return self.tail[i - self.base_n]
        return self._store_ref.rows_to_host(i, i + 1)[0]

    def __iter__(self):
        for lo in range(0, self.base_n, 1 << 16):             # chunked device-to-host copies
            yield from self._store_ref.rows_to_host(lo, min(self.base_n, lo + (1 << 16)))
        yield from self.tail


# Worst-case |bf16-operand score - fp32 score| per unit of |q| |x|: each operand is rounded to 8 significant
# bits (unit roundoff 2^-8), products are exact, so
# |delta| <= (2*2^-8 + 2^-16) * sum|q_i x_i| <= (2^-7 + 2^-16) * |q| |x|  (Cauchy-Schwarz) plus the fp32
# accumulation error of both sums (<= 3 * 768 * 2^-24).  Used by the two-stage / large-k routes, multiplied by
# the store's largest row norm; `vq_search_exact` computes a tighter, per-query bound on the device from the
# actual rounding errors (|q^ - q|, max |x^ - x|).
BF16_SCORE_EPS = 2.0 ** -7 + 2.0 ** -16 + 1.4e-4
MAX_TENSOR_K, MAX_TENSOR_LD = 64, 768        # limits of scan_mma_bf16_kernel (csrc/scan_mma.cu)
MAX_BATCH = 8192                             # queries per launch chain of the facade (64 query tiles)


def exact_search(scanner: Scanner, st: DeviceStore, q_dev: torch.Tensor, k: int):
    """The exact fp32 top-k of the store at tensor-core speed.  k <= 128 and ld <= 768: ONE pass
    (`vq_search_exact`: scan of the bf16 copy that gathers everything within the operand-rounding bound of
    the running k-th best, fp32 re-score, exact by construction); larger k / wider rows: `large_k_search`.
    Returns ([b,k] f32, [b,k] i32, overflow [b] i32 device tensor).  The caller re-runs queries
    whose overflow flag is set (mass ties) with `exact_fallback` after its device->host read, so no sync
    is added here."""
    if st.ld > MAX_TENSOR_LD or not scanner.lib.vq_search_exact_supported(st.n, st.dim, st.ld, q_dev.shape[0], k):
        return large_k_search(scanner, st, q_dev, k)
    return scanner.exact(st, q_dev, k, _lib.NORM_EPS)


def two_stage_search(scanner: Scanner, st: DeviceStore, q_dev: torch.Tensor, k: int, path: str = "auto",
                     max_row_norm: float | None = None):
    """(Superseded by `exact_search`; kept as the k_cand experiment route.)
    Exact top-k from a bf16 scan, ONE C-ABI call (`vq_search_two_stage`): (1) the tensor-core
    scan of the bf16 copy selects k' candidates, (2) they are re-scored exactly from the fp32 copy,
    (3) the result is *certified*: every row outside the candidate set has exact score <=
    (k'-th bf16 score) + eps, so if the exact k-th score is above that bound nothing was missed.
    Returns ([b,k] f32, [b,k] i32, uncertified [b] bool device tensor or None); the caller re-runs
    uncertified queries (near-duplicate heavy data) with `exact_fallback` after its device->host
    read, so no extra sync is added here."""
    # candidates per query: 32 for k <= 16, else max(2k, k+22), capped at the register top-k of the tensor
    # kernel.  Where the cap bites (k > 42) the certificate mostly fails and the collect pass of
    # `resolve_uncertified` delivers the exact answer — still two tensor-core passes instead of the
    # fp32 FMA scan.
    kc = int(os.environ.get("VQ_KCAND", 0)) or (32 if k <= 16 else max(2 * k, k + 22))
    kc = max(min(st.n, kc, MAX_TENSOR_K), min(k, st.n))
    if max_row_norm is None:
        max_row_norm = st.max_row_norm()
    if k > MAX_TENSOR_K or st.ld > MAX_TENSOR_LD:
        return large_k_search(scanner, st, q_dev, k, max_row_norm)
    s_hi, rows, bad = scanner.two_stage(st.bf16, st.f32, st.n, st.dim, q_dev, k, kc, _lib.NORM_EPS,
                                        BF16_SCORE_EPS * max_row_norm)
    return s_hi, rows, (bad if kc < st.n else None)


def exact_fallback(scanner: Scanner, st: DeviceStore, q_dev: torch.Tensor, k: int, idx: torch.Tensor):
    """fp32 FMA scan for the queries `idx` (last resort: always exact, any k)."""
    return scanner.scan(st.f32, st.n, st.dim, q_dev[idx].contiguous(), k, _lib.NORM_EPS, "fma")


COLLECT_CAP = 4096
LARGE_K_CAP = 16384


def large_k_search(scanner: Scanner, st: DeviceStore, q_dev: torch.Tensor, k: int, max_row_norm: float | None = None):
    """Exact top-k beyond the register lists of the tensor kernel (64 < k <= 256; BASELINE config 4:
    k = 100), in two tensor-core collect passes:
      1. over a STRIDED SAMPLE of the store (a compact copy of every `stride`-th row, cached with the
         store) with thresholds derived from the sample itself (k-th largest per-tile maximum): at
         least k rows are gathered and re-scored in fp32; the k-th exact score among them, s_k, is
         reached by k real rows, so it is a lower bound of the k-th best of the whole store;
      2. over the whole store with threshold s_k - eps: every row of the true top-k is gathered
         (about k * stride rows per query), all are re-scored in fp32, the best k are returned.
    Exact by construction; nothing here synchronises with the host (the two passes can be captured in a
    CUDA graph).  Returns (scores, rows, overflow [b] i32): queries whose gather overflowed must be re-run
    with `exact_fallback`; stores that are too small or too wide for the tensor kernel are answered by the
    fp32 FMA scan right away (overflow all zero)."""
    tile = 128 if st.ld <= 512 else 64
    # the sample must hold k + 1 full tiles (the derived threshold is the k-th largest tile maximum); the
    # second pass gathers ~k * stride rows per query, which has to stay well inside LARGE_K_CAP
    stride = min(64, LARGE_K_CAP // (3 * k), st.n // ((k + 1) * tile)) if k <= 256 else 0
    if st.bf16 is None or st.ld > MAX_TENSOR_LD or stride < 2:
        s, r = scanner.scan(st.f32, st.n, st.dim, q_dev, k, _lib.NORM_EPS, "fma")
        return s, r, torch.zeros((q_dev.shape[0],), dtype=torch.int32, device=q_dev.device)
    if max_row_norm is None:
        max_row_norm = st.max_row_norm_cached()
    smp_f32, smp_bf16 = st.sample_f32(stride)       # rows 0, stride, 2*stride, ...
    s1, _, over1 = scanner.collect(smp_bf16, smp_f32, smp_f32.shape[0], st.dim, q_dev, k, None, LARGE_K_CAP)
    thr = s1[:, k - 1] - BF16_SCORE_EPS * max_row_norm
    thr = torch.where(over1 > 0, torch.full_like(thr, float("-inf")), thr)      # incomplete sample answer: no bound
    # second pass: only the best candidates by bf16 score and those within eps of the k-th are re-scored (exact_finish)
    return scanner.collect(st.bf16, st.f32, st.n, st.dim, q_dev, k, thr, LARGE_K_CAP, bounds=st.bounds)


def resolve_uncertified(scanner: Scanner, st: DeviceStore, q_dev: torch.Tensor, k: int, idx: torch.Tensor,
                        s_two_stage: torch.Tensor, max_row_norm: float | None = None):
    """Exact top-k for the queries `idx` whose two-stage result was not certified (near-duplicate
    heavy data).  The k-th exact score the two-stage pass returned is a lower bound s_k of the true
    k-th best, so every row of the true top-k has bf16-operand score >= s_k - eps: a second
    tensor-core pass gathers exactly those rows (`vq_search_collect`), re-scores ALL of them from
    the fp32 copy and keeps the best k.  Only queries with more than COLLECT_CAP such rows (mass
    duplicates) go to the fp32 FMA scan.  Returns (scores [m,k] f32, rows [m,k] i32) device tensors."""
    if max_row_norm is None:
        max_row_norm = st.max_row_norm()
    qs = q_dev[idx].contiguous()
    thr = s_two_stage[idx, k - 1] - BF16_SCORE_EPS * max_row_norm
    s, r, over = scanner.collect(st.bf16, st.f32, st.n, st.dim, qs, k, thr, COLLECT_CAP)
    over_h = torch.nonzero(over).flatten()
    if len(over_h):
        sf, rf = exact_fallback(scanner, st, qs, k, over_h)
        s[over_h], r[over_h] = sf, rf
    return s, r


class B200FlatIndex:
    """Exact cosine search over a device-resident frame-embedding matrix."""

    def __init__(self, device=None, store_dtype: str = "bf16", rescore: bool = True, path: str = "auto"):
        """store_dtype 'bf16' + rescore (the default) = exact mode: an fp32 master copy plus a bf16 scan copy;
        the tensor-core scan of the bf16 copy gathers every row that can belong to the exact top-k and the
        answer is re-scored from the fp32 copy (`vq_search_exact`) — the ids of the fp32 scan and fp32 scores
        within 1e-5 of the reference, at 6 bytes per element.
        store_dtype 'fp32' = the fp32 copy only, scanned by the fp32 FMA kernel (4 bytes per element, same
        results, HBM-bound up to batch 8 and FMA-bound beyond).
        store_dtype 'bf16' without rescore = the bf16 copy only: approximate scores (operands rounded to
        bf16), 2 bytes per element."""
        self._embeddings = EmbeddingList()
        self.metadata: List[Dict] = []
        self.video_hashes: Dict = {}
        self.device = _require_cuda(device)
        self.store_dtype = "bf16" if store_dtype in ("bf16", "bfloat16") else "fp32"
        self.rescore = bool(rescore)
        self.path = path
        self._store: DeviceStore | None = None
        self._scanner = Scanner(self.device)
        self._lock = threading.RLock()
        self.search_times: List[float] = []
        self.stats = {"exact_queries": 0, "overflow_queries": 0}   # overflowing ones (mass ties) re-run on the fp32 FMA scan

    # ------------------------------------------------------------------ attribute surface
    @property
    def embeddings(self) -> EmbeddingList:
        return self._embeddings

    @embeddings.setter
    def embeddings(self, value):
        # handlers rebind the attribute: `system.index.embeddings = []` (routes.py:979,1014)
        with self._lock:
            if isinstance(value, LazyRows):
                self._embeddings = value
                return
            self._embeddings = value if isinstance(value, EmbeddingList) else EmbeddingList(value)
            self._embeddings.valid_prefix = 0

    # ------------------------------------------------------------------ ingest
    def add_frame(self, embedding: np.ndarray, video_name: str, timestamp: float):
        """video_search_overhaul.py:31-38."""
        self._embeddings.append(np.asarray(embedding).astype(np.float32))
        self.metadata.append({
            'video_name': video_name,
            'timestamp': timestamp,
            'frame_id': len(self._embeddings) - 1,
        })

    def add_frames(self, embeddings, video_names, timestamps):
        """Bulk form of `add_frame` (same bookkeeping, one upload).  A CUDA tensor (the encoder's output,
        video_search_overhaul.py:224-228 before its `.cpu().numpy()`) is appended to the device matrix
        directly — no host round trip; `embeddings` then becomes a `DeviceRows` view."""
        base = len(self._embeddings)
        if isinstance(embeddings, torch.Tensor) and embeddings.is_cuda:
            with self._lock:
                self._adopt_device_rows(embeddings.shape[-1])
                self._sync()
                self._store.append(embeddings.reshape(-1, embeddings.shape[-1]), _lib.NORM_NONE)
                self._embeddings.base_n = self._store.n          # (the host tail was uploaded by _sync: it is part of the base now)
                self._embeddings.tail = EmbeddingList()
        else:
            if isinstance(embeddings, torch.Tensor):
                embeddings = embeddings.detach().cpu().numpy()
            embeddings = np.asarray(embeddings, dtype=np.float32)
            self._embeddings.extend(list(embeddings))
        for i, (vn, ts) in enumerate(zip(video_names, timestamps)):
            self.metadata.append({'video_name': vn, 'timestamp': ts, 'frame_id': base + i})

    def _new_store(self, dim: int) -> DeviceStore:
        bf = self.store_dtype == "bf16"
        return DeviceStore(dim, self.device, keep_fp32=(not bf) or self.rescore, keep_bf16=bf)

    def _adopt_device_rows(self, dim: int):
        """Switch `embeddings` to the device-resident form (rows already listed on the host are uploaded first)."""
        if isinstance(self._embeddings, DeviceRows):
            return
        if isinstance(self._embeddings, LazyRows):
            raise NotImplementedError("device-resident appends to a raw-loaded index: call materialise() first")
        self._sync()
        if self._store is None:
            self._store = self._new_store(int(dim))
        self._embeddings = DeviceRows(self._store)

    def adopt_store(self, store: DeviceStore, metadata: List[Dict]):
        """Serve an existing device-resident store (e.g. built by the encoder pipeline) through the drop-in
        surface without copying it; `metadata` is the parallel list of the reference's per-frame dicts."""
        if len(metadata) != store.n:
            raise ValueError(f"{len(metadata)} metadata entries for {store.n} rows")
        with self._lock:
            self._store = store
            self.store_dtype = "bf16" if store.bf16 is not None else "fp32"
            self.rescore = store.f32 is not None
            self._embeddings = DeviceRows(store)
            self.metadata = metadata

    def _sync(self):
        emb = self._embeddings
        if isinstance(emb, LazyRows):            # stored rows are already on the device; only the tail moves
            st, tail = self._store, emb.tail
            keep = min(tail.valid_prefix, st.n - emb.base_n, len(tail))
            if keep == len(tail) and st.n == len(emb):
                return
            st.truncate(emb.base_n + keep)
            for s0 in range(keep, len(tail), 1 << 16):
                st.append(np.stack(tail[s0: s0 + (1 << 16)]).astype(np.float32, copy=False), _lib.NORM_NONE)
            tail.valid_prefix = len(tail)
            return
        n = len(emb)
        if n == 0:
            if self._store is not None:
                self._store.truncate(0)
            emb.valid_prefix = 0
            return
        dim = int(np.asarray(emb[0]).shape[-1])
        if self._store is None or self._store.dim != dim:
            self._store = self._new_store(dim)
            emb.valid_prefix = 0
        st = self._store
        keep = min(emb.valid_prefix, st.n, n)
        if keep == n and st.n == n:
            return
        st.truncate(keep)
        chunk = 1 << 16
        for s in range(keep, n, chunk):
            e = min(n, s + chunk)
            st.append(np.stack(emb[s:e]).astype(np.float32, copy=False), _lib.NORM_NONE)
        emb.valid_prefix = n

    # ------------------------------------------------------------------ search
    def search_arrays(self, queries, k: int = 5):
        """Batched search returning ([b,k'] scores float32, [b,k'] rows int32) numpy arrays,
        k' = min(k, N).  `queries` may be numpy or a torch tensor (host or device)."""
        with self._lock:
            self._sync()
            n = 0 if self._store is None else self._store.n
            if n == 0:
                b = 1 if np.ndim(queries) == 1 else len(queries)
                return np.zeros((b, 0), np.float32), np.zeros((b, 0), np.int32)
            st = self._store
            q_all = as_device_queries(queries, st.dim, self.device)
            kk = min(int(k), n)
            if q_all.shape[0] > MAX_BATCH:          # one launch covers at most #SMs query tiles: chunk beyond
                parts = [self._search_chunk(st, q_all[i: i + MAX_BATCH].contiguous(), kk)
                         for i in range(0, q_all.shape[0], MAX_BATCH)]
                return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])
            return self._search_chunk(st, q_all, kk)

    def _search_chunk(self, st, q, kk):
        if self.store_dtype != "bf16":
            s, r = self._scanner.scan(st.f32, st.n, st.dim, q, kk, _lib.NORM_EPS, self.path)
            return s.cpu().numpy(), r.cpu().numpy()
        if not self.rescore:
            s, r = self._scanner.scan(st.bf16, st.n, st.dim, q, kk, _lib.NORM_EPS, self.path)
            return s.cpu().numpy(), r.cpu().numpy()
        s, r, over = exact_search(self._scanner, st, q, kk)
        s_h, r_h = s.cpu().numpy(), r.cpu().numpy()
        self.stats["exact_queries"] += int(q.shape[0])
        over_h = np.nonzero(over.cpu().numpy())[0]
        if len(over_h):                      # mass ties: more rows within the error bound than the gather holds
            self.stats["overflow_queries"] += len(over_h)
            idx = torch.from_numpy(over_h).to(self.device)
            s2, r2 = exact_fallback(self._scanner, st, q, kk, idx)
            s_h[over_h], r_h[over_h] = s2.cpu().numpy(), r2.cpu().numpy()
        return s_h, r_h

    def search(self, query_embedding: np.ndarray, k: int = 5) -> List[Dict]:
        """video_search_overhaul.py:40-64: list of metadata copies + 'score', best first."""
        if not self._embeddings:
            return []
        scores, rows = self.search_arrays(query_embedding, k)
        return self._hits(scores[0].tolist(), rows[0].tolist())

    def _hits(self, scores: list, rows: list) -> List[Dict]:
        """Metadata copies + 'score' (Python float), best first (video_search_overhaul.py:58-62).  Works on
        plain Python lists: one `.tolist()` per batch instead of a numpy scalar conversion per hit."""
        md = self.metadata
        return [{**md[r], 'score': s} for s, r in zip(scores, rows) if r >= 0]

    def search_batch(self, queries, k: int = 5) -> List[List[Dict]]:
        """One launch for the whole batch (replaces the per-query loop of routes.py:627-634)."""
        if not self._embeddings:
            return [[] for _ in range(len(queries))]
        scores, rows = self.search_arrays(queries if isinstance(queries, torch.Tensor) else np.asarray(queries), k)
        return [self._hits(sb, rb) for sb, rb in zip(scores.tolist(), rows.tolist())]

    # ------------------------------------------------------------------ persistence
    def save_to_disk(self, cache_path: Path):
        """Same pickle as the reference (video_search_overhaul.py:66-85) so caches interchange."""
        try:
            cache_data = {
                'embeddings': [np.asarray(e) for e in self._embeddings],
                'metadata': self.metadata,
                'video_hashes': self.video_hashes,
                'version': '1.0',
            }
            with open(cache_path, 'wb') as f:
                pickle.dump(cache_data, f)
            logger.info("Saved %d embeddings to %s", len(self._embeddings), cache_path)
            return True
        except Exception as e:  # noqa: BLE001 — the reference swallows everything here (:83-85)
            logger.error("Failed to save cache: %s", e)
            return False

    def load_from_disk(self, cache_path: Path) -> bool:
        """video_search_overhaul.py:87-106."""
        try:
            cache_path = Path(cache_path)
            if not cache_path.exists():
                return False
            with open(cache_path, 'rb') as f:
                cache_data = pickle.load(f)
            self.embeddings = cache_data.get('embeddings', [])
            self.metadata = cache_data.get('metadata', [])
            self.video_hashes = cache_data.get('video_hashes', {})
            logger.info("Loaded %d embeddings from %s", len(self._embeddings), cache_path)
            return True
        except Exception as e:  # noqa: BLE001 — (:104-106)
            logger.error("Failed to load cache: %s", e)
            return False

    # ------------------------------------------------------------------ raw persistence
    def save_raw(self, path) -> bool:
        """Write the device matrices as they are + metadata to a raw store directory (rawstore.py)."""
        from . import rawstore
        with self._lock:
            self._sync()
            if self._store is None or self._store.n == 0:
                raise ValueError("nothing to save: the index is empty")
            st = self._store
            w = rawstore.RawWriter(str(path), "flat", {"n": st.n, "dim": st.dim, "ld": st.ld,
                                                      "store_dtype": self.store_dtype, "rescore": self.rescore})
            st.save_raw_arrays(w)
            w.put_objects({"metadata": self.metadata, "video_hashes": self.video_hashes})
            w.close()
        return True

    def load_raw(self, path, verify: bool = True) -> bool:
        """Load a raw store: memmap + chunked copies into the device matrix; `embeddings` becomes a
        `LazyRows` view (no per-row Python objects).  Raises ValueError on a damaged store."""
        from . import rawstore
        attrs, arrays, objects = rawstore.open_raw(str(path), "flat", verify)
        with self._lock:
            self.store_dtype = attrs["store_dtype"]
            self.rescore = bool(attrs["rescore"])
            self._store = DeviceStore.from_raw_arrays(int(attrs["dim"]), arrays, self.device)
            self._embeddings = LazyRows(arrays.get("rows_f32"), arrays.get("rows_bf16"), int(attrs["dim"]))
            self.metadata = objects["metadata"]
            self.video_hashes = objects["video_hashes"]
        return True

    def materialise(self):
        """Turn a raw-loaded index back into the reference's list form (one array per frame)."""
        with self._lock:
            if isinstance(self._embeddings, LazyRows):
                rows = EmbeddingList(list(self._embeddings))
                self._embeddings = rows
                self._embeddings.valid_prefix = len(rows) if self._store is not None and self._store.n == len(rows) else 0

    # ------------------------------------------------------------------ introspection
    @property
    def last_scan_path(self) -> str:
        return self._scanner.last_path
