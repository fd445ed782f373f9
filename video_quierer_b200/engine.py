"""Device-side plumbing shared by the index facades: the frame-embedding store (a contiguous
device matrix with amortised growth) and thin wrappers that hand torch tensors to the C-ABI
as raw pointers.  torch is used for memory, streams and distributed only — every kernel
on the search path is ours (libvqsearch.so).
"""

from __future__ import annotations

import threading

import numpy as np
import torch

from . import _lib

_DT = {"fp32": _lib.F32, "f32": _lib.F32, "float32": _lib.F32, "bf16": _lib.BF16, "bfloat16": _lib.BF16}
_PATH = {"auto": _lib.SCAN_AUTO, "fma": _lib.SCAN_FMA, "gemv": _lib.SCAN_FMA, "mma": _lib.SCAN_MMA,
         "gemm": _lib.SCAN_MMA, "fma32": _lib.SCAN_FMA32}


def _require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("video_quierer_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def padded_ld(dim: int, dtype: str = "fp32") -> int:
    """Row stride of a store: multiple of 64 elements (covers both the fp32 and bf16 rules)."""
    return (dim + 63) // 64 * 64


class Workspace:
    """A grow-only device scratch buffer (256-byte aligned by the caching allocator)."""

    def __init__(self, device):
        self.device = device
        self.buf = None

    def get(self, nbytes: int) -> torch.Tensor:
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=self.device)
        return self.buf


class DeviceStore:
    """Contiguous [capacity, ld] device matrix of frame embeddings (+ optional bf16 twin).

    Replaces the reference's Python list of per-frame arrays that is re-stacked for every
    query (video_search_overhaul.py:27,46).  Rows are stored as given (the reference never
    re-normalises stored rows on the exact path) or normalised on ingest for HNSW
    (hnsw.py:157).
    """

    def __init__(self, dim: int, device=None, keep_fp32: bool = True, keep_bf16: bool = False,
                 capacity: int = 0):
        self.device = _require_cuda(device)
        self.dim = int(dim)
        self.ld = padded_ld(dim)
        self.n = 0
        self.keep_fp32, self.keep_bf16 = keep_fp32, keep_bf16
        self.f32 = None
        self.bf16 = None
        self.lib = _lib.load()
        # operand-rounding bounds of the exact search, on the device: [max |bf16 row|, max |bf16 row - fp32 row|]
        # (folded in by `append`; monotone — rows that are removed later only leave the bound looser)
        self.bounds = None
        if capacity:
            self._reserve(capacity)

    # ------------------------------------------------------------------ capacity
    def capacity(self) -> int:
        t = self.f32 if self.f32 is not None else self.bf16
        return 0 if t is None else t.shape[0]

    def _reserve(self, rows: int):
        cap = self.capacity()
        if rows <= cap:
            return
        new_cap = max(rows, 2 * cap, 1024)
        with torch.cuda.device(self.device):
            if self.keep_fp32:
                t = torch.empty((new_cap, self.ld), dtype=torch.float32, device=self.device)
                if self.n:
                    t[: self.n].copy_(self.f32[: self.n])
                self.f32 = t
            if self.keep_bf16:
                t = torch.empty((new_cap, self.ld), dtype=torch.bfloat16, device=self.device)
                if self.n:
                    t[: self.n].copy_(self.bf16[: self.n])
                self.bf16 = t

    # ------------------------------------------------------------------ ingest
    def append(self, rows, norm: int = _lib.NORM_NONE):
        """Append fp32 rows ([m, dim] numpy or torch, host or device) through `vq_ingest_rows`."""
        if isinstance(rows, np.ndarray):
            src = torch.from_numpy(np.ascontiguousarray(rows, dtype=np.float32))
        else:
            src = rows.detach().to(torch.float32).contiguous()
        if src.dim() == 1:
            src = src[None, :]
        if src.shape[1] != self.dim:
            raise ValueError(f"embedding dimension {src.shape[1]} != store dimension {self.dim}")
        m = src.shape[0]
        if m == 0:
            return
        self._reserve(self.n + m)
        with torch.cuda.device(self.device):
            src = src.to(self.device, non_blocking=False)
            st = _stream(self.device)
            if self.keep_fp32:
                dst = self.f32[self.n: self.n + m]
                _lib.check(self.lib.vq_ingest_rows(_ptr(src), m, self.dim, self.dim, _ptr(dst), _lib.F32, self.ld,
                                                   norm, st), "vq_ingest_rows")
            if self.keep_bf16:
                dst = self.bf16[self.n: self.n + m]
                _lib.check(self.lib.vq_ingest_rows(_ptr(src), m, self.dim, self.dim, _ptr(dst), _lib.BF16, self.ld,
                                                   norm, st), "vq_ingest_rows")
            self._fold_bounds(self.n, self.n + m)
        self.n += m
        self._sample = None

    def _fold_bounds(self, lo: int, hi: int):
        """`vq_store_bounds` over rows [lo, hi) of the twins (no-op unless both copies are kept)."""
        if self.f32 is None or self.bf16 is None or hi <= lo:
            return
        with torch.cuda.device(self.device):
            if self.bounds is None:
                self.bounds = torch.zeros(2, dtype=torch.float32, device=self.device)
            _lib.check(self.lib.vq_store_bounds(_ptr(self.f32[lo:hi]), _ptr(self.bf16[lo:hi]), hi - lo, self.ld,
                                                _ptr(self.bounds), _stream(self.device)), "vq_store_bounds")

    # ------------------------------------------------------------------ raw persistence (rawstore.py)
    def save_raw_arrays(self, writer):
        """Chunked device-to-host copies of the [n, ld] matrices into the writer's memmaps."""
        from . import rawstore
        for name, t, dt, item in (("rows_f32", self.f32, "float32", 4), ("rows_bf16", self.bf16, "uint16", 2)):
            if t is None:
                continue
            dst = writer.create(name, dt, (self.n, self.ld))
            for lo, hi in rawstore.chunks(self.n, self.ld * item):
                chunk = t[lo:hi]
                if dt == "uint16":
                    dst[lo:hi] = chunk.view(torch.int16).cpu().numpy().view(np.uint16)
                else:
                    dst[lo:hi] = chunk.cpu().numpy()

    @classmethod
    def from_raw_arrays(cls, dim: int, arrays, device=None):
        """Rebuild a store from `rows_f32` / `rows_bf16` memmaps: chunked host-to-device copies straight
        into the device matrices (rows are stored exactly as they sat in HBM: no ingest kernel)."""
        from . import rawstore
        f, b = arrays.get("rows_f32"), arrays.get("rows_bf16")
        ref = f if f is not None else b
        if ref is None:
            raise ValueError("raw store holds neither rows_f32 nor rows_bf16")
        n, ld = int(ref.shape[0]), int(ref.shape[1])
        st = cls(dim, device, keep_fp32=f is not None, keep_bf16=b is not None, capacity=n)
        if ld != st.ld:
            raise ValueError(f"raw store row stride {ld} != {st.ld} expected for dimension {dim}")
        for src, dst, item in ((f, st.f32, 4), (b, st.bf16, 2)):
            if src is None:
                continue
            for lo, hi in rawstore.chunks(n, ld * item):
                host = np.array(src[lo:hi])               # private writable copy of the read-only memmap slice
                if item == 2:
                    dst[lo:hi].view(torch.int16).copy_(torch.from_numpy(host.view(np.int16)))
                else:
                    dst[lo:hi].copy_(torch.from_numpy(host))
        st.n = n
        st._fold_bounds(0, n)
        return st

    def truncate(self, n: int):
        self.n = min(self.n, max(0, int(n)))
        self._sample = None
        if self.n == 0 and self.bounds is not None:
            self.bounds.zero_()

    def sample_f32(self, stride: int) -> torch.Tensor:
        """Compact (fp32, bf16) copies of every `stride`-th row (rows 0, stride, 2*stride, ...), cached
        until the store changes: the k-th best exact score found inside this sample is a lower bound of
        the k-th best of the whole store (used by the large-k search to seed its collect pass)."""
        key = (self.n, int(stride))
        if getattr(self, "_sample", None) is None or self._sample[0] != key:
            with torch.cuda.device(self.device):
                f = self.f32[: self.n: stride].contiguous()
                b = self.bf16[: self.n: stride].contiguous() if self.bf16 is not None else None
                self._sample = (key, f, b)
        return self._sample[1], self._sample[2]

    def max_row_norm(self) -> float:
        """Largest row norm of the store (host float; synchronises).  1.0 if no bounds are tracked."""
        if self.bounds is None:
            return 1.0
        b = self.bounds.cpu()
        return max(float(b[0] + b[1]), 1e-30)           # |x| <= |x^| + |x^ - x|

    def max_row_norm_cached(self) -> float:
        """`max_row_norm` read once per store size (the large-k route needs it on the host to scale its
        threshold; reading it per search would put a device synchronisation on the search path)."""
        c = getattr(self, "_mrn", None)
        if c is None or c[0] != self.n:
            c = (self.n, self.max_row_norm())
            self._mrn = c
        return c[1]

    def view(self, dtype: str = "fp32") -> torch.Tensor:
        t = self.f32 if _DT[dtype] == _lib.F32 else self.bf16
        if t is None:
            raise RuntimeError(f"store keeps no {dtype} copy")
        return t[: self.n]

    def rows_to_host(self, start: int = 0, stop: int | None = None) -> np.ndarray:
        return self.view("fp32")[start:stop, : self.dim].cpu().numpy()


class Scanner:
    """Exact scan + fused top-k through `vq_scan_topk` (and the two-stage bf16 + re-score mode)."""

    def __init__(self, device=None):
        self.device = _require_cuda(device)
        self.lib = _lib.load()
        self.ws = Workspace(self.device)
        self.lock = threading.Lock()
        self.last_path = ""
        self.last_launches = 0

    def scan(self, mat: torch.Tensor, n: int, dim: int, queries: torch.Tensor, k: int,
             norm: int = _lib.NORM_EPS, path: str = "auto"):
        """mat: [>=n, ld] fp32 or bf16 device tensor; queries: [b, dim] fp32 device tensor.
        Returns (scores [b,k] fp32, rows [b,k] int32) device tensors; empty slots have row -1."""
        dt = _lib.F32 if mat.dtype == torch.float32 else _lib.BF16
        ld = mat.stride(0)
        b = queries.shape[0]
        with torch.cuda.device(self.device):
            out_s = torch.empty((b, k), dtype=torch.float32, device=self.device)
            out_r = torch.empty((b, k), dtype=torch.int32, device=self.device)
            if b == 0:
                return out_s, out_r
            p = _PATH[path]
            need = self.lib.vq_scan_workspace_bytes(n, dim, ld, dt, b, k, p)
            with self.lock:
                ws = self.ws.get(need)
                rc = self.lib.vq_scan_topk(_ptr(mat), n, dim, ld, dt, _ptr(queries), b, k, norm, _ptr(out_s),
                                           _ptr(out_r), _ptr(ws), ws.numel(), p, _stream(self.device))
                _lib.check(rc, "vq_scan_topk")
                self.last_path = _lib.last_scan_path()
                self.last_launches = _lib.last_launch_count()
        return out_s, out_r

    def two_stage(self, bf16: torch.Tensor, f32: torch.Tensor, n: int, dim: int, queries: torch.Tensor, k: int,
                  k_cand: int, norm: int = _lib.NORM_EPS, score_eps: float = 2.0 ** -7 + 1e-4):
        """`vq_search_two_stage`: tensor-core scan of the bf16 copy for k_cand candidates, exact fp32
        re-score, best k, per-query certificate.  Returns (scores [b,k] f32, rows [b,k] i32,
        uncertified [b] i32) device tensors."""
        ld = bf16.stride(0)
        if f32.stride(0) != ld:
            raise ValueError("bf16 and fp32 copies must share the row stride")
        b = queries.shape[0]
        with torch.cuda.device(self.device):
            out_s = torch.empty((b, k), dtype=torch.float32, device=self.device)
            out_r = torch.empty((b, k), dtype=torch.int32, device=self.device)
            bad = torch.zeros((b,), dtype=torch.int32, device=self.device)
            if b == 0:
                return out_s, out_r, bad
            need = self.lib.vq_search_two_stage_workspace_bytes(n, dim, ld, b, k_cand)
            with self.lock:
                ws = self.ws.get(need)
                rc = self.lib.vq_search_two_stage(_ptr(bf16), _ptr(f32), n, dim, ld, _ptr(queries), b, k, k_cand, norm,
                                                  float(score_eps), _ptr(out_s), _ptr(out_r), _ptr(bad), _ptr(ws),
                                                  ws.numel(), _stream(self.device))
                _lib.check(rc, "vq_search_two_stage")
                self.last_path = _lib.last_scan_path()
                self.last_launches = _lib.last_launch_count()
        return out_s, out_r, bad

    def exact(self, st: "DeviceStore", queries: torch.Tensor, k: int, norm: int = _lib.NORM_EPS, stats: torch.Tensor | None = None):
        """`vq_search_exact`: the exact fp32 top-k in ONE tensor-core pass over the bf16 copy (+ fp32 re-score
        of the gathered candidates).  Returns (scores [b,k] f32, rows [b,k] i32, overflow [b] i32) device
        tensors; a query whose overflow flag is set (mass ties) must be re-run on the fp32 FMA scan.
        stats: optional [b, 2] int32 device tensor (rows gathered by the scan, rows re-scored)."""
        if st.bf16 is None or st.f32 is None or st.bounds is None:
            raise RuntimeError("exact search needs a store that keeps both the fp32 and the bf16 copy")
        b = queries.shape[0]
        with torch.cuda.device(self.device):
            out_s = torch.empty((b, k), dtype=torch.float32, device=self.device)
            out_r = torch.empty((b, k), dtype=torch.int32, device=self.device)
            over = torch.zeros((b,), dtype=torch.int32, device=self.device)
            if b == 0:
                return out_s, out_r, over
            need = self.lib.vq_search_exact_workspace_bytes(st.n, st.dim, st.ld, b, k)
            with self.lock:
                ws = self.ws.get(need)
                rc = self.lib.vq_search_exact(_ptr(st.bf16), _ptr(st.f32), st.n, st.dim, st.ld, _ptr(queries), b, k, norm,
                                              _ptr(st.bounds), _ptr(out_s), _ptr(out_r), _ptr(over), _ptr(stats), _ptr(ws), ws.numel(),
                                              _stream(self.device))
                _lib.check(rc, "vq_search_exact")
                self.last_path = _lib.last_scan_path()
                self.last_launches = _lib.last_launch_count()
        return out_s, out_r, over

    def collect(self, bf16: torch.Tensor, f32: torch.Tensor, n: int, dim: int, queries: torch.Tensor, k: int,
                thresholds: torch.Tensor, cap: int = 4096, norm: int = _lib.NORM_EPS, bounds: torch.Tensor | None = None):
        """`vq_search_collect`: gather every row whose bf16 score reaches thresholds[q], re-score them exactly
        (all of them, or with the store's `bounds` only those that can still belong to the top-k), best k.
        Returns (scores [b,k] f32, rows [b,k] i32, overflow [b] i32)."""
        ld = bf16.stride(0)
        b = queries.shape[0]
        with torch.cuda.device(self.device):
            out_s = torch.empty((b, k), dtype=torch.float32, device=self.device)
            out_r = torch.empty((b, k), dtype=torch.int32, device=self.device)
            over = torch.zeros((b,), dtype=torch.int32, device=self.device)
            if b == 0:
                return out_s, out_r, over
            if thresholds is not None:
                thresholds = thresholds.to(torch.float32).contiguous()
            need = self.lib.vq_search_collect_workspace_bytes(n, dim, ld, b, cap)
            with self.lock:
                ws = self.ws.get(need)
                rc = self.lib.vq_search_collect(_ptr(bf16), _ptr(f32), n, dim, ld, _ptr(queries), b, k, norm, _ptr(thresholds),
                                                cap, _ptr(bounds), _ptr(out_s), _ptr(out_r), _ptr(over), _ptr(ws), ws.numel(),
                                                _stream(self.device))
                _lib.check(rc, "vq_search_collect")
                self.last_path = _lib.last_scan_path()
                self.last_launches = _lib.last_launch_count()
        return out_s, out_r, over

    def rescore(self, f32: torch.Tensor, n: int, dim: int, queries_norm_padded: torch.Tensor,
                cand_rows: torch.Tensor, k: int):
        b, kc = cand_rows.shape
        with torch.cuda.device(self.device):
            out_s = torch.empty((b, k), dtype=torch.float32, device=self.device)
            out_r = torch.empty((b, k), dtype=torch.int32, device=self.device)
            tmp = torch.empty((b, kc), dtype=torch.float32, device=self.device)
            rc = self.lib.vq_rescore_topk(_ptr(f32), n, dim, f32.stride(0), _ptr(queries_norm_padded), b,
                                          _ptr(cand_rows), kc, k, _ptr(out_s), _ptr(out_r), _ptr(tmp), tmp.numel() * 4,
                                          _stream(self.device))
            _lib.check(rc, "vq_rescore_topk")
        return out_s, out_r

    def normalise_padded(self, queries: torch.Tensor, ld: int, norm: int) -> torch.Tensor:
        """[b, dim] -> normalised, zero-padded [b, ld] fp32 (vq_ingest_rows on the queries)."""
        b, dim = queries.shape
        with torch.cuda.device(self.device):
            out = torch.empty((b, ld), dtype=torch.float32, device=self.device)
            _lib.check(self.lib.vq_ingest_rows(_ptr(queries), b, dim, dim, _ptr(out), _lib.F32, ld, norm,
                                               _stream(self.device)), "vq_ingest_rows")
        return out

    def merge(self, scores: torch.Tensor, rows: torch.Tensor, offsets, k_out: int, g_stride: int = 0):
        """scores/rows: [g, b, k_in] views whose shard stride is `g_stride` elements (0 = dense);
        offsets: int64 [g] device tensor or None → ([b,k] f32, [b,k] i64)."""
        g, b, k_in = scores.shape
        with torch.cuda.device(self.device):
            out_s = torch.empty((b, k_out), dtype=torch.float32, device=self.device)
            out_r = torch.empty((b, k_out), dtype=torch.int64, device=self.device)
            rc = self.lib.vq_topk_merge(_ptr(scores), _ptr(rows), g, g_stride, b, k_in,
                                        _ptr(offsets), k_out, _ptr(out_s), _ptr(out_r), _stream(self.device))
            _lib.check(rc, "vq_topk_merge")
        return out_s, out_r


def as_device_queries(q, dim: int, device) -> torch.Tensor:
    """Accept numpy / torch, 1-D or 2-D, any float dtype → [b, dim] fp32 device tensor."""
    if isinstance(q, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(q, dtype=np.float32))
    elif isinstance(q, torch.Tensor):
        t = q.detach().to(torch.float32)
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(q), dtype=np.float32))
    if t.dim() == 1:
        t = t[None, :]
    if t.dim() != 2 or t.shape[1] != dim:
        raise ValueError(f"query shape {tuple(t.shape)} does not match dimension {dim}")
    return t.to(device).contiguous()
