"""Peer-memory exchange of the shard/merge layer.

`PeerExchange` owns this rank's exchange window (include/vq_search.h: vq_peer_window_*) and the
mapped windows of all peers of a process group on ONE node; `exchange_merge` is the fused
push -> wait -> merge kernel that replaces `all_gather_into_tensor` + `vq_topk_merge`: the local
top-k of every rank is stored straight into the peers' HBM over NVLink and merged as soon as the
same queries of all ranks have landed.  The reference has no counterpart (single process).

There is no CPU path: construction raises without CUDA.  `LocalWindows` builds the same windows
for several *simulated* ranks inside one process on one GPU (plain device buffers instead of IPC
mappings, one stream per rank) — the protocol test that runs on a single-GPU box.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _launch(lib, windows_dev, world, rank, b_max, k_max, scores, rows, offsets, k_out, status, stream):
    b, k = scores.shape
    assert scores.dtype == torch.float32 and rows.dtype == torch.int32
    assert scores.is_contiguous() and rows.is_contiguous() and rows.shape == scores.shape
    out_s = torch.empty((b, k_out), dtype=torch.float32, device=scores.device)
    out_r = torch.empty((b, k_out), dtype=torch.int64, device=scores.device)
    rc = lib.vq_peer_exchange_merge(_ptr(windows_dev), world, rank, b_max, k_max, _ptr(scores), _ptr(rows), b, k,
                                    _ptr(offsets), k_out, _ptr(out_s), _ptr(out_r), _ptr(status),
                                    C.c_void_p(stream.cuda_stream))
    _lib.check(rc, "vq_peer_exchange_merge")
    return out_s, out_r


class _PeerWindows:
    """One exchange window per rank, mapped by all peers of a process group on ONE node (CUDA IPC)."""

    def __init__(self, device, group, nbytes: int):
        if not torch.cuda.is_available():
            raise RuntimeError("peer exchange windows need CUDA (there is no CPU path)")
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._local = C.c_void_p()
        self._peers: list[Optional[C.c_void_p]] = []
        self._closed = False
        with torch.cuda.device(self.device):
            # every rank takes part in every collective below whatever fails locally, and all ranks
            # agree on the outcome — a refused IPC mapping must not leave the others in a barrier
            err = None
            handle = C.create_string_buffer(64)
            try:
                _lib.check(self.lib.vq_peer_window_create(nbytes, C.byref(self._local), handle), "vq_peer_window_create")
            except _lib.VQError as e:
                err = e
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle.raw) if err is None else None, group=group)
            ptrs = []
            for r, h in enumerate(handles):
                if r == self.rank or h is None or err is not None:
                    self._peers.append(None)
                    ptrs.append(self._local.value or 0)
                    continue
                p = C.c_void_p()
                try:
                    _lib.check(self.lib.vq_peer_window_open(h, C.byref(p)), f"vq_peer_window_open(rank {r})")
                except _lib.VQError as e:
                    err = e
                    p = None
                self._peers.append(p)
                ptrs.append(p.value if p is not None else 0)
            ok = torch.tensor([0 if (err is not None or any(h is None for h in handles)) else 1],
                              dtype=torch.int32, device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                self.close()
                raise _lib.VQError(f"peer windows could not be mapped on every rank ({err or 'failure on another rank'})")
            self.windows = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
            self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
            torch.cuda.synchronize(self.device)
        dist.barrier(group=group)          # every window is zeroed and mapped before the first push

    def check(self):
        """Raises if a wait timed out since the windows were created (synchronises the device)."""
        ep, err = C.c_uint32(), C.c_uint32()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.vq_peer_window_status(self._local, C.byref(ep), C.byref(err)), "vq_peer_window_status")
        if err.value:
            raise _lib.VQError(f"peer exchange: a peer did not arrive in time (rank {self.rank}, epoch {ep.value})")
        return int(ep.value)

    def close(self):
        if self._closed:
            return
        self._closed = True
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for p in self._peers:
                if p is not None:
                    self.lib.vq_peer_window_close(p)
            self._peers = []
            if dist.is_initialized():
                dist.barrier(group=self.group)          # nobody still maps the window we are about to free
            if self._local.value is not None:
                self.lib.vq_peer_window_destroy(self._local)
            self._local = C.c_void_p()


class PeerExchange(_PeerWindows):
    """Fused exchange + merge of the per-shard top-k (vq_peer_exchange_merge)."""

    def __init__(self, device, group=None, b_max: int = 1024, k_max: int = 16):
        self.b_max, self.k_max = int(b_max), int(k_max)
        lib = _lib.load()
        super().__init__(device, group, lib.vq_peer_window_bytes(dist.get_world_size(group), self.b_max, self.k_max))

    def fits(self, b: int, k: int) -> bool:
        return b <= self.b_max and k <= self.k_max

    def exchange_merge(self, scores: torch.Tensor, rows: torch.Tensor, offsets: Optional[torch.Tensor], k_out: int):
        """scores/rows [b,k] (local rows, best first) -> (scores [b,k_out] f32, global rows [b,k_out] i64),
        identical on every rank.  Collective: same shapes, same order on all ranks, one stream."""
        with torch.cuda.device(self.device):
            return _launch(self.lib, self.windows, self.world, self.rank, self.b_max, self.k_max, scores, rows,
                           offsets, k_out, self.status, torch.cuda.current_stream(self.device))


def slice_range(b: int, world: int, rank: int):
    """Rows of a [b, dim] query batch that `rank` ingests (vq_peer_allgather_rows)."""
    per = -(-b // world)
    return min(b, rank * per), min(b, (rank + 1) * per)


def _launch_rows(lib, windows_dev, world, rank, b_max, ld_max, slice_, b, dim, status, stream):
    assert slice_.dtype == torch.float32 and slice_.is_contiguous()
    lo, hi = slice_range(b, world, rank)
    assert tuple(slice_.shape) == (hi - lo, dim), (tuple(slice_.shape), hi - lo, dim)
    out = torch.empty((b, dim), dtype=torch.float32, device=slice_.device)
    rc = lib.vq_peer_allgather_rows(_ptr(windows_dev), world, rank, b_max, ld_max, _ptr(slice_) if hi > lo else None,
                                    b, dim, _ptr(out), _ptr(status), C.c_void_p(stream.cuda_stream))
    _lib.check(rc, "vq_peer_allgather_rows")
    return out


class PeerRowGather(_PeerWindows):
    """All-gather of the query batch over peer memory (vq_peer_allgather_rows): every rank ingests only
    its `slice_range` of the host batch over PCIe, the slices travel over NVLink."""

    def __init__(self, device, group=None, b_max: int = 1024, ld_max: int = 512):
        self.b_max, self.ld_max = int(b_max), int(ld_max)
        lib = _lib.load()
        super().__init__(device, group, lib.vq_peer_rows_window_bytes(dist.get_world_size(group), self.b_max, self.ld_max))

    def allgather_rows(self, slice_: torch.Tensor, b: int):
        """slice_: this rank's rows [hi-lo, dim] fp32 -> the whole batch [b, dim] on every rank (collective)."""
        with torch.cuda.device(self.device):
            return _launch_rows(self.lib, self.windows, self.world, self.rank, self.b_max, self.ld_max, slice_, b,
                                slice_.shape[1], self.status, torch.cuda.current_stream(self.device))


class LocalWindows:
    """`world` simulated ranks on ONE GPU: one zeroed device buffer per rank as its window and one
    stream per rank; the `world` kernels wait for each other exactly like ranks on different GPUs
    (they must be co-resident: keep world * ceil(b/8) CTAs well below the GPU's capacity)."""

    def __init__(self, world: int, device, b_max: int, k_max: int, rows_ld: int = 0):
        """rows_ld > 0: windows for vq_peer_allgather_rows (k_max unused) instead of the top-k exchange."""
        if not torch.cuda.is_available():
            raise RuntimeError("LocalWindows needs CUDA (there is no CPU path)")
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.world, self.b_max, self.k_max, self.rows_ld = world, b_max, k_max, rows_ld
        nbytes = self.lib.vq_peer_rows_window_bytes(world, b_max, rows_ld) if rows_ld else \
            self.lib.vq_peer_window_bytes(world, b_max, k_max)
        self.bufs = [torch.zeros(nbytes, dtype=torch.uint8, device=self.device) for _ in range(world)]
        self.windows = torch.tensor([t.data_ptr() for t in self.bufs], dtype=torch.int64, device=self.device)
        self.status = torch.zeros(world, dtype=torch.int32, device=self.device)
        self.streams = [torch.cuda.Stream(self.device) for _ in range(world)]

    def exchange_merge_all(self, scores, rows, offsets, k_out):
        """scores[r]/rows[r]: the local candidates of simulated rank r -> list of per-rank outputs."""
        cur = torch.cuda.current_stream(self.device)
        outs = []
        for r in range(self.world):
            self.streams[r].wait_stream(cur)
            with torch.cuda.stream(self.streams[r]):
                outs.append(_launch(self.lib, self.windows, self.world, r, self.b_max, self.k_max, scores[r], rows[r],
                                    offsets, k_out, self.status[r:r + 1], self.streams[r]))
        for st in self.streams:
            cur.wait_stream(st)
        return outs

    def allgather_rows_all(self, slices, b: int):
        """slices[r]: the query rows simulated rank r ingests -> list of per-rank [b, dim] outputs."""
        cur = torch.cuda.current_stream(self.device)
        outs = []
        for r in range(self.world):
            self.streams[r].wait_stream(cur)
            with torch.cuda.stream(self.streams[r]):
                outs.append(_launch_rows(self.lib, self.windows, self.world, r, self.b_max, self.rows_ld, slices[r], b,
                                         slices[r].shape[1], self.status[r:r + 1], self.streams[r]))
        for st in self.streams:
            cur.wait_stream(st)
        return outs
