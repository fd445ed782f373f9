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


class PeerExchange:
    """Exchange windows of a process group whose ranks sit on one node (one GPU per rank)."""

    def __init__(self, device, group=None, b_max: int = 1024, k_max: int = 16):
        if not torch.cuda.is_available():
            raise RuntimeError("PeerExchange needs CUDA (there is no CPU path)")
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.b_max, self.k_max = int(b_max), int(k_max)
        self._local = C.c_void_p()
        self._peers: list[Optional[C.c_void_p]] = []
        self._closed = False
        with torch.cuda.device(self.device):
            # every rank takes part in every collective below whatever fails locally, and all ranks
            # agree on the outcome — a refused IPC mapping must not leave the others in a barrier
            err = None
            nbytes = self.lib.vq_peer_window_bytes(self.world, self.b_max, self.k_max)
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

    def fits(self, b: int, k: int) -> bool:
        return b <= self.b_max and k <= self.k_max

    def exchange_merge(self, scores: torch.Tensor, rows: torch.Tensor, offsets: Optional[torch.Tensor], k_out: int):
        """scores/rows [b,k] (local rows, best first) -> (scores [b,k_out] f32, global rows [b,k_out] i64),
        identical on every rank.  Collective: same shapes, same order on all ranks, one stream."""
        with torch.cuda.device(self.device):
            return _launch(self.lib, self.windows, self.world, self.rank, self.b_max, self.k_max, scores, rows,
                           offsets, k_out, self.status, torch.cuda.current_stream(self.device))

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


class LocalWindows:
    """`world` simulated ranks on ONE GPU: one zeroed device buffer per rank as its window and one
    stream per rank; the `world` kernels wait for each other exactly like ranks on different GPUs
    (they must be co-resident: keep world * ceil(b/8) CTAs well below the GPU's capacity)."""

    def __init__(self, world: int, device, b_max: int, k_max: int):
        if not torch.cuda.is_available():
            raise RuntimeError("LocalWindows needs CUDA (there is no CPU path)")
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.world, self.b_max, self.k_max = world, b_max, k_max
        nbytes = self.lib.vq_peer_window_bytes(world, b_max, k_max)
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
