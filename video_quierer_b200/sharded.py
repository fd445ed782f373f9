"""Shard/merge layer — one process per GPU over torch.distributed.

The frame-embedding store is row-sharded: rank r owns the contiguous rows
``[r*ceil(N/G), min(N, (r+1)*ceil(N/G)))``.  A search is
  1. every rank scans its shard for a local top-k   (vq_scan_topk / vq_hnsw_search),
  2+3. on CUDA ranks of one node (`exchange="peer"`, the default there): ONE fused kernel per rank
     (`peer.PeerExchange`, vq_peer_exchange_merge) stores the local candidates straight into every
     peer's HBM over NVLink, waits per CTA for the same queries of the other ranks and merges
     ``g*k -> k`` with the shard offsets added — no NCCL call on the data path;
     otherwise (`exchange="collective"`: gloo, several nodes, or peer mapping unavailable):
  2. ONE all-gather of the packed ``[scores | rows]`` candidate block (8*b*k bytes per rank,
     latency-bound on NVSwitch),
  3. an on-device merge ``g*k -> k`` that adds the shard offsets (vq_topk_merge reads the
     packed gather buffer in place through its shard stride).
The reference has no counterpart (single process).  The collective plumbing is independent of
CUDA so that it can be exercised with the gloo backend on CPU: `local_search` and `merge` are
injectable; the defaults are the CUDA kernels and raise if no GPU is present.
"""

from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist


def shard_range(n_total: int, world: int, rank: int):
    """Contiguous row range of `rank`."""
    per = -(-n_total // world) if world > 0 else n_total
    lo = min(n_total, rank * per)
    hi = min(n_total, (rank + 1) * per)
    return lo, hi


def shard_offsets(n_total: int, world: int):
    return [shard_range(n_total, world, r)[0] for r in range(world)]


def pack_candidates(scores: torch.Tensor, rows: torch.Tensor) -> torch.Tensor:
    """[b,k] fp32 + [b,k] int32 → one int32 buffer [2, b, k] (scores bit-cast)."""
    return torch.stack((scores.contiguous().view(torch.int32), rows.contiguous().to(torch.int32)), dim=0)


class ShardedSearcher:
    """Row-sharded exact/ANN search across the ranks of a process group."""

    def __init__(self, local_search: Callable, n_total: int, group=None,
                 merge: Optional[Callable] = None, device=None, exchange: str = "auto"):
        """local_search(queries[b,dim], k) -> (scores [b,k] fp32, rows [b,k] int32) on this
        rank's shard (local row numbers, -1 = empty slot).
        exchange: "peer" (fused NVLink push + merge kernel; raises if the windows cannot be mapped),
        "collective" (all-gather + merge) or "auto" (peer on CUDA ranks with the NCCL backend and no
        injected merge, falling back to the collective — with a message on stderr — if the IPC
        mapping is refused)."""
        if exchange not in ("auto", "peer", "collective"):
            raise ValueError(f"exchange must be auto|peer|collective, got {exchange!r}")
        self.exchange = exchange
        self._peer = None
        self._peer_failed = False
        self._broken = False
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_total = int(n_total)
        self.local_search = local_search
        self.device = device
        self._merge = merge
        self._offsets = None
        self._scanner = None
        self._gather_buf = None

    # default merge = the CUDA kernel, reading the packed gather buffer in place
    def _cuda_merge(self, gathered: torch.Tensor, k_out: int):
        from .engine import Scanner
        if self._scanner is None:
            self._scanner = Scanner(self.device)
            self._offsets = torch.tensor(shard_offsets(self.n_total, self.world), dtype=torch.int64,
                                         device=gathered.device)
        g, _, b, k = gathered.shape
        scores = gathered[:, 0].view(torch.float32)          # [g, b, k] view, shard stride 2*b*k
        rows = gathered[:, 1]
        return self._scanner.merge(scores, rows, self._offsets, k_out, g_stride=2 * b * k)

    def _peer_exchange(self, s: torch.Tensor, r: torch.Tensor):
        """The PeerExchange serving [b,k] candidates, (re)built collectively when the shape outgrows it;
        None = use the collective route."""
        if self.exchange == "collective" or self._merge is not None or self.world == 1 or not s.is_cuda:
            return None
        if self.exchange == "auto" and (self._peer_failed or dist.get_backend(self.group) != "nccl"):
            return None
        b, k = s.shape
        if self._peer is not None and self._peer.fits(b, k):
            return self._peer
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("peer exchange windows must be sized before CUDA-graph capture (run one eager search)")
        from .peer import PeerExchange
        try:
            if self._peer is not None:
                old, self._peer = self._peer, None
                b, k = max(b, old.b_max), max(k, old.k_max)
                old.close()
            self._peer = PeerExchange(s.device, self.group, b_max=b, k_max=k)
            if self._offsets is None:
                self._offsets = torch.tensor(shard_offsets(self.n_total, self.world), dtype=torch.int64, device=s.device)
        except Exception as e:  # noqa: BLE001 — IPC refused (container policy, no P2P): all ranks see the same
            if self.exchange == "peer":
                raise
            import sys
            print(f"[sharded] peer windows unavailable ({type(e).__name__}: {e}); using all-gather + merge", file=sys.stderr)
            self._peer_failed = True
            self._peer = None
        return self._peer

    def check(self):
        """Raises if the peer exchange recorded a timeout (synchronises); no-op on the collective route."""
        if self._peer is not None:
            try:
                self._peer.check()
            except Exception:
                self._broken = True
                raise

    @property
    def status(self) -> Optional[torch.Tensor]:
        """Device int32[1], sticky: non-zero once a peer missed an exchange (its shard's candidates were then
        replaced by empty slots, so every result since is PARTIAL).  None on the collective route.  Consumers
        that replay a captured step copy it to the host together with the results (`search_checked` does)."""
        return None if self._peer is None else self._peer.status

    def search_checked(self, queries: torch.Tensor, k: int):
        """`search` + device-to-host copy of the results AND of the exchange status in the same
        synchronisation: raises instead of returning partial results when a peer did not arrive in time.
        Returns (scores [b,k] float32, global rows [b,k] int64) numpy arrays.  After a failure the searcher
        refuses further searches until `reset()` has been called on EVERY rank."""
        s, r = self.search(queries, k)
        st = self.status
        s_h, r_h = s.cpu().numpy(), r.cpu().numpy()
        if st is not None and int(st.cpu().item()) != 0:
            self._broken = True
            from ._lib import VQError
            raise VQError(f"shard exchange on rank {self.rank}: a peer did not deliver its candidates in time; the merged "
                          "top-k would be partial.  Call reset() on every rank to rebuild the exchange windows.")
        return s_h, r_h

    def reset(self):
        """Collective: drop the exchange windows (their epoch counters may be out of step after a rank missed
        an exchange); the next search rebuilds them with all ranks taking part."""
        self.close()
        self._broken = False
        self._peer_failed = False

    def close(self):
        if self._peer is not None:
            self._peer.close()
            self._peer = None

    def search(self, queries: torch.Tensor, k: int):
        """Returns (scores [b,k] fp32, global rows [b,k] int64), identical on every rank.  Everything is
        stream-ordered (capturable); a peer that misses the exchange is reported through `status` — use
        `search_checked` (or read `status` with the results) wherever the results leave the device."""
        if self._broken:
            raise RuntimeError("the shard exchange failed earlier (a peer timed out): call reset() on every rank first")
        s, r = self.local_search(queries, k)
        if self.world == 1 and self._merge is None:
            return s, r.to(torch.int64)                       # one shard: nothing to exchange or merge
        px = self._peer_exchange(s, r)
        if px is not None:
            return px.exchange_merge(s.contiguous(), r.contiguous().to(torch.int32), self._offsets, k)
        packed = pack_candidates(s, r)                        # [2, b, k] int32
        if self.world == 1:
            gathered = packed[None]
        else:
            # flat (world*2, b, k) buffer: the concatenated form every backend accepts
            shape = (self.world * packed.shape[0],) + tuple(packed.shape[1:])
            if self._gather_buf is None or tuple(self._gather_buf.shape) != shape or \
                    self._gather_buf.device != packed.device:
                self._gather_buf = torch.empty(shape, dtype=torch.int32, device=packed.device)
            dist.all_gather_into_tensor(self._gather_buf, packed, group=self.group)
            gathered = self._gather_buf.view((self.world,) + tuple(packed.shape))
        merge = self._merge or self._cuda_merge
        return merge(gathered, k)
