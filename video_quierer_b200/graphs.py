"""CUDA-graph capture of a fixed-shape search step.

A search step is a short chain of small launches (query prep -> scan -> merge -> re-score ->
merge [-> all-gather -> merge]); at batch 1-32 the launch gaps and the Python/ctypes overhead
are comparable to the scan itself.  `GraphedSearch` captures the whole chain once for a fixed
(batch, dim) shape and replays it with a single launch — the serving path for steady traffic
(the micro-batcher pads to a few fixed batch sizes).  Everything captured is stream-ordered on
torch's capture stream; the C-ABI only enqueues work, so it is capture-safe.
"""

from __future__ import annotations

from typing import Callable

import torch


class GraphedSearch:
    def __init__(self, fn: Callable, batch: int, dim: int, device, warmup: int = 3, stream=None, host_io: bool = False,
                 ingest: Callable | None = None):
        """fn(queries [batch, dim] fp32 device tensor) -> tensor or tuple of tensors (None entries allowed).
        stream: the stream the graph is replayed on (None = whatever stream is current at the call).
        host_io: the host<->device copies are part of the captured step — `host_q` (pinned, [batch, dim]) is
        the request slot the batcher fills, `host_out` the pinned twins of the outputs; one replay = H2D of
        the queries + search + D2H of the results, with no per-step Python work beyond the launch.
        ingest: replaces the H2D copy of `host_q`: a callable that enqueues whatever brings the step's queries
        onto the device (e.g. H2D of this rank's slice + `peer.PeerRowGather.allgather_rows`) and returns the
        [batch, dim] device tensor; it is captured with the step (implies pinned result slots like host_io)."""
        self.device = torch.device(device)
        self.stream = stream
        self._fn, self._ingest = fn, ingest          # the captured nodes point into whatever these closures own
        self.q = torch.zeros((batch, dim), dtype=torch.float32, device=self.device)
        host_io = host_io or ingest is not None
        self.host_q = torch.zeros((batch, dim), dtype=torch.float32).pin_memory() if (host_io and ingest is None) else None
        self.host_out = None
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):               # sizes workspaces, caches tensor maps / attributes
                probe = fn(ingest() if ingest is not None else self.q)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        if host_io:                                       # pinned memory cannot be allocated while capturing
            outs = probe if isinstance(probe, (tuple, list)) else (probe,)
            self.host_out = tuple(None if o is None else torch.empty(tuple(o.shape), dtype=o.dtype).pin_memory() for o in outs)
        del probe
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            if ingest is not None:
                self.q = ingest()
            elif host_io:
                self.q.copy_(self.host_q, non_blocking=True)
            self.out = fn(self.q)
            if host_io:
                outs = self.out if isinstance(self.out, (tuple, list)) else (self.out,)
                for h, o in zip(self.host_out, outs):
                    if h is not None:
                        h.copy_(o, non_blocking=True)

    def replay(self):
        if self.stream is None:
            self.graph.replay()
        else:
            with torch.cuda.stream(self.stream):
                self.graph.replay()
        return self.out

    def __call__(self, queries: torch.Tensor):
        """queries: [batch, dim] fp32, host (pinned for async copy) or device.  The returned tensors
        are the graph's static outputs: consume or copy them before the next call (on `stream`)."""
        if self.stream is None:
            self.q.copy_(queries, non_blocking=True)
        else:
            with torch.cuda.stream(self.stream):
                self.q.copy_(queries, non_blocking=True)
        return self.replay()


class PipelinedSearch:
    """`depth` search lanes, each with its own stream, workspace and captured step; consecutive batches
    go to consecutive lanes.  A step is one bandwidth/tensor-bound scan between latency-bound kernels
    (query prep, threshold bootstrap, final selection + re-score, shard exchange): with two batches in
    flight the head and tail of one overlap the scan of the other.  Every lane needs its OWN scanner
    (workspace) and, when sharded, its own exchange windows: `make_fn(lane)` must return a step function
    that shares nothing mutable with the other lanes (the store is read-only and shared)."""

    def __init__(self, make_fn: Callable, batch: int, dim: int, device, depth: int = 2, graph: bool = True):
        self.device = torch.device(device)
        self.depth = depth
        self.streams = [torch.cuda.Stream(self.device) for _ in range(depth)]
        self.fns = [make_fn(i) for i in range(depth)]
        self.lanes = [GraphedSearch(self.fns[i], batch, dim, device, stream=self.streams[i]) if graph else None
                      for i in range(depth)]
        self.q = [None if graph else torch.zeros((batch, dim), dtype=torch.float32, device=self.device) for _ in range(depth)]
        self.done = [torch.cuda.Event() for _ in range(depth)]
        self.n = 0

    def submit(self, queries: torch.Tensor):
        """Enqueue one batch on the next lane; returns (lane, outputs).  The outputs are valid once
        `self.done[lane]` has completed and until the lane is used again (depth submissions later)."""
        lane = self.n % self.depth
        self.n += 1
        st = self.streams[lane]
        st.wait_stream(torch.cuda.current_stream(self.device))
        if self.lanes[lane] is not None:
            out = self.lanes[lane](queries)
        else:
            with torch.cuda.stream(st):
                self.q[lane].copy_(queries, non_blocking=True)
                out = self.fns[lane](self.q[lane])
        self.done[lane].record(st)
        return lane, out

    def drain(self):
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            cur.wait_stream(st)
