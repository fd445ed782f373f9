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
    def __init__(self, fn: Callable, batch: int, dim: int, device, warmup: int = 3):
        """fn(queries [batch, dim] fp32 device tensor) -> tensor or tuple of tensors."""
        self.device = torch.device(device)
        self.q = torch.zeros((batch, dim), dtype=torch.float32, device=self.device)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):               # sizes workspaces, caches tensor maps / attributes
                fn(self.q)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = fn(self.q)

    def __call__(self, queries: torch.Tensor):
        """queries: [batch, dim] fp32, host (pinned for async copy) or device.  The returned tensors
        are the graph's static outputs: consume or copy them before the next call."""
        self.q.copy_(queries, non_blocking=True)
        self.graph.replay()
        return self.out
