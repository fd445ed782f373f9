"""One simulated rank (world = 1) driving vq_peer_exchange_merge on one GPU: the kernel's own cost
(push into its own window, flag, stage, merge) without a peer to wait for — the form ncu can capture
(under ncu kernels are serialised, so simulated ranks that wait for each other would time out)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_quierer_b200.peer import LocalWindows  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda", 0)
lw = LocalWindows(1, dev, a.batch, max(a.k, 16))
s = torch.sort(torch.randn((a.batch, a.k), device=dev), dim=1, descending=True).values.contiguous()
r = torch.randint(0, 1 << 20, (a.batch, a.k), dtype=torch.int32, device=dev)
off = torch.zeros(1, dtype=torch.int64, device=dev)
for _ in range(3):
    lw.exchange_merge_all([s], [r], off, a.k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    out = lw.exchange_merge_all([s], [r], off, a.k)
e1.record()
torch.cuda.synchronize()
assert int(lw.status.sum()) == 0 and torch.equal(out[0][0], s)
print(f"peer_exchange_merge world=1 batch={a.batch} k={a.k}: {e0.elapsed_time(e1) / a.iters * 1e3:.1f} us per launch (incl. stream fork/join)")
