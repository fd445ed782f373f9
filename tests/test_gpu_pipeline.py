"""GPU test of `graphs.PipelinedSearch`: two search steps in flight on two streams (own workspace per lane,
shared read-only store) return exactly what the same steps return one after the other."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("graph", [True, False])
def test_two_lanes_match_sequential(built_lib, graph):
    from video_quierer_b200 import engine
    from video_quierer_b200.flat_index import exact_search
    from video_quierer_b200.graphs import PipelinedSearch
    from video_quierer_b200.utils import synth
    dev = torch.device("cuda", 0)
    n, dim, b, k = 60_000, 512, 200, 10
    store = engine.DeviceStore(dim, dev, keep_fp32=True, keep_bf16=True)
    store.append(torch.from_numpy(synth.gauss(n, dim, seed=5)).to(dev))
    rng = np.random.default_rng(11)
    batches = [torch.from_numpy(rng.standard_normal((b, dim), dtype=np.float32)).pin_memory() for _ in range(7)]

    ref_scanner = engine.Scanner(dev)
    ref = []
    for q in batches:
        s, r, bad = exact_search(ref_scanner, store, q.to(dev), k)
        assert int(bad.sum()) == 0
        ref.append((s.cpu(), r.cpu()))

    def make_fn(lane):
        sc = engine.Scanner(dev)                       # lane-private workspace
        return lambda qq: exact_search(sc, store, qq, k)

    pipe = PipelinedSearch(make_fn, b, dim, dev, depth=2, graph=graph)
    got = [None] * len(batches)
    pending = {}
    for i, q in enumerate(batches):
        lane = pipe.n % pipe.depth
        if lane in pending:                            # harvest the lane's previous step before reusing it
            j, (s, r, bad) = pending.pop(lane)
            pipe.done[lane].synchronize()
            got[j] = (s.cpu(), r.cpu())
        lane, out = pipe.submit(q)
        pending[lane] = (i, out)
    for lane, (j, (s, r, bad)) in pending.items():
        pipe.done[lane].synchronize()
        got[j] = (s.cpu(), r.cpu())
    pipe.drain()
    torch.cuda.synchronize()
    for j, ((s, r), (rs, rr)) in enumerate(zip(got, ref)):
        assert torch.equal(r, rr), f"batch {j}: rows differ"
        assert torch.equal(s, rs), f"batch {j}: scores differ"
