"""World-size-2 (and 3) CPU tests of the shard/merge layer over the gloo backend.

The collective plumbing of `ShardedSearcher` (row ranges, candidate packing, one all-gather,
shard offsets, rank order) is backend-independent; on CPU the per-shard scan and the merge are
played by the oracle (tests may use it), on the GPU box they are the CUDA kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import compare, exact
from video_quierer_b200.sharded import ShardedSearcher, pack_candidates, shard_offsets, shard_range
from video_quierer_b200.utils import synth


def test_shard_ranges_cover_and_are_contiguous():
    for n in (0, 1, 7, 1000, 1_000_003):
        for g in (1, 2, 3, 8):
            r = [shard_range(n, g, i) for i in range(g)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(g - 1))
            assert shard_offsets(n, g) == [lo for lo, _ in r]
            assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= -(-n // g)


def test_pack_roundtrip():
    s = torch.randn(3, 5)
    r = torch.randint(0, 100, (3, 5), dtype=torch.int32)
    p = pack_candidates(s, r)
    assert p.dtype == torch.int32 and tuple(p.shape) == (2, 3, 5)
    assert torch.equal(p[0].view(torch.float32), s) and torch.equal(p[1], r)


def _oracle_merge(n_total, world):
    offs = shard_offsets(n_total, world)

    def merge(gathered: torch.Tensor, k: int):
        g, _, b, kk = gathered.shape
        scores = gathered[:, 0].contiguous().view(torch.float32).numpy()
        rows = gathered[:, 1].numpy().astype(np.int64)
        out_s = np.full((b, k), -np.inf, np.float32)
        out_r = np.full((b, k), -1, np.int64)
        for q in range(b):
            cand = [(-(scores[s, q, j]), rows[s, q, j] + offs[s]) for s in range(g) for j in range(kk) if rows[s, q, j] >= 0]
            cand.sort()
            for i, (ns, r) in enumerate(cand[:k]):
                out_s[q, i], out_r[q, i] = -ns, r
        return torch.from_numpy(out_s), torch.from_numpy(out_r)
    return merge


def _worker(rank, world, port, n, dim, k, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        store = synth.gauss(n, dim, seed=3)
        queries = np.random.default_rng(4).standard_normal((6, dim), dtype=np.float32)
        lo, hi = shard_range(n, world, rank)
        shard = store[lo:hi]

        def local_search(q, kk):
            kk_eff = min(kk, len(shard))
            rows, scores = exact.exact_search_batch(shard, q.numpy(), kk_eff) if len(shard) else (np.zeros((len(q), 0), np.int64), np.zeros((len(q), 0)))
            s = np.full((len(q), kk), -np.inf, np.float32); r = np.full((len(q), kk), -1, np.int32)
            s[:, :kk_eff] = scores; r[:, :kk_eff] = rows
            return torch.from_numpy(s), torch.from_numpy(r)

        ss = ShardedSearcher(local_search, n, merge=_oracle_merge(n, world))
        s, r = ss.search(torch.from_numpy(queries), k)
        ro, so = exact.exact_search_batch(store, queries, min(k, n))
        bad = compare.check_topk_batch(r.numpy()[:, : min(k, n)], s.numpy()[:, : min(k, n)], ro, so)
        # checked read-out: on the collective route there is no exchange status, results come back as numpy; a
        # searcher whose exchange failed refuses to search until every rank has reset it
        assert ss.status is None
        s2, r2 = ss.search_checked(torch.from_numpy(queries), k)
        assert np.array_equal(r2, r.numpy()) and np.array_equal(s2, s.numpy())
        ss._broken = True
        try:
            ss.search(torch.from_numpy(queries), k)
            refused = False
        except RuntimeError:
            refused = True
        ss.reset()
        s3, r3 = ss.search_checked(torch.from_numpy(queries), k)
        assert refused and np.array_equal(r3, r.numpy())
        ret[rank] = (len(bad), ss.world, ss.rank)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,n,k", [(2, 1001, 10), (3, 64, 10), (2, 5, 10)])
def test_sharded_search_equals_unsharded(world, n, k):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, 32, k, ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank in range(world):
        bad, w, r = ret[rank]
        assert bad == 0 and w == world and r == rank
