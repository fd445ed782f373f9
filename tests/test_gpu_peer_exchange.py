"""GPU tests of the fused peer-memory exchange + merge kernel (vq_peer_exchange_merge).

1. protocol on ONE GPU: `world` simulated ranks (one window + one stream each, `peer.LocalWindows`) push
   into each other's windows and wait for each other exactly like ranks on different GPUs; results are
   checked against a CPU merge of the same candidates (score desc, global row asc) for several epochs,
   ragged batches, k_out < k, empty slots and exact score ties across shards;
2. the same kernel behind `ShardedSearcher(exchange="peer")` on 2 real GPUs over CUDA IPC, one process per
   GPU (skipped on a single-GPU box), eager and replayed from a CUDA graph, against the all-gather route.
"""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cpu_merge(scores, rows, offsets, k_out):
    """scores/rows: [g, b, k] numpy -> ([b,k_out] f32, [b,k_out] i64)."""
    g, b, k = scores.shape
    out_s = np.full((b, k_out), -np.inf, np.float32)
    out_r = np.full((b, k_out), -1, np.int64)
    for q in range(b):
        cand = [(-float(scores[s, q, j]), int(rows[s, q, j]) + int(offsets[s]))
                for s in range(g) for j in range(k) if rows[s, q, j] >= 0]
        cand.sort()
        for i, (ns, r) in enumerate(cand[:k_out]):
            out_s[q, i], out_r[q, i] = -ns, r
    return out_s, out_r


def _local_lists(rng, g, b, k, n_local, ties=False, holes=False):
    """Per-shard sorted candidate lists like a local search returns them (best first, -1 padded)."""
    scores = rng.standard_normal((g, b, k)).astype(np.float32)
    if ties:                                          # the same score values on every shard
        scores[:] = np.round(scores[0:1] * 4) / 4
    rows = np.stack([np.stack([rng.choice(n_local, size=k, replace=False) for _ in range(b)]) for _ in range(g)]).astype(np.int32)
    # order every list by (score desc, row asc)
    for s in range(g):
        for q in range(b):
            order = sorted(range(k), key=lambda j: (-scores[s, q, j], rows[s, q, j]))
            scores[s, q], rows[s, q] = scores[s, q, order], rows[s, q, order]
    if holes:                                         # shards that found fewer than k rows
        for s in range(g):
            for q in range(0, b, 3):
                cut = int(rng.integers(0, k))
                scores[s, q, cut:] = -np.inf
                rows[s, q, cut:] = -1
    return scores, rows


@pytest.mark.parametrize("world,b,k,k_out", [(2, 37, 10, 10), (4, 128, 10, 10), (8, 64, 10, 7), (8, 200, 32, 32), (3, 5, 100, 100)])
def test_simulated_ranks_protocol(built_lib, world, b, k, k_out):
    from video_quierer_b200.peer import LocalWindows
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(world * 1000 + b)
    n_local = 5000
    offsets = np.arange(world, dtype=np.int64) * n_local
    lw = LocalWindows(world, dev, b_max=max(b, 256), k_max=max(k, 16))
    off_d = torch.from_numpy(offsets).to(dev)
    for epoch in range(5):                            # both parities, flags reused
        bb = b if epoch != 2 else max(1, b // 2)      # a smaller batch in between (fewer CTAs)
        s, r = _local_lists(rng, world, bb, k, n_local, ties=(epoch == 3), holes=(epoch == 4))
        sd = [torch.from_numpy(s[i]).to(dev) for i in range(world)]
        rd = [torch.from_numpy(r[i]).to(dev) for i in range(world)]
        outs = lw.exchange_merge_all(sd, rd, off_d, k_out)
        torch.cuda.synchronize()
        assert int(lw.status.sum()) == 0, "a simulated rank timed out"
        ref_s, ref_r = _cpu_merge(s, r, offsets, k_out)
        for i, (os_, or_) in enumerate(outs):
            assert np.array_equal(or_.cpu().numpy(), ref_r), f"rank {i} epoch {epoch}: rows differ"
            assert np.array_equal(os_.cpu().numpy(), ref_s), f"rank {i} epoch {epoch}: scores differ"


@pytest.mark.parametrize("world,b,dim", [(2, 37, 512), (8, 1024, 512), (8, 5, 64), (4, 130, 768), (3, 64, 32)])
def test_simulated_ranks_rows_allgather(built_lib, world, b, dim):
    """vq_peer_allgather_rows: every simulated rank contributes its slice of the query batch and ends up
    with the whole batch, bit for bit, over several epochs (both parities) and a smaller batch in between."""
    from video_quierer_b200.peer import LocalWindows, slice_range
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu").manual_seed(world * 100 + b)
    lw = LocalWindows(world, dev, b_max=max(b, 64), k_max=0, rows_ld=max(dim, 512))
    for epoch in range(5):
        bb = b if epoch != 2 else max(1, b // 3)
        q = torch.randn((bb, dim), generator=g)
        slices = []
        for r in range(world):
            lo, hi = slice_range(bb, world, r)
            slices.append(q[lo:hi].contiguous().to(dev))
        outs = lw.allgather_rows_all(slices, bb)
        torch.cuda.synchronize()
        assert int(lw.status.sum()) == 0, "a simulated rank timed out"
        for r, o in enumerate(outs):
            assert torch.equal(o.cpu(), q), f"rank {r} epoch {epoch}"


def test_missing_peer_times_out_instead_of_hanging(built_lib):
    """Only rank 0 of 2 launches: its wait must give up (VQ_PEER_TIMEOUT_MS) and flag the error."""
    import subprocess
    import sys
    code = (
        "import torch, numpy as np\n"
        "from video_quierer_b200.peer import LocalWindows, _launch\n"
        "dev = torch.device('cuda', 0)\n"
        "lw = LocalWindows(2, dev, 64, 16)\n"
        "s = torch.zeros((8, 10), device=dev); r = torch.zeros((8, 10), dtype=torch.int32, device=dev)\n"
        "_launch(lw.lib, lw.windows, 2, 0, 64, 16, s, r, None, 10, lw.status[0:1], torch.cuda.current_stream())\n"
        "torch.cuda.synchronize()\n"
        "assert int(lw.status[0]) == 1\n"
        "hdr = lw.bufs[0][:16].view(torch.int32).cpu().numpy()\n"
        "assert hdr[2] == 1, hdr\n"
        "print('TIMED_OUT_OK')\n")
    env = dict(os.environ, VQ_PEER_TIMEOUT_MS="200")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert "TIMED_OUT_OK" in out.stdout, out.stdout + out.stderr


# ----------------------------------------------------------------------------- 2 real GPUs over CUDA IPC
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ipc_worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["VQ_PEER_TIMEOUT_MS"] = "1000"
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from video_quierer_b200 import _lib, engine
        from video_quierer_b200.graphs import GraphedSearch
        from video_quierer_b200.sharded import ShardedSearcher, shard_range
        n, dim, k, b = 40_000, 512, 10, 96
        g = torch.Generator(device="cpu").manual_seed(3)
        full = torch.nn.functional.normalize(torch.randn((n, dim), generator=g), dim=1)
        queries = torch.randn((b, dim), generator=g).to(dev)
        lo, hi = shard_range(n, world, rank)
        store = engine.DeviceStore(dim, dev, keep_fp32=True)
        store.append(full[lo:hi].to(dev))
        scanner = engine.Scanner(dev)

        def local(q, kk):
            return scanner.scan(store.f32, store.n, dim, q, kk, _lib.NORM_EPS, "auto")

        peer = ShardedSearcher(local, n, device=dev, exchange="peer")
        coll = ShardedSearcher(local, n, device=dev, exchange="collective")
        for it in range(4):
            ps, pr = peer.search(queries, k)
            cs, cr = coll.search(queries, k)
            assert torch.equal(pr, cr) and torch.equal(ps, cs), f"rank {rank} iteration {it}: peer != collective"
        peer.check()
        # a smaller batch (fewer CTAs) and a larger k (window rebuilt collectively)
        ps, pr = peer.search(queries[:5], 40)
        cs, cr = coll.search(queries[:5], 40)
        assert torch.equal(pr, cr) and torch.equal(ps, cs)
        # replayed from a CUDA graph
        gs = GraphedSearch(lambda qq: peer.search(qq, k), b, dim, dev)
        cs, cr = coll.search(queries, k)
        for it in range(6):
            gs_s, gs_r = gs(queries)
            torch.cuda.synchronize()
            assert torch.equal(gs_r, cr) and torch.equal(gs_s, cs), f"rank {rank} replay {it}"
        peer.check()
        # ingest side: every rank contributes its slice of the batch, all end up with the whole batch
        from video_quierer_b200.peer import PeerRowGather, slice_range
        rg = PeerRowGather(dev, None, b_max=128, ld_max=dim)
        for bb in (b, 7, b):
            lo_q, hi_q = slice_range(bb, world, rank)
            got = rg.allgather_rows(queries[lo_q:hi_q].contiguous(), bb)
            assert torch.equal(got, queries[:bb]), f"rank {rank}: gathered queries differ (b={bb})"
        gsl = GraphedSearch(lambda qq: peer.search(qq, k), b, dim, dev,
                            ingest=lambda: rg.allgather_rows(queries[slice_range(b, world, rank)[0]:slice_range(b, world, rank)[1]].contiguous(), b))
        for it in range(3):
            gsl.replay()
            torch.cuda.synchronize()
            assert torch.equal(gsl.host_out[1], cr.cpu()) and torch.equal(gsl.host_out[0], cs.cpu()), f"rank {rank} sliced replay {it}"
        rg.check()
        rg.close()
        # against the whole store on one GPU
        whole = (queries / (queries.norm(dim=1, keepdim=True) + 1e-10)) @ full.to(dev).T
        top = torch.topk(whole, k, dim=1).indices
        assert (top == cr).float().mean().item() > 0.999          # (near-ties may swap neighbours)
        # the exact single-pass search as the local search (what bench.py shards), checked read-out
        st2 = engine.DeviceStore(dim, dev, keep_fp32=True, keep_bf16=True)
        st2.append(full[lo:hi].to(dev))
        ex = ShardedSearcher(lambda q, kk: scanner.exact(st2, q, kk)[:2], n, device=dev, exchange="peer")
        es, er = ex.search_checked(queries, k)
        assert np.array_equal(er, cr.cpu().numpy()) and np.allclose(es, cs.cpu().numpy(), rtol=1e-6, atol=1e-7)
        ex.close()
        # a late peer: rank 1 skips one exchange -> rank 0 must RAISE instead of returning a partial top-k, refuse
        # further searches, and work again after reset() on both ranks (failure surfaced on the data path)
        if rank == 0:
            try:
                peer.search_checked(queries, k)
                raise AssertionError("a missing peer went unnoticed")
            except _lib.VQError as e:
                assert "did not deliver" in str(e)
            try:
                peer.search(queries, k)
                raise AssertionError("search after a failed exchange must be refused")
            except RuntimeError:
                pass
        else:
            import time
            time.sleep(3.0)
        peer.reset()
        ps, pr = peer.search_checked(queries, k)
        assert np.array_equal(pr, cr.cpu().numpy()) and np.array_equal(ps, cs.cpu().numpy())
        peer.close()
        ret[rank] = "ok"
    except Exception as e:  # noqa: BLE001
        import traceback
        ret[rank] = "".join(traceback.format_exception(e))
    finally:
        dist.destroy_process_group()


def _hnsw_shard_worker(rank, world, port, ret):
    """North-star: 'HNSW is partitioned as per-shard sub-graphs searched in parallel and merged the same way' —
    one sub-graph per GPU, the same fused exchange + merge kernel, no host round trip on the search path."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from video_quierer_b200.hnsw_index import B200HNSWIndex
        from video_quierer_b200.sharded import ShardedSearcher, shard_range
        from video_quierer_b200.utils import synth
        n, dim, k, b = 24_000, 128, 10, 200
        full = synth.clip_like(n, dim, seed=5)
        queries = torch.from_numpy(synth.clip_like(b, dim, seed=6, n_store=n)).to(dev)
        lo, hi = shard_range(n, world, rank)
        h = B200HNSWIndex(dimension=dim, M=16, ef_construction=200, ef_search=128, max_M=16, device=dev)
        h.add_batch(list(full[lo:hi]), list(range(lo, hi)))
        h.build()
        sh = ShardedSearcher(h.as_local_search(), n, device=dev, exchange="peer")
        s, r = sh.search_checked(queries, k)
        assert int(h.last_overflow.sum()) == 0
        qn = queries.cpu().numpy()
        qn = qn / np.linalg.norm(qn, axis=1, keepdims=True)
        truth = np.argsort(-(qn @ full.T), axis=1)[:, :k]
        recall = np.mean([len(set(r[i]) & set(truth[i])) / k for i in range(b)])
        # the unsharded index at the same M / ef on this rank's GPU as the bar (union of sub-graph results >= it)
        whole = B200HNSWIndex(dimension=dim, M=16, ef_construction=200, ef_search=128, max_M=16, device=dev)
        whole.add_batch(list(full), list(range(n)))
        _, wr = whole.search_arrays(qn, k)
        recall_whole = np.mean([len(set(wr[i]) & set(truth[i])) / k for i in range(b)])
        assert recall >= recall_whole - 0.01 and recall > 0.9, (recall, recall_whole)
        assert np.all(np.diff(s, axis=1) <= 1e-7)                          # merged best first
        gathered = [None] * world
        dist.all_gather_object(gathered, r.tolist())
        assert gathered[0] == gathered[rank]                               # identical on every rank
        sh.close()
        ret[rank] = "ok"
    except Exception as e:  # noqa: BLE001
        import traceback
        ret[rank] = "".join(traceback.format_exception(e))
    finally:
        dist.destroy_process_group()


def test_two_gpu_hnsw_subgraphs(built_lib):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs on one node")
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_hnsw_shard_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret.get(r) == "ok" for r in range(world)), dict(ret)


def test_two_gpu_ipc_exchange(built_lib):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs on one node")
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_ipc_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret.get(r) == "ok" for r in range(world)), dict(ret)
