#!/usr/bin/env python3
"""Contract benchmark: exact frame-embedding search (BASELINE.json configs 2 / 4 / 5).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2|4|5] [--data gauss|clip]
                    [--batch B] [--impl reference]

Workload (config.workload), synthetic unit-norm frame embeddings generated on the device from fixed seeds
(S-gauss: iid; S-clip: clustered like temporally adjacent frames — SURVEY.md 8(d)):
  --config 2 (default)  1,000,000 x 512, k = 10, query batch 1024 (the batch QPS peaks at; batches 1 and 32
                        are swept in the same line, and the other data distribution at the headline batch)
  --config 4            10,000,000 x 768, k = 100, query batch 1024 (N in {2, 4, 8}; fits one GPU as well)
  --config 5            100,000,000 x 512, k = 10, 4096-query batches (needs >= 4 GPUs: 38 GB per GPU at N = 8)
A *step* is one pass of the hot path over one batch: L2-normalise the queries -> exact inner-product scan
(tensor-core scan of the bf16 copy that gathers every row within the computed rounding bound of the running
k-th best) -> fp32 re-score -> top-k -> (N > 1) exchange + merge over NVLink peer memory.  With --gpus N the
SAME store is row-sharded over N ranks (strong scaling, one process per GPU); rank 0 prints ONE JSON line.

  value      QPS with the query batch already resident in HBM (CUDA events, max over ranks)
  e2e        the same step with HOST (pinned) query buffers: H2D + search + D2H captured per step
  e2e_api    (N = 1) the same batch through the drop-in facade B200FlatIndex.search_batch: host numpy in,
             list of per-hit metadata dicts out (what /api/search/batch hands to FastAPI)
  sustained  the device-resident step again for >= 2 s right after the contract region (power-capped clocks)
  roofline   the scan kernel's algorithmic bytes / flops over its own CUDA-event time vs MEASURED_PEAKS.json
  parity     64 queries of the batch answered by the timed path vs the fp32 FMA scan of every shard merged on
             the host (outside the timed regions); a mismatch fails the run
  cpu_baseline  the reference's CPU search on this box's host cores (bounded sample)

`--impl reference` times the UNMODIFIED reference (baseline/_ref/video_search_overhaul.py, copied from
/root/reference by __graft_entry__.build(); SimpleVideoIndex.search in a Python loop exactly like
src/api/routes.py:627-634) on the host cores; if that copy is absent, the oracle port (numpy restatement).
"""

from __future__ import annotations

import argparse
import contextlib
import json
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    2: {"rows": 1_000_000, "dim": 512, "k": 10, "batch": 1024, "label": "BASELINE config 2"},
    4: {"rows": 10_000_000, "dim": 768, "k": 100, "batch": 1024, "label": "BASELINE config 4"},
    5: {"rows": 100_000_000, "dim": 512, "k": 10, "batch": 4096, "label": "BASELINE config 5"},
}
N_ROWS, DIM, K_TOP = 1_000_000, 512, 10
LABEL = "BASELINE config 2"
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def metric_name():
    return f"search QPS @k={K_TOP} ({N_ROWS // 1_000_000}M x {DIM} frames, exact scan)"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops", 0)), "measured", float(d.get("bf16_tflops_sustained", 0) or d.get("bf16_tflops", 0))
    return 6650.0, 1590.0, "fallback", 1400.0


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region: ONE long-running
    `nvidia-smi -lms 200` per rank (the profiling recipe's clocks line), read when the run ends."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:  # noqa: BLE001
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=6)
        except Exception:  # noqa: BLE001
            self.proc.kill()
            out = ""
        self.rows = [[c.strip() for c in ln.split(",")] for ln in out.strip().splitlines() if ln.strip()]
        if not self.rows:                      # nothing came through the pipe: one direct query, better than none
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows = [[c.strip() for c in out.strip().split(",")]]
            except Exception:  # noqa: BLE001
                pass

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ----------------------------------------------------------------------------- synthetic data
def host_rows(kind: str, n: int, dim: int, seed: int, n_store: int | None = None):
    from video_quierer_b200.utils import synth
    if kind == "clip":
        return synth.clip_like(n, dim, seed=seed, n_store=n_store)
    return synth.gauss(n, dim, seed=seed)


def device_block(kind, b0, blk, dim, dev, centres, seed_base=1000):
    """Rows [b0, b0 + blk) of the synthetic store: the same whatever the sharding."""
    import torch
    g = torch.Generator(device=dev).manual_seed(seed_base + b0 // blk)
    x = torch.randn((blk, dim), device=dev, generator=g)
    if kind == "clip":
        ids = torch.randint(0, centres.shape[0], (blk,), device=dev, generator=g)
        x = centres[ids] + (0.35 / dim ** 0.5) * x
    return x


def clip_centres(n_rows, dim, dev):
    import torch
    g = torch.Generator(device=dev).manual_seed(77)
    c = torch.randn((max(256, n_rows // 4096), dim), device=dev, generator=g)
    return c / c.norm(dim=1, keepdim=True)


def device_queries(kind, b, dim, dev, centres, seed=7):
    """[b, dim] fp32 queries on the HOST (pinned): S-gauss iid, S-clip drawn around the store's centres."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed + b)
    x = torch.randn((b, dim), generator=g)
    if kind == "clip":
        c = centres.cpu()
        ids = torch.randint(0, c.shape[0], (b,), generator=g)
        x = c[ids] + (0.35 / dim ** 0.5) * x
    return x.pin_memory()


# ----------------------------------------------------------------------------- reference arm / cpu baseline
def _load_reference():
    """The unmodified reference module from baseline/_ref (None if the copy did not travel)."""
    path = os.path.join(REF_DIR, "video_search_overhaul.py")
    if not os.path.exists(path):
        return None
    import importlib.util
    sys.dont_write_bytecode = True
    spec = importlib.util.spec_from_file_location("ref_video_search_overhaul", path)
    mod = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(mod)
    except Exception as e:  # noqa: BLE001 — e.g. cv2 missing: fall back to the port, say so
        print(f"[bench] reference import failed ({type(e).__name__}: {e}); timing the oracle port", file=sys.stderr)
        return None
    return mod


def _host_threads():
    import numpy as np  # noqa: F401
    info = {"cores": len(os.sched_getaffinity(0)), "OMP_NUM_THREADS": os.environ.get("OMP_NUM_THREADS"),
            "OPENBLAS_NUM_THREADS": os.environ.get("OPENBLAS_NUM_THREADS")}
    try:
        from threadpoolctl import threadpool_info
        info["blas_threads"] = max((p.get("num_threads", 0) for p in threadpool_info()), default=None)
    except Exception:  # noqa: BLE001
        pass
    return info


class CpuReference:
    """The reference's exact search on host cores over a (sample of the) workload's store."""

    def __init__(self, data: str, max_rows: int = 1_000_000):
        import numpy as np
        self.np = np
        self.sample_rows = min(N_ROWS, max_rows)
        self.store = host_rows(data, self.sample_rows, DIM, seed=0)
        self.queries = host_rows(data, 64, DIM, seed=1, n_store=self.sample_rows) * np.float32(3.0)
        self.ref = _load_reference()
        self.index = None

    def build_index(self):
        """SimpleVideoIndex filled through add_frame (video_search_overhaul.py:31-38): a Python list of rows."""
        if self.ref is not None and self.index is None:
            idx = self.ref.SimpleVideoIndex()
            for i in range(self.sample_rows):
                idx.add_frame(self.store[i], "v.mp4", float(i))
            self.index = idx
        return self.index

    def as_shipped(self, n_queries: int, k: int):
        """Loop of SimpleVideoIndex.search like routes.py:627-634 (re-stacks the matrix per query, :46)."""
        idx = self.build_index()
        t0 = time.perf_counter()
        for i in range(n_queries):
            idx.search(self.queries[i % len(self.queries)], k)
        return time.perf_counter() - t0

    def charitable(self, n_queries: int, k: int):
        """Same arithmetic with the matrix stacked once (oracle port of :49-56)."""
        from oracle import exact
        t0 = time.perf_counter()
        for i in range(n_queries):
            exact.exact_search(self.store, self.queries[i % len(self.queries)], k)
        return time.perf_counter() - t0

    def scale(self):
        """QPS over the sample -> QPS over the whole store (both searches are linear in the rows)."""
        return self.sample_rows / N_ROWS


def cpu_baseline_leg(data: str, budget_s: float = 14.0):
    cr = CpuReference(data)
    threads = _host_threads()
    cr.charitable(1, K_TOP)                                  # warm BLAS threads
    t_c = cr.charitable(4, K_TOP) / 4
    n_c = max(4, min(64, int(budget_s * 0.3 / max(t_c, 1e-6))))
    el_c = cr.charitable(n_c, K_TOP)
    out = {"unit": "queries/s", "cores": threads["cores"], "threads": threads,
           "charitable": {"value": n_c / el_c * cr.scale(), "sample": f"{n_c} single-query searches, np.dot + argsort with the "
                          f"{cr.sample_rows}x{DIM} fp32 matrix stacked ONCE (the reference re-stacks per query)"}}
    if cr.ref is not None:
        t0 = time.perf_counter()
        cr.build_index()
        t_build = time.perf_counter() - t0
        n_s = 3
        el_s = cr.as_shipped(n_s, K_TOP)
        out.update({"value": n_s / el_s * cr.scale(), "kind": "reference",
                    "sample": f"{n_s} queries of the batch through the UNMODIFIED SimpleVideoIndex.search "
                              f"(baseline/_ref/video_search_overhaul.py:40-64, per-query np.vstack of {cr.sample_rows} rows + np.dot + "
                              f"argsort, looped like src/api/routes.py:627-634) in {el_s:.1f}s; index build {t_build:.1f}s not counted"
                              + (f"; QPS scaled by {cr.scale():.3f} to the {N_ROWS}-row store" if cr.scale() != 1 else "")})
    else:
        out.update({"value": out["charitable"]["value"], "kind": "port", "sample": out["charitable"]["sample"]})
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: the reference arm must get all host
    # threads, so re-exec once without it (BLAS reads the variable when numpy loads)
    if os.environ.get("OMP_NUM_THREADS") and not os.environ.get("VQ_REF_REEXEC"):
        env = {k: v for k, v in os.environ.items() if k not in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS")}
        env["VQ_REF_REEXEC"] = "1"
        os.dup2(args._stdout_fd, 1)
        os.execve(sys.executable, [sys.executable] + sys.argv, env)
    cr = CpuReference(args.data)
    threads = _host_threads()
    shipped = cr.ref is not None
    run = cr.as_shipped if shipped else cr.charitable
    if shipped:
        cr.build_index()
    else:
        cr.charitable(1, K_TOP)
    t0 = time.perf_counter()
    run(1, K_TOP)
    t_query = time.perf_counter() - t0
    # bounded sample: as many queries per step (<= 4) as keep the whole run near two minutes
    per_step = max(1, min(4, int(120.0 / max((args.steps + args.warmup) * t_query, 1e-9))))
    for _ in range(args.warmup):
        run(per_step, K_TOP)
    el = 0.0
    for _ in range(args.steps):
        el += run(per_step, K_TOP)
    qps = args.steps * per_step / el * cr.scale()
    t_ch = cr.charitable(4, K_TOP) / 4
    what = (f"{per_step} queries/step of the batch-{args.batch} workload through the UNMODIFIED reference "
            f"(baseline/_ref/video_search_overhaul.py:40-64 SimpleVideoIndex.search: per-query np.vstack of {cr.sample_rows} rows "
            "+ np.dot + argsort, looped like src/api/routes.py:627-634)") if shipped else \
           (f"{per_step} queries/step of the batch-{args.batch} workload, sequential np.dot + argsort per query like "
            "src/api/routes.py:627-634, matrix pre-stacked (oracle port; baseline/_ref absent)")
    if cr.scale() != 1:
        what += f"; measured on the first {cr.sample_rows} rows, QPS scaled by {cr.scale():.3f} (the search is linear in the rows)"
    line = {"impl": "reference", "metric": metric_name(), "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": _config(args.batch, args.gpus, "cpu-reference" if shipped else "cpu-oracle-port", args.data),
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads["cores"], "threads": threads,
                             "kind": "reference" if shipped else "port", "sample": what,
                             "charitable": {"value": 1.0 / t_ch * cr.scale(), "sample": "same arithmetic, matrix stacked once"}},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def _config(batch, gpus, path, data, exact=True):
    elem = 2 if exact else 4
    return {"workload": f"exact cosine top-{K_TOP}: {N_ROWS}x{DIM} frame store ({'S-clip clustered' if data == 'clip' else 'S-gauss iid'}), "
                        f"query batch {batch} ({LABEL}), row-sharded over {gpus} GPU(s)",
            "n_rows": N_ROWS, "dim": DIM, "k": K_TOP, "batch": batch, "data_kind": data,
            "store_dtype": "fp32 master + bf16 scan copy (one tensor-core pass gathers every row within the computed rounding "
                           "bound of the k-th best, exact fp32 re-score)" if exact else "fp32",
            "scan_path": path,
            "l2": "scanned shard > 1.5 x the 126 MB L2: it evicts itself between steps" if N_ROWS // gpus * DIM * elem > 190e6
                  else "scanned shard could sit in L2: consecutive steps scan different identical copies of it (>= 500 MB in rotation)",
            "parallelism": f"rows/{gpus}"}


# ----------------------------------------------------------------------------- parity (outside the timed regions)
def compare_topk(rows, scores, rows_ref, scores_ref, tol=1e-5):
    """North-star rule: ids identical except where the scores involved tie within `tol` (relative), scores
    within `tol` relative.  Returns the number of queries that violate it."""
    import numpy as np
    bad = 0
    for r, s, rr, sr in zip(rows, scores, rows_ref, scores_ref):
        ok = np.allclose(s, sr, rtol=tol, atol=tol * 1e-2)
        if ok and not np.array_equal(r, rr):
            for i in np.nonzero(r != rr)[0]:
                # a differing id is fine only if it sits in a run of scores that tie with the reference's score there
                tie = np.abs(sr - sr[i]) <= tol * max(abs(float(sr[i])), 1e-30)
                if r[i] not in rr[tie]:
                    ok = False
                    break
        bad += 0 if ok else 1
    return bad


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from video_quierer_b200 import _lib, engine
    from video_quierer_b200.flat_index import exact_fallback, exact_search
    from video_quierer_b200.sharded import ShardedSearcher, shard_offsets, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    exact = args.mode == "exact"

    # ---- synthetic store shard, generated on the device (fixed seed per 64k-row block)
    lo, hi = shard_range(N_ROWS, world, rank)
    elem = 2 if exact else 4
    need_gb = (hi - lo) * engine.padded_ld(DIM) * (6 if exact else 4) / 2 ** 30
    free_gb = torch.cuda.mem_get_info(dev)[0] / 2 ** 30
    if need_gb > free_gb - 6:
        raise SystemExit(f"[bench] rank {rank}: the {hi - lo}-row shard needs {need_gb:.0f} GB of HBM, {free_gb:.0f} GB free: use more GPUs")
    # L2 rule: the rows a step scans must not be L2-resident from the step before.  A shard larger than
    # 1.5 x L2 evicts itself; a smaller one (8-way sharding of config 2: 128 MB) is kept in several identical
    # copies and consecutive steps scan different copies, so that the working set between two uses of a copy
    # exceeds L2 (126 MB) several times — no flush kernel inside the loop, one timing method for every N.
    shard_bytes = (hi - lo) * DIM * elem
    n_copies = 1 if shard_bytes > 190e6 else int(500e6 // max(shard_bytes, 1)) + 1
    centres = clip_centres(N_ROWS, DIM, dev)

    def fill(st_c, kind):
        st_c.truncate(0)
        blk = 1 << 16
        for b0 in range(lo // blk * blk, hi, blk):           # block b0 is the same whatever the sharding
            x = device_block(kind, b0, blk, DIM, dev, centres)
            s, e = max(lo, b0), min(hi, b0 + blk)
            st_c.append(x[s - b0: e - b0], _lib.NORM_PLAIN)  # kernel (a): L2-normalise on ingest

    stores = []
    for _c in range(n_copies):
        stores.append(engine.DeviceStore(DIM, dev, keep_fp32=True, keep_bf16=exact, capacity=hi - lo))
        fill(stores[-1], args.data)
    store = stores[0]

    # ---- search lanes: `depth` steps in flight, each lane with its own stream, workspace and (N > 1)
    # exchange windows; the store copies are shared.  Depth 1 = strictly one step after the other.
    class Lane:
        def __init__(self, idx):
            self.idx = idx
            self.stream = torch.cuda.Stream(dev) if args.pipeline > 1 else None
            self.scanner = engine.Scanner(dev)
            self.bad = None            # overflow flags of the lane's most recent local search ([b] int32, device)
            self.copy = 0              # which copy of the shard the next search scans
            self.searcher = ShardedSearcher(self.local_search, N_ROWS, device=dev, exchange=args.exchange)

        def local_search(self, q, k):
            st_i = stores[self.copy]
            if exact:
                s, r, self.bad = exact_search(self.scanner, st_i, q, k)
                return s, r
            return self.scanner.scan(st_i.f32, st_i.n, DIM, q, k, _lib.NORM_EPS, args.path)

        def search(self, q, k):
            # one shard: the local search IS the answer (int32 rows); several: exchange + merge (int64 rows)
            return self.local_search(q, k) if world == 1 else self.searcher.search(q, k)

    lanes = [Lane(0)]
    scanner = lanes[0].scanner

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def on(stream):
        return torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()

    def parity_check(dev_q, n_check=64):
        """The timed path's answer for the first queries of the batch against an independent one: fp32 FMA scan
        (`vq_scan_topk`, the reference's arithmetic) of EVERY rank's shard, the per-shard lists merged on the host
        by (score desc, global row asc).  All ranks take part; every rank compares its own copy of the answer."""
        m = min(n_check, dev_q.shape[0])
        q = dev_q[:m].contiguous()
        lanes[0].copy = 0
        s, r = lanes[0].search(q, K_TOP)
        over = int(lanes[0].bad.sum().item()) if lanes[0].bad is not None else 0
        fs, fr = scanner.scan(store.f32, store.n, DIM, q, K_TOP, _lib.NORM_EPS, "fma")
        torch.cuda.synchronize()
        if world > 1:
            gs = [torch.empty_like(fs) for _ in range(world)]
            gr = [torch.empty_like(fr) for _ in range(world)]
            dist.all_gather(gs, fs)
            dist.all_gather(gr, fr)
            offs = shard_offsets(N_ROWS, world)
            cs = np.concatenate([t.cpu().numpy() for t in gs], axis=1)
            cr_ = np.concatenate([np.where(t.cpu().numpy() >= 0, t.cpu().numpy().astype(np.int64) + o, -1) for t, o in zip(gr, offs)], axis=1)
            cs = np.where(cr_ >= 0, cs, -np.inf)
            order = np.lexsort((cr_, -cs), axis=1)[:, :K_TOP]
            ref_s, ref_r = np.take_along_axis(cs, order, axis=1), np.take_along_axis(cr_, order, axis=1)
        else:
            ref_s, ref_r = fs.cpu().numpy(), fr.cpu().numpy().astype(np.int64)
        bad = compare_topk(r.cpu().numpy().astype(np.int64), s.cpu().numpy(), ref_r, ref_s)
        t = torch.tensor([bad, over], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return {"checked": m, "mismatches": int(t[0]), "overflowed_queries": int(t[1]),
                "against": "fp32 FMA scan (vq_scan_topk) of every shard, per-shard lists merged on the host by (score desc, global row asc)",
                "rule": "ids identical except ties within 1e-5, scores within 1e-5 relative; checked on every rank"}

    def measure(B, steps, warmup, with_e2e, data=None, sustain_s=0.0, with_parity=False):
        host_q = device_queries(data or args.data, B, DIM, dev, centres)
        dev_q = host_q.to(dev)
        for _ in range(max(warmup, 3)):
            lanes[0].search(dev_q, K_TOP)
        launches = scanner.last_launches + (1 if world > 1 else 0)   # + exchange/merge of the shard layer
        path = scanner.last_path
        barrier()
        parity = parity_check(dev_q) if with_parity else None
        barrier()
        # pipelining needs lane-private exchange windows; the NCCL route serialises on one communicator
        depth = args.pipeline if (world == 1 or lanes[0].searcher._peer is not None) else 1
        while len(lanes) < depth:
            lanes.append(Lane(len(lanes)))
            lanes[-1].search(dev_q, K_TOP)
        barrier()
        # slot j = (lane j % depth, store copy j % n_copies); consecutive steps take consecutive slots
        n_slots = math.lcm(n_copies, depth)            # every lane meets every copy: the rotation never shortens
        slots = [(lanes[j % depth], j % n_copies) for j in range(n_slots)]
        # the serving path for a fixed batch shape: the whole step captured once as a CUDA graph
        graphs = None
        if not args.no_graph:
            try:
                from video_quierer_b200.graphs import GraphedSearch
                graphs = []
                for ln, c in slots:
                    ln.copy = c
                    gs = GraphedSearch(lambda qq, ln=ln: ln.search(qq, K_TOP) + (ln.bad,), B, DIM, dev, stream=ln.stream)
                    gs.q.copy_(dev_q)
                    graphs.append(gs)
            except Exception as e:  # noqa: BLE001 — e.g. a collective that refuses capture: run eagerly
                print(f"[bench] CUDA graph capture unavailable ({type(e).__name__}: {e}); eager launches", file=sys.stderr)
                graphs = None
            barrier()
        # e2e: the same step with its host<->device copies captured too (one replay = H2D of the pinned
        # request slot + search + D2H of ids, scores and overflow flags into pinned result slots)
        hgraphs = None
        ingest_mode = ["full"]
        if graphs is not None and with_e2e and not args.no_host_graph:
            try:
                hgraphs = []
                # N > 1 on the peer route: every rank ingests only ITS slice of the batch over PCIe and the
                # slices are all-gathered over NVLink (vq_peer_allgather_rows) — 8 ranks pulling the whole batch
                # from host memory at once capped the step near 0.25 ms
                sliced = world > 1 and lanes[0].searcher._peer is not None and args.ingest == "slice"
                if sliced:
                    from video_quierer_b200.peer import PeerRowGather, slice_range
                    q_lo, q_hi = slice_range(B, world, rank)
                    for ln in lanes[:depth]:
                        if getattr(ln, "rowgather", None) is None or ln.rowgather.b_max < B:
                            ln.rowgather = PeerRowGather(dev, None, b_max=B, ld_max=DIM)
                for ln, c in slots:
                    ln.copy = c
                    ingest = None
                    if sliced:
                        h_slice = host_q[q_lo:q_hi].clone().pin_memory()     # the request slot of this rank's slice
                        d_slice = torch.empty((q_hi - q_lo, DIM), dtype=torch.float32, device=dev)

                        def ingest(ln=ln, h_slice=h_slice, d_slice=d_slice):
                            d_slice.copy_(h_slice, non_blocking=True)
                            return ln.rowgather.allgather_rows(d_slice, B)
                    hg = GraphedSearch(lambda qq, ln=ln: ln.search(qq, K_TOP) + (ln.bad,), B, DIM, dev, stream=ln.stream,
                                       host_io=True, ingest=ingest)
                    if hg.host_q is not None:
                        hg.host_q.copy_(host_q)        # the request slot of this step's batch
                    hgraphs.append(hg)
                ingest_mode[0] = "slice" if sliced else "full"
            except Exception as e:  # noqa: BLE001
                print(f"[bench] host-io graph capture unavailable ({type(e).__name__}: {e}); explicit copies", file=sys.stderr)
                hgraphs = None
            barrier()
        step_no = [0]
        cur_stream = torch.cuda.current_stream(dev)

        def fork():
            for ln in lanes[:depth]:
                if ln.stream is not None:
                    ln.stream.wait_stream(cur_stream)

        def join():
            for ln in lanes[:depth]:
                if ln.stream is not None:
                    cur_stream.wait_stream(ln.stream)

        def step_device():
            j = step_no[0] % n_slots
            step_no[0] += 1
            if graphs is not None:
                return graphs[j].replay()
            ln, c = slots[j]
            ln.copy = c
            with on(ln.stream):
                return ln.search(dev_q, K_TOP)

        # -- dominant kernel alone: the library's own events around the scan kernel, sampled right before
        # AND right after the timed region (clocks under a long tensor-bound run sag with the power cap)
        kern = []

        def sample_kernel(times):
            lib.vq_profile_enable(1)
            for it_k in range(times):
                lanes[0].copy = it_k % n_copies
                lanes[0].local_search(dev_q, K_TOP)
                kern.append(lib.vq_profile_last_kernel_ms())
            lib.vq_profile_enable(0)

        sample_kernel(min(steps, 10))
        barrier()
        fork()
        for _ in range(max(warmup, n_slots)):           # W untimed steps through the very path that is timed
            step_device()                               # (every captured graph is replayed at least once)
        join()
        barrier()
        # -- timed region 1: device-resident queries, CUDA events on the launch stream (the lanes fork from
        # it after the first event and join it before the second)
        # steps are queued back to back; with a collective inside the step the queue is drained every
        # `sync_every` steps (measured on 8 GPUs: an unbounded queue of graph replays that contain an NCCL
        # all-gather runs 3x slower per step than the same replays with a shallow queue)
        nccl_inside = world > 1 and lanes[0].searcher._peer is None
        sync_every = int(os.environ.get("VQ_BENCH_SYNC_EVERY", "0")) or (4 if nccl_inside else 1 << 30)
        # Peer route, N > 1: the host keeps at most `cap` steps (one per lane) enqueued ahead of the GPU.
        # With the whole run enqueued at once the lanes of different ranks drift apart — rank A ahead on
        # lane 0, rank B ahead on lane 1 — and every exchange then waits for a different straggler (seen on
        # 8 GPUs: 3.8 M QPS free-running in a run whose host-gated e2e loop reached 4.4 M; same session,
        # final kernel: 3.85 / 4.39 / 4.47 M QPS at 6 / 4 / 3 steps ahead, 2 GPUs: 2.07 / 2.19 M at 6 / 3 —
        # tools/inflight8.sh, tools/inflight2.sh).  A serving process bounds its queue the same way.
        cap = int(os.environ.get("VQ_BENCH_INFLIGHT", "0")) or (depth if (world > 1 and not nccl_inside) else 0)

        def timed_steps(n_steps):
            done_evs = []
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            fork()
            for i in range(n_steps):
                if cap and i >= cap:
                    done_evs[i - cap].synchronize()
                ln_i = slots[step_no[0] % n_slots][0]
                step_device()
                if cap:
                    ev = torch.cuda.Event()
                    ev.record(ln_i.stream if ln_i.stream is not None else cur_stream)
                    done_evs.append(ev)
                if (i + 1) % sync_every == 0 and i + 1 < n_steps:
                    torch.cuda.synchronize()
            join()
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        dev_ms = timed_steps(steps)
        n_cool = len(kern)
        sample_kernel(min(steps, 10))
        med = lambda v: (sorted(v)[len(v) // 2] if v else None)  # noqa: E731
        kmed = med([v for v in kern[:n_cool] if v > 0])          # before the timed region: the kernel timed alone
        khot = med([v for v in kern[n_cool:] if v > 0])          # right after it: the board at its power cap
        torch.cuda.synchronize()
        sustained = None
        if sustain_s > 0:
            n_sus = max(steps, int(math.ceil(sustain_s * 1e3 / (dev_ms / steps))))
            sus_ms = timed_steps(n_sus)
            sustained = {"value": B * n_sus / (sus_ms * 1e-3), "unit": "queries/s", "steps": n_sus, "seconds": sus_ms * 1e-3,
                         "ms_per_step": sus_ms / n_sus}

        e2e_ms = None
        e2e_host = None
        n_fallback = [0]
        if with_e2e:
            # -- timed region 2: end to end with HOST (pinned) buffers: H2D + search + D2H (+ fallback), every
            # step.  `depth` steps are in flight: a lane's result is harvested (event wait, overflow check,
            # exact fallback) right before the lane is reused.
            class Out:
                def __init__(self):
                    self.rows = torch.empty((B, K_TOP), dtype=torch.int64 if world > 1 else torch.int32).pin_memory()
                    self.scores = torch.empty((B, K_TOP), dtype=torch.float32).pin_memory()
                    self.bad = torch.zeros((B,), dtype=torch.int32).pin_memory()
                    self.ev = torch.cuda.Event()
                    self.pending = None          # (lane, copy, device queries, device scores) of the step in flight
            outs = [Out() for _ in range(depth)]
            host_t = [0.0]

            def harvest(o):
                if o.pending is None:
                    return
                ln, c, q, s = o.pending
                o.pending = None
                o.ev.synchronize()
                if ln.bad is not None and world == 1 and bool(o.bad.any()):
                    # queries whose gather overflowed (mass ties): fp32 FMA scan, inside the timing
                    idx = torch.nonzero(o.bad).flatten()
                    with on(ln.stream):
                        sf, rf = exact_fallback(ln.scanner, stores[c], q, K_TOP, idx.to(dev))
                        o.scores[idx] = sf.cpu()
                        o.rows[idx] = rf.cpu().to(o.rows.dtype)
                    n_fallback[0] += len(idx)

            def step_e2e():
                j = step_no[0] % n_slots
                step_no[0] += 1
                ln, c = slots[j]
                o = outs[ln.idx]
                t_a = time.perf_counter()
                harvest(o)                                     # the lane's previous step
                host_t[0] += time.perf_counter() - t_a         # host blocked on the GPU (plus the overflow check)
                ln.copy = c
                with on(ln.stream):
                    if hgraphs is not None:
                        hg = hgraphs[j]
                        hg.replay()                            # H2D + search + D2H, all inside the captured step
                        o.scores, o.rows = hg.host_out[0], hg.host_out[1]
                        if hg.host_out[2] is not None:
                            o.bad = hg.host_out[2]
                        o.ev.record()
                        o.pending = (ln, c, hg.q, hg.out[0])
                        return
                    if graphs is not None:
                        res = graphs[j](host_q)                # H2D of this step's inputs + one graph launch
                        q = graphs[j].q
                    else:
                        q = host_q.to(dev, non_blocking=True)  # H2D of this step's inputs
                        res = ln.search(q, K_TOP) + (ln.bad,)
                    s, r, bad = res
                    o.scores.copy_(s, non_blocking=True)       # D2H of the step's result
                    o.rows.copy_(r, non_blocking=True)
                    if bad is not None:
                        o.bad.copy_(bad, non_blocking=True)    # + its per-query overflow flags
                    o.ev.record()
                o.pending = (ln, c, q, s)

            def drain():
                for o in outs:
                    harvest(o)
                torch.cuda.synchronize()

            for _ in range(3):
                step_e2e()
            drain()
            barrier()
            host_t[0] = 0.0
            t0 = time.perf_counter()
            fork()
            for _ in range(steps):
                step_e2e()
            t1 = time.perf_counter()
            drain()
            barrier()
            e2e_ms = (time.perf_counter() - t0) * 1e3
            # where the host thread of this rank spent the loop: blocked on results vs issuing work
            e2e_host = {"wait_ms_per_step": host_t[0] / steps * 1e3, "issue_ms_per_step": ((t1 - t0) - host_t[0]) / steps * 1e3}
            t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        unc = max((int(ln.bad.sum().item()) if ln.bad is not None else 0) for ln in lanes[:depth])
        for ln in lanes:
            ln.searcher.check()                         # a peer-exchange wait that timed out voids the run
            if getattr(ln, "rowgather", None) is not None:
                ln.rowgather.check()
        # where a step's time goes, one step at a time (eager, CUDA events, max over ranks): the scan kernel alone, the
        # whole local search (prep + bootstrap + scan + finish) and the same plus the shard exchange + merge
        breakdown = None
        if world > 1 and with_e2e:
            def timed_eager(fn, n_it=10):
                fn()
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for it_b in range(n_it):
                    lanes[0].copy = it_b % n_copies
                    fn()
                e1.record()
                barrier()
                t = torch.tensor([e0.elapsed_time(e1) / n_it], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                return float(t.item())
            t_local = timed_eager(lambda: lanes[0].local_search(dev_q, K_TOP))
            t_full = timed_eager(lambda: lanes[0].search(dev_q, K_TOP))
            step_ms = dev_ms / steps
            breakdown = {"scan_kernel_ms": kmed, "local_search_ms": t_local, "local_search_plus_exchange_ms": t_full,
                         "pipelined_step_ms": step_ms,
                         "bound": "scan kernel" if kmed and kmed >= 0.75 * step_ms else
                                  ("latency-bound kernels around the scan (prep, bootstrap, finish)" if t_full - t_local < 0.5 * (t_local - (kmed or 0)) else "shard exchange + merge"),
                         "note": "eager, one step in flight; the pipelined step overlaps everything but the scan kernels of consecutive steps"}
        # rows gathered / re-scored per query by the exact search (its out_stats), one eager call
        gathered = None
        if exact and K_TOP <= 64:
            stats = torch.zeros((B, 2), dtype=torch.int32, device=dev)
            scanner.exact(stores[0], dev_q, K_TOP, stats=stats)
            sh = stats.float().mean(dim=0).cpu()
            gathered = {"rows_gathered_per_query": float(sh[0]), "rows_rescored_per_query": float(sh[1])}
        return {"B": B, "steps": steps, "dev_ms": dev_ms, "e2e_ms": e2e_ms, "kernel_ms": kmed, "kernel_hot_ms": khot, "launches": launches, "path": path,
                "graph": graphs is not None, "host_graph": hgraphs is not None, "ingest": ingest_mode[0], "overflow": unc, "depth": depth, "cap": cap,
                "fallbacks": n_fallback[0], "e2e_host": e2e_host, "sustained": sustained, "parity": parity, "gathered": gathered, "breakdown": breakdown,
                "host_q": host_q, "data": data or args.data}

    def measure_api(host_q, steps):
        """The drop-in call a maintainer makes: B200FlatIndex.search_batch with host numpy queries
        -> list of per-hit metadata dicts.  The facade adopts the bench store (no second copy); wall clock."""
        from video_quierer_b200.flat_index import B200FlatIndex
        idx = B200FlatIndex(device=dev)
        idx.adopt_store(store, [{"video_name": f"v{i >> 10}.mp4", "timestamp": float(i & 1023), "frame_id": i} for i in range(store.n)])
        qn = host_q.numpy()
        for _ in range(2):
            idx.search_batch(qn, K_TOP)
        t0 = time.perf_counter()
        for _ in range(steps):
            hits = idx.search_batch(qn, K_TOP)
        el = time.perf_counter() - t0
        t1 = time.perf_counter()
        for _ in range(steps):
            idx.search_arrays(qn, K_TOP)
        el_arr = time.perf_counter() - t1
        assert len(hits) == len(qn) and len(hits[0]) == K_TOP and "score" in hits[0][0]
        return {"value": len(qn) * steps / el, "unit": "queries/s", "ms_per_step": el / steps * 1e3, "steps": steps,
                "arrays_only_ms_per_step": el_arr / steps * 1e3,
                "what": "B200FlatIndex.search_batch(host numpy [B, dim], k) -> list of B lists of k metadata dicts (+ 'score'), synchronous, "
                        "one step in flight; arrays_only = search_arrays (same call without building the dicts)"}

    def hnsw_sharded_leg(n_queries=8192, efs=(64, 128, 256)):
        """North-star: 'HNSW is partitioned as per-shard sub-graphs searched in parallel and merged the same way'.
        Every rank builds the sub-graph of ITS shard of the store (rows already on the device), the per-shard beams are
        exchanged + merged by the same peer-memory kernel as the exact path; recall@10 against the sharded exact search."""
        from video_quierer_b200.hnsw_index import B200HNSWIndex
        st0 = stores[0]
        h = B200HNSWIndex(dimension=DIM, M=16, ef_construction=200, ef_search=64, max_M=16, device=dev)
        t0 = time.perf_counter()
        ok = 1
        try:
            h.add_device_rows(st0.f32[: st0.n, :DIM], level_seed=rank)
            h.build()
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001 — every rank must learn about it before the first collective of the leg
            print(f"[bench] rank {rank}: HNSW shard build failed ({type(e).__name__}: {e})", file=sys.stderr)
            ok = 0
        t_ok = torch.tensor([ok], dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
        if int(t_ok.item()) == 0:
            return {"error": "HNSW shard build failed on a rank (see stderr)"}
        build_s = time.perf_counter() - t0
        q = device_queries(args.data, n_queries, DIM, dev, centres, seed=11).to(dev)
        lanes[0].copy = 0
        truth = torch.cat([lanes[0].search(q[s0:s0 + 1024].contiguous(), 10)[1] for s0 in range(0, n_queries, 1024)]).cpu().numpy()
        sh = ShardedSearcher(h.as_local_search(), N_ROWS, device=dev, exchange=args.exchange)
        runs = []
        for ef in efs:
            h.ef_search = ef
            sh.search(q, 10)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s, r = sh.search(q, 10)
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1), float(h.last_overflow.sum().item())], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            rows = r.cpu().numpy()
            rec = float(np.mean([len(set(rows[i]) & set(truth[i])) / 10 for i in range(n_queries)]))
            runs.append({"ef": ef, "recall@10": round(rec, 4), "qps": round(n_queries / (float(t[0]) * 1e-3)), "ms": round(float(t[0]), 3),
                         "visited_overflow_queries": int(t[1])})
        sh.check()
        sh.close()
        return {"workload": f"HNSW per-shard sub-graphs, M=16 ef_construction=200 max_M=16, {N_ROWS}x{DIM} ({args.data}) over {world} GPU(s), "
                            f"{n_queries} queries, k=10; recall vs the sharded exact search", "build_s_per_shard": round(build_s, 2), "runs": runs}

    with ClockSampler(local) as clocks:
        main = measure(args.batch, args.steps, args.warmup, True, sustain_s=args.sustain, with_parity=True)
        sweep = []
        if not args.no_sweep:
            for B in (1, 32, 1024):
                if B != args.batch:
                    sweep.append(measure(B, min(args.steps, 30), 3, False))
        api = measure_api(main["host_q"], max(3, min(args.steps, 20))) if (world == 1 and exact and not args.no_api) else None
        hnsw_sharded = None
        if world > 1 and exact and args.config == 2 and not args.no_hnsw:
            try:
                hnsw_sharded = hnsw_sharded_leg()
            except Exception as e:  # noqa: BLE001 — the headline line must survive a failure of the extra leg
                hnsw_sharded = {"error": f"{type(e).__name__}: {e}"[:300]}
        if not args.no_sweep and args.config == 2:
            # the other distribution at the headline batch: same store shape, rows regenerated in place
            other = "gauss" if args.data == "clip" else "clip"
            for st_c in stores:
                fill(st_c, other)
            sweep.append(measure(args.batch, min(args.steps, 30), 3, False, data=other, with_parity=True))

    parity_failed = any(m["parity"] is not None and m["parity"]["mismatches"] for m in [main] + sweep)
    if parity_failed:
        print(f"[bench] PARITY FAILURE: {[m['parity'] for m in [main] + sweep if m['parity']]}", file=sys.stderr)

    if rank == 0:
        hbm_peak, tf_peak, peak_src, tf_sustained = _peaks()
        n_local = hi - lo
        kname = "scan_mma_bf16_kernel<exact>" if exact and K_TOP <= 64 else ("scan_mma_bf16_kernel<collect>" if exact else "scan_fma_kernel")

        traffic_tab = {}
        tp = os.path.join(ROOT, "profiles", "r02_traffic.json")
        if os.path.exists(tp) and exact:
            traffic_tab = json.load(open(tp)).get("scan_mma_bf16_kernel", {})

        def roof(m):
            r = roof_kernel(m)
            if r is None:
                return None
            step_ms = m["dev_ms"] / m["steps"]
            if r["bound"] == "tensor":
                # the timed region is a long tensor-bound run at the board's power cap (clocks.reasons): the same
                # kernel sampled right after it, and the whole step, are set against the SUSTAINED cuBLAS figure
                if m["kernel_hot_ms"]:
                    hot = r["algorithmic_flops"] / (m["kernel_hot_ms"] * 1e-3) / 1e12
                    r["sustained"] = {"kernel_ms": m["kernel_hot_ms"], "achieved": hot, "peak": tf_sustained,
                                      "frac": hot / tf_sustained, "peak_source": peak_src + " (bf16 sustained)"}
                r["whole_step_frac"] = r["algorithmic_flops"] / (step_ms * 1e-3) / 1e12 / tf_sustained
            else:
                # the same algorithmic bytes over the WHOLE step (all launches, steps pipelined as timed)
                r["whole_step_frac"] = r["algorithmic_bytes"] / (step_ms * 1e-3) / 1e9 / hbm_peak
            r["kernel_vs_step"] = ("kernel_ms is the scan kernel timed ALONE; with %d steps in flight the latency-bound kernels of a step "
                                   "overlap the neighbouring scans, so ms_per_step can be below kernel_ms" % m["depth"]) if m["depth"] > 1 else None
            return r

        def roof_kernel(m):
            if m["kernel_ms"] is None:
                return None
            # DRAM bytes per launch from the committed ncu --set full capture of this kernel / shard / batch class
            traffic = traffic_tab.get(f"rows{n_local}_b{'le128' if m['B'] <= 128 else m['B']}_{m['data']}")
            bytes_ = n_local * store.ld * elem + min(m["B"], 128) * DIM * elem
            flops = 2.0 * m["B"] * n_local * DIM
            gbs = bytes_ / (m["kernel_ms"] * 1e-3) / 1e9
            tfs = flops / (m["kernel_ms"] * 1e-3) / 1e12
            if exact and flops / bytes_ > tf_peak * 1e12 / (hbm_peak * 1e9):
                return {"bound": "tensor", "achieved": tfs, "peak": tf_peak, "unit": "TFLOP/s", "frac": tfs / tf_peak,
                        "traffic": traffic, "kernel": kname, "kernel_ms": m["kernel_ms"],
                        "peak_source": peak_src + " (bf16 burst)", "algorithmic_flops": flops}
            return {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                    "traffic": traffic, "kernel": kname, "kernel_ms": m["kernel_ms"], "peak_source": peak_src + " (copy bandwidth)",
                    "algorithmic_bytes": bytes_}

        B = main["B"]
        cpu = cpu_baseline_leg(args.data) if world == 1 and not args.no_cpu else None
        line = {
            "metric": metric_name(), "value": B * args.steps / (main["dev_ms"] * 1e-3), "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": main["dev_ms"] / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16 scan + f32 rescore" if exact else "f32", "data": "synthetic",
            "config": dict(_config(B, world, main["path"], args.data, exact), launch="cuda-graph" if main["graph"] else "eager",
                           **({"exchange": "peer-memory push + merge kernel over NVLink (vq_peer_exchange_merge)"
                               if lanes[0].searcher._peer is not None else "nccl all-gather + vq_topk_merge"} if world > 1 else {}),
                           steps_in_flight=main["depth"],
                           **({"max_steps_enqueued": main["cap"]} if main["cap"] else {}),
                           e2e_launch="H2D + search + D2H captured in one CUDA graph per step" if main["host_graph"]
                           else "explicit pinned copies around the step",
                           **({"e2e_ingest": "each rank copies 1/N of the batch from pinned host memory, slices all-gathered "
                                             "over NVLink peer memory (vq_peer_allgather_rows)" if main["ingest"] == "slice"
                               else "each rank copies the whole batch from pinned host memory"} if world > 1 else {})),
            "clocks": clocks.summary(),
            "e2e": {"value": B * args.steps / (main["e2e_ms"] * 1e-3), "unit": "queries/s",
                    "h2d_bytes_per_step": B * DIM * 4 * (1 if main["ingest"] == "slice" or world == 1 else world),
                    "d2h_bytes_per_step": B * K_TOP * (12 if world > 1 else 8) + (B * 4 if exact else 0),
                    "host_thread_rank0": main["e2e_host"], "fallback_queries": main["fallbacks"]},
            "gpu_launches": main["launches"] * args.steps,
            "roofline": roof(main),
            "parity": main["parity"],
            "overflowed_queries_per_batch": main["overflow"],
            "exact_search": main["gathered"],
            "sweep": [{"batch": m["B"], "data_kind": m["data"], "value": m["B"] * m["steps"] / (m["dev_ms"] * 1e-3),
                       "ms_per_step": m["dev_ms"] / m["steps"], "roofline": roof(m),
                       "overflowed_queries_per_batch": m["overflow"], "exact_search": m["gathered"],
                       **({"parity": m["parity"]} if m["parity"] is not None else {})} for m in sweep],
        }
        if main["sustained"] is not None:
            line["sustained"] = main["sustained"]
        if main["breakdown"] is not None:
            line["step_breakdown"] = main["breakdown"]
        if api is not None:
            line["e2e_api"] = api
        if hnsw_sharded is not None:
            line["hnsw_sharded"] = hnsw_sharded
        if world == 1 and not args.no_hnsw and args.config == 2:
            # BASELINE config 3 next to the headline: HNSW M=16 ef_construction=200 on 1M x 512 clustered
            # (CLIP-like) rows, ef_search 64-256, recall@10 against the exact scan, QPS and the
            # achieved gather bandwidth from the kernel's own evaluation / expansion counters
            try:
                from types import SimpleNamespace
                from tools import bench_hnsw
                del stores[:]
                torch.cuda.empty_cache()
                h = bench_hnsw.measure(SimpleNamespace(n=N_ROWS, dim=DIM, queries=10000, kind="clip", efs="64,128,256",
                                                       search_dtype="fp32"), dev)
                line["hnsw"] = {"workload": "HNSW M=16 ef_construction=200 max_M=16, 1M x 512 clustered rows, 10000 queries, k=10",
                                "build_s": h["build_s"], "hbm_peak_GBps": hbm_peak, "runs": h["runs"]}
            except Exception as e:  # noqa: BLE001 — the headline line must survive a failure of the extra leg
                line["hnsw"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if parity_failed:
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS),
                    help="BASELINE.json config: 2 = 1M x 512 k=10 (contract workload), 4 = 10M x 768 k=100, 5 = 100M x 512 batch 4096")
    ap.add_argument("--data", default="clip", choices=["gauss", "clip"],
                    help="synthetic distribution (SURVEY.md 8(d)): S-clip clustered (default for the headline) or S-gauss iid")
    ap.add_argument("--batch", type=int, default=0,
                    help="query batch of the headline line (default: the config's; config 2 names 1, 32 and 1024, the others are swept)")
    ap.add_argument("--path", default="auto")
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "collective"],
                    help="N > 1: fused NVLink peer-memory push+merge kernel (auto/peer) or NCCL all-gather + merge")
    ap.add_argument("--pipeline", type=int, default=0,
                    help="search steps in flight (each on its own stream with its own workspace / exchange windows); default 4 on "
                         "one GPU (the end-to-end loop with host buffers: 1.22 / 1.30 / 1.29 M QPS at 3 / 4 / 6, the device-resident "
                         "loop does not care: tools/pipeline_ab.sh), 3 with an exchange in the step (measured at N = 2 / 4 / 8)")
    ap.add_argument("--sustain", type=float, default=2.0, help="seconds of extra device-resident steps after the contract region (0 = off)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-api", action="store_true", help="skip the e2e_api leg (drop-in facade, dicts out)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the batch 1/32/1024 sweep")
    ap.add_argument("--no-hnsw", action="store_true", help="skip the HNSW (BASELINE config 3) leg")
    ap.add_argument("--ingest", default="slice", choices=["slice", "full"],
                    help="e2e at N > 1: every rank copies its slice of the host batch + NVLink all-gather (slice), "
                         "or every rank copies the whole batch (full)")
    ap.add_argument("--no-host-graph", action="store_true",
                    help="e2e: explicit H2D/D2H copies around the captured step instead of capturing them with it")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--mode", default="exact", choices=["exact", "fp32"],
                    help="exact: bf16 tensor scan + exact fp32 re-score (default); fp32: FMA scan of the fp32 store")
    ap.add_argument("--rows", type=int, default=0, help="store rows (experiments only: overrides the config's)")
    args = ap.parse_args()
    if args.pipeline <= 0:
        args.pipeline = 4 if int(os.environ.get('WORLD_SIZE', '1')) == 1 else 3
    cfg = CONFIGS[args.config]
    g = globals()
    g["N_ROWS"], g["DIM"], g["K_TOP"], g["LABEL"] = args.rows or cfg["rows"], cfg["dim"], cfg["k"], cfg["label"]
    if args.rows and args.rows != cfg["rows"]:
        g["LABEL"] = f"experiment: {cfg['label']} shape with --rows {args.rows}"
    args.batch = args.batch or cfg["batch"]
    # the contract is ONE JSON line on stdout: libraries that print there (NCCL's version banner, ...)
    # are redirected to stderr for the whole run and the line is written to the real stdout at the end
    args._stdout_fd = os.dup(1)
    real_stdout = os.fdopen(os.dup(args._stdout_fd), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    real_stdout.flush()


if __name__ == "__main__":
    main()
