"""Runs the UNMODIFIED reference application code (baseline/_ref, copied from /root/reference by
__graft_entry__.build()) with either its stock index classes or the B200 drop-ins and prints what the HTTP /
orchestrator layer answers, as one JSON line.  Executed by tests/test_gpu_dropin.py in a subprocess whose cwd
is a scratch directory (the reference resolves "videos/", "static/" and config.json relative to it).

    python dropin_driver.py live  <repo_root> stock|b200
    python dropin_driver.py orch  <repo_root> stock|b200
"""
import asyncio
import json
import os
import pickle
import sys
import types

import numpy as np

mode, root, which = sys.argv[1], sys.argv[2], sys.argv[3]
REF = os.path.join(root, "baseline", "_ref")
sys.path.insert(0, root)
sys.dont_write_bytecode = True
os.environ.setdefault("HF_HUB_OFFLINE", "1")
os.environ.setdefault("TRANSFORMERS_OFFLINE", "1")

from video_quierer_b200.utils import synth  # noqa: E402


def live():
    """server.py -> src/api/routes.py -> video_search_overhaul.py."""
    sys.path.insert(0, REF)
    n, dim = 5000, 512
    rows = synth.clip_like(n, dim, seed=3)
    # keyword queries of the non-CLIP encoder (video_search_overhaul.py:297-322) light up dims 0/10/20/30:
    # plant rows that answer them so that the top-k is not a set of near-ties around 0
    for j, d in enumerate((0, 10, 20, 30)):
        for i in range(40):
            rows[100 * j + i, d] += 0.5 + 0.01 * i
    rows /= np.linalg.norm(rows, axis=1, keepdims=True)
    os.makedirs("videos", exist_ok=True)
    os.makedirs("static", exist_ok=True)
    with open(os.path.join(REF, "static", "index.html")) as f, open("static/index.html", "w") as g:
        g.write(f.read())
    with open(os.path.join(REF, "config.json")) as f, open("config.json", "w") as g:
        g.write(f.read())
    meta = [{"video_name": f"v{i // 250}.mp4", "timestamp": float(i % 250) * 0.5, "frame_id": i} for i in range(n)]
    with open("videos/video_search_cache.pkl", "wb") as f:
        pickle.dump({"embeddings": [r for r in rows.astype(np.float32)], "metadata": meta, "video_hashes": {}, "version": "1.0"}, f)
    import video_search_overhaul as vso
    if which == "b200":
        from video_quierer_b200.flat_index import B200FlatIndex
        vso.SimpleVideoIndex = B200FlatIndex               # the one-line change
    import server                                          # noqa: F401  (the reference FastAPI app)
    from fastapi.testclient import TestClient
    import src.api.routes as routes
    if which == "b200":
        # INTEGRATION.md section 1, second change: one batched launch behind /api/search/batch.  The handler body
        # is reference code; only its per-query loop target is swapped, through the system object it already uses.
        from video_quierer_b200.scheduler import BatchSearchScheduler
    out = {}
    with TestClient(server.app) as client:
        system = routes.get_video_search_system()
        out["index_class"] = type(system.index).__name__
        out["use_clip"] = bool(system.processor.use_clip)
        out["n"] = len(system.index.embeddings)
        out["single"] = {}
        for q, k in (("bright car", 3), ("dark phone app", 5), ("football goal", 10), ("vehicle", 50)):
            r = client.post("/api/search", json={"query": q, "k": k, "use_cache": True})
            body = r.json()
            out["single"][f"{q}|{k}"] = {"status": r.status_code, "keys": sorted(body.keys()), "results": body.get("results"),
                                         "from_cache": body.get("from_cache"), "performance": body.get("performance")}
        r = client.post("/api/search/batch", json={"queries": ["bright car", "phone", "football goal", "dark vehicle"], "k": 4})
        out["batch"] = {"status": r.status_code, "body": r.json()}
        out["k100"] = client.post("/api/search", json={"query": "car", "k": 100}).status_code
        out["blank"] = client.post("/api/search", json={"query": "   ", "k": 3}).status_code
        if which == "b200":
            sched = BatchSearchScheduler(system)
            out["scheduler_batch"] = sched.batch_response(["bright car", "phone", "football goal", "dark vehicle"], 4)
            out["scan_path"] = system.index.last_scan_path
        # a handler that mutates the index through its attribute surface (routes.py:754-762 pattern), then search again
        idx = system.index
        first = out["single"]["bright car|3"]["results"][0]["frame_id"]
        idx.embeddings.pop(first)
        idx.metadata.pop(first)
        r = client.post("/api/search", json={"query": "bright car", "k": 3})
        out["after_pop"] = r.json().get("results")
    print("DROPIN_JSON " + json.dumps(out, default=float))


def orch():
    """src/video_search_system.py (the designed-but-unwired orchestrator)."""
    sys.path.insert(0, os.path.join(REF, "src"))
    for name in ("aioredis", "redis"):                      # hard imports of storage/cache.py:14-15, not installed
        sys.modules.setdefault(name, types.ModuleType(name))
    import yaml
    from utils.config import get_default_config
    cfg = get_default_config()
    cfg["cache"]["enable_cache"] = False
    with open("config.yaml", "w") as f:
        yaml.safe_dump(cfg, f)
    import video_search_system as vss
    if which == "b200":
        from video_quierer_b200.hnsw_index import B200HNSWIndex
        vss.OptimizedHNSWIndex = B200HNSWIndex              # the import swap

    class FakeExtractor:                                    # the real one downloads CLIP (feature_extractor.py:76-77)
        batch_size = 8
        def __init__(self, *a, **k): pass
        def extract_text_features(self, text):
            rng = np.random.default_rng(abs(hash(text)) % (2 ** 31))
            return rng.standard_normal(512).astype(np.float32)
        def get_stats(self): return {}
    vss.FeatureExtractor = FakeExtractor
    import random
    random.seed(0)
    s = vss.VideoSearchSystem("config.yaml")
    n_vid, per = 40, 30
    rows = synth.clip_like(n_vid * per, 512, seed=5)
    ids = []
    for v in range(n_vid):
        for i in range(per):
            nid = f"vid{v}_{i}"
            ids.append(nid)
            s.video_metadata[nid] = {"video_id": f"vid{v}", "timestamp": float(i), "frame_number": i, "video_path": f"/x/vid{v}.mp4"}
        s.video_metadata[f"video_vid{v}"] = {"video_id": f"vid{v}", "path": f"/x/vid{v}.mp4", "duration": float(per), "frame_count": per,
                                             "indexed_at": 0.0, "file_size": 1234}
    s.index.add_batch(list(rows), ids)
    queries = synth.clip_like(24, 512, seed=6, n_store=n_vid * per)
    out = {"index_class": type(s.index).__name__, "size": s.index.size(), "queries": []}
    for q in queries:
        res = asyncio.run(s.search(q, k=3, use_cache=True))
        out["queries"].append({"keys": sorted(res.keys()), "result_keys": sorted(res["results"][0].keys()) if res["results"] else [],
                               "video_ids": [r["video_id"] for r in res["results"]],
                               "scores": [float(r["score"]) for r in res["results"]],
                               "from_cache": res.get("from_cache"), "performance": {k: v for k, v in res.get("performance", {}).items()
                                                                                    if k in ("results_count", "total_results_found")}})
    again = asyncio.run(s.search(queries[0], k=3, use_cache=True))
    out["second_call_from_cache"] = again.get("from_cache")
    out["batch_len"] = len(asyncio.run(s.search_batch([q for q in queries[:5]], k=3)))
    out["stats_keys"] = sorted(s.index.get_stats().keys())
    out["health"] = asyncio.run(s.health_check()).get("status") if hasattr(s, "health_check") else None
    s.index.thread_pool.shutdown()
    print("DROPIN_JSON " + json.dumps(out, default=float))


if __name__ == "__main__":
    live() if mode == "live" else orch()
