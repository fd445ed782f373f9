"""The reference's OWN application code (baseline/_ref = /root/reference copied unmodified by
__graft_entry__.build()) driven with the B200 drop-ins swapped in, against the same code with its stock
index classes:

  * server.py + src/api/routes.py under fastapi.testclient.TestClient, SimpleVideoIndex -> B200FlatIndex
    (routes.py:589-645: POST /api/search, POST /api/search/batch, 422 for k > 50, 400 for a blank query);
  * src/video_search_system.py (2k over-fetch + one-hit-per-video dedup, :297-342), OptimizedHNSWIndex ->
    B200HNSWIndex.

Each arm runs in its own subprocess with a scratch cwd (tests/dropin_driver.py)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def _run(mode, which, tmp_path):
    cwd = tmp_path / f"{mode}_{which}"
    cwd.mkdir()
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_driver.py"), mode, ROOT, which],
                         cwd=str(cwd), capture_output=True, text=True, timeout=900)
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("DROPIN_JSON ")]
    assert lines, (out.stdout[-3000:], out.stderr[-3000:])
    return json.loads(lines[-1][len("DROPIN_JSON "):])


def _same_hits(a, b, id_key):
    """Same hits in the same order, except that neighbours whose scores tie within 1e-5 may swap."""
    assert len(a) == len(b)
    sa, sb = np.array([h["score"] for h in a]), np.array([h["score"] for h in b])
    assert np.allclose(sa, sb, rtol=1e-5, atol=1e-7)
    for i, (x, y) in enumerate(zip(a, b)):
        if x[id_key] != y[id_key]:
            tie = np.abs(sb - sb[i]) <= 1e-5 * max(abs(sb[i]), 1e-30)
            assert x[id_key] in {h[id_key] for h, t in zip(b, tie) if t}, (i, x, y)
        else:
            assert {k: v for k, v in x.items() if k != "score"} == {k: v for k, v in y.items() if k != "score"}


@pytest.fixture(scope="module")
def need_ref():
    if not os.path.isdir(REF):
        pytest.skip("baseline/_ref absent (run __graft_entry__.build() where /root/reference exists)")


def test_reference_server_with_b200_flat_index(built_lib, need_ref, tmp_path):
    stock = _run("live", "stock", tmp_path)
    b200 = _run("live", "b200", tmp_path)
    assert stock["index_class"] == "SimpleVideoIndex" and b200["index_class"] == "B200FlatIndex"
    assert stock["use_clip"] is False and b200["use_clip"] is False and stock["n"] == b200["n"] == 5000
    assert b200["scan_path"] == "scan_mma_bf16<exact>+finish"
    for key, s in stock["single"].items():
        g = b200["single"][key]
        assert g["status"] == s["status"] == 200 and g["keys"] == s["keys"], key
        assert g["from_cache"] == s["from_cache"] and g["performance"] == s["performance"], key
        _same_hits(g["results"], s["results"], "frame_id")
        assert all(isinstance(h["score"], float) and "formatted_time" in h for h in g["results"])
    sb, gb = stock["batch"], b200["batch"]
    assert gb["status"] == sb["status"] == 200
    assert {k: v for k, v in gb["body"].items() if k != "results"} == {k: v for k, v in sb["body"].items() if k != "results"}
    for x, y in zip(gb["body"]["results"], sb["body"]["results"]):
        assert x["query"] == y["query"] and x["count"] == y["count"]
        _same_hits(x["results"], y["results"], "frame_id")
    # the batched scheduler (one launch for the whole batch) answers exactly what the handler's loop answers
    sch = b200["scheduler_batch"]
    assert sch["query_count"] == sb["body"]["query_count"] and sch["total_results"] == sb["body"]["total_results"]
    for x, y in zip(sch["results"], sb["body"]["results"]):
        assert x["query"] == y["query"] and x["count"] == y["count"]
        _same_hits(x["results"], y["results"], "frame_id")
    assert stock["k100"] == b200["k100"] == 422 and stock["blank"] == b200["blank"] == 400
    _same_hits(b200["after_pop"], stock["after_pop"], "frame_id")     # handlers mutate index.embeddings / .metadata directly


def test_reference_orchestrator_with_b200_hnsw_index(built_lib, need_ref, tmp_path):
    stock = _run("orch", "stock", tmp_path)
    b200 = _run("orch", "b200", tmp_path)
    assert stock["index_class"] == "OptimizedHNSWIndex" and b200["index_class"] == "B200HNSWIndex"
    assert stock["size"] == b200["size"] == 1200
    same_first = same_all = 0
    for s, g in zip(stock["queries"], b200["queries"]):
        assert g["keys"] == s["keys"] and g["result_keys"] == s["result_keys"]
        assert g["from_cache"] == s["from_cache"] and g["performance"] == s["performance"]      # 2k over-fetch, same counts
        same_first += g["video_ids"][:1] == s["video_ids"][:1]
        same_all += g["video_ids"] == s["video_ids"]
        for vid, sc in zip(g["video_ids"], g["scores"]):
            if vid in s["video_ids"]:
                # the per-video hit may be another frame of the same video only if the reference's graph missed the best one
                assert sc >= s["scores"][s["video_ids"].index(vid)] - 1e-5
    n = len(stock["queries"])
    # different graphs (GPU batch build vs the reference's incremental inserts): ANN results agree on almost every query,
    # and where they differ the B200 index found a better (closer) frame — asserted above
    assert same_first >= 0.9 * n and same_all >= 0.75 * n, (same_first, same_all, n)
    assert b200["second_call_from_cache"] == stock["second_call_from_cache"]
    assert b200["batch_len"] == stock["batch_len"] == 5
    assert b200["stats_keys"] == stock["stats_keys"] and b200["health"] == stock["health"]
