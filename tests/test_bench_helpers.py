"""CPU tests of bench.py's host-side helpers: the parity comparison rule that gates every bench line, the workload
table and the reference arm's plumbing (no GPU, no timing)."""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def test_compare_topk_rule():
    s = np.array([[0.9, 0.8, 0.7, 0.6]], np.float32)
    r = np.array([[4, 3, 2, 1]], np.int64)
    assert bench.compare_topk(r, s, r, s) == 0
    # scores within 1e-5 relative are fine, beyond are not
    assert bench.compare_topk(r, s * np.float32(1 + 5e-6), r, s) == 0
    assert bench.compare_topk(r, s * np.float32(1 + 5e-4), r, s) == 1
    # a different id is a mismatch ...
    r2 = r.copy(); r2[0, 2] = 9
    assert bench.compare_topk(r2, s, r, s) == 1
    # ... unless it sits in a run of scores that tie within 1e-5 (two rows swapped)
    st = np.array([[0.9, 0.8, 0.8 * (1 - 2e-6), 0.6]], np.float32)
    rs = np.array([[4, 2, 3, 1]], np.int64)
    assert bench.compare_topk(rs, st, r, st) == 0
    # every query counts once
    assert bench.compare_topk(np.vstack([r2, r2, r]), np.vstack([s, s, s]), np.vstack([r, r, r]), np.vstack([s, s, s])) == 2


def test_workload_table_matches_baseline_json():
    import json
    cfg = json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"]
    assert bench.CONFIGS[2]["rows"] == 1_000_000 and bench.CONFIGS[2]["dim"] == 512 and bench.CONFIGS[2]["k"] == 10 and "1M×512" in cfg[1]
    assert bench.CONFIGS[4]["rows"] == 10_000_000 and bench.CONFIGS[4]["dim"] == 768 and bench.CONFIGS[4]["k"] == 100 and "10M×768" in cfg[3]
    assert bench.CONFIGS[5]["rows"] == 100_000_000 and bench.CONFIGS[5]["batch"] == 4096 and "100M×512" in cfg[4]
    bench.N_ROWS, bench.DIM, bench.K_TOP, bench.LABEL = 1_000_000, 512, 10, "BASELINE config 2"
    c = bench._config(1024, 8, "x", "clip")
    assert c["n_rows"] == 1_000_000 and c["batch"] == 1024 and c["parallelism"] == "rows/8" and "S-clip" in c["workload"]
    assert "identical copies" in c["l2"]                   # a 128 MB shard would sit in L2: copies rotate
    assert "evicts itself" in bench._config(1024, 1, "x", "gauss")["l2"]
    assert bench.metric_name() == "search QPS @k=10 (1M x 512 frames, exact scan)"


def test_reference_arm_uses_the_vendored_reference_when_present():
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    mod = bench._load_reference()
    if not os.path.exists(os.path.join(ref_dir, "video_search_overhaul.py")):
        assert mod is None                                  # falls back to the oracle port and says so in the line
        return
    idx = mod.SimpleVideoIndex()
    rng = np.random.default_rng(0)
    x = rng.standard_normal((50, 16)).astype(np.float32)
    for i, v in enumerate(x):
        idx.add_frame(v, "v.mp4", float(i))
    hits = idx.search(x[7], 3)
    assert hits[0]["frame_id"] == 7 and sorted(hits[0]) == ["frame_id", "score", "timestamp", "video_name"]
    # the vendored copy is byte-identical to the reference where both are visible (authoring container)
    src = "/root/reference/video_search_overhaul.py"
    if os.path.exists(src):
        assert open(src, "rb").read() == open(os.path.join(ref_dir, "video_search_overhaul.py"), "rb").read()


def test_traffic_table_keys_match_the_capture_names():
    """profiles/r02_traffic.json (what bench.py reports as roofline.traffic) is keyed by tools/ncu_summarise.py from the
    capture file names: the keys bench.py looks up must be the ones the summariser writes, and the committed table must
    carry the shapes the driver's runs use (1M rows at N = 1, its 2- / 4- / 8-way shards)."""
    import json
    spec2 = importlib.util.spec_from_file_location("ncu_summarise", os.path.join(ROOT, "tools", "ncu_summarise.py"))
    summ = importlib.util.module_from_spec(spec2)
    spec2.loader.exec_module(summ)
    assert summ.traffic_key("r02b_scan_exact_b1024_clip") == ("scan_mma_bf16_kernel", "rows1000000_b1024_clip")
    assert summ.traffic_key("r02b_scan_exact_b1024_gauss") == ("scan_mma_bf16_kernel", "rows1000000_b1024_gauss")
    assert summ.traffic_key("r02b_scan_exact_b32_clip") == ("scan_mma_bf16_kernel", "rows1000000_ble128_clip")
    assert summ.traffic_key("r02b_scan_exact_b1024_shard125k") == ("scan_mma_bf16_kernel", "rows125000_b1024_clip")
    assert summ.traffic_key("r02b_exact_finish_b1024_clip") == ("exact_finish_kernel", "rows1000000_b1024_clip")
    tab = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))["scan_mma_bf16_kernel"]
    for rows in (1_000_000, 500_000, 250_000, 125_000):
        t = tab["rows%d_b1024_clip" % rows]
        assert 1.0 <= t / (rows * 512 * 2) < 1.35           # DRAM traffic of the scan: the algorithmic bytes + gather buffers
