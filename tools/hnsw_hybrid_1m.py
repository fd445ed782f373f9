#!/usr/bin/env python3
"""Developer analysis: where does the GPU-built HNSW graph lose recall against the reference at 1M clustered rows?
Loads the reference's graph (oracle/_build/ref_graph_clip1m.npz, built by the C++ restatement on the same rows), builds
the GPU graph, and searches both plus the two hybrids (upper layers of one + layer 0 of the other) with the GPU kernel."""
import json, os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tools.hnsw_recall_at_scale import gpu_truth
from video_quierer_b200.hnsw_index import B200HNSWIndex, DeviceGraph
from video_quierer_b200.utils import synth

n, dim, nq = 1_000_000, 512, 1000
ref = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_build", "ref_graph_clip1m.npz"))
store = synth.clip_like(n, dim, seed=0)
queries = synth.clip_like(nq, dim, seed=1, n_store=n)
truth = gpu_truth(store, queries)
sel = sys.argv[1] if len(sys.argv) > 1 else "diverse"
random.seed(0)
h = B200HNSWIndex(dimension=dim, M=16, ef_construction=200, ef_search=64, max_M=16, select=sel)
h.add_batch(list(store), list(range(n)))
h.build()
assert np.array_equal(np.asarray(h._level_list, dtype=np.int32), ref["levels"]), "level streams differ"
ours = h._graph
dev = h.device
refg = DeviceGraph.from_numpy(ref["levels"], ref["adj0"], ref["upper_off"], ref["upper_adj"], int(ref["entry"]), int(ref["max_level"]), dev)


def recall(g, label):
    h._graph = g
    out = {}
    for ef in (64, 128, 256):
        h.ef_search = ef
        _, rows = h.search_arrays(queries, 10)
        out[ef] = round(float(np.mean([len(set(rows[i]) & set(truth[i])) / 10 for i in range(nq)])), 4)
    deg = (g.adj0 >= 0).sum(dim=1).float()
    print(json.dumps({"graph": label, "recall": out, "evals": float(h.last_stats[:, 0].mean()), "deg0_mean": float(deg.mean()),
                      "deg0_lt16": float((deg < 16).float().mean())}), flush=True)


recall(refg, "reference graph (searched by the GPU kernel)")
recall(ours, f"gpu {sel}")
recall(DeviceGraph(ours.levels, refg.adj0, ours.upper_off, ours.upper_adj, ours.entry, ours.max_level), f"upper gpu-{sel} + layer0 reference")
recall(DeviceGraph(refg.levels, ours.adj0, refg.upper_off, refg.upper_adj, refg.entry, refg.max_level), f"upper reference + layer0 gpu-{sel}")
