#!/bin/bash
python - <<'PY'
import sys, os, random, numpy as np, torch
sys.path.insert(0, os.getcwd())
from video_quierer_b200.hnsw_index import B200HNSWIndex
from video_quierer_b200.utils import synth
store = synth.clip_like(100000, 512, seed=0)
for sel in ("diverse", "incremental", "closest"):
    random.seed(0)
    h = B200HNSWIndex(dimension=512, M=16, ef_construction=200, ef_search=64, max_M=16, select=sel)
    h.add_batch(list(store), list(range(100000)))
    h.build()
    g = h._graph
    np.savez_compressed(f"gpurun_out/r2i_graph100k_{sel}.npz", levels=g.levels.cpu().numpy(), adj0=g.adj0.cpu().numpy(),
                        upper_off=g.upper_off.cpu().numpy(), upper_adj=g.upper_adj.cpu().numpy(), entry=g.entry, max_level=g.max_level)
    print("saved", sel, flush=True)
PY
timeout 600 python tools/hnsw_recall_at_scale.py --n 1000000 --kinds clip --select diverse --max-candidates 95
