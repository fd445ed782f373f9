#!/bin/bash
python tools/hnsw_recall_at_scale.py --n 1000000 --kinds clip --select diverse --max-candidates 63
python tools/hnsw_recall_at_scale.py --n 1000000 --kinds clip --select closest
python tools/hnsw_recall_at_scale.py --n 1000000 --kinds clip --select diverse --max-candidates 32
