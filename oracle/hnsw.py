"""ORACLE — test infrastructure only, never imported by the product path.

CPU restatement of the reference's pure-Python HNSW (reference ``src/indexes/hnsw.py``):
level draw (:68-74), layer beam search (:76-121), top-M "heuristic" selection (:123-148),
insert with bidirectional link + prune + reverse-edge discard (:150-229), and query
(:238-280 / :488-528 — the "optimized" subclass always falls through to the plain layer
search because entry lists have length 1, :445).

Nodes are dense ints in insertion order; external ids are the caller's business.  The
distance is the reference's ``1.0 - np.dot(a, b)`` on 1-D arrays so fp32 values are
bit-identical to the reference's.

Parity status: **pinned** — `tests/golden/make_golden.py` builds graphs with the unmodified
reference (same ``random.seed``) and `tests/test_oracle_hnsw.py` requires this restatement
to reproduce those graphs edge for edge and the reference's search results id for id.
"""

from __future__ import annotations

import heapq
import math
import random
from dataclasses import dataclass

import numpy as np


@dataclass
class GraphArrays:
    """Dense device-format graph shared with the CUDA path."""
    levels: np.ndarray      # int32 [N]
    adj0: np.ndarray        # int32 [N, M0]  layer-0 neighbours, -1 padded
    upper_off: np.ndarray   # int32 [N]      first upper slot of a node, -1 when level 0
    upper_adj: np.ndarray   # int32 [S, M]   slot(node, lv) = upper_off[node] + lv - 1
    entry: int
    max_level: int


class OracleHNSW:
    def __init__(self, dimension=512, M=16, ef_construction=200, ef_search=50, max_M=16,
                 level_generation_factor=1.0 / math.log(2.0)):
        self.dimension = dimension
        self.M = M
        self.max_M = max_M
        self.ef_construction = ef_construction
        self.ef_search = ef_search
        self.mL = level_generation_factor
        self.vec = []            # node -> normalised vector
        self.level_of = []       # node -> top level
        self.links = []          # links[lv][node] -> set(node)
        self.entry = None
        self.dist_evals = 0

    # ---------------------------------------------------------------- primitives
    def _dist(self, a, b):
        """hnsw.py:59-66."""
        self.dist_evals += 1
        return 1.0 - np.dot(a, b)

    def draw_level(self) -> int:
        """hnsw.py:68-74 — uses the *global* ``random`` stream like the reference."""
        return int(-math.log(random.uniform(0, 1)) * self.mL)

    def _layer(self, lv):
        while len(self.links) <= lv:
            self.links.append({})
        return self.links[lv]

    def search_layer(self, q, entries, ef, lv):
        """hnsw.py:76-121: min-heap of candidates, bounded max-heap of results, stop when the
        best candidate is strictly worse than the worst kept (:103), admit when not full or
        strictly better than the worst kept (:113)."""
        layer = self.links[lv] if lv < len(self.links) else {}
        seen = set()
        cand, best = [], []
        for e in entries:
            d = self._dist(q, self.vec[e])
            heapq.heappush(cand, (d, e))
            heapq.heappush(best, (-d, e))
            seen.add(e)
        while cand:
            d, u = heapq.heappop(cand)
            if best and d > -best[0][0]:
                break
            for v in layer.get(u, ()):
                if v in seen:
                    continue
                seen.add(v)
                dv = self._dist(q, self.vec[v])
                if len(best) < ef or dv < -best[0][0]:
                    heapq.heappush(cand, (dv, v))
                    heapq.heappush(best, (-dv, v))
                    if len(best) > ef:
                        heapq.heappop(best)
        return [(-nd, v) for nd, v in best]

    @staticmethod
    def closest(cands, m):
        """hnsw.py:123-148: despite its name, plain closest-M (sorted by (distance, id))."""
        if len(cands) <= m:
            return [v for _, v in cands]
        return [v for _, v in sorted(cands)[:m]]

    # ---------------------------------------------------------------- build
    def add(self, vector, level=None):
        """hnsw.py:150-229.  `level` may be injected to replay a known level sequence."""
        v = vector / np.linalg.norm(vector)           # :157 (no epsilon; zero → NaN row)
        node = len(self.vec)
        self.vec.append(v)
        lvl = self.draw_level() if level is None else int(level)
        self.level_of.append(lvl)
        for lv in range(lvl + 1):
            self._layer(lv)[node] = set()
        if self.entry is None:
            self.entry = node
            return node
        top = self.level_of[self.entry]
        cur = [self.entry]
        for lv in range(max(top, lvl), lvl, -1):       # :175  greedy descent, ef = 1
            cur = [u for _, u in self.search_layer(v, cur, 1, lv)]
        for lv in range(min(lvl, top), -1, -1):        # :183
            found = self.search_layer(v, cur, self.ef_construction, lv)
            cur = [u for _, u in found]
            cap = self.M if lv > 0 else self.max_M     # :191
            layer = self.links[lv]
            for nb in self.closest(found, cap):
                layer[node].add(nb)
                layer[nb].add(node)
                conns = list(layer[nb])
                if len(conns) > cap:                   # :203 prune the neighbour
                    scored = [(self._dist(self.vec[nb], self.vec[c]), c) for c in conns]
                    keep = self.closest(scored, cap)
                    old = layer[nb]
                    layer[nb] = set(keep)
                    for c in old:                      # :221-223 drop the reverse edge too
                        if c not in keep:
                            layer[c].discard(nb)
        if lvl > top:                                  # :226
            self.entry = node
        return node

    # ---------------------------------------------------------------- query
    def search(self, query, k=5, ef_search=None):
        """hnsw.py:238-280: returns [(distance, node)] ascending (ties by node)."""
        if self.entry is None:
            return []
        ef = self.ef_search if ef_search is None else ef_search
        q = query / np.linalg.norm(query)              # :250 (no epsilon)
        cur = [self.entry]
        for lv in range(self.level_of[self.entry], 0, -1):
            cur = [u for _, u in self.search_layer(q, cur, 1, lv)]
        found = self.search_layer(q, cur, max(ef, k), 0)
        return sorted(found)[:k]

    # ---------------------------------------------------------------- export
    def to_arrays(self) -> GraphArrays:
        return links_to_arrays(self.links, self.level_of, self.entry, self.M, self.max_M)

    def store(self) -> np.ndarray:
        return np.stack(self.vec).astype(np.float32) if self.vec else np.zeros((0, self.dimension), np.float32)


def links_to_arrays(links, level_of, entry, M, max_M) -> GraphArrays:
    n = len(level_of)
    levels = np.asarray(level_of, dtype=np.int32)
    adj0 = np.full((n, max_M), -1, np.int32)
    upper_off = np.full(n, -1, np.int32)
    slots = 0
    for u in range(n):
        if levels[u] > 0:
            upper_off[u] = slots
            slots += int(levels[u])
    upper_adj = np.full((max(slots, 1), M), -1, np.int32)
    for lv, layer in enumerate(links):
        for u, nbrs in layer.items():
            nb = sorted(nbrs)
            if lv == 0:
                adj0[u, :len(nb)] = nb
            else:
                upper_adj[upper_off[u] + lv - 1, :len(nb)] = nb
    return GraphArrays(levels, adj0, upper_off, upper_adj,
                       -1 if entry is None else int(entry),
                       int(levels[entry]) if entry is not None else 0)


def search_arrays(store: np.ndarray, g: GraphArrays, query: np.ndarray, k: int, ef_search: int,
                  normalise: bool = True):
    """hnsw.py:238-280 on the dense graph format.  Returns (found, dist_evals, expansions)
    with found = [(distance, node)] ascending."""
    if g.entry < 0:
        return [], 0, 0
    q = query / np.linalg.norm(query) if normalise else query
    evals = hops = 0

    def nbrs(u, lv):
        row = g.adj0[u] if lv == 0 else g.upper_adj[g.upper_off[u] + lv - 1]
        return [int(v) for v in row if v >= 0]

    def layer(entries, ef, lv):
        nonlocal evals, hops
        seen = set(entries)
        cand, best = [], []
        for e in entries:
            d = 1.0 - np.dot(q, store[e]); evals += 1
            heapq.heappush(cand, (d, e)); heapq.heappush(best, (-d, e))
        while cand:
            d, u = heapq.heappop(cand)
            if best and d > -best[0][0]:
                break
            hops += 1
            for v in nbrs(u, lv):
                if v in seen:
                    continue
                seen.add(v)
                dv = 1.0 - np.dot(q, store[v]); evals += 1
                if len(best) < ef or dv < -best[0][0]:
                    heapq.heappush(cand, (dv, v)); heapq.heappush(best, (-dv, v))
                    if len(best) > ef:
                        heapq.heappop(best)
        return [(-nd, v) for nd, v in best]

    cur = [g.entry]
    for lv in range(g.max_level, 0, -1):
        cur = [u for _, u in layer(cur, 1, lv)]
    found = layer(cur, max(ef_search, k), 0)
    return sorted(found)[:k], evals, hops
