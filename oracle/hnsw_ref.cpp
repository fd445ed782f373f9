// ORACLE — test infrastructure only (tests/ and bench baselines); never linked into or called by the product.
//
// C++ restatement of the reference's pure-Python HNSW, src/indexes/hnsw.py, fast enough to give the reference's
// recall at 100k - 1M rows where the Python build (6-16 ms per insert) is infeasible:
//   level draw      :68-74   (levels are INJECTED: the wrapper replays Python's `random` stream)
//   _search_layer   :76-121  (min-heap of candidates, bounded max-heap of results, strict > stop :103, strict < admit :113)
//   selection       :123-148 (plain closest-M despite its name; all candidates when there are <= M)
//   add             :150-229 (ef=1 descent :174-180, connect :183-199, prune + reverse-edge discard :202-223,
//                             entry-point update :226-227)
//   search          :238-280 (ef = max(ef_search, k) :264, sorted (distance, id) :269)
// To reproduce the reference's graphs EDGE FOR EDGE the restatement also reproduces what the Python code
// inherits from CPython and NumPy:
//   * `1.0 - np.dot(a, b)` on float32 vectors is OpenBLAS sdot; the wrapper hands over the cblas_sdot of the very
//     OpenBLAS NumPy loaded, so every distance is bit-identical;
//   * neighbour sets are Python `set`s of ints and are iterated in hash-table order (`for neighbor in neighbors`,
//     `list(self.graph[lv][neighbor])`): PySet below follows Objects/setobject.c (linear probes 9, perturb shift 5,
//     fill*5 >= mask*3 resize to used*4, dummies on discard) so the iteration order is CPython's;
//   * results of a layer search are returned in heapq's ARRAY order (`[(-d, id) for d, id in dynamic_list]`), which
//     is the order neighbours are linked and pruned in when there are <= M candidates: Heap below is heapq's
//     _siftdown/_siftup verbatim in behaviour.
// Pinned by tests/test_oracle_hnsw_ref.py: the golden graphs built by the UNMODIFIED reference (10k x 512 clustered
// and iid, M=16; 1.5k and 800-row graphs with M != max_M) are reproduced edge for edge, searches id for id.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

namespace {

typedef float (*SdotFn)(int64_t, const float*, int64_t, const float*, int64_t);

// ---------------------------------------------------------------------------- CPython set of non-negative ints
struct PySet {
    static constexpr int64_t kUnused = -1, kDummy = -2;     // keys >= 0 are active (hash(int) == int)
    std::vector<int64_t> table;
    size_t mask = 7, fill = 0, used = 0;
    PySet() : table(8, kUnused) {}

    static void insert_clean(std::vector<int64_t>& t, size_t mask, int64_t key) {
        size_t perturb = (size_t)key, i = (size_t)key & mask;
        while (true) {
            if (t[i] == kUnused) { t[i] = key; return; }
            if (i + 9 <= mask) {
                for (size_t j = 1; j <= 9; ++j)
                    if (t[i + j] == kUnused) { t[i + j] = key; return; }
            }
            perturb >>= 5;
            i = (i * 5 + 1 + perturb) & mask;
        }
    }
    void resize(size_t minused) {
        size_t newsize = 8;
        while (newsize <= minused) newsize <<= 1;
        std::vector<int64_t> nt(newsize, kUnused);
        for (int64_t k : table)
            if (k >= 0) insert_clean(nt, newsize - 1, k);
        table.swap(nt);
        mask = newsize - 1;
        fill = used;
    }
    bool contains(int64_t key) const {
        size_t perturb = (size_t)key, i = (size_t)key & mask;
        while (true) {
            const size_t probes = (i + 9 <= mask) ? 9 : 0;
            for (size_t j = 0; j <= probes; ++j) {
                const int64_t e = table[i + j];
                if (e == kUnused) return false;
                if (e == key) return true;
            }
            perturb >>= 5;
            i = (i * 5 + 1 + perturb) & mask;
        }
    }
    void add(int64_t key) {                                  // set_add_entry
        size_t perturb = (size_t)key, i = (size_t)key & mask;
        int64_t freeslot = -1;
        while (true) {
            const size_t probes = (i + 9 <= mask) ? 9 : 0;
            for (size_t j = 0; j <= probes; ++j) {
                const int64_t e = table[i + j];
                if (e == kUnused) {
                    if (freeslot >= 0) { table[(size_t)freeslot] = key; ++used; return; }
                    table[i + j] = key;
                    ++fill; ++used;
                    if (fill * 5 >= mask * 3) resize(used > 50000 ? used * 2 : used * 4);
                    return;
                }
                if (e == key) return;
                if (e == kDummy) freeslot = (int64_t)(i + j);
            }
            perturb >>= 5;
            i = (i * 5 + 1 + perturb) & mask;
        }
    }
    void discard(int64_t key) {                              // set_discard_entry: the slot becomes a dummy
        size_t perturb = (size_t)key, i = (size_t)key & mask;
        while (true) {
            const size_t probes = (i + 9 <= mask) ? 9 : 0;
            for (size_t j = 0; j <= probes; ++j) {
                const int64_t e = table[i + j];
                if (e == kUnused) return;
                if (e == key) { table[i + j] = kDummy; --used; return; }
            }
            perturb >>= 5;
            i = (i * 5 + 1 + perturb) & mask;
        }
    }
    template <typename F> void for_each(F&& f) const {       // iteration = table order
        for (int64_t k : table) if (k >= 0) f((int)k);
    }
};

// ---------------------------------------------------------------------------- heapq on (float32, int) tuples
typedef std::pair<float, int> Item;
inline bool lt(const Item& a, const Item& b) { return a.first != b.first ? a.first < b.first : a.second < b.second; }
struct Heap {
    std::vector<Item> h;
    void siftdown(size_t startpos, size_t pos) {
        const Item newitem = h[pos];
        while (pos > startpos) {
            const size_t parentpos = (pos - 1) >> 1;
            if (lt(newitem, h[parentpos])) { h[pos] = h[parentpos]; pos = parentpos; continue; }
            break;
        }
        h[pos] = newitem;
    }
    void siftup(size_t pos) {
        const size_t endpos = h.size(), startpos = pos;
        const Item newitem = h[pos];
        size_t childpos = 2 * pos + 1;
        while (childpos < endpos) {
            const size_t rightpos = childpos + 1;
            if (rightpos < endpos && !lt(h[childpos], h[rightpos])) childpos = rightpos;
            h[pos] = h[childpos];
            pos = childpos;
            childpos = 2 * pos + 1;
        }
        h[pos] = newitem;
        siftdown(startpos, pos);
    }
    void push(Item x) { h.push_back(x); siftdown(0, h.size() - 1); }
    Item pop() {
        Item last = h.back();
        h.pop_back();
        if (h.empty()) return last;
        Item ret = h[0];
        h[0] = last;
        siftup(0);
        return ret;
    }
};

struct Index {
    int dim, M, max_M, ef_c;
    SdotFn sdot;
    const float* vec = nullptr;                              // [n, dim] normalised rows (borrowed)
    std::vector<int> level_of;
    std::vector<std::vector<PySet*>> links;                  // links[lv][node] (nullptr: node not in layer)
    int entry = -1;
    int64_t n = 0;
    uint64_t dist_evals = 0;
    std::vector<uint32_t> seen_mark;                         // visited set: epoch marks
    uint32_t epoch = 0;

    ~Index() { for (auto& l : links) for (PySet* s : l) delete s; }
    float dist(const float* a, const float* b) { ++dist_evals; return 1.0f - sdot(dim, a, 1, b, 1); }
    const float* row(int u) const { return vec + (size_t)u * dim; }

    // hnsw.py:76-121; returns dynamic_list in heapq ARRAY order as (distance, id)
    void search_layer(const float* q, const std::vector<int>& entries, int ef, int lv, std::vector<Item>& out) {
        if (++epoch == 0) { std::fill(seen_mark.begin(), seen_mark.end(), 0u); epoch = 1; }
        Heap cand, best;
        for (int e : entries) {
            const float d = dist(q, row(e));
            cand.push({d, e});
            best.push({-d, e});
            seen_mark[e] = epoch;
        }
        const std::vector<PySet*>* layer = lv < (int)links.size() ? &links[lv] : nullptr;
        while (!cand.h.empty()) {
            const Item cur = cand.pop();
            if (!best.h.empty() && cur.first > -best.h[0].first) break;
            const PySet* nb = (layer && cur.second < (int)layer->size()) ? (*layer)[cur.second] : nullptr;
            if (!nb) continue;
            nb->for_each([&](int v) {
                if (seen_mark[v] == epoch) return;
                seen_mark[v] = epoch;
                const float dv = dist(q, row(v));
                if ((int)best.h.size() < ef || dv < -best.h[0].first) {
                    cand.push({dv, v});
                    best.push({-dv, v});
                    if ((int)best.h.size() > ef) best.pop();
                }
            });
        }
        out.clear();
        for (const Item& it : best.h) out.push_back({-it.first, it.second});
    }
    // hnsw.py:123-148 (note: sorts the caller's list in place when it is longer than M, like candidates.sort())
    static void closest(std::vector<Item>& c, int m, std::vector<int>& sel) {
        sel.clear();
        if ((int)c.size() <= m) { for (auto& it : c) sel.push_back(it.second); return; }
        std::sort(c.begin(), c.end(), lt);
        for (int i = 0; i < m; ++i) sel.push_back(c[i].second);
    }
    PySet*& slot(int lv, int node) {
        if ((int)links.size() <= lv) links.resize(lv + 1);
        if ((int)links[lv].size() <= node) links[lv].resize(node + 1, nullptr);
        return links[lv][node];
    }
    // hnsw.py:150-229 with the level injected
    void add(int node, int lvl) {
        level_of.push_back(lvl);
        seen_mark.push_back(0);
        for (int lv = 0; lv <= lvl; ++lv) slot(lv, node) = new PySet();
        ++n;
        if (entry < 0) { entry = node; return; }
        const float* v = row(node);
        const int top = level_of[entry];
        std::vector<int> cur{entry}, sel, keep;
        std::vector<Item> found, scored;
        for (int lv = std::max(top, lvl); lv > lvl; --lv) {
            search_layer(v, cur, 1, lv, found);
            cur.clear();
            for (auto& it : found) cur.push_back(it.second);
        }
        for (int lv = std::min(lvl, top); lv >= 0; --lv) {
            search_layer(v, cur, ef_c, lv, found);
            cur.clear();
            for (auto& it : found) cur.push_back(it.second);           // :187 (before the in-place sort of :138)
            const int cap = lv > 0 ? M : max_M;
            closest(found, cap, sel);
            for (int nb : sel) {
                slot(lv, node)->add(nb);
                PySet* ns = slot(lv, nb);
                ns->add(node);
                if ((int)ns->used > cap) {
                    scored.clear();
                    ns->for_each([&](int c) { scored.push_back({dist(row(nb), row(c)), c}); });
                    closest(scored, cap, keep);
                    PySet* fresh = new PySet();
                    for (int c : keep) fresh->add(c);                   // set(selected_for_neighbor)
                    ns->for_each([&](int c) {                           // old_connections
                        if (std::find(keep.begin(), keep.end(), c) == keep.end()) slot(lv, c)->discard(nb);
                    });
                    delete ns;
                    slot(lv, nb) = fresh;
                }
            }
        }
        if (lvl > top) entry = node;
    }
    // hnsw.py:238-280; q already normalised by the wrapper (numpy arithmetic)
    int search(const float* q, int k, int ef_search, float* out_d, int* out_id) {
        if (entry < 0) return 0;
        std::vector<int> cur{entry};
        std::vector<Item> found;
        for (int lv = level_of[entry]; lv > 0; --lv) {
            search_layer(q, cur, 1, lv, found);
            cur.clear();
            for (auto& it : found) cur.push_back(it.second);
        }
        search_layer(q, cur, std::max(ef_search, k), 0, found);
        std::sort(found.begin(), found.end(), lt);
        const int m = std::min<int>(k, (int)found.size());
        for (int i = 0; i < m; ++i) { out_d[i] = found[i].first; out_id[i] = found[i].second; }
        return m;
    }
};

}  // namespace

extern "C" {

void* href_create(int dim, int M, int max_M, int ef_construction, void* sdot_fn) {
    Index* ix = new Index();
    ix->dim = dim; ix->M = M; ix->max_M = max_M; ix->ef_c = ef_construction;
    ix->sdot = (SdotFn)sdot_fn;
    return ix;
}
void href_destroy(void* h) { delete (Index*)h; }
// rows [n, dim] normalised float32 (stays owned by the caller, must outlive the index), levels [n]
void href_build(void* h, const float* rows, const int* levels, int64_t n) {
    Index* ix = (Index*)h;
    ix->vec = rows;
    for (int64_t i = 0; i < n; ++i) ix->add((int)i, levels[i]);
}
int64_t href_size(void* h) { return ((Index*)h)->n; }
int href_entry(void* h) { return ((Index*)h)->entry; }
uint64_t href_dist_evals(void* h) { return ((Index*)h)->dist_evals; }
// neighbours of `node` at `lv`, ascending, into out (capacity cap); returns the count (-1: node not in that layer)
int href_neighbours(void* h, int lv, int node, int* out, int cap) {
    Index* ix = (Index*)h;
    if (lv >= (int)ix->links.size() || node >= (int)ix->links[lv].size() || !ix->links[lv][node]) return -1;
    std::vector<int> v;
    ix->links[lv][node]->for_each([&](int c) { v.push_back(c); });
    std::sort(v.begin(), v.end());
    const int m = std::min<int>((int)v.size(), cap);
    std::memcpy(out, v.data(), (size_t)m * sizeof(int));
    return (int)v.size();
}
// b normalised queries -> ids / distances [b, k] (-1 / +inf padded); returns total distance evaluations of the searches
uint64_t href_search(void* h, const float* queries, int b, int k, int ef_search, float* out_d, int* out_id) {
    Index* ix = (Index*)h;
    const uint64_t before = ix->dist_evals;
    for (int i = 0; i < b; ++i) {
        float* d = out_d + (size_t)i * k;
        int* id = out_id + (size_t)i * k;
        const int m = ix->search(queries + (size_t)i * ix->dim, k, ef_search, d, id);
        for (int j = m; j < k; ++j) { d[j] = __builtin_inff(); id[j] = -1; }
    }
    return ix->dist_evals - before;
}
// differential test hook for PySet: ops[i] = key (add) or -(key+1) (discard); writes the iteration order, returns its length
int href_pyset_order(const int64_t* ops, int n_ops, int64_t* out, int cap) {
    PySet s;
    for (int i = 0; i < n_ops; ++i) { if (ops[i] >= 0) s.add(ops[i]); else s.discard(-(ops[i] + 1)); }
    int m = 0;
    s.for_each([&](int k) { if (m < cap) out[m] = k; ++m; });
    return m;
}

}  // extern "C"
