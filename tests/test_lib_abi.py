"""CPU-only: the C-ABI library builds, loads and exports every symbol include/vq_search.h declares."""
import ctypes
import os
import re

from video_quierer_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "vq_search.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(vq_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported(built_lib):
    names = _declared_symbols()
    assert len(names) >= 12
    lib = ctypes.CDLL(built_lib)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in vq_search.h but not exported"


def test_binding_covers_header(built_lib):
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    lib = _lib.load()
    assert lib.vq_abi_version() == 1
    assert isinstance(lib.vq_last_error(), bytes)


def test_argument_validation_without_gpu(built_lib):
    """Argument errors are reported before any CUDA call, so they are testable on CPU."""
    lib = _lib.load()
    rc = lib.vq_scan_topk(None, 10, 512, 500, 0, None, 1, 10, 1, None, None, None, 0, 0, None)
    assert rc == -1 and b"ld" in lib.vq_last_error()
    rc = lib.vq_scan_topk(None, 10, 512, 512, 7, None, 1, 10, 1, None, None, None, 0, 0, None)
    assert rc == -1 and b"store_dtype" in lib.vq_last_error()
    rc = lib.vq_scan_topk(None, 10, 512, 512, 0, None, 1, 0, 1, None, None, None, 0, 0, None)
    assert rc == -1
    assert lib.vq_scan_workspace_bytes(1000000, 512, 512, 0, 32, 10, 0) > 0


def test_two_stage_argument_validation_without_gpu(built_lib):
    lib = _lib.load()
    # k_cand < k
    rc = lib.vq_search_two_stage(None, None, 1000, 512, 512, None, 4, 10, 5, 1, 0.004, None, None, None, None, 0, None)
    assert rc == -1 and b"k_cand" in lib.vq_last_error()
    # bf16 stores need ld % 64 == 0
    rc = lib.vq_search_two_stage(None, None, 1000, 500, 520, None, 4, 10, 32, 1, 0.004, None, None, None, None, 0, None)
    assert rc == -1 and b"ld" in lib.vq_last_error()
    # b == 0 is a no-op, NULL pointers with b > 0 are rejected
    assert lib.vq_search_two_stage(None, None, 1000, 512, 512, None, 0, 10, 32, 1, 0.004, None, None, None, None, 0, None) == 0
    rc = lib.vq_search_two_stage(None, None, 1000, 512, 512, None, 4, 10, 32, 1, 0.004, None, None, None, None, 0, None)
    assert rc == -1 and b"NULL" in lib.vq_last_error()
    assert lib.vq_search_two_stage_workspace_bytes(1000000, 512, 512, 1024, 32) > (1 << 20)


def test_exact_search_argument_validation_without_gpu(built_lib):
    """vq_search_exact / vq_store_bounds / vq_search_collect check their arguments before any CUDA call."""
    lib = _lib.load()
    one = ctypes.c_void_p(256)                   # non-NULL, 256-byte aligned placeholder: validation must fail before any dereference

    def call(n=1000, dim=512, ld=512, b=4, k=10, norm=1, bounds=one, ws=one):
        return lib.vq_search_exact(one, one, n, dim, ld, one, b, k, norm, bounds, one, one, one, None, ws, 1 << 30, None)
    assert call(k=0) == -1 and call(k=129) == -1 and b"k" in lib.vq_last_error()
    assert call(ld=520) == -1 and b"ld" in lib.vq_last_error()              # bf16 stores need ld % 64 == 0
    assert call(norm=7) == -1
    assert call(bounds=None) == -1 and b"NULL" in lib.vq_last_error()         # the rounding bounds are not optional
    assert call(n=0) == -1
    assert call(b=0) == 0                                                     # empty batch: nothing to do
    assert lib.vq_search_exact_supported(1_000_000, 512, 512, 1024, 10) == 1
    assert lib.vq_search_exact_supported(1_000_000, 512, 512, 1024, 64) == 1
    assert lib.vq_search_exact_supported(10_000_000, 768, 768, 1024, 100) == 0    # k > 64 on a large store: two-pass route
    assert lib.vq_search_exact_supported(20_000, 768, 768, 12, 100) == 1          # listless mode where the gather cannot overflow
    assert lib.vq_search_exact_supported(1_000_000, 1024, 1024, 8, 10) == 0       # rows wider than the tensor-memory budget
    assert lib.vq_search_exact_workspace_bytes(1_000_000, 512, 512, 1024, 10) > (32 << 20)    # 1024 x 4096 gather slots
    assert lib.vq_store_bounds(one, one, 10, 500, one, None) == -1
    assert lib.vq_store_bounds(one, one, 10, 512, None, None) == -1
    assert lib.vq_store_bounds(None, None, 0, 512, one, None) == 0
    rc = lib.vq_search_collect(one, one, 1000, 512, 512, one, 4, 10, 1, None, 8, None, one, one, one, one, 1 << 30, None)
    assert rc == -1 and b"cap" in lib.vq_last_error()                         # cap < k


def test_no_product_import_of_oracle_or_baseline():
    """The product package must never import the oracle or the vendored reference (test / baseline infrastructure)."""
    pkg = os.path.join(ROOT, "video_quierer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+(oracle|baseline)\b", src, flags=re.M), f
                assert "baseline/_ref" not in src and "hnsw_ref" not in src, f


def test_peer_exchange_argument_validation_without_gpu(built_lib):
    """The shard-exchange entry points check their arguments before any CUDA call."""
    lib = _lib.load()
    # window size: header + 2 parities x world x b_max x k_max lines of 16 bytes
    assert lib.vq_peer_window_bytes(8, 1024, 16) >= 256 + 2 * 8 * 1024 * 16 * 16
    assert lib.vq_peer_window_bytes(0, 1024, 16) == 0
    one = ctypes.c_void_p(1)                     # non-NULL placeholders: validation must fail before any dereference
    args = dict(world=2, rank=0, b_max=64, k_max=16, b=8, k=10, k_out=10)

    def call(**kw):
        a = dict(args, **kw)
        return lib.vq_peer_exchange_merge(one, a["world"], a["rank"], a["b_max"], a["k_max"], one, one, a["b"], a["k"],
                                          None, a["k_out"], one, one, None, None)
    assert call(rank=2) == -1 and b"rank" in lib.vq_last_error()
    assert call(world=33) == -1
    assert call(b=65) == -1 and b"window" in lib.vq_last_error()
    assert call(k=17) == -1
    assert call(k_out=0) == -1
    assert lib.vq_peer_exchange_merge(None, 2, 0, 64, 16, one, one, 8, 10, None, 10, one, one, None, None) == -1
    assert call(b=0) == 0                        # empty batch: nothing to do
    assert lib.vq_peer_window_open(None, None) == -1


def test_peer_classes_need_cuda():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from video_quierer_b200.peer import LocalWindows
    with pytest.raises(RuntimeError, match="no CPU path"):
        LocalWindows(2, "cpu", 8, 16)


def test_sharded_searcher_exchange_argument():
    import pytest
    from video_quierer_b200.sharded import ShardedSearcher
    with pytest.raises(ValueError, match="exchange"):
        ShardedSearcher(lambda q, k: None, 10, exchange="smoke-signals")
    s = ShardedSearcher(lambda q, k: None, 10, exchange="collective")
    s.check()                                    # no peer windows: a no-op
    s.close()


def test_peer_rows_allgather_argument_validation_without_gpu(built_lib):
    lib = _lib.load()
    # header + 2 parities x world x ceil(b/world) rows x dim/2 lines of 16 bytes
    assert lib.vq_peer_rows_window_bytes(8, 1024, 512) >= 256 + 2 * 8 * 128 * 256 * 16
    assert lib.vq_peer_rows_window_bytes(8, 1024, 511) == 0            # odd row length: two floats per line
    one = ctypes.c_void_p(1)

    def call(world=2, rank=0, b_max=64, ld_max=512, b=8, dim=512, slice_=one):
        return lib.vq_peer_allgather_rows(one, world, rank, b_max, ld_max, slice_, b, dim, one, None, None)
    assert call(rank=5) == -1 and b"rank" in lib.vq_last_error()
    assert call(b=65) == -1 and b"window" in lib.vq_last_error()
    assert call(dim=514) == -1
    assert call(dim=511) == -1
    assert call(slice_=None) == -1 and b"slice" in lib.vq_last_error()
    assert call(b=0) == 0


def test_query_slices_cover_the_batch():
    from video_quierer_b200.peer import slice_range
    for b in (0, 1, 5, 37, 1024, 4096):
        for world in (1, 2, 3, 8):
            r = [slice_range(b, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == b
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            per = -(-b // world) if b else 0
            assert all(hi - lo <= per for lo, hi in r)
