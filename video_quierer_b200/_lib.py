"""ctypes binding of libvqsearch.so (the C-ABI declared in include/vq_search.h).

There is deliberately NO fallback: if the shared object is missing or a call fails, an
exception is raised.  The product path never routes through numpy or the oracle.
"""

from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libvqsearch.so")

F32, BF16 = 0, 1
NORM_NONE, NORM_EPS, NORM_PLAIN = 0, 1, 2
SCAN_AUTO, SCAN_FMA, SCAN_MMA, SCAN_FMA32 = 0, 1, 2, 3

_vp, _i32, _i64, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t

# name -> (restype, argtypes); must list every symbol include/vq_search.h declares
SIGNATURES = {
    "vq_abi_version": (_i32, []),
    "vq_last_error": (C.c_char_p, []),
    "vq_last_scan_path": (C.c_char_p, []),
    "vq_last_launch_count": (_i32, []),
    "vq_l2_normalize": (_i32, [_vp, _i64, _i32, _i32, _i32, _vp]),
    "vq_ingest_rows": (_i32, [_vp, _i64, _i32, _i32, _vp, _i32, _i32, _i32, _vp]),
    "vq_scan_workspace_bytes": (_sz, [_i64, _i32, _i32, _i32, _i32, _i32, _i32]),
    "vq_scan_topk": (_i32, [_vp, _i64, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _sz, _i32, _vp]),
    "vq_topk_merge": (_i32, [_vp, _vp, _i32, _i64, _i32, _i32, _vp, _i32, _vp, _vp, _vp]),
    "vq_peer_window_bytes": (_sz, [_i32, _i32, _i32]),
    "vq_peer_window_create": (_i32, [_sz, C.POINTER(C.c_void_p), C.c_char_p]),
    "vq_peer_window_open": (_i32, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "vq_peer_window_close": (_i32, [_vp]),
    "vq_peer_window_destroy": (_i32, [_vp]),
    "vq_peer_window_status": (_i32, [_vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "vq_peer_exchange_merge": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _vp]),
    "vq_peer_rows_window_bytes": (_sz, [_i32, _i32, _i32]),
    "vq_peer_allgather_rows": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _vp, _vp, _vp]),
    "vq_profile_enable": (_i32, [_i32]),
    "vq_profile_last_kernel_ms": (C.c_float, []),
    "vq_rescore_topk": (_i32, [_vp, _i64, _i32, _i32, _vp, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _sz, _vp]),
    "vq_search_two_stage_workspace_bytes": (_sz, [_i64, _i32, _i32, _i32, _i32]),
    "vq_search_two_stage": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _i32, _i32, _i32, _i32, C.c_float,
                                   _vp, _vp, _vp, _vp, _sz, _vp]),
    "vq_search_exact_supported": (_i32, [_i64, _i32, _i32, _i32, _i32]),
    "vq_search_exact_workspace_bytes": (_sz, [_i64, _i32, _i32, _i32, _i32]),
    "vq_search_exact": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "vq_store_bounds": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "vq_search_collect_workspace_bytes": (_sz, [_i64, _i32, _i32, _i32, _i32]),
    "vq_search_collect": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _i32, _i32, _i32, _vp, _i32, _vp,
                                 _vp, _vp, _vp, _vp, _sz, _vp]),
    "vq_hnsw_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "vq_hnsw_search": (_i32, [_vp, _i64, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32,
                              _vp, _i32, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _sz, _vp]),
    "vq_hnsw_layer_workspace_bytes": (_sz, [_i64, _i32, _i32, _i32, _i32, _i32]),
    "vq_hnsw_build_layer": (_i32, [_vp, _i64, _i32, _i32, _i32, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _sz, _vp]),
}


class VQError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared object (once).  Raises ImportError with the build command if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found — build it with `python -m video_quierer_b200.build` "
            "(there is no CPU fallback for the search path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.vq_abi_version() != 1:
        raise ImportError(f"libvqsearch ABI {lib.vq_abi_version()} != 1")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().vq_last_error().decode("utf-8", "replace")
        raise VQError(f"{what} failed (code {rc}): {msg}")


def last_scan_path() -> str:
    return load().vq_last_scan_path().decode()


def last_launch_count() -> int:
    return int(load().vq_last_launch_count())
