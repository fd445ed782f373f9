"""Native raw store format.

The reference persists its indexes as pickles of Python lists / dicts
(video_search_overhaul.py:66-106, src/indexes/hnsw.py:306-380).  Those stay readable and writable
(`save_to_disk` / `load_from_disk`, `save` / `load` of the facades) so existing caches interchange,
but a pickle of 100M per-row arrays is hopeless.  The raw format is a directory

    header.json     {"format": "vq-raw", "version": 1, "kind": ..., "arrays": {name: {file, dtype, shape,
                     bytes, digest}}, "attrs": {...}}
    <name>.bin      one C-contiguous little-endian array per entry, exactly as it sits in HBM
                    (rows.f32 / rows.bf16 = the [n, ld] device matrices, adj0 / upper_adj / ... = the graph)
    objects.pkl     the small host-side Python objects (external ids, metadata dicts), pickled

so that loading is `np.memmap` + chunked host-to-device copies straight into the device matrix (no
per-row Python objects, no re-normalisation), and saving is chunked device-to-host copies into a
memmap.  `digest` is a SHA-256 over the array's size and a strided sample of its bytes (hashing 100 GB
on every load would cost more than the load); a mismatch raises `ValueError` like the reference's
checksum sidecar does (hnsw.py:353-357).

This module is pure numpy (no CUDA): the device copies live in `engine.DeviceStore.save_raw/load_raw`.
"""

from __future__ import annotations

import hashlib
import json
import os
import pickle
from typing import Dict, Iterable, Tuple

import numpy as np

FORMAT, VERSION = "vq-raw", 1
_DTYPES = {"float32": np.float32, "uint16": np.uint16, "int32": np.int32, "int64": np.int64}
_SAMPLE_BLOCK = 1 << 16          # bytes per sampled block of the digest
_SAMPLE_BLOCKS = 256             # blocks per array (first, last and evenly strided in between)


def sampled_digest(buf: np.ndarray) -> str:
    """SHA-256 of (byte length, <= 256 evenly spaced 64 KiB blocks incl. the first and the last)."""
    flat = buf.reshape(-1).view(np.uint8)
    n = flat.shape[0]
    h = hashlib.sha256(str(n).encode())
    if n <= _SAMPLE_BLOCK * _SAMPLE_BLOCKS:
        h.update(flat.tobytes())
        return h.hexdigest()
    last = n - _SAMPLE_BLOCK
    for i in range(_SAMPLE_BLOCKS):
        off = (last * i) // (_SAMPLE_BLOCKS - 1)
        h.update(flat[off: off + _SAMPLE_BLOCK].tobytes())
    return h.hexdigest()


class RawWriter:
    """Create a raw store directory; arrays are filled chunk by chunk through writable memmaps."""

    def __init__(self, path: str, kind: str, attrs: Dict | None = None):
        self.path = path
        os.makedirs(path, exist_ok=True)
        self.header = {"format": FORMAT, "version": VERSION, "kind": kind, "arrays": {}, "attrs": dict(attrs or {})}
        self._open: Dict[str, np.memmap] = {}

    def create(self, name: str, dtype: str, shape: Tuple[int, ...]) -> np.ndarray:
        if dtype not in _DTYPES:
            raise ValueError(f"unsupported dtype {dtype}")
        shape = tuple(int(x) for x in shape)
        fn = f"{name}.bin"
        nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(_DTYPES[dtype]).itemsize
        self.header["arrays"][name] = {"file": fn, "dtype": dtype, "shape": list(shape), "bytes": nbytes}
        full = os.path.join(self.path, fn)
        if nbytes == 0:
            open(full, "wb").close()
            arr = np.zeros(shape, dtype=_DTYPES[dtype])
        else:
            arr = np.memmap(full, dtype=_DTYPES[dtype], mode="w+", shape=shape)
        self._open[name] = arr
        return arr

    def put(self, name: str, array: np.ndarray):
        array = np.ascontiguousarray(array)
        dst = self.create(name, str(array.dtype), array.shape)
        if array.size:
            dst[...] = array

    def put_objects(self, objects):
        with open(os.path.join(self.path, "objects.pkl"), "wb") as f:
            pickle.dump(objects, f, protocol=pickle.HIGHEST_PROTOCOL)
        self.header["objects"] = "objects.pkl"

    def close(self):
        for name, arr in self._open.items():
            if isinstance(arr, np.memmap):
                arr.flush()
            self.header["arrays"][name]["digest"] = sampled_digest(arr)
        self._open.clear()
        tmp = os.path.join(self.path, "header.json.tmp")
        with open(tmp, "w") as f:
            json.dump(self.header, f, indent=1)
        os.replace(tmp, os.path.join(self.path, "header.json"))     # the header appears last: a torn save is not loadable


def open_raw(path: str, kind: str | None = None, verify: bool = True):
    """-> (attrs, {name: read-only memmap}, objects or None).  Raises ValueError on a damaged store."""
    hp = os.path.join(path, "header.json")
    if not os.path.exists(hp):
        raise FileNotFoundError(hp)
    with open(hp) as f:
        header = json.load(f)
    if header.get("format") != FORMAT or header.get("version") != VERSION:
        raise ValueError(f"{path}: not a {FORMAT} v{VERSION} store")
    if kind is not None and header.get("kind") != kind:
        raise ValueError(f"{path}: store kind {header.get('kind')!r}, expected {kind!r}")
    arrays = {}
    for name, meta in header["arrays"].items():
        full = os.path.join(path, meta["file"])
        shape = tuple(meta["shape"])
        dt = _DTYPES[meta["dtype"]]
        if os.path.getsize(full) != meta["bytes"]:
            raise ValueError(f"{full}: size {os.path.getsize(full)} != {meta['bytes']} (truncated store)")
        arr = np.memmap(full, dtype=dt, mode="r", shape=shape) if meta["bytes"] else np.zeros(shape, dtype=dt)
        if verify and sampled_digest(arr) != meta.get("digest"):
            raise ValueError(f"{full}: digest mismatch (corrupted store)")
        arrays[name] = arr
    objects = None
    if header.get("objects"):
        with open(os.path.join(path, header["objects"]), "rb") as f:
            objects = pickle.load(f)
    return header["attrs"], arrays, objects


def chunks(n_rows: int, row_bytes: int, target_bytes: int = 256 << 20) -> Iterable[Tuple[int, int]]:
    """Row ranges of about `target_bytes` each (the staging granularity of the device copies)."""
    step = max(1, target_bytes // max(1, row_bytes))
    for lo in range(0, n_rows, step):
        yield lo, min(n_rows, lo + step)
