"""CPU tests of the raw store container (rawstore.py): layout, round trip, damage detection."""
import json
import os

import numpy as np
import pytest

from video_quierer_b200 import rawstore


def _write(path):
    w = rawstore.RawWriter(path, "flat", {"n": 5, "dim": 3})
    rows = w.create("rows_f32", "float32", (5, 64))
    rows[:] = np.arange(5 * 64, dtype=np.float32).reshape(5, 64)
    w.put("rows_bf16", (np.arange(5 * 64) % 65536).astype(np.uint16).reshape(5, 64))
    w.put("empty", np.zeros((0, 16), np.int32))
    w.put_objects({"metadata": [{"video_name": "a.mp4"}] * 5})
    w.close()


def test_roundtrip_and_layout(tmp_path):
    p = str(tmp_path / "store")
    _write(p)
    hdr = json.load(open(os.path.join(p, "header.json")))
    assert hdr["format"] == "vq-raw" and hdr["kind"] == "flat" and set(hdr["arrays"]) == {"rows_f32", "rows_bf16", "empty"}
    assert os.path.getsize(os.path.join(p, "rows_f32.bin")) == 5 * 64 * 4      # exactly the HBM image
    attrs, arrays, objects = rawstore.open_raw(p, "flat")
    assert attrs == {"n": 5, "dim": 3}
    assert isinstance(arrays["rows_f32"], np.memmap) and arrays["rows_f32"].shape == (5, 64)
    assert float(arrays["rows_f32"][4, 63]) == 5 * 64 - 1 and arrays["rows_bf16"].dtype == np.uint16
    assert arrays["empty"].shape == (0, 16) and len(objects["metadata"]) == 5
    with pytest.raises(ValueError):
        rawstore.open_raw(p, "hnsw")                                           # wrong kind


def test_damage_is_detected(tmp_path):
    p = str(tmp_path / "store")
    _write(p)
    with open(os.path.join(p, "rows_f32.bin"), "r+b") as f:
        f.seek(100); f.write(b"\xff\xff\xff\xff")
    with pytest.raises(ValueError, match="digest"):
        rawstore.open_raw(p)
    rawstore.open_raw(p, verify=False)                                         # explicit opt-out still loads
    with open(os.path.join(p, "rows_bf16.bin"), "ab") as f:
        f.write(b"\0")
    with pytest.raises(ValueError, match="truncated|size"):
        rawstore.open_raw(p, verify=False)
    with pytest.raises(FileNotFoundError):
        rawstore.open_raw(str(tmp_path / "missing"))


def test_sampled_digest_covers_both_ends_of_large_arrays():
    a = np.zeros(40 << 20, np.uint8)
    d0 = rawstore.sampled_digest(a)
    a[-1] = 1
    assert rawstore.sampled_digest(a) != d0
    a[-1] = 0; a[0] = 1
    assert rawstore.sampled_digest(a) != d0
    assert list(rawstore.chunks(10, 4, target_bytes=16)) == [(0, 4), (4, 8), (8, 10)]


def test_lazy_rows_view():
    from video_quierer_b200.flat_index import LazyRows
    f = np.arange(12, dtype=np.float32).reshape(3, 4)
    lr = LazyRows(f, None, dim=3)
    assert len(lr) == 3 and bool(lr) and lr[1].tolist() == [4.0, 5.0, 6.0] and lr[-1].tolist() == [8.0, 9.0, 10.0]
    lr.append(np.ones(3, np.float32))
    assert len(lr) == 4 and lr[3].tolist() == [1.0, 1.0, 1.0] and len(list(lr)) == 4
    with pytest.raises(NotImplementedError):
        lr.pop(0)
    b = (np.array([[1.5, -2.0]], np.float32).view(np.uint32) >> 16).astype(np.uint16)
    assert LazyRows(None, b, dim=2)[0].tolist() == [1.5, -2.0]                 # bf16 widened exactly
