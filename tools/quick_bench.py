#!/usr/bin/env python3
"""Developer micro-benchmark of the scan paths (CUDA events on the launch stream).
Not the contract bench (that is /bench.py) — used to tune kernels between rounds."""
import argparse
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from video_quierer_b200 import _lib, engine


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--batches", default="1,2,4,8,16,32")
    ap.add_argument("--dtypes", default="fp32,bf16")
    ap.add_argument("--paths", default="fma")
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn((a.n, a.dim), device=dev, generator=g)
    x = x / x.norm(dim=1, keepdim=True)
    st = engine.DeviceStore(a.dim, dev, keep_fp32=True, keep_bf16=True)
    st.append(x)
    del x
    sc = engine.Scanner(dev)
    lib = _lib.load()
    lib.vq_profile_enable(1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for dt in a.dtypes.split(","):
        mat = st.view(dt)
        bytes_pass = st.n * st.ld * (4 if dt == "fp32" else 2)
        for path in a.paths.split(","):
            for b in [int(v) for v in a.batches.split(",")]:
                q = torch.randn((b, a.dim), device=dev, generator=g)
                try:
                    for _ in range(3):
                        sc.scan(mat, st.n, st.dim, q, a.k, _lib.NORM_EPS, path)
                except Exception as e:  # noqa: BLE001
                    print(json.dumps({"dtype": dt, "path": path, "b": b, "error": str(e)[:200]}))
                    continue
                torch.cuda.synchronize()
                ts = []
                ks = []
                for _ in range(a.iters):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    sc.scan(mat, st.n, st.dim, q, a.k, _lib.NORM_EPS, path)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                    ks.append(lib.vq_profile_last_kernel_ms())
                ts.sort()
                med = ts[len(ts) // 2]
                print(json.dumps({"dtype": dt, "path": sc.last_path, "b": b, "ms_med": round(med, 4),
                                  "ms_min": round(ts[0], 4), "kernel_ms": round(sorted(ks)[len(ks) // 2], 4), "GBps_kernel": round(bytes_pass / sorted(ks)[len(ks) // 2] / 1e6, 1),
                                  "qps": round(b / med * 1e3, 1), "launches": sc.last_launches,
                                  "tflops": round(2.0 * b * st.n * a.dim / med / 1e9, 2)}), flush=True)


if __name__ == "__main__":
    main()
