#!/usr/bin/env python3
"""Bring-up check of the tcgen05 scan path against a torch fp32 reference of the same bf16 data.
Run under `timeout` on the GPU box: a mis-programmed mbarrier hangs instead of failing."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_quierer_b200 import _lib, engine


def check(n, dim, b, k, seed=0):
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn((n, dim), device=dev, generator=g)
    x = x / x.norm(dim=1, keepdim=True)
    q = torch.randn((b, dim), device=dev, generator=g)
    st = engine.DeviceStore(dim, dev, keep_fp32=False, keep_bf16=True)
    st.append(x)
    sc = engine.Scanner(dev)
    # queries pre-normalised and rounded to bf16 on the host side of the call, so the kernel's
    # fp32->bf16 conversion is exact and the fp64 reference sees identical operands
    qn = (q / (q.norm(dim=1, keepdim=True) + 1e-10)).to(torch.bfloat16).to(torch.float32)
    s, r = sc.scan(st.view("bf16"), st.n, st.dim, qn, k, _lib.NORM_NONE, "mma")
    torch.cuda.synchronize()
    xb = st.view("bf16")[:, :dim].to(torch.float32)
    ref = (qn.double() @ xb.double().T)
    ref32 = qn @ xb.T
    print('  torch fp32 matmul vs fp64 max rel:', ((ref32 - ref).abs() / ref.abs().clamp_min(1e-3)).max().item())
    rs, rr = torch.topk(ref, min(k, n), dim=1)
    kk = min(k, n)
    ids_ok = (r[:, :kk].long() == rr).float().mean().item()
    err = ((s[:, :kk].double() - rs).abs() / rs.abs().clamp_min(1e-6)).max().item()
    print(json.dumps({"n": n, "dim": dim, "b": b, "k": k, "path": sc.last_path, "id_match": round(ids_ok, 4),
                      "max_rel_err": err}), flush=True)
    return ids_ok > 0.99 and err < 1e-4


if __name__ == "__main__":
    ok = True
    for (n, dim, b, k) in [(128, 64, 1, 1), (256, 64, 3, 4), (1000, 128, 17, 10), (5000, 512, 32, 10),
                           (20000, 512, 130, 10), (100000, 512, 1024, 10), (4097, 256, 5, 32)]:
        ok &= check(n, dim, b, k)
    print("MMA_CHECK", "PASS" if ok else "FAIL")
