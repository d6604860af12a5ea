"""HBM throughput of the decode projections (csrc/decode.cu gemv_kernel) at OpenVLA-7B shapes, M = 1.
Weights rotate over enough copies to exceed the 126 MB L2; launches are back to back on one stream (as in a decode step, so
the programmatic-dependent-launch prologue overlap is part of the number).  usage (GPU box): python tools/gemv_bench.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from roboticattack_b200 import _lib

L = _lib.lib()
M = int(os.environ.get("GEMV_M", "1"))
shapes = [("qkv+norm", 12288, 4096, True, False), ("o+resid", 4096, 4096, False, False), ("gate|up+norm+swiglu", 22016, 4096, True, True),
          ("down+resid", 4096, 11008, False, False), ("lm_head+norm", 32064, 4096, True, False)]
for name, N, K, norm, swiglu in shapes:
    copies = max(2, int(400e6 // (N * K * 2)) + 1)
    W = [(torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16() for _ in range(copies)]
    A = torch.randn(M, K, device="cuda").bfloat16()
    nw = torch.ones(K, device="cuda", dtype=torch.bfloat16)
    cols = N // 2 if swiglu else N
    out = torch.empty(M, cols, device="cuda", dtype=torch.bfloat16)
    R = torch.zeros(M, cols, device="cuda", dtype=torch.bfloat16)
    resid = None if (norm or swiglu) else R
    st = _lib.cur_stream()

    def run(i):
        _lib.check(L.vla_gemv_bf16(_lib.ptr(A), K, _lib.ptr(nw) if norm else None, 1e-6, _lib.ptr(W[i % copies]), K, _lib.ptr(out), cols, M, N, K,
                                   _lib.ptr(resid) if resid is not None else None, cols, 0, int(swiglu), st))
    for i in range(2 * copies):
        run(i)
    torch.cuda.synchronize()
    n = 20 * copies
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    print(f"{name:22s} M={M} N={N:6d} K={K:6d}: {us:7.2f} us  {N * K * 2 / us / 1e6:6.2f} TB/s  ({copies} weight copies)")
