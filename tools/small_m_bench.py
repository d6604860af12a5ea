"""tcgen05 GEMM at the small-M shapes of the batch-1 attack iteration / predict_action's prefill (M = 289 Llama rows) and of the
supervised rows of the last decoder layer (M = 32): us per launch, weight-streaming rate and useful TFLOP/s, weights rotating over
> 400 MB of copies.  usage (GPU box): python tools/small_m_bench.py [M]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from roboticattack_b200 import _lib

L = _lib.lib()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 289
shapes = [("o", 4096, 4096), ("down", 4096, 11008), ("d(gu).Wgu^T", 4096, 22016), ("d(qkv).Wqkv^T", 4096, 12288), ("qkv", 12288, 4096),
          ("gate|up", 22016, 4096)]
for name, N, K in shapes:
    copies = max(2, int(400e6 // (N * K * 2)) + 1)
    W = [(torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16() for _ in range(copies)]
    A = torch.randn(M, K, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ep = _lib.GemmEpilogue()
    st = _lib.cur_stream()

    def run(i):
        _lib.check(L.vla_gemm_bf16_tn_ex(_lib.ptr(A), K, _lib.ptr(W[i % copies]), K, _lib.ptr(out), N, M, N, K, ctypes.byref(ep), st))
    for i in range(2 * copies):
        run(i)
    torch.cuda.synchronize()
    n = 10 * copies
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    print(f"{name:14s} M={M} N={N:6d} K={K:6d}: {us:7.2f} us  {N * K * 2 / us / 1e6:5.2f} TB/s of weights  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s")
