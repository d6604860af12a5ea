"""Times every GEMM variant (and cuBLAS) on the ViT / projector shapes of one attack iteration at bs = 8, back to back
with L2-resident operands (the in-step situation: activations come from the previous kernel, weights are 2-9 MB)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roboticattack_b200 import _lib
L = _lib.lib()

SHAPES = [  # (name, M, N, K)
    ("dino qkv", 2088, 3072, 1024), ("dino proj", 2088, 1024, 1024), ("dino fc1", 2088, 4096, 1024), ("dino fc2", 2088, 1024, 4096),
    ("dino d(qkv)", 2088, 1024, 3072),
    ("sig qkv", 2048, 3456, 1152), ("sig proj", 2048, 1152, 1152), ("sig fc1", 2048, 4304, 1152), ("sig fc2", 2048, 1152, 4304),
    ("sig d(qkv)", 2048, 1152, 3456),
    ("proj fc1", 2048, 8704, 2176), ("proj fc2", 2048, 4096, 8704), ("llama o", 2304, 4096, 4096),
]


def t(fn, iters=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


L.vla_gemm_set_autotune(0)
for name, M, N, K in SHAPES:
    A = torch.randn(M, K, device="cuda").bfloat16()
    W = torch.randn(N, K, device="cuda").bfloat16()
    Np = (N + 7) // 8 * 8
    out = torch.empty(M, Np, device="cuda", dtype=torch.bfloat16)
    fl = 2.0 * M * N * K
    row = f"{name:12s} M={M} N={N} K={K}:"
    for ctas, bn in [(1, 128), (1, 256), (2, 128), (2, 256)]:
        _lib.check(L.vla_gemm_set_mode(ctas, bn))
        us = t(lambda: _lib.check(L.vla_gemm_bf16_tn(_lib.ptr(A), K, _lib.ptr(W), K, _lib.ptr(out), Np, M, N, K, None, None, None, 0, 0,
                                                     None, 0, _lib.cur_stream())))
        row += f"  ({ctas},{bn}) {us:6.1f}us {fl / us / 1e6:5.0f}TF"
    _lib.check(L.vla_gemm_set_mode(0, 0))
    o2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    us = t(lambda: torch.matmul(A, W.t(), out=o2))
    row += f"  cuBLAS {us:6.1f}us {fl / us / 1e6:5.0f}TF"
    print(row, flush=True)

# ---- with the in-step epilogues and cold (HBM-resident) weights: 32 weight copies are cycled so that no launch finds its W in L2
print("--- epilogue + cold-weight variants (autotune off, variant pinned) ---")
for name, M, N, K, kind in [("dino fc1 gelu", 2088, 4096, 1024, "gelu"), ("dino fc2 ls+res", 2088, 1024, 4096, "res"),
                            ("dino proj ls+res", 2088, 1024, 1024, "res"), ("dino qkv bias", 2088, 3072, 1024, "bias"),
                            ("sig fc1 gelu", 2048, 4304, 1152, "gelu"), ("sig fc2 res", 2048, 1152, 4304, "res")]:
    NW = 32
    A = torch.randn(M, K, device="cuda").bfloat16()
    Ws = [torch.randn(N, K, device="cuda").bfloat16() for _ in range(NW)]
    Np = (N + 7) // 8 * 8
    out = torch.empty(M, Np, device="cuda", dtype=torch.bfloat16)
    pre = torch.empty(M, Np, device="cuda", dtype=torch.bfloat16)
    bias = torch.randn(N, device="cuda").bfloat16()
    gamma = torch.randn(N, device="cuda").bfloat16()
    resid = torch.randn(M, Np, device="cuda").bfloat16()
    fl = 2.0 * M * N * K
    row = f"{name:17s} M={M} N={N} K={K}:"
    for ctas, bn in [(1, 128), (1, 256), (2, 128), (2, 256)]:
        _lib.check(L.vla_gemm_set_mode(ctas, bn))
        for cold in (False, True):
            it = [0]
            def call():
                W = Ws[it[0] % NW] if cold else Ws[0]
                it[0] += 1
                if kind == "gelu":
                    args = (_lib.ptr(bias), None, None, 0, 1, _lib.ptr(pre))
                elif kind == "res":
                    args = (_lib.ptr(bias), _lib.ptr(gamma), _lib.ptr(resid), Np, 0, None)
                else:
                    args = (_lib.ptr(bias), None, None, 0, 0, None)
                _lib.check(L.vla_gemm_bf16_tn(_lib.ptr(A), K, _lib.ptr(W), K, _lib.ptr(out), Np, M, N, K, *args, 0, _lib.cur_stream()))
            us = t(call, iters=64)
            row += f"  ({ctas},{bn}){'c' if cold else 'w'} {us:5.1f}us"
    _lib.check(L.vla_gemm_set_mode(0, 0))
    print(row, flush=True)
