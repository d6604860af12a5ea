"""Launches one GEMM shape with a chosen epilogue a few times (ncu target).  usage: one_gemm.py M N K kind ctas bn"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roboticattack_b200 import _lib
L = _lib.lib()
M, N, K = (int(v) for v in sys.argv[1:4])
kind = sys.argv[4]
ctas, bn = int(sys.argv[5]), int(sys.argv[6])
L.vla_gemm_set_autotune(0)
_lib.check(L.vla_gemm_set_mode(ctas, bn))
A = torch.randn(M, K, device="cuda").bfloat16()
W = torch.randn(N, K, device="cuda").bfloat16()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
pre = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
bias = torch.randn(N, device="cuda").bfloat16()
gamma = torch.randn(N, device="cuda").bfloat16()
resid = torch.randn(M, N, device="cuda").bfloat16()
args = {"plain": (None, None, None, 0, 0, None), "bias": (_lib.ptr(bias), None, None, 0, 0, None),
        "gelu": (_lib.ptr(bias), None, None, 0, 1, _lib.ptr(pre)),
        "res": (_lib.ptr(bias), _lib.ptr(gamma), _lib.ptr(resid), N, 0, None)}[kind]
for _ in range(4):
    _lib.check(L.vla_gemm_bf16_tn(_lib.ptr(A), K, _lib.ptr(W), K, _lib.ptr(out), N, M, N, K, *args, 0, _lib.cur_stream()))
torch.cuda.synchronize()
