"""Runs the attention forward + backward of one shape a few times (ncu target).  usage: one_attn.py B N H hd causal impl"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roboticattack_b200 import _lib
L = _lib.lib()
B, N, H, hd, causal, impl = (int(v) for v in sys.argv[1:7])
D = H * hd
qkv = torch.randn(B * N, 3 * D, device="cuda").bfloat16()
do = torch.randn(B * N, D, device="cuda").bfloat16()
o = torch.empty(B * N, D, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(B, H, N, device="cuda")
delta = torch.empty(B, H, N, device="cuda")
dqkv = torch.empty_like(qkv)
_lib.check(L.vla_attention_set_impl(impl))
for _ in range(3):
    _lib.check(L.vla_attention_fwd(_lib.ptr(qkv), _lib.ptr(o), _lib.ptr(lse), None, B, N, H, hd, causal, _lib.cur_stream()))
    _lib.check(L.vla_attention_bwd(_lib.ptr(qkv), _lib.ptr(o), _lib.ptr(do), _lib.ptr(lse), _lib.ptr(delta), _lib.ptr(dqkv), None, B, N, H, hd,
                                   causal, _lib.cur_stream()))
torch.cuda.synchronize()
