"""Times the attention kernels (legacy vs tcgen05) on the three shapes of the step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roboticattack_b200 import _lib
L = _lib.lib()

def t(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

for (B, N, H, hd, causal) in [(8, 288, 32, 128, 1), (8, 261, 16, 64, 0), (8, 256, 16, 72, 0)]:
    D = H * hd
    qkv = torch.randn(B * N, 3 * D, device="cuda").bfloat16()
    do = torch.randn(B * N, D, device="cuda").bfloat16()
    o = torch.empty(B * N, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, H, N, device="cuda")
    delta = torch.empty(B, H, N, device="cuda")
    dqkv = torch.empty_like(qkv)
    fl = 4.0 * B * H * N * N * hd * (0.5 if causal else 1.0)
    for impl in (0, 3):
        L.vla_attention_set_impl(impl)
        f = lambda: _lib.check(L.vla_attention_fwd(_lib.ptr(qkv), _lib.ptr(o), _lib.ptr(lse), None, B, N, H, hd, causal, _lib.cur_stream()))
        bwd = lambda: _lib.check(L.vla_attention_bwd(_lib.ptr(qkv), _lib.ptr(o), _lib.ptr(do), _lib.ptr(lse), _lib.ptr(delta), _lib.ptr(dqkv), None, B, N, H, hd, causal, _lib.cur_stream()))
        tf, tb = t(f), t(bwd)
        print(f"N={N} H={H} hd={hd} causal={causal} impl={impl}: fwd {tf:.1f} us ({fl/tf/1e6:.0f} TFLOP/s)  bwd {tb:.1f} us ({2.5*fl/tb/1e6:.0f} TFLOP/s)")

# optional in-kernel timeline of the backward (build with VLA_NVCC_EXTRA=-DVLA_ATTN_TIMING): %globaltimer stamps of CTA 0
import ctypes
try:
    fn = L.vla_attn_dbg_read
except AttributeError:
    fn = None
if fn is not None:
    buf = (ctypes.c_ulonglong * 128)()
    fn(buf)
    for mode in (0, 1):
        st = [buf[mode * 64 + i] for i in range(64)]
        t0 = st[0]
        print("mode", mode, " ".join(f"{i}:{(v - t0)}" for i, v in enumerate(st) if v))
