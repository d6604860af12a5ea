"""Times the norm kernels on the shapes of the step (rotating over 16 buffer sets so that inputs come from HBM)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roboticattack_b200 import _lib
L = _lib.lib()
NB = 16

def t(fn, iters=64):
    for i in range(8): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

for name, M, d in [("llama rms", 2304, 4096), ("dino ln", 2088, 1024), ("siglip ln", 2048, 1152)]:
    xs = [torch.randn(M, d, device="cuda").bfloat16() for _ in range(NB)]
    dys = [torch.randn(M, d, device="cuda").bfloat16() for _ in range(NB)]
    drs = [torch.randn(M, d, device="cuda").bfloat16() for _ in range(NB)]
    w = torch.randn(d, device="cuda").bfloat16(); b = torch.randn(d, device="cuda").bfloat16()
    y = torch.empty(M, d, device="cuda", dtype=torch.bfloat16)
    mean = torch.empty(M, device="cuda"); rstd = torch.ones(M, device="cuda")
    S = _lib.cur_stream
    if "rms" in name:
        f = t(lambda i: _lib.check(L.vla_rmsnorm_fwd(_lib.ptr(xs[i % NB]), _lib.ptr(w), _lib.ptr(y), _lib.ptr(rstd), M, d, 1e-6, S())))
        bw = t(lambda i: _lib.check(L.vla_rmsnorm_bwd(_lib.ptr(dys[i % NB]), _lib.ptr(xs[i % NB]), _lib.ptr(w), _lib.ptr(rstd), _lib.ptr(drs[i % NB]), _lib.ptr(y), M, d, S())))
    else:
        f = t(lambda i: _lib.check(L.vla_layernorm_fwd(_lib.ptr(xs[i % NB]), _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), _lib.ptr(mean), _lib.ptr(rstd), M, d, 1e-6, S())))
        bw = t(lambda i: _lib.check(L.vla_layernorm_bwd(_lib.ptr(dys[i % NB]), _lib.ptr(xs[i % NB]), _lib.ptr(w), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(drs[i % NB]), _lib.ptr(y), M, d, S())))
    mb = M * d * 2 / 1e6
    print(f"{name:10s} M={M} d={d}: fwd {f:5.1f} us ({2 * mb / f / 1e3 * 1e3:5.0f} GB/s)   bwd {bw:5.1f} us ({4 * mb / bw / 1e3 * 1e3:5.0f} GB/s)")
