"""Forward-attention check of one implementation against torch on a list of shapes (bring-up tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roboticattack_b200 import _lib
L = _lib.lib()
impl = int(sys.argv[1]) if len(sys.argv) > 1 else 1
_lib.check(L.vla_attention_set_impl(impl))
for (B, N, H, hd, causal) in [(1, 64, 1, 128, 0), (1, 64, 1, 64, 0), (1, 128, 1, 64, 0), (1, 192, 1, 64, 0), (1, 261, 1, 64, 0), (2, 261, 4, 64, 0),
                              (1, 261, 1, 128, 0), (1, 288, 1, 128, 1), (8, 288, 32, 128, 1), (8, 261, 16, 64, 0), (8, 256, 16, 72, 0)]:
    D = H * hd
    g = torch.Generator(device="cuda").manual_seed(N + hd)
    qkv = torch.randn(B * N, 3 * D, device="cuda", generator=g).bfloat16()
    o = torch.zeros(B * N, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(B, H, N, device="cuda")
    _lib.check(L.vla_attention_fwd(_lib.ptr(qkv), _lib.ptr(o), _lib.ptr(lse), None, B, N, H, hd, causal, _lib.cur_stream()))
    torch.cuda.synchronize()
    x = qkv.float().view(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)
    s = (x[0] @ x[1].transpose(-1, -2)) * hd ** -0.5
    if causal:
        s = s.masked_fill(~torch.ones(N, N, dtype=torch.bool, device="cuda").tril(), float("-inf"))
    ref = (torch.softmax(s, -1) @ x[2]).transpose(1, 2).reshape(B * N, D)
    err = (o.float() - ref).abs()
    rows = err.view(B, N, H, hd).amax(dim=(0, 2, 3))
    bad = (rows > 0.02).nonzero().flatten().tolist()
    print(f"B={B} N={N} H={H} hd={hd} causal={causal}: max err {err.max().item():.4f}  lse err {(lse - torch.logsumexp(s, -1)).abs().max().item():.4f}"
          f"  bad rows: {bad[:6]}...{bad[-3:] if bad else ''} ({len(bad)})", flush=True)
