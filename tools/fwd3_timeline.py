"""In-kernel timeline of the tc3 forward (build with VLA_NVCC_EXTRA=-DVLA_ATTN_TIMING): stamps of CTA 0, softmax thread 0."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roboticattack_b200 import _lib
L = _lib.lib()
B, N, H, hd, causal = 8, 288, 32, 128, 1
D = H * hd
qkv = torch.randn(B * N, 3 * D, device="cuda").bfloat16()
o = torch.empty(B * N, D, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(B, H, N, device="cuda")
_lib.check(L.vla_attention_set_impl(3))
for _ in range(3):
    _lib.check(L.vla_attention_fwd(_lib.ptr(qkv), _lib.ptr(o), _lib.ptr(lse), None, B, N, H, hd, causal, _lib.cur_stream()))
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 64)()
L.vla_attn_fwd3_dbg_read(buf)
t0 = buf[0]
for i in range(7):
    row = [buf[i * 8 + k] - t0 for k in range(6)]
    print("item", i, "start %d | pass0 end +%d | exchange +%d | pass1 end +%d | bar_out +%d | epilogue end +%d" % (row[0], row[1] - row[0], row[2] - row[1], row[3] - row[2], row[4] - row[3], row[5] - row[4]))
