"""GPU parity tests of the decode-step kernels (csrc/decode.cu) through the C ABI, against fp32 torch maths with the bf16
rounding points of the eager Llama layer (transformers modeling_llama.py: LlamaRMSNorm, LlamaMLP, apply_rotary_pos_emb,
eager attention with a KV cache).  Outputs are bf16: 1 ulp = 2^-8 relative to the tensor scale; sums over K = 4096..11008
in a different order than torch's move a value by a fraction of that, hence <= 1.01 ulp of the scale for single roundings
and 2 ulp where two roundings stack (residual, SwiGLU)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from roboticattack_b200 import _lib  # noqa: E402

BF16_ULP = 2.0 ** -8
L = None


def setup_module(module):
    global L
    L = _lib.lib()


def close(got, ref, ulps, what):
    got, ref = got.float().cpu(), ref.float().cpu()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert torch.isfinite(got).all(), f"{what}: non-finite output"
    scale = ref.abs().max().item() + 1e-12
    err = (got - ref).abs().max().item()
    assert err <= ulps * BF16_ULP * scale, f"{what}: max err {err:.4g} vs scale {scale:.4g} ({err / scale / BF16_ULP:.2f} ulp)"


def rb(x):
    return x.bfloat16().float()


def gemv(A, W, norm_w=None, eps=1e-6, resid=None, out_f32=False, swiglu=False):
    M, K = A.shape
    N = W.shape[0]
    cols = N // 2 if swiglu else N
    out = torch.full((M, cols), float("nan"), device="cuda", dtype=torch.float32 if out_f32 else torch.bfloat16)
    _lib.check(L.vla_gemv_bf16(_lib.ptr(A), A.stride(0), _lib.ptr(norm_w) if norm_w is not None else None, eps, _lib.ptr(W), W.stride(0),
                               _lib.ptr(out), cols, M, N, K, _lib.ptr(resid) if resid is not None else None,
                               resid.stride(0) if resid is not None else 0, int(out_f32), int(swiglu), _lib.cur_stream()))
    return out


@pytest.mark.parametrize("M", [1, 2, 3, 4])
@pytest.mark.parametrize("N,K", [(4096, 4096), (4096, 11008), (12288, 4096), (1000, 136), (6, 64), (32064, 4096)])
def test_gemv_plain_and_residual(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N + K)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    R = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    ref = A.float() @ W.float().t()
    close(gemv(A, W), ref.bfloat16(), 1.01, f"gemv {M}x{N}x{K}")
    f32 = gemv(A, W, out_f32=True)
    assert torch.equal(f32, rb(f32)), "fp32 output must hold bf16 values (logits are a bf16 tensor upstream)"
    close(f32, rb(ref), 1.01, f"gemv f32 {M}x{N}x{K}")
    close(gemv(A, W, resid=R), rb(R.float() + rb(ref)), 2.0, f"gemv resid {M}x{N}x{K}")


@pytest.mark.parametrize("M", [1, 4])
@pytest.mark.parametrize("N,K", [(12288, 4096), (520, 136), (8, 4096)])
def test_gemv_fused_rmsnorm(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    X = (torch.randn(M, K, device="cuda", generator=g) * 3.0).bfloat16()
    nw = (1.0 + 0.1 * torch.randn(K, device="cuda", generator=g)).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    eps = 1e-6
    x = X.float()
    normed = rb(nw.float() * rb(x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)))   # LlamaRMSNorm: weight * hidden.to(bf16)
    # the row's rstd may differ by an fp32 ulp from torch's (summation order), which can flip single bf16 roundings of the
    # normalised row; the product averages them out
    close(gemv(X, W, norm_w=nw, eps=eps), (normed @ W.float().t()).bfloat16(), 1.5, f"gemv rmsnorm {M}x{N}x{K}")


@pytest.mark.parametrize("M", [1, 3])
@pytest.mark.parametrize("F,K", [(11008, 4096), (128, 72), (192, 4096)])
def test_gemv_swiglu(M, F, K):
    g = torch.Generator(device="cuda").manual_seed(M + F + K)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    Wg = (torch.randn(F, K, device="cuda", generator=g) * 2.0 / K ** 0.5).bfloat16()
    Wu = (torch.randn(F, K, device="cuda", generator=g) * 2.0 / K ** 0.5).bfloat16()
    # the engine's packing: interleaved groups of 64 gate rows and 64 up rows
    W = torch.stack([Wg.view(F // 64, 64, K), Wu.view(F // 64, 64, K)], dim=1).reshape(2 * F, K).contiguous()
    gt, up = rb(A.float() @ Wg.float().t()), rb(A.float() @ Wu.float().t())
    ref = rb(rb(torch.nn.functional.silu(gt)) * up)   # LlamaMLP: act_fn(gate_proj(x)) * up_proj(x), each a bf16 tensor
    close(gemv(A, W, swiglu=True), ref, 3.0, f"gemv swiglu {M}x{F}x{K}")


def rope_tables(Lmax, hd, base=10000.0):
    inv = 1.0 / (base ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))
    ang = torch.arange(Lmax, dtype=torch.float32)[:, None] * inv[None, :]
    return ang.cos().cuda().contiguous(), ang.sin().cuda().contiguous()


def rope_ref(x, cos, sin):
    """x [..., hd] (bf16 values in fp32) at one position; cos/sin [hd/2]; HF rotate_half with bf16 products."""
    half = x.shape[-1] // 2
    c, s = torch.cat([cos, cos]), torch.cat([sin, sin])
    rot = torch.cat([-x[..., half:], x[..., :half]], dim=-1)
    return rb(rb(x * c) + rb(rot * s))


@pytest.mark.parametrize("B,Lc,pos,H,hd", [(1, 300, 281, 32, 128), (2, 64, 0, 4, 128), (3, 40, 39, 3, 64), (1, 50, 17, 2, 32),
                                            (4, 200, 131, 8, 128), (1, 700, 650, 4, 128), (2, 400, 333, 2, 64), (1, 330, 320, 2, 128),
                                            (1, 330, 319, 2, 128)])   # > 320 cached rows: more than one round of loads per warp
def test_attention_decode_fused_rope(B, Lc, pos, H, hd):
    g = torch.Generator(device="cuda").manual_seed(B + Lc + pos)
    qkv = torch.randn(B * Lc, 3 * H * hd, device="cuda", generator=g).bfloat16()
    cos, sin = rope_tables(Lc, hd)
    before = qkv.clone()
    o = torch.full((B, H * hd), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.check(L.vla_attention_decode(_lib.ptr(qkv), _lib.ptr(o), _lib.ptr(cos), _lib.ptr(sin), B, Lc, pos, H, hd, _lib.cur_stream()))
    torch.cuda.synchronize()
    x = before.float().view(B, Lc, 3, H, hd)
    q = rope_ref(x[:, pos, 0], cos[pos], sin[pos])            # [B, H, hd]
    k = x[:, : pos + 1, 1].clone()                             # cached keys are post-RoPE already
    k[:, pos] = rope_ref(x[:, pos, 1], cos[pos], sin[pos])
    v = x[:, : pos + 1, 2]
    # the new row's q and k are rotated in place (bit-exact: same products, same roundings); nothing else is touched
    after = qkv.float().view(B, Lc, 3, H, hd)
    assert torch.equal(after[:, pos, 0], q) and torch.equal(after[:, pos, 1], k[:, pos])
    mask = torch.ones(B, Lc, 3, H, hd, dtype=torch.bool, device="cuda")
    mask[:, pos, :2] = False
    assert torch.equal(after[mask], x[mask])
    s = torch.einsum("bhd,bjhd->bhj", q, k) / hd ** 0.5
    p = rb(torch.softmax(s, dim=-1))
    ref = torch.einsum("bhj,bjhd->bhd", p, v).reshape(B, H * hd)
    close(o, ref.bfloat16(), 1.5, f"attention decode B{B} L{Lc} pos{pos} H{H} hd{hd}")
