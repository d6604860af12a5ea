"""Every tcgen05 GEMM kernel variant x every fused epilogue, at the shapes of one OpenVLA-7B attack iteration (bs 8, T 33).

The step's dominant kernel has 4 tile variants -- (CTAs per tile, BLOCK_N) in {(1,128), (1,256), (2,128), (2,256)}; the
autotuner picks one per shape by timing -- and 7 epilogue modes.  Here each variant is pinned with `vla_gemm_set_mode` and
every epilogue is compared with fp32 torch maths that rounds to bf16 exactly where the reference's eager ops materialise a
bf16 tensor (nn.Linear output, rotate_half products, silu(gate), ...; transformers modeling_llama.py apply_rotary_pos_emb /
LlamaMLP, timm Mlp / LayerScale as wired by prismatic/extern/hf/modeling_prismatic.py:78-123,146-158).

Shapes (M = 8 x 288 Llama rows, 8 x 261 DINOv2 rows, 8 x 256 SigLIP rows):
  RoPE        2304 x 12288 x 4096   q|k|v projection
  SwiGLU      2304 x 22016 x 4096   gate|up projection (interleaved [gate 64 | up 64] weight rows)
  SwiGLU-bwd  2304 x 11008 x 4096   d(act) = dX . W_down  -> d(gate|up)
  GELU-bwd    2088 x  4096 x 1024   DINOv2 d(fc2), and the ragged SigLIP width 2048 x 4304 x 1152
  delta       2304 x  4096 x 4096   dO = d(x_mid) . W_o with rowsum(dO * O) per head
  GELU        2048 x  4304 x 1152   SigLIP fc1 (+ bias, saved pre-activation), ragged N
  general     2048 x  1152 x 4304   SigLIP fc2 (+ bias, residual); 2048 x 1024 x 640 patch embed with row remap + pos_embed
  plain       2304 x  4096 x 12288  d(q|k|v) . W_qkv ; fp32 logits 32 x 32064 x 4096
"""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu

from roboticattack_b200 import _lib  # noqa: E402

BF16_ULP = 2.0 ** -8
VARIANTS = [(1, 128), (1, 256), (2, 128), (2, 256)]
L = None


def setup_module(module):
    global L
    L = _lib.lib()


def teardown_module(module):
    _lib.lib().vla_gemm_set_mode(0, 0)


def rbf(x):
    return x.bfloat16().float()


def rand(shape, scale, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).bfloat16()


def close(got, ref, ulps, what, mean_ulps=0.25, mag=None):
    """Element-wise: |got - ref| <= ulps x (the bf16 spacing at the element's own magnitude, 2^-7 |ref|) + fp32 accumulation
    noise (2^-16 of the tensor scale); a result that rounds the other way at a bf16 tie is 1 spacing off.  Where the output is
    a SUM of bf16 terms (rotary products, residual add) a flipped rounding of a term shows at the term's magnitude, not the
    (possibly cancelling) sum's: `mag` = element-wise magnitude of the largest term.  Plus a bound on the mean error in units
    of 2^-8 x the tensor scale."""
    got, ref = got.float(), ref.float()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    assert torch.isfinite(got).all(), f"{what}: non-finite output"
    scale = ref.abs().max().item() + 1e-12
    err = (got - ref).abs()
    m = ref.abs() if mag is None else torch.maximum(ref.abs(), mag.float().abs())
    tol = ulps * (2.0 ** -7 * m + 2.0 ** -16 * scale)
    excess = (err - tol).max().item()
    assert excess <= 0, f"{what}: worst excess over {ulps} bf16 spacings: {excess:.4g} (max err {err.max().item():.4g}, scale {scale:.4g})"
    assert err.mean().item() <= mean_ulps * BF16_ULP * scale, f"{what}: mean err {err.mean().item() / scale / BF16_ULP:.3f} ulp of scale"


def gemm(A, W, out, M, N, K, ldc=None, **ep_fields):
    ep = _lib.GemmEpilogue()
    keep = []
    for k, v in ep_fields.items():
        if isinstance(v, torch.Tensor):
            keep.append(v)
            v = v.data_ptr()
        setattr(ep, k, v)
    _lib.check(L.vla_gemm_bf16_tn_ex(_lib.ptr(A), A.stride(0), _lib.ptr(W), W.stride(0), _lib.ptr(out), ldc or out.stride(0), M, N, K,
                                     ctypes.byref(ep), _lib.cur_stream()), "vla_gemm_bf16_tn_ex")
    torch.cuda.synchronize()


def lin(A, W):
    return A.float() @ W.float().t()


@pytest.fixture(params=VARIANTS, ids=lambda v: f"ctas{v[0]}_n{v[1]}")
def variant(request):
    _lib.check(L.vla_gemm_set_mode(*request.param))
    yield request.param
    L.vla_gemm_set_mode(0, 0)


def test_plain_and_f32_out(variant):
    M, N, K = 2304, 4096, 12288
    A, W = rand((M, K), 0.5, 1), rand((N, K), K ** -0.5, 2)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    gemm(A, W, out, M, N, K)
    close(out, rbf(lin(A, W)), 1.01, "plain d(qkv)")
    M, N, K = 32, 32064, 4096                       # lm_head on the supervised rows, fp32 logits holding bf16 values
    A, W = rand((M, K), 0.5, 3), rand((N, K), K ** -0.5, 4)
    of = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32)
    gemm(A, W, of, M, N, K, out_f32=1)
    assert torch.equal(of, rbf(of))
    close(of, rbf(lin(A, W)), 1.01, "fp32 logits")


def test_rope_epilogue(variant):
    M, N, K, Lseq, heads = 2304, 12288, 4096, 288, 32
    A, W = rand((M, K), 0.5, 5), rand((N, K), K ** -0.5, 6)
    inv = 1.0 / (10000 ** (torch.arange(0, 128, 2, dtype=torch.float32) / 128))
    fr = torch.outer(torch.arange(Lseq, dtype=torch.float32), inv)
    cos, sin = rbf(fr.cos()).cuda().contiguous(), rbf(fr.sin()).cuda().contiguous()
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    gemm(A, W, out, M, N, K, pair_mode=1, rope_cos=cos, rope_sin=sin, rope_L=Lseq, rope_cols=2 * 4096)
    y = rbf(lin(A, W)).view(M, 3, heads, 128)
    pos = torch.arange(M, device="cuda") % Lseq
    c, s = cos[pos][:, None, None, :], sin[pos][:, None, None, :]
    qk = y[:, :2]
    x1, x2 = qk[..., :64], qk[..., 64:]
    r1 = rbf(x1 * c) + rbf(-x2 * s)                 # q * cos + rotate_half(q) * sin, each product a bf16 tensor
    r2 = rbf(x2 * c) + rbf(x1 * s)
    ref = torch.cat([torch.cat([r1, r2], -1), y[:, 2:]], 1).reshape(M, N)
    big = torch.maximum(x1.abs(), x2.abs())
    mag = torch.cat([torch.cat([big, big], -1), y[:, 2:].abs()], 1).reshape(M, N)
    close(out, rbf(ref), 2.0, "rope q|k|v", mag=mag)
    close(out[:, 8192:], rbf(lin(A, W))[:, 8192:], 1.01, "v passes through")


def interleave_gate_up(gate, up):
    """[.., F] x 2 -> [.., 2F] in groups [gate 64 | up 64] (layout of the packed gate|up projection)."""
    F = gate.shape[-1]
    return torch.stack([gate.reshape(*gate.shape[:-1], F // 64, 64), up.reshape(*up.shape[:-1], F // 64, 64)], -2).reshape(*gate.shape[:-1], 2 * F)


def test_swiglu_forward_epilogue(variant):
    M, F, K = 2304, 11008, 4096
    A = rand((M, K), 0.5, 7)
    Wg, Wu = rand((F, K), K ** -0.5, 8), rand((F, K), K ** -0.5, 9)
    W = interleave_gate_up(Wg.t().contiguous(), Wu.t().contiguous()).t().contiguous()      # rows interleaved in groups of 64
    raw = torch.full((M, 2 * F), float("nan"), device="cuda", dtype=torch.bfloat16)
    act = torch.full((M, F), float("nan"), device="cuda", dtype=torch.bfloat16)
    gemm(A, W, raw, M, 2 * F, K, pair_mode=2, act_out=act, ld_act=F)
    g, u = rbf(lin(A, Wg)), rbf(lin(A, Wu))
    close(raw, interleave_gate_up(g, u), 1.01, "raw gate|up")
    gk = raw.float().view(M, F // 64, 2, 64)       # the activation is an exact function of the stored bf16 gate / up
    g2, u2 = gk[:, :, 0].reshape(M, F), gk[:, :, 1].reshape(M, F)
    ref = rbf(rbf(torch.nn.functional.silu(g2)) * u2)
    close(act, ref, 1.5, "silu(gate) * up", mean_ulps=0.05)


def test_swiglu_backward_epilogue(variant):
    M, F, K = 2304, 11008, 4096
    dX, Wd_t = rand((M, K), 0.5, 10), rand((F, K), K ** -0.5, 11)          # d(act) = dX . W_down  (W_down^T stored [F, K])
    g, u = rand((M, F), 1.0, 12), rand((M, F), 1.0, 13)
    aux = interleave_gate_up(g, u).contiguous()
    out = torch.full((M, 2 * F), float("nan"), device="cuda", dtype=torch.bfloat16)
    gemm(dX, Wd_t, out, M, F, K, ldc=2 * F, aux_mode=2, aux=aux, ldaux=2 * F)
    dact = rbf(lin(dX, Wd_t))
    gf, uf = g.float(), u.float()
    sig = torch.sigmoid(gf)
    silu_b = rbf(gf * sig)
    ds = rbf(dact * uf)                              # autograd of bf16(silu) * up: d(silu) is a bf16 tensor
    dgate = rbf(ds * (sig * (1 + gf * (1 - sig))))
    dup = rbf(dact * silu_b)
    close(out, interleave_gate_up(dgate, dup), 2.5, "d(gate|up)", mean_ulps=0.1)


@pytest.mark.parametrize("M,N,K", [(2088, 4096, 1024), (2048, 4304, 1152)])
def test_gelu_backward_epilogue(variant, M, N, K):
    dY, Wt = rand((M, K), 0.5, 14), rand((N, K), K ** -0.5, 15)
    pre = rand((M, N), 1.0, 16)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    gemm(dY, Wt, out, M, N, K, aux_mode=1, aux=pre, ldaux=N)
    x = pre.float()
    dg = 0.5 * (1 + torch.erf(x / 2 ** 0.5)) + x * torch.exp(-0.5 * x * x) / (2 * torch.pi) ** 0.5
    close(out, rbf(rbf(lin(dY, Wt)) * dg), 2.0, f"gelu backward {M}x{N}x{K}", mean_ulps=0.1)


def test_delta_epilogue(variant):
    M, N, K, Lseq = 2304, 4096, 4096, 288
    dXm, Wo_t = rand((M, K), 0.5, 17), rand((N, K), K ** -0.5, 18)
    O = rand((M, N), 0.7, 19)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    delta = torch.full((M // Lseq, N // 128, Lseq), float("nan"), device="cuda", dtype=torch.float32)
    gemm(dXm, Wo_t, out, M, N, K, aux=O, ldaux=N, delta_out=delta, delta_L=Lseq)
    dO = rbf(lin(dXm, Wo_t))
    close(out, dO, 1.01, "dO")
    ref = (out.float() * O.float()).view(M // Lseq, Lseq, N // 128, 128).sum(-1).permute(0, 2, 1)   # exact function of the stored dO
    torch.testing.assert_close(delta, ref.contiguous(), rtol=1e-4, atol=1e-4)


def test_gelu_forward_ragged(variant):
    M, N, K = 2048, 4304, 1152                      # SigLIP fc1
    A, W, b = rand((M, K), 0.5, 20), rand((N, K), K ** -0.5, 21), rand((N,), 0.2, 22)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    pre = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    gemm(A, W, out, M, N, K, bias=b, act=1, preact_out=pre)
    close(pre, rbf(lin(A, W) + b.float()), 1.01, "fc1 pre-activation")
    close(out, rbf(torch.nn.functional.gelu(pre.float())), 1.01, "gelu(fc1)", mean_ulps=0.05)


def test_general_epilogue_residual_layerscale_and_row_remap(variant):
    M, N, K = 2048, 1152, 4304                      # SigLIP fc2 + bias + residual (K ragged: 4304 = 67 * 64 + 16)
    A, W, b = rand((M, K), 0.5, 23), rand((N, K), K ** -0.5, 24), rand((N,), 0.2, 25)
    r, gm = rand((M, N), 1.0, 26), rand((N,), 1.0, 27)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    gemm(A, W, out, M, N, K, bias=b, gamma=gm, resid=r, ldr=N)
    x = rbf(lin(A, W) + b.float())
    x = rbf(x * gm.float())
    close(out, rbf(r.float() + x), 1.5, "fc2 + LayerScale + residual", mag=torch.maximum(r.float().abs(), x.abs()))
    # patch embed: conv-as-GEMM rows [8 x 256] written past the 5 prefix tokens of each sample, + pos_embed broadcast over the batch
    Bn, P, npre, d, Kp = 8, 256, 5, 1024, 640
    A, W, b = rand((Bn * P, Kp), 0.5, 28), rand((d, Kp), Kp ** -0.5, 29), rand((d,), 0.2, 30)
    pos = rand((P, d), 0.5, 31)
    tok = torch.zeros(Bn * (P + npre), d, device="cuda", dtype=torch.bfloat16)
    gemm(A, W, tok, Bn * P, d, Kp, bias=b, resid=pos, ldr=d, resid_mod=P, out_group=P, out_stride=P + npre, out_offset=npre)
    ref = rbf(pos.float()[None] + rbf(lin(A, W) + b.float()).view(Bn, P, d))
    got = tok.view(Bn, P + npre, d)
    assert got[:, :npre].abs().max().item() == 0, "prefix-token rows must not be written"
    close(got[:, npre:], ref, 1.5, "patch embed + pos_embed, remapped rows",
          mag=torch.maximum(pos.float().abs()[None].expand_as(ref), rbf(lin(A, W) + b.float()).view(Bn, P, d).abs()))


@pytest.mark.parametrize("block_n", [64, 32])
@pytest.mark.parametrize("M", [1, 8, 32, 100])
def test_narrow_tile_variants_for_small_m(block_n, M):
    """The 64- / 32-wide single-CTA tiles of the M = batch GEMMs (greedy decode, supervised rows of the last decoder layer):
    plain, general (bias + residual, fp32 logits) and the cache-row remap of the decode's q|k|v projection."""
    _lib.check(L.vla_gemm_set_mode(1, block_n))
    try:
        K, N = 4096, 4096
        A, W = rand((M, K), 0.5, 40 + M), rand((N, K), K ** -0.5, 41)
        out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
        gemm(A, W, out, M, N, K)
        close(out, rbf(lin(A, W)), 1.01, f"plain M={M}")
        b, r = rand((N,), 0.2, 42), rand((M, N), 1.0, 43)
        gemm(A, W, out, M, N, K, bias=b, resid=r, ldr=N)
        x = rbf(lin(A, W) + b.float())
        close(out, rbf(r.float() + x), 1.5, f"bias + residual M={M}", mag=torch.maximum(r.float().abs(), x.abs()))
        V = 32064
        Wv = rand((V, K), K ** -0.5, 44)
        of = torch.full((M, V), float("nan"), device="cuda", dtype=torch.float32)
        gemm(A, Wv, of, M, V, K, out_f32=1)
        assert torch.equal(of, rbf(of))
        close(of, rbf(lin(A, Wv)), 1.01, f"fp32 logits M={M}")
        # decode: row r of the result lands in cache row r * Lc + pos
        Lc, pos, N3 = 40, 17, 3 * 1024
        W3 = rand((N3, K), K ** -0.5, 45)
        cache = torch.zeros(M * Lc, N3, device="cuda", dtype=torch.bfloat16)
        gemm(A, W3, cache, M, N3, K, out_group=1, out_stride=Lc, out_offset=pos)
        got = cache.view(M, Lc, N3)
        close(got[:, pos], rbf(lin(A, W3)), 1.01, f"cache-row remap M={M}")
        assert got[:, :pos].abs().max().item() == 0 and got[:, pos + 1:].abs().max().item() == 0
    finally:
        L.vla_gemm_set_mode(0, 0)


@pytest.mark.parametrize("kind", ["plain", "resid", "rope", "dual_shapes"])
def test_weight_tiles_requested_before_the_wait(variant, kind):
    """`w_constant`: the producer warp requests the W tiles of the first pipeline fill before griddepcontrol.wait (they do not
    depend on the previous kernel) and adds the A tiles after it.  Same tiles, same accumulation order: results are
    bit-identical to the plain launch, for one tile per CTA (short K, fewer k-blocks than ring slots) and for many."""
    shapes = {"plain": [(2304, 4096, 4096), (300, 512, 192), (128, 256, 64)], "resid": [(2048, 1152, 4304)], "rope": [(2304, 12288, 4096)],
              "dual_shapes": [(2088, 1024, 1024), (32, 32064, 4096)]}[kind]
    for (M, N, K) in shapes:
        A, W = rand((M, K), 0.5, 90), rand((N, K), K ** -0.5, 91)
        kw = {}
        if kind == "resid":
            kw = dict(bias=rand((N,), 0.2, 92), resid=rand((M, N), 1.0, 93), ldr=N)
        if kind == "rope":
            inv = 1.0 / (10000 ** (torch.arange(0, 128, 2, dtype=torch.float32) / 128))
            fr = torch.outer(torch.arange(288, dtype=torch.float32), inv)
            kw = dict(pair_mode=1, rope_cos=rbf(fr.cos()).cuda().contiguous(), rope_sin=rbf(fr.sin()).cuda().contiguous(), rope_L=288, rope_cols=8192)
        torch.cuda.synchronize()                      # W really is a constant of the stream by now
        base = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
        gemm(A, W, base, M, N, K, **kw)
        for _ in range(3):
            # a kernel right before the GEMM keeps the device busy while the GEMM's CTAs start and prefetch
            busy = torch.randn(1 << 24, device="cuda").sin_()
            got = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
            gemm(A, W, got, M, N, K, w_constant=1, **kw)
            assert torch.equal(got, base), f"{kind} {M}x{N}x{K}: prefetching W changed the result"
        del busy
