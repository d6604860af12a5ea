"""Shapes of the models on the hot path and the constants of the attack front end.

Reference: ``prismatic/extern/hf/configuration_prismatic.py:15-45,72-140`` (model ids, image size, LLM dims),
``VLAAttacker/white_patch/UADA.py:56-57`` (normalisation constants), ``prismatic/vla/action_tokenizer.py:28-36``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Tuple

IGNORE_INDEX = -100
PAD_TOKEN_ID = 32000
EOS_TOKEN_ID = 2
VOCAB_TOKENIZER = 32000          # LlamaTokenizer vocab (action ids are the last 256 of it)
N_ACTION_BINS = 256
ACTION_TOKEN_BEGIN_IDX = VOCAB_TOKENIZER - (N_ACTION_BINS + 1)   # 31743
ACTION_LOGIT_LO = 31744          # logits[..., 31744:32000] are the 256 action classes (UADA.py:384)
ACTION_LOGIT_HI = 32000
ZERO_ACTION_TOKEN = 31872

# index 0 = DINOv2 (ImageNet stats rounded to bf16), index 1 = SigLIP  (UADA.py:56-57)
NORM_MEAN = ((0.484375, 0.455078125, 0.40625), (0.5, 0.5, 0.5))
NORM_STD = ((0.228515625, 0.2236328125, 0.224609375), (0.5, 0.5, 0.5))


@dataclass(frozen=True)
class ViTConfig:
    dim: int
    depth: int
    heads: int
    mlp_hidden: int
    num_prefix: int          # cls + register tokens prepended after the position embedding
    layerscale: bool
    img: int = 224
    patch: int = 14
    ln_eps: float = 1e-6

    @property
    def head_dim(self) -> int:
        return self.dim // self.heads

    @property
    def grid(self) -> int:
        return self.img // self.patch

    @property
    def num_patches(self) -> int:
        return self.grid * self.grid

    @property
    def tokens(self) -> int:
        return self.num_patches + self.num_prefix

    @property
    def blocks_used(self) -> int:
        """get_intermediate_layers(n={depth-2}) takes the output of block index depth-2 (modeling_prismatic.py:85-87)."""
        return self.depth - 1


@dataclass(frozen=True)
class LlamaConfig:
    hidden: int
    layers: int
    heads: int
    ffn: int
    vocab: int = 32064
    rms_eps: float = 1e-6
    rope_theta: float = 10000.0

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads


@dataclass(frozen=True)
class OpenVLAConfig:
    dino: ViTConfig
    siglip: ViTConfig
    llm: LlamaConfig
    name: str = "custom"

    @property
    def img(self) -> int:
        return self.dino.img

    @property
    def num_patches(self) -> int:
        return self.dino.num_patches

    @property
    def vision_dim(self) -> int:
        return self.dino.dim + self.siglip.dim

    @property
    def proj_hidden(self) -> int:
        return 4 * self.vision_dim


def openvla_7b() -> OpenVLAConfig:
    """OpenVLA-7B: DINOv2-L/14 reg4 + SigLIP-so400m/14 @224, Llama-2-7B."""
    return OpenVLAConfig(
        dino=ViTConfig(dim=1024, depth=24, heads=16, mlp_hidden=4096, num_prefix=5, layerscale=True),
        siglip=ViTConfig(dim=1152, depth=27, heads=16, mlp_hidden=4304, num_prefix=0, layerscale=False),
        llm=LlamaConfig(hidden=4096, layers=32, heads=32, ffn=11008),
        name="openvla-7b",
    )


def tiny(img: int = 56, llm_layers: int = 2, vit_depth: int = 3) -> OpenVLAConfig:
    """Same structure, toy widths: same head dims (64 / 72 / 128) so every kernel variant is exercised."""
    return OpenVLAConfig(
        dino=ViTConfig(dim=128, depth=vit_depth, heads=2, mlp_hidden=512, num_prefix=5, layerscale=True, img=img),
        siglip=ViTConfig(dim=144, depth=vit_depth, heads=2, mlp_hidden=536, num_prefix=0, layerscale=False, img=img),
        llm=LlamaConfig(hidden=256, layers=llm_layers, heads=2, ffn=704),
        name=f"tiny-{img}",
    )


def flops_per_sample(cfg: OpenVLAConfig, text_len: int, supervised_rows: int = 8) -> dict:
    """Algorithmic FLOPs of one attack iteration for one sample:
    F_iter = 2*F_lin + 3*F_att (forward + input-gradient for linears, forward + 2x for attention; no dW)."""
    def vit(c: ViTConfig):
        n = c.tokens
        lin = c.blocks_used * 2 * n * (4 * c.dim * c.dim + 2 * c.dim * c.mlp_hidden)
        att = c.blocks_used * 4 * n * n * c.dim
        pe = 2 * c.num_patches * (3 * c.patch * c.patch) * c.dim
        return lin + pe, att
    dl, da = vit(cfg.dino)
    sl, sa = vit(cfg.siglip)
    vd, ph, hd = cfg.vision_dim, cfg.proj_hidden, cfg.llm.hidden
    proj = 2 * cfg.num_patches * (vd * ph + ph * hd + hd * hd)
    L = text_len + cfg.num_patches
    llm_lin = cfg.llm.layers * 2 * L * (4 * hd * hd + 3 * hd * cfg.llm.ffn)
    llm_att = cfg.llm.layers * 4 * L * L * hd / 2
    head = 2 * supervised_rows * hd * cfg.llm.vocab
    f_lin = dl + sl + proj + llm_lin + head
    f_att = da + sa + llm_att
    # what the engine executes: the last decoder layer's o_proj + MLP run on the supervised rows only (engine.cu, exact)
    saved = 2 * (L - supervised_rows) * (hd * hd + 3 * hd * cfg.llm.ffn)
    return {"f_lin": f_lin, "f_att": f_att, "fwd": f_lin + f_att, "iter": 2 * f_lin + 3 * f_att,
            "iter_executed": 2 * (f_lin - saved) + 3 * f_att}


def gemm_shape_table(cfg: OpenVLAConfig, batch: int, text_len: int, supervised_rows: int):
    """The tcgen05 GEMM launches of one attack iteration as the engine issues them (csrc/engine.cu), as tuples
    (count, M, N, K, extra bf16 elements read by the fused epilogue, extra bf16 elements written, bytes per output element).
    Used for the roofline's algorithmic bytes (every operand and every result once) and FLOPs."""
    rows = []
    B, P = batch, cfg.num_patches
    kpad = -(-3 * cfg.dino.patch ** 2 // 64) * 64
    for v in (cfg.dino, cfg.siglip):
        Mv, d, m, n = B * v.tokens, v.dim, v.mlp_hidden, v.blocks_used
        rows += [(1, B * P, d, kpad, P * d, 0, 2), (1, B * P, kpad, d, 0, 0, 2)]                     # patch embed fwd (+pos_embed) / bwd
        rows += [(n, Mv, 3 * d, d, 0, 0, 2), (n, Mv, d, d, Mv * d, 0, 2), (n, Mv, m, d, 0, Mv * m, 2), (n, Mv, d, m, Mv * d, 0, 2)]
        rows += [(n, Mv, m, d, Mv * m, 0, 2), (n, Mv, d, m, 0, 0, 2), (n, Mv, d, d, 0, 0, 2), (n, Mv, d, 3 * d, 0, 0, 2)]
    vd, ph, h, f, V = cfg.vision_dim, cfg.proj_hidden, cfg.llm.hidden, cfg.llm.ffn, cfg.llm.vocab
    MP = B * P
    rows += [(1, MP, ph, vd, 0, MP * ph, 2), (1, MP, h, ph, 0, MP * h, 2), (1, MP, h, h, 0, 0, 2)]
    rows += [(1, MP, h, h, MP * h, 0, 2), (1, MP, ph, h, MP * ph, 0, 2), (1, MP, vd, ph, 0, 0, 2)]
    L = text_len + P - 1
    ML, R, nl = B * L, B * supervised_rows, cfg.llm.layers
    rows += [(nl, ML, 3 * h, h, 0, 0, 2), (nl, ML, h, 3 * h, 0, 0, 2)]                               # q|k|v fwd (+RoPE), d(q|k|v) . W
    for cnt, M in ((nl - 1, ML), (1, R)):                                                           # last layer on the supervised rows
        rows += [(cnt, M, h, h, M * h, 0, 2), (cnt, M, 2 * f, h, 0, M * f, 2), (cnt, M, h, f, M * h, 0, 2)]
        # backward: d(act) = dX . W_down with the SwiGLU backward fused (reads gate|up, writes d(gate|up) = 2 x N columns)
        rows += [(cnt, M, f, h, 2 * M * f, M * f, 2), (cnt, M, h, 2 * f, 0, 0, 2), (cnt, M, h, h, M * h, 0, 2)]
    rows += [(1, R, V, h, 0, 0, 4), (1, R, h, V, 0, 0, 2)]
    return rows


def gemm_algorithmic_bytes(cfg: OpenVLAConfig, batch: int, text_len: int, supervised_rows: int) -> int:
    """Bytes the GEMMs of one attack iteration must move if every operand and result crosses HBM exactly once."""
    total = 0
    for cnt, M, N, K, rd, wr, ob in gemm_shape_table(cfg, batch, text_len, supervised_rows):
        total += cnt * (2 * (M * K + N * K + rd + wr) + ob * M * N)
    return int(total)


def config_from_hf(hf_config) -> OpenVLAConfig:
    """Engine configuration for an HF ``OpenVLAConfig`` (prismatic/extern/hf/configuration_prismatic.py:72-140): the LLM fields
    come from ``text_config`` (LlamaConfig), the vocabulary is padded to ``pad_to_multiple_of`` as the HF port does, the vision
    towers are the fused DINOv2-L/14-reg4 + SigLIP-so400m/14 pair (``timm_model_ids`` is checked when present)."""
    import dataclasses
    base = openvla_7b()
    ids = getattr(hf_config, "timm_model_ids", None)
    if ids is not None and list(ids) != ["vit_large_patch14_reg4_dinov2.lvd142m", "vit_so400m_patch14_siglip_224"]:
        raise ValueError(f"unsupported vision backbone {ids}: the engine implements the fused DINOv2 + SigLIP towers of OpenVLA")
    tc = getattr(hf_config, "text_config", None)
    if tc is None:
        return base
    get = (lambda k, d: tc.get(k, d)) if isinstance(tc, dict) else (lambda k, d: getattr(tc, k, d))
    vocab = int(get("vocab_size", base.llm.vocab))
    mult = int(getattr(hf_config, "pad_to_multiple_of", 64) or 1)
    vocab = -(-vocab // mult) * mult
    llm = dataclasses.replace(base.llm, hidden=int(get("hidden_size", base.llm.hidden)), layers=int(get("num_hidden_layers", base.llm.layers)),
                              heads=int(get("num_attention_heads", base.llm.heads)), ffn=int(get("intermediate_size", base.llm.ffn)),
                              vocab=vocab, rms_eps=float(get("rms_norm_eps", base.llm.rms_eps)),
                              rope_theta=float(get("rope_theta", base.llm.rope_theta)))
    return dataclasses.replace(base, llm=llm)
