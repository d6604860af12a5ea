"""timm 0.9.10 ``VisionTransformer`` forward as used by ``PrismaticVisionBackbone``
(``prismatic/extern/hf/modeling_prismatic.py:63-123``; model ids ``configuration_prismatic.py:36``).

timm is a third-party dependency absent from /root/reference and from this container (pinned ``timm==0.9.10``,
``pyproject.toml:45``); this file restates its published algorithm: PatchEmbed conv 14x14/14 -> ``_pos_embed``
(``no_embed_class=True``: pos-embed added to patch tokens, then [cls, reg x4] prepended) -> pre-LN blocks
``x = x + ls1(attn(norm1(x))); x = x + ls2(mlp(norm2(x)))`` with LayerNorm eps 1e-6, fused-qkv MHA through
``F.scaled_dot_product_attention`` and an erf-GELU MLP.  ``forward`` is monkey-patched to
``get_intermediate_layers(n={depth-2})``: the output of block depth-2, no final norm, prefix tokens stripped.

Pinning: tests/test_oracle_golden.py checks this restatement against ``transformers.Dinov2WithRegistersModel`` and
``transformers.SiglipVisionModel`` (independent implementations of the two model families) on shared random weights.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def vit_block(sd, p, x, heads, eps, layerscale):
    B, N, C = x.shape
    hd = C // heads
    h = F.layer_norm(x, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps)
    qkv = F.linear(h, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"])
    qkv = qkv.reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    a = F.scaled_dot_product_attention(q, k, v)
    a = a.transpose(1, 2).reshape(B, N, C)
    a = F.linear(a, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
    if layerscale:
        a = a * sd[p + "ls1.scale_factor"]
    x = x + a
    h = F.layer_norm(x, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps)
    h = F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
    h = F.gelu(h)
    h = F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    if layerscale:
        h = h * sd[p + "ls2.scale_factor"]
    return x + h


def vit_forward(sd, prefix, cfg, img, run_unused_last_block=False):
    """img [B,3,H,W] -> [B, num_patches, dim] (second-to-last block output)."""
    x = F.conv2d(img, sd[prefix + "patch_embed.proj.weight"], sd[prefix + "patch_embed.proj.bias"], stride=cfg.patch)
    x = x.flatten(2).transpose(1, 2)
    x = x + sd[prefix + "pos_embed"]
    if cfg.num_prefix:
        B = x.shape[0]
        x = torch.cat([sd[prefix + "cls_token"].expand(B, -1, -1), sd[prefix + "reg_token"].expand(B, -1, -1), x], dim=1)
    out = None
    for i in range(cfg.depth if run_unused_last_block else cfg.depth - 1):
        x = vit_block(sd, f"{prefix}blocks.{i}.", x, cfg.heads, cfg.ln_eps, cfg.layerscale)
        if i == cfg.depth - 2:
            out = x
    return out[:, cfg.num_prefix:]
