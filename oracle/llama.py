"""HF transformers 4.40.1 ``LlamaForCausalLM.forward(inputs_embeds, attention_mask, labels)`` restated.

Constructed at ``prismatic/extern/hf/modeling_prismatic.py:248-250`` and called at ``:404-415``; the body is a
third-party dependency (pinned ``transformers==4.40.1``, ``pyproject.toml:50``; 5.5.0 is installed here and is
what ``tests/golden/make_golden.py`` pins this restatement against).  RMSNorm in fp32 cast back before the
weight multiply; RoPE with ``position_ids = arange(L)`` and cos/sin cast to the activation dtype; SDPA with an
additive causal+key-padding mask; SwiGLU MLP; final norm; ``lm_head`` over all rows; ``logits.float()``;
shifted cross-entropy with ``ignore_index=-100`` averaged over non-ignored tokens of the whole batch.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def rms_norm(x, w, eps):
    dt = x.dtype
    h = x.to(torch.float32)
    var = h.pow(2).mean(-1, keepdim=True)
    h = h * torch.rsqrt(var + eps)
    return w * h.to(dt)


def rope_cos_sin(L, head_dim, theta, dtype, device=None):
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    pos = torch.arange(L, dtype=torch.float32)
    freqs = torch.outer(pos, inv_freq)
    emb = torch.cat([freqs, freqs], dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def rotate_half(x):
    x1 = x[..., : x.shape[-1] // 2]
    x2 = x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def llama_hidden(sd, prefix, cfg, inputs_embeds, attention_mask):
    """-> final-norm hidden states [B,L,hidden]."""
    B, L, C = inputs_embeds.shape
    H, hd = cfg.heads, cfg.head_dim
    dt = inputs_embeds.dtype
    cos, sin = rope_cos_sin(L, hd, cfg.rope_theta, dt)
    cos, sin = cos[None, None], sin[None, None]
    causal = torch.ones(L, L, dtype=torch.bool).tril()
    allowed = causal[None, None] & attention_mask.bool()[:, None, None, :]
    mask = torch.zeros(B, 1, L, L, dtype=dt).masked_fill(~allowed, torch.finfo(dt).min)
    h = inputs_embeds
    for i in range(cfg.layers):
        p = f"{prefix}model.layers.{i}."
        r = h
        x = rms_norm(h, sd[p + "input_layernorm.weight"], cfg.rms_eps)
        q = F.linear(x, sd[p + "self_attn.q_proj.weight"]).view(B, L, H, hd).transpose(1, 2)
        k = F.linear(x, sd[p + "self_attn.k_proj.weight"]).view(B, L, H, hd).transpose(1, 2)
        v = F.linear(x, sd[p + "self_attn.v_proj.weight"]).view(B, L, H, hd).transpose(1, 2)
        q = (q * cos) + (rotate_half(q) * sin)
        k = (k * cos) + (rotate_half(k) * sin)
        a = F.scaled_dot_product_attention(q, k, v, attn_mask=mask)
        a = a.transpose(1, 2).reshape(B, L, C)
        h = r + F.linear(a, sd[p + "self_attn.o_proj.weight"])
        r = h
        x = rms_norm(h, sd[p + "post_attention_layernorm.weight"], cfg.rms_eps)
        x = F.linear(F.silu(F.linear(x, sd[p + "mlp.gate_proj.weight"])) * F.linear(x, sd[p + "mlp.up_proj.weight"]),
                     sd[p + "mlp.down_proj.weight"])
        h = r + x
    return rms_norm(h, sd[prefix + "model.norm.weight"], cfg.rms_eps)


def causal_lm_loss(logits, labels):
    """HF shifted CE (``LlamaForCausalLM.forward``): logits fp32 [B,L,V], labels [B,L]."""
    shift_logits = logits[..., :-1, :].contiguous()
    shift_labels = labels[..., 1:].contiguous()
    return F.cross_entropy(shift_logits.view(-1, shift_logits.shape[-1]), shift_labels.view(-1), ignore_index=-100)


def llama_forward(sd, prefix, cfg, inputs_embeds, attention_mask, labels=None):
    h = llama_hidden(sd, prefix, cfg, inputs_embeds, attention_mask)
    logits = F.linear(h, sd[prefix + "lm_head.weight"]).float()
    loss = causal_lm_loss(logits, labels) if labels is not None else None
    return loss, logits
