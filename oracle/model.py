"""``OpenVLAForActionPrediction.forward`` (multimodal branch) restated over a flat HF-named state dict.

Follows ``prismatic/extern/hf/modeling_prismatic.py``: ``PrismaticVisionBackbone.forward`` :114-123 (split 6ch ->
DINOv2 / SigLIP, concat features), ``PrismaticProjector.forward`` :146-158 (fused variant fc1-GELU-fc2-GELU-fc3),
``PrismaticForConditionalGeneration.forward`` :362-415 (embed, splice after BOS, mask / label splice, LLM call),
weight init ``_init_weights`` :185-205.  Parameter names as in the HF checkpoint.

Pinning: tests/test_oracle_golden.py::test_multimodal_forward_matches_reference_model_class compares ``forward`` (vision features,
projector output, loss, logits) with the reference's own ``OpenVLAForActionPrediction.forward`` executed on the CPU
(tests/golden/make_golden_glue.py -> tests/golden/reference_golden_glue.npz).  Test infrastructure: never imported by the product.
"""
from __future__ import annotations

from collections import namedtuple

import torch
import torch.nn.functional as F

from roboticattack_b200.config import IGNORE_INDEX, OpenVLAConfig
from . import llama as _llama
from . import vit as _vit

from roboticattack_b200.weights import DINO, LM, PROJ, SIGLIP, param_shapes  # noqa: F401

Output = namedtuple("Output", ["loss", "logits"])


def vision_backbone(sd, cfg, pixel_values, run_unused_last_block=False):   # :114-123
    img, img_fused = torch.split(pixel_values, [3, 3], dim=1)
    a = _vit.vit_forward(sd, DINO, cfg.dino, img, run_unused_last_block)
    b = _vit.vit_forward(sd, SIGLIP, cfg.siglip, img_fused, run_unused_last_block)
    return torch.cat([a, b], dim=2)


def projector(sd, x):   # :146-158 (fused-backbone variant)
    x = F.gelu(F.linear(x, sd[PROJ + "fc1.weight"], sd[PROJ + "fc1.bias"]))
    x = F.gelu(F.linear(x, sd[PROJ + "fc2.weight"], sd[PROJ + "fc2.bias"]))
    return F.linear(x, sd[PROJ + "fc3.weight"], sd[PROJ + "fc3.bias"])


def splice(sd, cfg, proj, input_ids, attention_mask, labels):   # :380-401
    emb = F.embedding(input_ids, sd[LM + "model.embed_tokens.weight"])
    B, P = proj.shape[:2]
    x = torch.cat([emb[:, :1], proj, emb[:, 1:]], dim=1)
    m = torch.cat([attention_mask[:, :1], torch.ones(B, P, dtype=attention_mask.dtype), attention_mask[:, 1:]], dim=1)
    y = None
    if labels is not None:
        y = torch.cat([labels[:, :1], torch.full((B, P), IGNORE_INDEX, dtype=labels.dtype), labels[:, 1:]], dim=1)
    return x, m, y


def forward(sd, cfg, input_ids, attention_mask, pixel_values, labels=None, run_unused_last_block=False):
    """-> Output(loss, logits fp32 [B, L, V])  (:362-415)."""
    feats = vision_backbone(sd, cfg, pixel_values, run_unused_last_block)
    proj = projector(sd, feats)
    x, m, y = splice(sd, cfg, proj, input_ids, attention_mask, labels)
    loss, logits = _llama.llama_forward(sd, LM, cfg.llm, x, m, y)
    return Output(loss, logits)
